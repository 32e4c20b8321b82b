"""Request-free serving (psx_serve_start): the PS consumes pushes as they ARRIVE,
with no host request per step -- the reference's default async discipline
(examples/mnist/mnist_replica.py:198-205) and SyncReplicasOptimizer
(mnist_replica.py:109-113,148-162) on the device.  Checked bit for bit against
the oracle wherever the schedule is deterministic.

PS and workers share ONE process here, so the workers wait on the host
(``Client.wait_host``): a stream wait on a shard served by the same process is
refused by the library (it could deadlock behind a shared hardware channel).  The
stream-wait form runs where workers are separate processes, as in the reference:
tests/test_gpu_examples.py (tfrun)."""
import time

import numpy as np
import pytest

from oracle import ps_oracle as o
from tfmesos_b200 import psx

pytestmark = pytest.mark.gpu
F = np.float32
N = 79510


def _grad(w, r, n=N):
    return (np.random.default_rng(100 * w + r).standard_normal(n) * 0.1).astype(F)


class _Rig(object):
    def __init__(self, W, opt, lr, n=N, wire=psx.F32):
        import torch
        psx.init(0)
        self.torch = torch
        self.W, self.n = W, n
        self.shard = psx.Shard(0, n, opt, lr=lr, n_slots=W, wire=wire)
        self.init = np.random.default_rng(5).standard_normal(n).astype(F)
        self.shard.set_values(psx.VAR, self.init)
        self.clients = [psx.Client(self.shard.export(), 0, w) for w in range(W)]
        for w, c in enumerate(self.clients):
            self.shard.register_client(w, c.export())
        self.streams = [torch.cuda.Stream(device=0) for _ in range(W)]
        self.grads = [torch.zeros(n, device="cuda") for _ in range(W)]
        self.params = [torch.zeros(n, device="cuda") for _ in range(W)]
        self.step_host = [torch.zeros(1, dtype=torch.int64).pin_memory() for _ in range(W)]

    def push(self, w, seq, g, stamp=0):
        with self.torch.cuda.stream(self.streams[w]):
            self.grads[w].copy_(self.torch.from_numpy(g), non_blocking=False)
        self.clients[w].push_stamped(self.grads[w].data_ptr(), self.n, 0, psx.F32, seq, stamp,
                                     self.streams[w])

    def close(self):
        self.shard.serve_stop()
        for w, c in enumerate(self.clients):
            self.shard.unregister_client(w)
            c.close()
        self.shard.destroy()


def test_served_async_applies_each_push_on_arrival_bit_exact():
    """3 workers; pushes issued one at a time in a scrambled order, each waited for
    (stream-wait on the worker's own client block, no request to the PS): the
    result is the oracle's async schedule in exactly that order, the mirrored
    global_step is the one each apply produced."""
    rig = _Rig(3, psx.OPT_ADAM, 0.01)
    ref = o.CShard(N, o.ADAM, lr=0.01)
    ref.var[:] = rig.init
    try:
        rig.shard.serve_start(psx.MODE_ASYNC_ORDERED)
        seqs = [0, 0, 0]
        order = [2, 0, 1, 1, 2, 0, 0, 2, 1]
        for k, w in enumerate(order):
            seqs[w] += 1
            g = _grad(w, seqs[w])
            rig.push(w, seqs[w], g)
            c, st = rig.clients[w], rig.streams[w]
            st.synchronize()
            c.wait_host("applied", seqs[w])
            c.pull(rig.params[w].data_ptr(), N, 0, psx.F32, 0, st)
            c.read_step_async(rig.step_host[w].data_ptr(), st)
            st.synchronize()
            ref.round(g[None, :], o.ASYNC_ORDERED)
            assert int(rig.step_host[w][0]) == k + 1 == ref.step
            got = rig.params[w].cpu().numpy()
            assert np.array_equal(got.view(np.uint32), ref.var.view(np.uint32)), (k, w)
        stats = rig.shard.serve_stats()
        assert stats["served"] == len(order) and stats["global_step"] == len(order)
        # the host accessors pause and resume the loop
        assert np.array_equal(rig.shard.get_values(psx.M).view(np.uint32), ref.m.view(np.uint32))
        assert np.array_equal(rig.shard.get_values(psx.V).view(np.uint32), ref.v.view(np.uint32))
        st = rig.shard.state()
        assert st["global_step"] == len(order)
        assert F(st["beta1_power"]) == ref.b1p and F(st["beta2_power"]) == ref.b2p
        # ... and it keeps serving afterwards
        g = _grad(1, 99)
        rig.push(1, seqs[1] + 1, g)
        rig.streams[1].synchronize()
        rig.clients[1].wait_host("applied", seqs[1] + 1)
        ref.round(g[None, :], o.ASYNC_ORDERED)
        assert np.array_equal(rig.shard.get_values(psx.VAR).view(np.uint32), ref.var.view(np.uint32))
    finally:
        rig.close()


def test_served_async_concurrent_pushes_are_all_consumed():
    """Free-running: 4 workers push at will from their own streams; whatever the
    interleaving, every push is applied exactly once (SGD: the sum of the steps is
    order-independent up to rounding)."""
    W, R = 4, 6
    rig = _Rig(W, psx.OPT_SGD, 0.05)
    try:
        rig.shard.serve_start(psx.MODE_ASYNC_ORDERED, idle_sleep_us=50)
        total = np.zeros(N, np.float64)
        steps = [[] for _ in range(W)]
        for r in range(1, R + 1):
            for w in range(W):
                g = _grad(w, r)
                total += g.astype(np.float64) * 0.05
                rig.push(w, r, g)
            for w in range(W):
                rig.streams[w].synchronize()
                steps[w].append(rig.clients[w].wait_host("applied", r)["global_step"])
        stats = rig.shard.serve_stats()
        assert stats["served"] == W * R and stats["global_step"] == W * R
        flat = sorted(s for per in steps for s in per)
        assert flat == list(range(1, W * R + 1))          # every apply produced its own step
        got = rig.shard.get_values(psx.VAR)
        np.testing.assert_allclose(got, rig.init - total, rtol=0, atol=2e-5)
    finally:
        rig.close()


def _wait_stats(shard, key, value, timeout=20.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        st = shard.serve_stats()
        if st[key] >= value:
            return st
        time.sleep(0.01)
    raise AssertionError("%s never reached %d: %r" % (key, value, shard.serve_stats()))


def test_served_sync_replicas_first_two_of_three_by_arrival_late_one_dropped():
    """SyncReplicas on the device, replicas_to_aggregate = 2 of 3 workers: each
    round the two gradients that ARRIVE first are averaged (which two changes from
    round to round), the third -- pushed after the round's apply, so stamped with
    an older global_step -- is dropped as stale; every worker gets its token."""
    W, R = 3, 5
    rig = _Rig(W, psx.OPT_ADAM, 0.01)
    ref = o.CShard(N, o.ADAM, lr=0.01)
    ref.var[:] = rig.init
    try:
        rig.shard.serve_start(psx.MODE_SYNC_MEAN, replicas_to_aggregate=2)
        for r in range(1, R + 1):
            late = r % W                                   # a different straggler each round
            early = [w for w in range(W) if w != late]
            grads = {w: _grad(w, r) for w in range(W)}
            for w in reversed(early):                      # arrival order != slot order
                rig.push(w, r, grads[w], stamp=r - 1)
                rig.streams[w].synchronize()               # it HAS arrived
            for w in early:
                rig.clients[w].wait_host("tokens", r)
                rig.clients[w].pull(rig.params[w].data_ptr(), N, 0, psx.F32, 0, rig.streams[w])
                rig.streams[w].synchronize()
            ref.round(np.stack([grads[w] for w in early]), o.SYNC_MEAN)
            for w in early:
                got = rig.params[w].cpu().numpy()
                assert np.array_equal(got.view(np.uint32), ref.var.view(np.uint32)), (r, w)
            # the straggler pushes a gradient computed at global_step r-1: stale now
            rig.push(late, r, grads[late], stamp=r - 1)
            rig.streams[late].synchronize()
            assert rig.clients[late].poll()["tokens"] >= r            # its token is there already
            st = _wait_stats(rig.shard, "dropped", r)
            assert st["dropped"] == r and st["global_step"] == r and st["served"] == 2 * r
        assert np.array_equal(rig.shard.get_values(psx.VAR).view(np.uint32),
                              ref.var.view(np.uint32))
        assert np.array_equal(rig.shard.get_values(psx.M).view(np.uint32), ref.m.view(np.uint32))
        assert rig.shard.state()["global_step"] == R == ref.step
    finally:
        rig.close()


def test_served_sync_all_replicas_is_the_oracle_mean_and_bf16_wire():
    """replicas_to_aggregate = W over a bf16 wire: every round is the oracle's
    SYNC_MEAN over all slots (f32 master on the PS)."""
    import torch
    W, R = 2, 4
    rig = _Rig(W, psx.OPT_ADAM, 0.01, wire=psx.BF16)
    ref = o.CShard(N, o.ADAM, lr=0.01)
    ref.var[:] = rig.init
    try:
        rig.shard.serve_start(psx.MODE_SYNC_MEAN, replicas_to_aggregate=W)
        for r in range(1, R + 1):
            slots = []
            for w in range(W):
                g = _grad(w, r)
                rig.push(w, r, g, stamp=r - 1)             # f32 source, cast to bf16 by the push
                slots.append(o.bf16_to_f32(o.f32_to_bf16(g)))
            for w in range(W):
                rig.streams[w].synchronize()
                rig.clients[w].wait_host("tokens", r)
            ref.round(np.stack(slots), o.SYNC_MEAN)
        assert np.array_equal(rig.shard.get_values(psx.VAR).view(np.uint32),
                              ref.var.view(np.uint32))
        assert rig.shard.serve_stats()["dropped"] == 0
        del torch
    finally:
        rig.close()


def test_serve_stop_keeps_pushes_that_arrive_while_stopped():
    """A push that lands while the loop is stopped (accessor pause, checkpoint) stays
    flagged and counted; the restarted loop applies it."""
    rig = _Rig(2, psx.OPT_SGD, 0.05)
    try:
        rig.shard.serve_start(psx.MODE_ASYNC_ORDERED)
        rig.shard.serve_stop()
        g = _grad(0, 1)
        rig.push(0, 1, g)
        rig.streams[0].synchronize()
        assert rig.shard.serve_stats()["served"] == 0
        rig.shard.serve_start(psx.MODE_ASYNC_ORDERED)
        rig.clients[0].wait_host("applied", 1)
        with pytest.raises(RuntimeError, match="serves the shard itself"):
            rig.clients[0].wait_applied(1, rig.streams[0])      # refused, not a hang
        ref = o.CShard(N, o.SGD, lr=0.05)
        ref.var[:] = rig.init
        ref.round(g[None, :], o.ASYNC_ORDERED)
        assert np.array_equal(rig.shard.get_values(psx.VAR).view(np.uint32), ref.var.view(np.uint32))
    finally:
        rig.close()


def test_accessors_from_several_threads_pause_the_loop_one_at_a_time():
    """Endpoint handler threads call state() / get_values() concurrently while the
    shard is served (every worker asks for global_step at the end of training): the
    pause / resume of the loop is serialised per shard, nothing is lost."""
    import threading
    rig = _Rig(2, psx.OPT_SGD, 0.05)
    errors = []
    try:
        rig.shard.serve_start(psx.MODE_ASYNC_ORDERED, idle_sleep_us=20)

        def hammer():
            try:
                for _ in range(8):
                    st = rig.shard.state()
                    assert st["global_step"] >= 0
                    rig.shard.get_values(psx.VAR, 0, 16)
            except Exception as exc:           # noqa: BLE001 - reported below
                errors.append(repr(exc))

        threads = [threading.Thread(target=hammer) for _ in range(4)]
        for t in threads:
            t.start()
        total = np.zeros(N, np.float64)
        for r in range(1, 9):
            for w in range(2):
                g = _grad(w, r)
                total += g.astype(np.float64) * 0.05
                rig.push(w, r, g)
            for w in range(2):
                rig.streams[w].synchronize()
                rig.clients[w].wait_host("applied", r)
        for t in threads:
            t.join(60)
        assert not errors, errors
        assert rig.shard.serve_stats()["served"] == 16
        np.testing.assert_allclose(rig.shard.get_values(psx.VAR), rig.init - total, rtol=0, atol=2e-5)
    finally:
        rig.close()
