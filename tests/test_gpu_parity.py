"""GPU parity tests: the CUDA path, called through the C ABI (tfmesos_b200.psx
-> libpsx.so), against the CPU oracle on the same seeded inputs.

Bar: BIT-EXACT for every float32 value (the kernels use one IEEE rounding per
operation, like the oracle), bit-exact for the bf16 wire format, exact integers
for global_step / sequence numbers.  Sizes: the reference's own variable sizes
(7 850 / 79 510 / 400 000 elements) plus ragged and tiny cases;
the BASELINE full sizes (25.5M ResNet-50 bucket, 2e8 NMF W) are checked against
the oracle element for element as well -- the update is elementwise, so the C
oracle covers them in seconds.
"""
import ctypes

import numpy as np
import pytest

from oracle import ps_oracle as o
from tfmesos_b200 import engine, psx

pytestmark = pytest.mark.gpu
F = np.float32

MODES = [psx.MODE_ASYNC_ORDERED, psx.MODE_SUM, psx.MODE_SYNC_MEAN]


@pytest.fixture(scope="module", autouse=True)
def _init():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    psx.init(0)
    yield
    torch.cuda.synchronize()


def _torch():
    import torch
    return torch


def bits(a):
    return np.ascontiguousarray(a, F).view(np.uint32)


def assert_bits_equal(got, want, what):
    """Same bits everywhere; a NaN must meet a NaN (sign/payload of a NaN are not
    defined by IEEE-754 and differ between x86 SSE and the GPU)."""
    g, w = bits(got), bits(want)
    both_nan = np.isnan(np.ascontiguousarray(got, F).ravel()) & \
        np.isnan(np.ascontiguousarray(want, F).ravel())
    bad = np.nonzero((g.ravel() != w.ravel()) & ~both_nan)[0]
    assert bad.size == 0, "%s: %d of %d elements differ, first at %d: got %r want %r" % (
        what, bad.size, g.size, bad[0], got.ravel()[bad[0]], want.ravel()[bad[0]])


def run_rounds(n, opt, mode, W, rounds, wire=psx.F32, seed=0, lr=0.01, scales=None):
    torch = _torch()
    rng = np.random.default_rng(seed)
    shard = psx.Shard(0, n, opt, lr=lr, n_slots=W, wire=wire)
    ref = o.CShard(n, opt, lr=lr)
    init = rng.standard_normal(n).astype(F)
    shard.set_values(psx.VAR, init)
    ref.var[:] = init
    client = [psx.Client(shard.export(), 0, w) for w in range(W)]
    try:
        for r in range(rounds):
            scale = F(scales[r % len(scales)]) if scales else F(10.0 ** rng.integers(-5, 3))
            slots = (rng.standard_normal((W, n)) * scale).astype(F)
            dev = torch.from_numpy(slots).cuda()
            for w in range(W):
                client[w].push(dev[w].data_ptr(), n, seq=r + 1)
            shard.apply(mode, 0, W, wait_seq=r + 1)
            if wire == psx.BF16:
                slots = o.bf16_to_f32(o.f32_to_bf16(slots)).reshape(W, n)
            ref.round(slots, mode)
        torch.cuda.synchronize()
        return shard, ref, client
    except Exception:
        for c in client:
            c.close()
        shard.destroy()
        raise


def check_against(shard, ref, n, opt):
    assert_bits_equal(shard.get_values(psx.VAR), ref.var, "var")
    if opt == psx.OPT_ADAM:
        assert_bits_equal(shard.get_values(psx.M), ref.m, "m")
        assert_bits_equal(shard.get_values(psx.V), ref.v, "v")
    st = shard.state()
    assert st["global_step"] == ref.step
    if opt == psx.OPT_ADAM:
        assert F(st["beta1_power"]) == ref.b1p and F(st["beta2_power"]) == ref.b2p


@pytest.mark.parametrize("opt", [psx.OPT_SGD, psx.OPT_ADAM])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n,W", [(1, 1), (7, 3), (1023, 2), (7850, 2), (79510, 2),
                                 (79510, 5), (400000, 4), (1 << 20, 1)])
def test_apply_bit_exact_vs_oracle(opt, mode, n, W):
    shard, ref, clients = run_rounds(n, opt, mode, W, rounds=4, seed=n + W)
    try:
        check_against(shard, ref, n, opt)
        assert shard.state()["apply_seq"] == 4
    finally:
        for c in clients:
            c.close()
        shard.destroy()


@pytest.mark.parametrize("mode", MODES)
def test_apply_max_slots(mode):
    shard, ref, clients = run_rounds(5000, psx.OPT_ADAM, mode, psx.MAX_SLOTS, rounds=2, seed=99)
    try:
        check_against(shard, ref, 5000, psx.OPT_ADAM)
    finally:
        for c in clients:
            c.close()
        shard.destroy()


@pytest.mark.parametrize("opt", [psx.OPT_SGD, psx.OPT_ADAM])
def test_apply_special_values_and_denormals(opt):
    """zeros (v=0 -> 0/eps), denormal gradients, huge gradients (g*g overflows to
    inf -> update 0/inf), negative zero: same bits as the CPU."""
    torch = _torch()
    n = 4096
    vals = np.array([0.0, -0.0, 1e-45, -1e-40, 1e-38, 1e-20, 1.0, -1.0, 1e19, -3e38,
                     np.float32(2 ** -126), 65504.0], F)
    g = np.resize(vals, n).astype(F)
    init = np.resize(vals[::-1], n).astype(F)
    shard = psx.Shard(0, n, opt, lr=0.01, n_slots=1)
    ref = o.CShard(n, opt, lr=0.01)
    shard.set_values(psx.VAR, init)
    ref.var[:] = init
    c = psx.Client(shard.export(), 0, 0)
    try:
        dev = torch.from_numpy(g).cuda()
        for r in range(3):
            c.push(dev.data_ptr(), n, seq=r + 1)
            shard.apply(psx.MODE_SUM, 0, 1, wait_seq=r + 1)
            with np.errstate(all="ignore"):
                ref.round(g[None, :], o.SUM)
        check_against(shard, ref, n, opt)
    finally:
        c.close()
        shard.destroy()


@pytest.mark.parametrize("mode", MODES)
def test_bf16_wire_slots(mode):
    """BASELINE config #4: gradients pushed as bf16.  The landing slot holds
    RNE-rounded bf16 (checked bit for bit), the apply consumes exactly those."""
    n, W = 30001, 3
    shard, ref, clients = run_rounds(n, psx.OPT_ADAM, mode, W, rounds=3, wire=psx.BF16, seed=4)
    try:
        check_against(shard, ref, n, psx.OPT_ADAM)
    finally:
        for c in clients:
            c.close()
        shard.destroy()


def test_push_converts_f32_to_bf16_like_the_oracle():
    torch = _torch()
    n = 10007
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n)).astype(F)
    x[:4] = [np.inf, -np.inf, 0.0, -0.0]
    shard = psx.Shard(0, n, psx.OPT_SGD, n_slots=1, wire=psx.BF16)
    c = psx.Client(shard.export(), 0, 0)
    try:
        c.push(torch.from_numpy(x).cuda().data_ptr(), n)
        got = shard.get_values(psx.SLOT0)
        assert_bits_equal(got, o.bf16_to_f32(o.f32_to_bf16(x)), "bf16 slot")
        # pull with a bf16 destination
        shard.set_values(psx.VAR, x)
        out = torch.empty(n, dtype=torch.bfloat16, device="cuda")
        c.pull(out.data_ptr(), n, dtype=psx.BF16)
        torch.cuda.synchronize()
        got = out.view(torch.int16).cpu().numpy().view(np.uint16)
        assert np.array_equal(got, o.f32_to_bf16(x))
    finally:
        c.close()
        shard.destroy()


@pytest.mark.parametrize("off,n", [(0, 0), (0, 1), (1, 1), (3, 5), (4, 4096), (5, 4099),
                                   (1, 79509), (32, 1000), (0, 79510)])
def test_push_pull_subranges_ragged_and_unaligned(off, n):
    """One RecvTensor per variable in the reference = one (off, n) range here;
    unaligned ranges take the scalar path and must still be exact."""
    torch = _torch()
    N = 79510
    rng = np.random.default_rng(off + n)
    shard = psx.Shard(0, N, psx.OPT_SGD, n_slots=2)
    c = psx.Client(shard.export(), 0, 1)
    try:
        var = rng.standard_normal(N).astype(F)
        shard.set_values(psx.VAR, var)
        g = rng.standard_normal(max(n, 1)).astype(F)
        dev = torch.from_numpy(g).cuda()
        c.push(dev.data_ptr(), n, off=off)
        slot = shard.get_values(psx.SLOT0 + 1)
        want = np.zeros(N, F)
        want[off:off + n] = g[:n]
        assert_bits_equal(slot, want, "slot")
        assert not shard.get_values(psx.SLOT0).any()        # the other slot is untouched
        out = torch.full((max(n, 1) + 8,), -7.0, device="cuda")
        c.pull(out.data_ptr() + 16, n, off=off)             # 16 B in: still 16 B aligned
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        assert_bits_equal(got[4:4 + n], var[off:off + n], "pulled")
        assert (got[:4] == -7).all() and (got[4 + n:] == -7).all()   # no overrun
        if n > 1:                                           # misaligned destination pointer
            out.fill_(-7.0)
            c.pull(out.data_ptr() + 4, n, off=off)
            torch.cuda.synchronize()
            got = out.cpu().numpy()
            assert_bits_equal(got[1:1 + n], var[off:off + n], "pulled (unaligned dst)")
            assert got[0] == -7 and (got[1 + n:] == -7).all()
    finally:
        c.close()
        shard.destroy()


def test_argument_errors_are_reported_not_crashes():
    shard = psx.Shard(0, 100, psx.OPT_SGD, n_slots=2)
    c = psx.Client(shard.export(), 0, 0)
    try:
        with pytest.raises(RuntimeError, match="outside"):
            c.push(1, 101)
        with pytest.raises(RuntimeError, match="outside"):
            c.pull(1, 50, off=60)
        with pytest.raises(RuntimeError, match="slot range"):
            shard.apply(psx.MODE_SUM, 1, 2)
        with pytest.raises(RuntimeError, match="no region"):
            shard.get_values(psx.M)                         # SGD shard has no m
        with pytest.raises(RuntimeError):
            psx.Client(shard.export(), 0, 5)
        with pytest.raises(RuntimeError, match="empty"):
            psx.Shard(0, 0)
    finally:
        c.close()
        shard.destroy()


@pytest.mark.parametrize("opt", [psx.OPT_SGD, psx.OPT_ADAM])
@pytest.mark.parametrize("mode", MODES)
def test_fused_round_bit_exact_and_scatters_parameters(opt, mode):
    """psx_round: gather from the workers' buffers, reduce, apply, scatter -- the
    same numbers as push + apply + pull, and every worker ends up holding var."""
    torch = _torch()
    n, W = 79510, 3
    rng = np.random.default_rng(7)
    variables = [("hid_w", (784, 100)), ("hid_b", (100,)), ("sm_w", (100, 10)), ("sm_b", (10,))]
    optimizer = (engine.AdamOptimizer(0.01) if opt == psx.OPT_ADAM
                 else engine.GradientDescentOptimizer(0.005))
    cl = engine.LocalCluster(variables, 1, W, optimizer, fused=True)
    nb = cl.layout.bucket_nelem[0]
    ref = o.CShard(nb, opt, lr=optimizer.learning_rate)
    try:
        init = rng.standard_normal(nb).astype(F)
        cl.servers[(0, 0)].shard.set_values(psx.VAR, init)
        ref.var[:] = init
        for r in range(3):
            slots = (rng.standard_normal((W, nb)) * 0.1).astype(F)
            for w in range(W):
                cl.workers[w].grad_flat[0][:nb].copy_(torch.from_numpy(slots[w]))
            cl.round(mode)
            ref.round(slots, mode)
        torch.cuda.synchronize()
        assert_bits_equal(cl.servers[(0, 0)].shard.get_values(psx.VAR), ref.var, "var")
        for w in range(W):
            assert_bits_equal(cl.workers[w].param_flat[0][:nb].cpu().numpy(), ref.var,
                              "worker %d params" % w)
        assert cl.global_step() == ref.step
    finally:
        cl.close()


def test_local_cluster_mlp_two_ps_matches_oracle_per_variable():
    """mnist_replica.py's variables on 2 ps tasks (placement global_step->0,
    hid_w->1, hid_b->0, sm_w->1, sm_b->0), 2 workers, async-ordered Adam, 5
    rounds of seeded gradients: every variable equals the oracle bit for bit."""
    torch = _torch()
    variables = [("global_step", ()), ("hid_w", (784, 100)), ("hid_b", (100,)),
                 ("sm_w", (100, 10)), ("sm_b", (10,))]
    W = 2
    cl = engine.LocalCluster(variables, 2, W, engine.AdamOptimizer(0.01))
    assert dict(cl.layout.placement()) == o.replica_device_setter_placement(
        [n for n, _ in variables], 2)
    refs = [o.CShard(cl.layout.bucket_nelem[t], o.ADAM, lr=0.01) for t in range(2)]
    rng = np.random.default_rng(11)
    try:
        hw = o.truncated_normal(np.random.default_rng(1), (784, 100), 1.0 / 28)
        sw = o.truncated_normal(np.random.default_rng(2), (100, 10), 0.1)
        cl.set_variable("hid_w", hw)
        cl.set_variable("sm_w", sw)
        for name, val in (("hid_w", hw), ("sm_w", sw)):
            task, off, shape, numel = cl.layout.entries[name]
            refs[task].var[off:off + numel] = val.ravel()
        for r in range(5):
            slots = [np.zeros((W, refs[t].n), F) for t in range(2)]
            for w in range(W):
                for name, (task, off, shape, numel) in cl.layout.entries.items():
                    if name == "global_step":
                        continue
                    g = (rng.standard_normal(numel) * 0.05).astype(F)
                    cl.workers[w].grads[name].copy_(torch.from_numpy(g).view(shape))
                    slots[task][w, off:off + numel] = g
            cl.round(psx.MODE_ASYNC_ORDERED)
            for t in range(2):
                refs[t].round(slots[t], o.ASYNC_ORDERED)
        torch.cuda.synchronize()
        for name, (task, off, shape, numel) in cl.layout.entries.items():
            want = refs[task].var[off:off + numel].reshape(shape)
            assert_bits_equal(cl.get_variable(name), want, name)
            for w in range(W):
                assert_bits_equal(cl.workers[w].params[name].cpu().numpy(), want,
                                  "worker %d %s" % (w, name))
        assert cl.global_step() == 10              # async: one step per worker push
    finally:
        cl.close()


def test_striped_bucket_equals_unstriped():
    """Physical striping of one logical bucket over several shards does not
    change a single bit (the update is elementwise)."""
    torch = _torch()
    variables = [("W", (3000, 70))]
    rng = np.random.default_rng(3)
    init = rng.standard_normal((3000, 70)).astype(F)
    grads = (rng.standard_normal((3, 2, 3000, 70)) * 0.1).astype(F)
    results = []
    for devs in ([0], [[0, 0, 0, 0]]):
        cl = engine.LocalCluster(variables, 1, 2, engine.AdamOptimizer(0.01), ps_devices=devs)
        try:
            cl.set_variable("W", init)
            for r in range(3):
                for w in range(2):
                    cl.workers[w].grads["W"].copy_(torch.from_numpy(grads[r, w]))
                cl.round(psx.MODE_SYNC_MEAN)
            torch.cuda.synchronize()
            results.append((cl.get_variable("W"), cl.workers[1].params["W"].cpu().numpy(),
                            len(cl.topo.shards)))
        finally:
            cl.close()
    assert results[0][2] == 1 and results[1][2] == 4
    assert_bits_equal(results[1][0], results[0][0], "striped var")
    assert_bits_equal(results[1][1], results[0][1], "striped worker params")


@pytest.mark.parametrize("n,opt", [(25_557_032, psx.OPT_SGD), (25_557_032, psx.OPT_ADAM),
                                   (200_000_000, psx.OPT_ADAM)])
def test_full_size_buckets_element_for_element(n, opt):
    """BASELINE sizes: ResNet-50's 25 557 032 parameters and the scaled NMF's
    W = 1e6 x 200.  Elementwise update -> the C oracle checks all of it."""
    torch = _torch()
    W = 2
    lr = 0.01
    shard = psx.Shard(0, n, opt, lr=lr, n_slots=W)
    clients = [psx.Client(shard.export(), 0, w) for w in range(W)]
    try:
        gen = torch.Generator(device="cuda").manual_seed(n % 1000)
        init = torch.randn(n, device="cuda", generator=gen)
        grads = torch.randn(W, n, device="cuda", generator=gen) * 0.01
        torch.cuda.synchronize()
        shard.set_values(psx.VAR, init.cpu().numpy())
        ref = o.CShard(n, opt, lr=lr)
        ref.var[:] = init.cpu().numpy()
        for r in range(2):
            for w in range(W):
                clients[w].push(grads[w].data_ptr(), n, seq=r + 1)
            shard.apply(psx.MODE_SUM, 0, W, wait_seq=r + 1)
            ref.round(grads.cpu().numpy(), o.SUM)
        out = torch.empty(n, device="cuda")
        clients[0].pull(out.data_ptr(), n, wait_seq=0)
        torch.cuda.synchronize()
        assert_bits_equal(out.cpu().numpy(), ref.var, "pulled var")
        check_against(shard, ref, n, opt)
        # size-independent properties: lr = 0 leaves var untouched; a second pull is idempotent
        shard.set_hyper(0.0)
        before = shard.get_values(psx.VAR)
        shard.apply(psx.MODE_SUM, 0, W)
        if opt == psx.OPT_SGD:
            assert_bits_equal(shard.get_values(psx.VAR), before, "lr=0")
        out2 = torch.empty(n, device="cuda")
        clients[1].pull(out2.data_ptr(), n)
        torch.cuda.synchronize()
        assert torch.equal(out2, torch.from_numpy(shard.get_values(psx.VAR)).cuda())
    finally:
        for c in clients:
            c.close()
        shard.destroy()


def test_training_parity_mnist_mlp_with_device_gradients():
    """End to end on the MNIST-replica workload (mnist_replica.py:124-157):
    gradients computed on the GPU by torch autograd (TF32 off), PS rounds on the
    CUDA path, against the all-CPU oracle run (numpy gradients + oracle Adam) on
    the same seeded synthetic batches.  Float tolerance (gradients differ in
    summation order): rtol 2e-3 on the parameters after 20 global steps."""
    torch = _torch()
    torch.backends.cuda.matmul.allow_tf32 = False
    variables = [("global_step", ()), ("hid_w", (784, 100)), ("hid_b", (100,)),
                 ("sm_w", (100, 10)), ("sm_b", (10,))]
    W = 2
    cl = engine.LocalCluster(variables, 1, W, engine.AdamOptimizer(0.01))
    names = ["hid_w", "hid_b", "sm_w", "sm_b"]
    try:
        ref = {"hid_w": o.truncated_normal(np.random.default_rng(1), (784, 100), 1.0 / 28),
               "hid_b": np.zeros(100, F),
               "sm_w": o.truncated_normal(np.random.default_rng(2), (100, 10), 0.1),
               "sm_b": np.zeros(10, F)}
        ref_shard = o.Shard(cl.layout.bucket_nelem[0], o.ADAM, lr=0.01)
        for k in names:
            cl.set_variable(k, ref[k])
            _, off, _, numel = cl.layout.entries[k]
            ref_shard.var[off:off + numel] = ref[k].ravel()
        for w in cl.workers:
            w.pull()
        rngs = [np.random.default_rng(1234 + w) for w in range(W)]
        for r in range(10):
            slots = np.zeros((W, ref_shard.n), F)
            for w in range(W):
                x = rngs[w].random((100, 784)).astype(F)
                y = np.eye(10, dtype=F)[rngs[w].integers(0, 10, 100)]
                # oracle side
                cur = {}
                for k in names:
                    _, off, shape, numel = cl.layout.entries[k]
                    cur[k] = ref_shard.var[off:off + numel].reshape(shape)
                _, *gs = o.mlp_grads(cur["hid_w"], cur["hid_b"], cur["sm_w"], cur["sm_b"], x, y)
                for k, g in zip(names, gs):
                    _, off, _, numel = cl.layout.entries[k]
                    slots[w, off:off + numel] = g.ravel()
                # device side
                wk = cl.workers[w]
                ps = [wk.params[k].detach().clone().requires_grad_(True) for k in names]
                tx, ty = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
                h = torch.relu(tx @ ps[0] + ps[1])
                p = torch.softmax(h @ ps[2] + ps[3], 1)
                loss = -(ty * torch.log(torch.clamp(p, 1e-10, 1.0))).sum()
                loss.backward()
                for k, t in zip(names, ps):
                    wk.grads[k].copy_(t.grad)
            cl.round(psx.MODE_ASYNC_ORDERED)
            ref_shard.round(slots, o.ASYNC_ORDERED)
        torch.cuda.synchronize()
        assert cl.global_step() == 20 == ref_shard.step
        for k in names:
            _, off, shape, numel = cl.layout.entries[k]
            want = ref_shard.var[off:off + numel].reshape(shape)
            np.testing.assert_allclose(cl.get_variable(k), want, rtol=2e-3, atol=2e-4)
    finally:
        cl.close()


def test_launch_counter_counts_this_librarys_kernels():
    torch = _torch()
    shard = psx.Shard(0, 4096, psx.OPT_SGD, n_slots=1)
    c = psx.Client(shard.export(), 0, 0)
    shard.register_client(0, c.export())       # the pull below waits on this client's mirror
    try:
        g = torch.zeros(4096, device="cuda")
        before = psx.launch_count()
        c.push(g.data_ptr(), 4096, seq=1)
        shard.apply(psx.MODE_SUM, 0, 1, wait_seq=1)
        c.pull(g.data_ptr(), 4096, wait_seq=1)
        torch.cuda.synchronize()
        assert psx.launch_count() - before == 3
    finally:
        c.close()
        shard.destroy()


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("mode", [psx.MODE_ASYNC_ORDERED, psx.MODE_SUM])
def test_bf16_wire_end_to_end_f32_master(fused, mode):
    """BASELINE config #4: workers hold bf16 gradients and bf16 parameters, the PS
    keeps the f32 master (var, m, v).  The master equals the oracle fed with the
    bf16-rounded gradients bit for bit; what the workers receive is exactly
    RNE-bf16(master)."""
    torch = _torch()
    n, W = 50000, 3
    variables = [("flat", (n,))]
    cl = engine.LocalCluster(variables, 1, W, engine.AdamOptimizer(0.01), wire=psx.BF16,
                             fused=fused)
    ref = o.CShard(n, o.ADAM, lr=0.01)
    rng = np.random.default_rng(13)
    try:
        init = rng.standard_normal(n).astype(F)
        cl.set_variable("flat", init)
        ref.var[:] = init
        assert cl.workers[0].grad_flat[0].dtype == torch.bfloat16
        for r in range(3):
            g32 = (rng.standard_normal((W, n)) * 0.1).astype(F)
            gb = o.f32_to_bf16(g32).reshape(W, n)
            for w in range(W):
                t = torch.from_numpy(gb[w].view(np.int16)).cuda().view(torch.bfloat16)
                cl.workers[w].grad_flat[0][:n].copy_(t)
            cl.round(mode)
            ref.round(o.bf16_to_f32(gb).reshape(W, n), mode)
        torch.cuda.synchronize()
        assert_bits_equal(cl.get_variable("flat"), ref.var, "f32 master")
        want = o.f32_to_bf16(ref.var)
        for w in range(W):
            got = cl.workers[w].param_flat[0][:n].view(torch.int16).cpu().numpy().view(np.uint16)
            assert np.array_equal(got, want), "worker %d bf16 parameters" % w
    finally:
        cl.close()


@pytest.mark.parametrize("fused", [False, True])
def test_torchrun_cluster_single_rank_batched_and_stepwise_rounds(fused):
    """engine.TorchrunCluster at world size 1 (what bench.py drives): a round
    issued as ONE psx_batch call and a round issued call by call (the timed
    variant) give the same bits as the oracle."""
    torch = _torch()

    class NullTimer(object):
        def start(self, s):
            pass

        def stop(self, s):
            pass

    variables = [("W", (3000, 40)), ("H", (40, 500))]
    cl = engine.TorchrunCluster(variables, 2, engine.AdamOptimizer(0.01),
                                placement={"W": 0, "H": 1}, stripes=3, fused=fused, device=0)
    refs = [o.CShard(cl.layout.bucket_nelem[t], o.ADAM, lr=0.01) for t in range(2)]
    rng = np.random.default_rng(17)
    try:
        for r in range(4):
            slots = []
            for t in range(2):
                g = (rng.standard_normal(refs[t].n) * 0.1).astype(F)
                cl.worker.grad_flat[t][:refs[t].n].copy_(torch.from_numpy(g))
                slots.append(g)
            torch.cuda.synchronize()
            cl.round(psx.MODE_SUM, NullTimer() if r % 2 else None)
            for t in range(2):
                refs[t].round(slots[t][None, :], o.SUM)
        cl.barrier()
        for t in range(2):
            got = cl.worker.param_flat[t][:refs[t].n].cpu().numpy()
            assert_bits_equal(got, refs[t].var, "ps task %d" % t)
    finally:
        cl.close()


def test_shard_larger_than_2_pow_31_elements():
    """Maximum sizes: a shard of 2^31 + 4099 elements (8.6 GB per region) -- every
    index computation is 64-bit.  SGD, one worker, one round, checked against the
    oracle over the WHOLE range in 256 Mi-element windows (elementwise update, so
    windows are independent), including the window that straddles 2^31."""
    torch = _torch()
    n = (1 << 31) + 4099
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * (1 << 30):
        pytest.skip("needs ~35 GB of free HBM")
    lr = 0.5
    shard = psx.Shard(0, n, psx.OPT_SGD, lr=lr, n_slots=1)
    c = psx.Client(shard.export(), 0, 0)
    try:
        gen = torch.Generator(device="cuda").manual_seed(31)
        init = torch.randn(n, device="cuda", generator=gen)
        grad = torch.randn(n, device="cuda", generator=gen)
        psx.copy(0, shard.ptr(psx.VAR), init.data_ptr(), n * 4)
        c.push(grad.data_ptr(), n, seq=1)
        shard.apply(psx.MODE_SUM, 0, 1, wait_seq=1)
        torch.cuda.synchronize()
        lib = o.c_lib()
        win = 1 << 28
        for lo in range(0, n, win):
            cnt = min(win, n - lo)
            want = init[lo:lo + cnt].cpu().numpy()
            g = grad[lo:lo + cnt].cpu().numpy()
            lib.psx_oracle_sgd(want.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                               g.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), cnt, lr)
            got = shard.get_values(psx.VAR, lo, cnt)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "window at %d" % lo
        assert shard.state()["global_step"] == 1
    finally:
        c.close()
        shard.destroy()


@pytest.mark.parametrize("opt", [psx.OPT_SGD, psx.OPT_ADAM])
def test_row_block_data_parallel_partial_applies(opt):
    """Row-block data parallelism over an embedding-like variable (the scaled NMF's
    W: each worker holds a block of rows of R, so its dW is non-zero only in its own
    rows): every worker pushes and pulls ONLY its block, the PS runs one
    psx_apply_range per block with that block's owner as the single slot, and only
    the last launch closes the round.  Equals ONE oracle apply of the assembled
    gradient, bit for bit; beta powers / global_step advance once."""
    torch = _torch()
    rows, rank, W = 1003, 36, 4                      # ragged: 1003 rows over 4 workers
    n = rows * rank
    bounds = [(rows * k // W) * rank for k in range(W + 1)]
    shard = psx.Shard(0, n, opt, lr=0.01, n_slots=W)
    ref = o.CShard(n, opt, lr=0.01)
    rng = np.random.default_rng(23)
    init = rng.standard_normal(n).astype(F)
    shard.set_values(psx.VAR, init)
    ref.var[:] = init
    clients = [psx.Client(shard.export(), 0, w) for w in range(W)]
    for w, c in enumerate(clients):
        shard.register_client(w, c.export())
    try:
        for r in range(1, 4):
            full = (rng.standard_normal(n) * 0.1).astype(F)
            dev = torch.from_numpy(full).cuda()
            for w in range(W):
                lo, hi = bounds[w], bounds[w + 1]
                clients[w].push(dev.data_ptr() + lo * 4, hi - lo, off=lo, seq=r)
            for w in range(W):
                lo, hi = bounds[w], bounds[w + 1]
                shard.apply_range(psx.MODE_SUM, w, 1, lo, hi - lo, finish=(w == W - 1),
                                  wait_seq=r)
            ref.round(full[None, :], o.SUM)
        out = torch.zeros(n, device="cuda")
        for w in range(W):
            lo, hi = bounds[w], bounds[w + 1]
            clients[w].pull(out.data_ptr() + lo * 4, hi - lo, off=lo, wait_seq=3)
        torch.cuda.synchronize()
        check_against(shard, ref, n, opt)
        assert_bits_equal(out.cpu().numpy(), ref.var, "pulled blocks")
        assert shard.state()["apply_seq"] == 3 and shard.state()["global_step"] == 3
    finally:
        for c in clients:
            c.close()
        shard.destroy()


@pytest.mark.parametrize("fused", [False, True])
def test_captured_round_replays_bit_exact(fused):
    """A PS round captured in a CUDA graph (counted rendez-vous: constant stream-wait
    values, consumed counters) and replayed N times equals N oracle rounds."""
    torch = _torch()
    variables = [("w", (500, 64)), ("b", (64,))]
    cl = engine.TorchrunCluster(variables, 1, engine.AdamOptimizer(0.01), stripes=2,
                                fused=fused, device=0)
    n = cl.layout.bucket_nelem[0]
    ref = o.CShard(n, o.ADAM, lr=0.01)
    rng = np.random.default_rng(29)
    try:
        g = (rng.standard_normal(n) * 0.1).astype(F)
        cl.worker.grad_flat[0][:n].copy_(torch.from_numpy(g))
        torch.cuda.synchronize()
        cl.round(psx.MODE_SUM)                       # one ordinary round first
        ref.round(g[None, :], o.SUM)
        try:
            graph = cl.capture_round(psx.MODE_SUM)
        except Exception as exc:                     # capture of stream memops unsupported
            pytest.skip("stream capture of the round failed: %s" % str(exc)[:200])
        for _ in range(5):
            with torch.cuda.stream(cl.worker_stream):
                graph.replay()
            ref.round(g[None, :], o.SUM)
        cl.round(psx.MODE_SUM)                       # and ordinary rounds still work after
        ref.round(g[None, :], o.SUM)
        cl.barrier()
        got = cl.worker.param_flat[0][:n].cpu().numpy()
        assert_bits_equal(got, ref.var, "params after replays")
        st = next(iter(cl.servers.values())).shard.state()
        assert st["global_step"] == ref.step == 7
    finally:
        cl.close()
