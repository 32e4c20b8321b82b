"""Checkpoint / restore of PS shards: a run resumed from a checkpoint continues
bit-identically to the run that was never interrupted."""
import numpy as np
import pytest

from tfmesos_b200 import checkpoint, engine, psx

pytestmark = pytest.mark.gpu
F = np.float32
VARS = [("hid_w", (784, 100)), ("hid_b", (100,)), ("sm_w", (100, 10)), ("sm_b", (10,))]


def _rounds(cl, rng, k):
    import torch
    for _ in range(k):
        for w in cl.workers:
            for name, (task, off, shape, numel) in cl.layout.entries.items():
                w.grads[name].copy_(torch.from_numpy(
                    (rng.standard_normal(numel) * 0.1).astype(F)).view(shape))
        cl.round(psx.MODE_ASYNC_ORDERED)
    torch.cuda.synchronize()


def _snapshot(cl):
    out = {}
    for key, ps in cl.servers.items():
        out[key] = (ps.shard.get_values(psx.VAR), ps.shard.get_values(psx.M),
                    ps.shard.get_values(psx.V), ps.shard.state())
    return out


@pytest.mark.parametrize("stripes", [1, 3])
def test_resume_is_bit_identical(tmp_path, stripes):
    psx.init(0)
    devs = [[0] * stripes, [0] * stripes]
    a = engine.LocalCluster(VARS, 2, 2, engine.AdamOptimizer(0.01), ps_devices=devs)
    try:
        a.set_variable("hid_w", np.random.default_rng(1).standard_normal((784, 100)).astype(F))
        _rounds(a, np.random.default_rng(5), 3)
        path = str(tmp_path / "ckpt")
        checkpoint.save(a, path)
        _rounds(a, np.random.default_rng(6), 2)
        want = _snapshot(a)
    finally:
        a.close()
    b = engine.LocalCluster(VARS, 2, 2, engine.AdamOptimizer(0.01), ps_devices=devs)
    try:
        checkpoint.restore(b, path)
        assert b.global_step() == 6                        # 3 rounds x 2 async workers
        for w in b.workers:
            w.pull()
        _rounds(b, np.random.default_rng(6), 2)
        got = _snapshot(b)
    finally:
        b.close()
    assert sorted(got) == sorted(want)
    for key in want:
        for i in range(3):
            assert np.array_equal(got[key][i].view(np.uint32), want[key][i].view(np.uint32)), (key, i)
        g, w = got[key][3], want[key][3]
        assert g["global_step"] == w["global_step"] == 10
        assert F(g["beta1_power"]) == F(w["beta1_power"]) and F(g["beta2_power"]) == F(w["beta2_power"])


def test_restore_rejects_a_different_layout(tmp_path):
    psx.init(0)
    a = engine.LocalCluster(VARS, 1, 1, engine.AdamOptimizer(0.01))
    path = str(tmp_path / "ckpt")
    try:
        checkpoint.save(a, path)
    finally:
        a.close()
    b = engine.LocalCluster(VARS[:2], 1, 1, engine.AdamOptimizer(0.01))
    try:
        with pytest.raises(RuntimeError, match="does not match"):
            checkpoint.restore(b, path)
    finally:
        b.close()


def test_parameter_client_save_restore_through_endpoints(tmp_path):
    """The chief checkpoints the PS tasks through their endpoints and a fresh
    cluster restores them (Supervisor logdir stand-in)."""
    import threading

    import torch

    from tfmesos_b200 import train as tf

    def start(ports):
        spec = {"ps": ["127.0.0.1:%d" % p for p in ports], "worker": ["127.0.0.1:1"]}
        servers = [tf.Server(spec, "ps", i) for i in range(len(ports))]
        for s in servers:
            t = threading.Thread(target=s.join)
            t.daemon = True
            t.start()
        return spec, servers

    import socket

    def free_ports(n):
        socks = [socket.socket() for _ in range(n)]
        for s in socks:
            s.bind(("127.0.0.1", 0))
        ports = [s.getsockname()[1] for s in socks]
        for s in socks:
            s.close()
        return ports

    variables = [("global_step", ()), ("w", (300, 20)), ("b", (20,))]
    path = str(tmp_path / "model")
    spec, servers = start(free_ports(2))
    try:
        sess = tf.ParameterClient(spec, variables, tf.AdamOptimizer(0.01), 0, device=0,
                                  init={"w": np.ones((300, 20), F)})
        for k in range(3):
            sess.grads["w"].fill_(0.5 + k)
            sess.grads["b"].fill_(-1.0)
            sess.minimize()
        files = sess.save(path)
        assert len(files) == 2
        want = {k: sess.read(k) for k in ("w", "b")}
        step = sess.global_step()
        sess.close()
    finally:
        for s in servers:
            s.endpoint.stop_event.set()
    spec2, servers2 = start(free_ports(2))
    try:
        sess2 = tf.ParameterClient(spec2, variables, tf.AdamOptimizer(0.01), 0, device=0)
        assert not np.array_equal(sess2.read("w"), want["w"])
        sess2.restore(path)
        assert sess2.global_step() == step == 3
        for k in ("w", "b"):
            assert np.array_equal(sess2.read(k).view(np.uint32), want[k].view(np.uint32))
            assert np.array_equal(sess2.params[k].cpu().numpy().view(np.uint32),
                                  want[k].view(np.uint32))
        sess2.close()
    finally:
        for s in servers2:
            s.endpoint.stop_event.set()
