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
