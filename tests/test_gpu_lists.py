"""Bucketed tensor lists (BASELINE config #4): one launch moves a whole list of
separately allocated tensors into / out of the flat PS bucket, TMA-staged or
with plain vector loads; byte-exact either way, including ragged and unaligned
tensors, and identical to the per-variable push/pull."""
import numpy as np
import pytest

from oracle import ps_oracle as o
from tfmesos_b200 import engine, psx

pytestmark = pytest.mark.gpu
F = np.float32


def resnet_like_shapes():
    # the size mix of torchvision resnet50: many tiny (64..2048) and a few huge tensors
    shapes = [("conv1", (64, 3, 7, 7)), ("bn1.w", (64,)), ("bn1.b", (64,))]
    for i, (c_in, c) in enumerate([(64, 64), (256, 128), (512, 256), (1024, 512)]):
        shapes += [("l%d.c1" % i, (c, c_in, 1, 1)), ("l%d.b1" % i, (c,)),
                   ("l%d.c2" % i, (c, c, 3, 3)), ("l%d.b2" % i, (c,)),
                   ("l%d.c3" % i, (4 * c, c, 1, 1)), ("l%d.b3" % i, (4 * c,))]
    shapes += [("fc.w", (1000, 2048)), ("fc.b", (1000,)), ("odd", (1237,)), ("one", (1,)),
               ("three", (3,))]
    return shapes


@pytest.mark.parametrize("tma", [True, False])
@pytest.mark.parametrize("stripes", [1, 3])
def test_list_push_apply_pull_matches_oracle(tma, stripes):
    import torch
    psx.init(0)
    shapes = resnet_like_shapes()
    cl = engine.LocalCluster(shapes, 1, 2, engine.GradientDescentOptimizer(0.1),
                             ps_devices=[[0] * stripes])
    nb = cl.layout.bucket_nelem[0]
    ref = o.CShard(nb, o.SGD, lr=0.1)
    rng = np.random.default_rng(8)
    try:
        grads, params, binds_g, binds_p = [], [], [], []
        for w in cl.workers:
            g = {n: torch.zeros(s, device="cuda") for n, s in shapes}
            p = {n: torch.full(s, -3.0, device="cuda") for n, s in shapes}
            grads.append(g)
            params.append(p)
            binds_g.append(engine.TensorListBinding(w, g))
            binds_p.append(engine.TensorListBinding(w, p))
        init = rng.standard_normal(nb).astype(F)
        for name, (task, off, shape, numel) in cl.layout.entries.items():
            cl.set_variable(name, init[off:off + numel].reshape(shape))
        ref.var[:] = 0
        for name, (task, off, shape, numel) in cl.layout.entries.items():
            ref.var[off:off + numel] = init[off:off + numel]
        for r in range(1, 3):
            slots = np.zeros((2, nb), F)
            for w in range(2):
                for name, (task, off, shape, numel) in cl.layout.entries.items():
                    g = rng.standard_normal(numel).astype(F)
                    slots[w, off:off + numel] = g
                    grads[w][name].copy_(torch.from_numpy(g).view(shape))
                binds_g[w].push(seq=r, tma=tma)
            for ps in cl.servers.values():
                ps.apply(psx.MODE_SUM, wait_seq=r)
            for w in range(2):
                binds_p[w].pull(wait_seq=r, tma=tma)
            ref.round(slots, o.SUM)
        torch.cuda.synchronize()
        for name, (task, off, shape, numel) in cl.layout.entries.items():
            want = ref.var[off:off + numel].reshape(shape)
            assert np.array_equal(cl.get_variable(name).view(np.uint32), want.view(np.uint32)), name
            for w in range(2):
                got = params[w][name].cpu().numpy()
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (w, name)
        for b in binds_g + binds_p:
            b.close()
    finally:
        cl.close()


@pytest.mark.parametrize("tma", [True, False])
def test_list_with_unaligned_tensor_addresses(tma):
    """Views that start 4 bytes into an allocation are not TMA-able: they take
    the element-wise path inside the same launch."""
    import torch
    psx.init(0)
    shapes = [("a", (5000,)), ("b", (4099,)), ("c", (70000,))]
    cl = engine.LocalCluster(shapes, 1, 1, engine.GradientDescentOptimizer(1.0))
    try:
        big = torch.arange(0, 100000, device="cuda", dtype=torch.float32)
        tensors = {"a": big[1:5001], "b": big[6000:10099], "c": big[20001:90001]}
        bind = engine.TensorListBinding(cl.workers[0], tensors)
        bind.push(seq=1, tma=tma)
        torch.cuda.synchronize()
        slot = cl.servers[(0, 0)].shard.get_values(psx.SLOT0)
        for name, (task, off, shape, numel) in cl.layout.entries.items():
            assert np.array_equal(slot[off:off + numel], tensors[name].cpu().numpy()), name
        # gaps between variables stay zero
        mask = np.ones(slot.size, bool)
        for name, (task, off, shape, numel) in cl.layout.entries.items():
            mask[off:off + numel] = False
        assert not slot[mask].any()
        bind.close()
    finally:
        cl.close()


def test_list_equals_per_variable_push():
    import torch
    psx.init(0)
    shapes = resnet_like_shapes()
    cl = engine.LocalCluster(shapes, 1, 2, engine.GradientDescentOptimizer(0.1))
    try:
        rng = np.random.default_rng(2)
        g = {}
        for name, (task, off, shape, numel) in cl.layout.entries.items():
            v = rng.standard_normal(numel).astype(F)
            g[name] = torch.from_numpy(v).view(shape).cuda()
            cl.workers[1].grads[name].copy_(g[name])
        bind = engine.TensorListBinding(cl.workers[0], g)
        bind.push(seq=1)                       # worker 0: list path
        cl.workers[1].push(seq=1)              # worker 1: flat per-bucket path
        torch.cuda.synchronize()
        sh = cl.servers[(0, 0)].shard
        assert np.array_equal(sh.get_values(psx.SLOT0), sh.get_values(psx.SLOT0 + 1))
        bind.close()
    finally:
        cl.close()


@pytest.mark.parametrize("to_shard", [True, False])
def test_ragged_tensor_first_in_a_list_larger_than_the_ring(to_shard):
    """Plain (ragged / unaligned) chunks interleaved with TMA chunks in a list big
    enough that every CTA owns many more chunks than the 4 ring stages: the ring
    stage and mbarrier phase must follow the count of TMA chunks, not the chunk
    index (with j-indexed phases this returned stale shared memory or
    hung).  72 MB list: ragged tensors first, between and after the large ones."""
    import torch
    psx.init(0)
    shapes = [("ragged_first", (4099,)), ("big0", (6_000_000,)), ("odd_mid", (1237,)),
              ("big1", (6_000_002,)), ("three", (3,)), ("big2", (6_000_001,)), ("tail", (7,))]
    cl = engine.LocalCluster(shapes, 1, 1, engine.GradientDescentOptimizer(0.1))
    try:
        gen = torch.Generator(device="cuda").manual_seed(3)
        src = {n: torch.randn(s, device="cuda", generator=gen) for n, s in shapes}
        if to_shard:
            bind = engine.TensorListBinding(cl.workers[0], src)
            bind.push(seq=1, tma=True)
            torch.cuda.synchronize()
            slot = cl.servers[(0, 0)].shard.get_values(psx.SLOT0)
            for name, (task, off, shape, numel) in cl.layout.entries.items():
                want = src[name].cpu().numpy().ravel()
                assert np.array_equal(slot[off:off + numel].view(np.uint32),
                                      want.view(np.uint32)), name
        else:
            for name, t in src.items():
                cl.set_variable(name, t.cpu().numpy())
            dst = {n: torch.full(s, -7.0, device="cuda") for n, s in shapes}
            bind = engine.TensorListBinding(cl.workers[0], dst)
            bind.pull(tma=True)
            torch.cuda.synchronize()
            for name in src:
                assert torch.equal(dst[name], src[name]), name
        bind.close()
    finally:
        cl.close()
