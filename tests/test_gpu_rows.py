"""Index-list (IndexedSlices) sparse push for embedding-like variables (SURVEY
8f-3; the NMF row blocks of examples/matrix_factorization.py:21-28,43-49): rows
pushed by several workers are merged in worker order on the PS, the optimizer is
applied once per touched row -- bit for bit the oracle's ``rows_round``."""
import numpy as np
import pytest

from oracle import ps_oracle as o
from tfmesos_b200 import psx

pytestmark = pytest.mark.gpu
F = np.float32


def _setup(opt, lr, n_rows, d, W, wire=psx.F32):
    import torch
    psx.init(0)
    shard = psx.Shard(0, n_rows * d, opt, lr=lr, n_slots=W, wire=wire)
    init = np.random.default_rng(11).standard_normal(n_rows * d).astype(F)
    shard.set_values(psx.VAR, init)
    clients = [psx.Client(shard.export(), 0, w) for w in range(W)]
    ref = o.Shard(n_rows * d, o.ADAM if opt == psx.OPT_ADAM else o.SGD, lr=lr)
    ref.var[:] = init
    return torch, shard, clients, ref


@pytest.mark.parametrize("opt", [psx.OPT_SGD, psx.OPT_ADAM])
@pytest.mark.parametrize("mode", [psx.MODE_SUM, psx.MODE_SYNC_MEAN])
@pytest.mark.parametrize("d", [200, 7])
def test_rows_from_three_workers_with_overlaps_match_oracle(opt, mode, d):
    n_rows, W, R = 4096, 3, 3
    torch, shard, clients, ref = _setup(opt, 0.05 if opt == psx.OPT_SGD else 0.01, n_rows, d, W)
    omode = o.SUM if mode == psx.MODE_SUM else o.SYNC_MEAN
    rng = np.random.default_rng(d)
    try:
        for r in range(1, R + 1):
            idx_lists, row_lists, keep = [], [], []
            for w in range(W):
                k = int(rng.integers(1, 900))
                idx = np.sort(rng.choice(n_rows, size=k, replace=False)).astype(np.int64)
                if w == 2:
                    idx = np.unique(np.concatenate([idx, idx_lists[0][:50], idx_lists[1][-30:]]))
                rows = (rng.standard_normal((idx.size, d)) * 0.1).astype(F)
                idx_lists.append(idx)
                row_lists.append(rows)
                ti, tr = torch.from_numpy(idx).cuda(), torch.from_numpy(rows).cuda()
                keep.append((ti, tr))
                clients[w].push_rows(ti.data_ptr(), tr.data_ptr(), idx.size, d, psx.F32, r)
            shard.apply_rows(mode, 0, W, d, wait_seq=r)
            torch.cuda.synchronize()
            o.rows_round(ref, d, idx_lists, row_lists, omode)
        got = shard.get_values(psx.VAR)
        assert np.array_equal(got.view(np.uint32), ref.var.view(np.uint32))
        if opt == psx.OPT_ADAM:
            assert np.array_equal(shard.get_values(psx.M).view(np.uint32), ref.m.view(np.uint32))
            assert np.array_equal(shard.get_values(psx.V).view(np.uint32), ref.v.view(np.uint32))
        st = shard.state()
        assert st["global_step"] == R == ref.step
        if opt == psx.OPT_ADAM:
            assert F(st["beta1_power"]) == ref.b1p and F(st["beta2_power"]) == ref.b2p
    finally:
        for c in clients:
            c.close()
        shard.destroy()


def test_nmf_row_blocks_as_index_lists_equal_the_dense_round():
    """Each worker owns a contiguous row block of W (the NMF data-parallel split):
    pushing the blocks as index lists gives exactly what ONE dense SUM round of
    the assembled gradient gives (disjoint rows: every sum has a single term)."""
    n_rows, d, W = 3000, 200, 4
    torch, shard, clients, ref = _setup(psx.OPT_ADAM, 0.01, n_rows, d, W)
    dense = o.CShard(n_rows * d, o.ADAM, lr=0.01)
    dense.var[:] = ref.var
    rng = np.random.default_rng(3)
    try:
        for r in range(1, 3):
            full = (rng.standard_normal((n_rows, d)) * 0.1).astype(F)
            keep = []
            for w in range(W):
                lo, hi = w * n_rows // W, (w + 1) * n_rows // W
                ti = torch.arange(lo, hi, dtype=torch.int64, device="cuda")
                tr = torch.from_numpy(full[lo:hi]).cuda()
                keep.append((ti, tr))
                clients[w].push_rows(ti.data_ptr(), tr.data_ptr(), hi - lo, d, psx.F32, r)
            shard.apply_rows(psx.MODE_SUM, 0, W, d, wait_seq=r)
            torch.cuda.synchronize()
            dense.round(full.reshape(1, -1), o.SUM)
        assert np.array_equal(shard.get_values(psx.VAR).view(np.uint32), dense.var.view(np.uint32))
        assert np.array_equal(shard.get_values(psx.V).view(np.uint32), dense.v.view(np.uint32))
    finally:
        for c in clients:
            c.close()
        shard.destroy()


def test_rows_that_do_not_fit_or_do_not_divide_are_refused():
    torch, shard, clients, _ = _setup(psx.OPT_SGD, 0.1, 64, 8, 1)
    try:
        ti = torch.arange(0, 64, dtype=torch.int64, device="cuda")
        tr = torch.zeros(64, 8, device="cuda")
        with pytest.raises(RuntimeError, match="does not divide"):
            clients[0].push_rows(ti.data_ptr(), tr.data_ptr(), 4, 7, psx.F32, 1)
        with pytest.raises(RuntimeError, match="do not fit"):
            clients[0].push_rows(ti.data_ptr(), tr.data_ptr(), 10_000, 8, psx.F32, 1)
    finally:
        clients[0].close()
        shard.destroy()
