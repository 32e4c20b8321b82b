"""The multi-rank parity harness's own pieces that need no GPU: its deterministic
gradient is the same bits in numpy (what the oracle consumes) and in torch (what
the GPU consumes), every case name is unique, and the roles it asks for exist."""
import numpy as np
import torch

from tests import multirank_parity as m


def test_gradient_hash_is_bit_identical_in_numpy_and_torch():
    for seed in (0, 17, m.seed_of(7, 3, 1)):
        a = m.grad_np(5, 200_005, seed)
        b = m.grad_torch(200_005, seed, "cpu").numpy()[5:]
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert np.all(np.abs(a) <= 0.1) and a.std() > 0.03
    # bf16 wire: torch's cast and the oracle's RNE agree on these values
    from oracle import ps_oracle as o
    g = m.grad_torch(4096, 9, "cpu")
    got = g.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(got, o.f32_to_bf16(m.grad_np(0, 4096, 9)))


def test_case_lists_are_unique_and_cover_every_path_and_entry():
    for world in (2, 4, 8):
        cases = m.build_cases({"small", "full", "nvls", "nvlsfull"}, world)
        names = [c.name for c in cases]
        assert len(names) == len(set(names))
        assert {c.path for c in cases} == {"fused", "staged", "nvls"}
        assert {c.entry for c in cases} >= {"round", "timer", "graph", "round_host"}
        assert {c.mode for c in cases} == {"sum", "mean", "async"}
        assert {c.wire for c in cases} == {"f32", "bf16"}
        for c in cases:
            if c.worker_ranks is not None:
                assert max(c.worker_ranks) < world
                flat = [r for rs in c.ps_ranks for r in (rs if isinstance(rs, list) else [rs])]
                assert not set(flat) & set(c.worker_ranks)      # PS GPUs host no worker
            if c.path == "nvls":
                assert c.rtol == (0.0 if world <= 2 else 2e-6) and c.mode != "async"
    assert any(c.ps_ranks for c in m.build_cases({"small"}, 4))
    assert not any(c.ps_ranks for c in m.build_cases({"small"}, 2))
