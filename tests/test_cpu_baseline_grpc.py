"""The gRPC-transport CPU baseline (oracle/cpu_ps_grpc.py) computes exactly what
the oracle computes, one RPC per variable per direction."""
import numpy as np

from oracle import cpu_ps_grpc
from oracle import ps_oracle as o

F = np.float32


def test_grpc_cpu_ps_matches_oracle_bit_for_bit():
    rng = np.random.default_rng(4)
    sizes = {"hid_w": 78400, "hid_b": 100, "sm_w": 1000, "sm_b": 10}
    ps = cpu_ps_grpc.GrpcCpuPs(sizes, opt_adam=True, lr=0.01)
    refs = {k: o.Shard(n, o.ADAM, lr=0.01) for k, n in sizes.items()}
    try:
        for k, n in sizes.items():
            init = rng.standard_normal(n).astype(F)
            ps.assign(k, init)
            refs[k].var[:] = init
        for _ in range(3):
            grads = {k: (rng.standard_normal(n) * 0.1).astype(F) for k, n in sizes.items()}
            got = ps.step(grads)
            for k in sizes:
                refs[k].round(grads[k][None, :], o.ASYNC_ORDERED)
                assert np.array_equal(got[k].view(np.uint32), refs[k].var.view(np.uint32)), k
    finally:
        ps.close()


def test_time_round_reports_a_rate():
    r = cpu_ps_grpc.time_round(200000, steps=2, warmup=1)
    assert r["value"] > 0 and r["unit"] == "GB/s" and "gRPC" in r["sample"]
