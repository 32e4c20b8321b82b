"""Raw PCIe ceilings for the e2e number: pinned H2D alone, D2H alone, and both at
once (copy engines, 0.8 GB each way = one worker's gradient / parameter set).

    python tools/bench_pcie.py                                    # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_pcie.py

Under torchrun EVERY rank runs the same copies at the same time (barrier before
each repetition), each bound to the cores of its GPU's NUMA node like bench.py:
that is the ceiling of e2e at N GPUs -- GPUs behind one PCIe switch / one socket
share an uplink and the host memory controllers.  Reports the slowest rank's
median, as bench.py's timing does."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

N = 200_200_000


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    try:
        from bench import bind_to_gpu_numa_node
        bind_to_gpu_numa_node(local)
    except Exception:
        pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    h_in = torch.empty(N, dtype=torch.float32).pin_memory()
    h_out = torch.empty(N, dtype=torch.float32).pin_memory()
    d_in = torch.empty(N, dtype=torch.float32, device="cuda")
    d_out = torch.ones(N, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=7):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(t.item())
        ts.sort()
        return ts[len(ts) // 2]

    def h2d():
        d_in.copy_(h_in, non_blocking=True)

    def d2h():
        h_out.copy_(d_out, non_blocking=True)

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    gb = N * 4 / 1e9
    out = {"gpus_active": world}
    for name, fn in (("h2d_alone", h2d), ("d2h_alone", d2h), ("h2d_and_d2h_together", both)):
        ms = timed(fn)
        out[name] = {"ms_slowest_rank_median": ms, "GBps_per_direction_per_gpu": gb / ms * 1e3}
    out["e2e_ceiling_GBps"] = world * 2 * gb / out["h2d_and_d2h_together"]["ms_slowest_rank_median"] * 1e3
    out["note"] = ("e2e moves 0.8 GB each way per rank and step; its floor is the 'together' "
                   "time, times (S+1)/S for the S-shard pipeline fill; e2e_ceiling_GBps is "
                   "bench.py's e2e value at that floor")
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
