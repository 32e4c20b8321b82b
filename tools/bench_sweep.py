"""BASELINE config #5: dense-gradient push/pull bandwidth sweep, 64 KB - 1 GB per
bucket, W = N workers (one per GPU), bucket striped over the N GPUs.

    python tools/bench_sweep.py                                   # N = 1
    python -m torch.distributed.run --nproc-per-node 2 ... tools/bench_sweep.py

Per size and path: >= 20 warm-up and 100 timed rounds (fewer for >= 256 MB), each
bracketed by CUDA events on the worker stream after an L2 flush, median / p10 /
p90 of the max over ranks.  GB/s = W * bytes * 2 / t (push + pull, all workers).
One JSON line per (size, path) on rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tfmesos_b200 import engine, psx  # noqa: E402

SIZES = [64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20, 64 << 20, 256 << 20, 1 << 30]


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sizes = [s for s in SIZES if s <= int(os.environ.get("SWEEP_MAX", 1 << 30))]
    for nbytes in sizes:
        n = nbytes // 4
        for path in ("fused", "staged"):
            cl = engine.TorchrunCluster([("bucket", (n,))], 1, engine.AdamOptimizer(0.01),
                                        fused=(path == "fused"), device=local)
            g = torch.Generator(device="cuda").manual_seed(7 + rank)
            cl.worker.grad_flat[0].copy_(torch.randn(cl.worker.grad_flat[0].numel(),
                                                     device="cuda", generator=g))
            iters = 100 if nbytes < (256 << 20) else 20
            for _ in range(20):
                cl.round(psx.MODE_SUM)
            cl.barrier()
            times = []
            for _ in range(iters):
                with torch.cuda.stream(cl.worker_stream):
                    flush.zero_()
                cl.barrier()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(cl.worker_stream)
                cl.round(psx.MODE_SUM)
                e1.record(cl.worker_stream)
                cl.barrier()
                t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                times.append(t.item())
            times.sort()
            med = times[len(times) // 2]
            if rank == 0:
                print(json.dumps({
                    "bucket_bytes": nbytes, "path": path, "workers": world,
                    "us_median": med * 1e3, "us_p10": times[len(times) // 10] * 1e3,
                    "us_p90": times[(9 * len(times)) // 10] * 1e3,
                    "push_pull_GBps": world * nbytes * 2 / (med * 1e-3) / 1e9,
                    "iters": iters}), flush=True)
            cl.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
