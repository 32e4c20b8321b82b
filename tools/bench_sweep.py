"""BASELINE config #5: dense-gradient push/pull bandwidth sweep, 64 KB - 1 GB per
bucket, at 1/2/4/8 GPUs.

    python tools/bench_sweep.py                                   # N = 1
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_sweep.py [args]

    --topology striped   W = N workers (one per GPU), bucket striped over the N GPUs
    --topology incast    W workers -> ONE PS: the PS shard alone on rank 0, workers
                         on ranks 1..W for every W in --incast-workers that fits
                         (SURVEY 8d cfg #5: W in {1, 2, 4, 7} against 1 PS)
    --paths fused,staged,nvls

Per size and path: >= 20 warm-up and 100 timed rounds (20 for >= 256 MB), each
bracketed by CUDA events after an L2 flush, median / p10 / p90 of the max over
ranks.  GB/s = W * bytes * 2 / t (push + pull, all workers).  One JSON line per
(size, path, W) on rank 0."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tfmesos_b200 import engine, psx  # noqa: E402

SIZES = [64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20, 64 << 20, 256 << 20, 1 << 30]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--topology", default="striped", choices=["striped", "incast"])
    ap.add_argument("--paths", default="fused,staged")
    ap.add_argument("--incast-workers", default="1,2,4,7")
    ap.add_argument("--max-bytes", type=int, default=int(os.environ.get("SWEEP_MAX", 1 << 30)))
    ap.add_argument("--min-bytes", type=int, default=0)
    ap.add_argument("--sizes", default=None, help="comma list of bucket sizes in bytes "
                    "(default: the 8 sizes 64 KB .. 1 GB)")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--wire", default="f32", choices=["f32", "bf16"])
    args = ap.parse_args()
    for k in ("paths", "incast_workers", "sizes"):     # ":" works as the list separator too
        if getattr(args, k):
            setattr(args, k, getattr(args, k).replace(":", ","))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    wire = psx.BF16 if args.wire == "bf16" else psx.F32
    esz = 2 if args.wire == "bf16" else 4
    if args.topology == "striped":
        shapes = [(None, None, world)]
    else:
        shapes = [([0], list(range(1, w + 1)), w)
                  for w in (int(x) for x in args.incast_workers.split(",")) if w + 1 <= world]
    sizes = [int(x) for x in args.sizes.split(",")] if args.sizes else SIZES
    for nbytes in [s for s in sizes if args.min_bytes <= s <= args.max_bytes]:
        n = nbytes // esz
        for ps_ranks, worker_ranks, W in shapes:
            for path in args.paths.split(","):
                if path == "nvls" and (world < 2 or args.topology != "striped" or wire != psx.F32):
                    continue
                cl = engine.TorchrunCluster([("bucket", (n,))], 1, engine.AdamOptimizer(0.01),
                                            path=path, device=local, wire=wire,
                                            ps_ranks=ps_ranks, worker_ranks=worker_ranks)
                if cl.worker is not None:
                    g = cl.worker.grad_flat[0]
                    g.copy_((torch.rand(g.numel(), device="cuda") - 0.5).to(g.dtype))
                st = cl.worker_stream if cl.worker is not None else cl.ps_stream
                iters = args.iters if nbytes < (256 << 20) else max(10, args.iters // 5)
                for _ in range(20):
                    cl.round(psx.MODE_SUM)
                cl.barrier()
                times = []
                for _ in range(iters):
                    with torch.cuda.stream(st):
                        flush.zero_()
                    cl.barrier()
                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(st)
                    cl.round(psx.MODE_SUM)
                    e1.record(st)
                    cl.barrier()
                    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64)
                    if world > 1:
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    times.append(t.item())
                times.sort()
                med = times[len(times) // 2]
                if rank == 0:
                    print(json.dumps({
                        "bucket_bytes": nbytes, "path": path, "topology": args.topology,
                        "gpus": world, "workers": W, "ps_gpus": 1 if ps_ranks else world,
                        "wire": args.wire,
                        "us_median": med * 1e3, "us_p10": times[len(times) // 10] * 1e3,
                        "us_p90": times[(9 * len(times)) // 10] * 1e3,
                        "push_pull_GBps": W * nbytes * 2 / (med * 1e-3) / 1e9,
                        "ps_port_GBps_per_direction": (W * nbytes / (med * 1e-3) / 1e9
                                                       if ps_ranks else None),
                        "iters": iters}), flush=True)
                cl.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
