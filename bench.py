#!/usr/bin/env python
"""bench.py -- PS push+pull GB/s and steps/sec of the B200 parameter-server path.

    python bench.py [--gpus N --steps K --warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                      # the CPU-PS path

Workload (BASELINE.json configs[2], the one the GB/s metric is meaningful on):
the scaled matrix-factorization parameter set -- W (1e6 x 200) on ps:0 and
H (200 x 1e3) on ps:1 (placement of examples/matrix_factorization.py:21-28),
f32, Adam fused into the PS reduction -- one PS round per step: every worker
PUSHes its dense gradient, the PS reduces + applies, every worker PULLs the new
parameters.  Gradients are synthetic (the model's fwd/bwd is outside the named
path).  One process per GPU; rank r is worker r and hosts the shards pinned to
GPU r (each bucket striped over all N GPUs unless --stripes says otherwise).

    value  = n_workers * n_params * 8 B / t_step      ("push+pull GB/s")
    e2e    = the same through TorchrunCluster.round_host(): gradients start in
             pinned HOST memory and parameters are read back to the host, both
             copies inside the timed region.
A second, small section times MNIST-replica training steps (mnist_replica.py's
MLP, fwd/bwd on the GPU by torch, Adam on the PS) and reports steps/sec.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NMF_ROWS, NMF_COLS, NMF_RANK = 1_000_000, 1_000, 200
WORKLOADS = {
    # name: (variables, ps_tasks, placement)
    "nmf_scaled": ([("W", (NMF_ROWS, NMF_RANK)), ("H", (NMF_RANK, NMF_COLS))], 2,
                   {"W": 0, "H": 1}),
    "nmf_reference": ([("W", (1000, 200)), ("H", (200, 1000))], 2, {"W": 0, "H": 1}),
    "resnet50_bucket": ([("flat", (25_557_032,))], 1, None),
    "mnist_mlp": ([("global_step", ()), ("hid_w", (784, 100)), ("hid_b", (100,)),
                   ("sm_w", (100, 10)), ("sm_b", (10,))], 1, None),
    "mnist_softmax": ([("W", (784, 10)), ("b", (10,)), ("global_step", ())], 1, None),
    # the other reading of BASELINE config #3: the PARAMETER is 1e6 x 1e3 (4 GB f32)
    "embedding_1e6x1e3": ([("P", (1_000_000, 1_000))], 1, None),
}
MODES = {"sum": 1, "async": 0, "mean": 2}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="nmf_scaled", choices=sorted(WORKLOADS))
    p.add_argument("--mode", default="sum", choices=sorted(MODES))
    p.add_argument("--path", default="fused", choices=["staged", "fused"],
                   help="staged = push kernel + apply kernel + pull kernel; "
                        "fused = one PS-side gather/apply/scatter kernel (psx_round)")
    p.add_argument("--stripes", type=int, default=None,
                   help="GPUs each bucket is striped over (default: all)")
    p.add_argument("--wire", default="f32", choices=["f32", "bf16"],
                   help="element type of the workers' gradient/parameter tensors "
                        "(the PS master is always f32; bf16 = BASELINE config #4)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-staged", action="store_true")
    p.add_argument("--no-mnist", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-sample-elems", type=int, default=100_000_000,
                   help="CPU arms: parameters per step (divided by the worker count)")
    return p.parse_args()


def n_params(workload):
    total = 0
    for _, shape in WORKLOADS[workload][0]:
        k = 1
        for d in shape:
            k *= d
        total += k
    return total


def bind_to_gpu_numa_node(index):
    """Run this rank on the cores next to its GPU so that the pinned staging
    buffers (first touch) and the copy-engine traffic stay on the GPU's socket."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


# ------------------------------------------------------------------ clocks ----
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.thread = [], None, None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read)
        self.thread.daemon = True
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, cell in zip(names, r[3:7]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class KernelTimer(object):
    """CUDA-event pairs around ONE kernel's launches, on its launching stream."""

    def __init__(self):
        self.pairs, self._cur = [], None

    def start(self, stream):
        import torch
        self._cur = torch.cuda.Event(enable_timing=True)
        self._cur.record(stream)

    def stop(self, stream):
        import torch
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        self.pairs.append((self._cur, e))

    def mean_ms(self):
        if not self.pairs:
            return None
        return sum(a.elapsed_time(b) for a, b in self.pairs) / len(self.pairs)


# --------------------------------------------------------------- CPU arms ----
def cpu_ps(args, steps, warmup):
    """The reference's CPU-PS path, best case (oracle/ps_oracle.c
    psx_oracle_cpu_ps_round: memcpy push, Eigen-style threaded apply on the PS
    host cores, memcpy pull), on a bounded sample of the workload."""
    from oracle import ps_oracle as o
    n_full = n_params(args.workload)
    W = max(1, args.gpus)           # N GPUs <-> N workers, as on the CUDA arm
    n = min(n_full, max(1_000_000, args.cpu_sample_elems // W))
    base = o.CpuPsBaseline(n, W, o.ADAM, lr=0.01)
    mode = {"sum": o.SUM, "async": o.ASYNC_ORDERED, "mean": o.SYNC_MEAN}[args.mode]
    threads = 0
    for _ in range(warmup):
        threads = base.round(mode)
    t0 = time.perf_counter()
    for _ in range(steps):
        threads = base.round(mode)
    dt = (time.perf_counter() - t0) / max(1, steps)
    gbs = W * n * 8 / dt / 1e9
    sample = ("%d of %d parameters (%.1f%%), %d worker, memcpy transport, %d rounds"
              % (n, n_full, 100.0 * n / n_full, W, steps))
    return {"value": gbs, "unit": "GB/s", "cores": int(threads), "kind": "port",
            "sample": sample, "ms_per_step_on_sample": dt * 1e3,
            "host_cores_online": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    cb = cpu_ps(args, steps, warmup)
    n_full = n_params(args.workload)
    line = {
        "impl": "reference",
        "metric": "ps_push_pull_GBps", "value": cb["value"], "unit": "GB/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": cb["ms_per_step_on_sample"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1, n_full),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the TF-0.12 PS path (TensorFlow itself cannot run "
                "here); the timed step is the bounded sample named in cpu_baseline.sample",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world, n_full):
    return {"workload": "%s: PS round (push+reduce/apply+pull) over %d f32 parameters, Adam(lr=0.01)"
                        % (args.workload, n_full),
            "discipline": args.mode, "path": args.path, "wire": args.wire,
            "parallelism": "%d workers (1/GPU), %d ps tasks, buckets striped over %s GPU(s)"
                           % (world, WORKLOADS[args.workload][1],
                              args.stripes if args.stripes else world),
            "l2": "per-step inputs (%.0f MB of gradients per worker) exceed the 126 MB L2"
                  % (n_full * 4 / 1e6) if n_full * 4 > 126e6 else "L2 flushed between steps"}


# ---------------------------------------------------------------- GPU arm ----
def mnist_section(torch, engine, psx, world, rank, dist, model="mlp", steps=100, warmup=10):
    """Full training steps on the reference's two MNIST models, synthetic
    [100,784] batches, forward/backward by torch on the worker GPU, one PS round
    per step with async-ordered applies (one global step per worker push):
      mlp      examples/mnist/mnist_replica.py:124-157  784-100-10, Adam(0.01)
      softmax  examples/mnist/mnist.py:44-55            784-10, SGD(0.005)
    Returns global steps/sec."""
    if model == "mlp":
        variables, ps_tasks, placement = WORKLOADS["mnist_mlp"]
        optimizer, names = engine.AdamOptimizer(0.01), ["hid_w", "hid_b", "sm_w", "sm_b"]
    else:
        variables, ps_tasks, placement = WORKLOADS["mnist_softmax"]
        optimizer, names = engine.GradientDescentOptimizer(0.005), ["W", "b"]
    cl = engine.TorchrunCluster(variables, ps_tasks, optimizer, placement=placement, stripes=1)
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    if rank == 0 and model == "mlp":
        import numpy as np
        rng = np.random.default_rng(1)
        cl.set_variable("hid_w", np.clip(rng.standard_normal((784, 100)), -2, 2) / 28)
        cl.set_variable("sm_w", np.clip(rng.standard_normal((100, 10)), -2, 2) / 10)
    cl.barrier()
    x = torch.rand(100, 784, device="cuda", generator=g)
    y = torch.nn.functional.one_hot(torch.randint(0, 10, (100,), device="cuda", generator=g),
                                    10).float()
    wk = cl.worker
    ws = cl.worker_stream

    def fwd_bwd():
        x.uniform_(0.0, 1.0)                    # a fresh synthetic batch every step
        ps = [wk.params[k].detach().requires_grad_(True) for k in names]
        if model == "mlp":
            h = torch.relu(x @ ps[0] + ps[1])
            p = torch.softmax(h @ ps[2] + ps[3], 1)
            loss = -(y * torch.log(torch.clamp(p, 1e-10, 1.0))).sum()
        else:
            p = torch.softmax(x @ ps[0] + ps[1], 1)
            loss = -(y * torch.log(p)).sum()
        grads = torch.autograd.grad(loss, ps)
        for k, gr in zip(names, grads):
            wk.grads[k].copy_(gr)

    graph = None

    whole = None

    def step():
        if whole is not None:
            with torch.cuda.stream(ws):
                whole.replay()
            return
        with torch.cuda.stream(ws):
            if graph is not None:
                graph.replay()
            else:
                fwd_bwd()
        cl.round(psx.MODE_ASYNC_ORDERED)

    with torch.cuda.stream(ws):
        wk.pull(0, ws)
    for _ in range(warmup):
        step()
    cl.barrier()
    # the worker's compute is launch-bound (20-odd tiny kernels): capture it once
    # in a CUDA graph; the PS round stays ordinary launches (its sequence numbers
    # change every step)
    whole = None
    try:
        # the whole step -- forward/backward AND the PS round (signal/push, counted
        # stream waits, apply, pull) -- as ONE graph: a replay is a training step
        whole = cl.capture_round(psx.MODE_ASYNC_ORDERED, pre=fwd_bwd)
    except Exception as exc:
        sys.stderr.write("mnist: whole-step graph capture failed (%s)\n" % str(exc)[:300])
        whole = None
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
    if whole is None:
        try:
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_, stream=ws):
                fwd_bwd()
            graph = g_
        except Exception as exc:                # keep measuring, eagerly
            sys.stderr.write("mnist: CUDA graph capture failed (%s), running eagerly\n" % exc)
            graph = None
    for _ in range(warmup):
        step()
    cl.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ws)
    for _ in range(steps):
        step()
    ev1.record(ws)
    cl.barrier()
    dt = ev0.elapsed_time(ev1) * 1e-3
    t = torch.tensor([dt], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n_par = n_params("mnist_mlp" if model == "mlp" else "mnist_softmax")
    out = {"global_steps_per_sec": steps * world / t.item(), "rounds": steps,
           "workers": world, "params": n_par,
           "model": ("784-100-10 MLP, batch 100, Adam 0.01, async-ordered" if model == "mlp"
                     else "784-10 softmax regression, batch 100, SGD 0.005, async-ordered"),
           "worker_compute": ("whole step (fwd/bwd + PS round) in one CUDA graph"
                              if whole is not None else
                              "CUDA graph" if graph is not None else "eager"),
           "push_pull_GBps": steps * world * n_par * 8 / t.item() / 1e9,
           "timing": "CUDA events on the worker stream over %d rounds, max over ranks" % steps}
    cl.close()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from tfmesos_b200 import engine, psx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    bind_to_gpu_numa_node(local_rank)
    if world > 1:
        # plumbing only (handle exchange, barriers, max-over-ranks): gloo over
        # loopback -- nothing on the measured path uses a collective library
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    mode = MODES[args.mode]
    variables, ps_tasks, placement = WORKLOADS[args.workload]
    n_full = n_params(args.workload)

    wire = psx.BF16 if args.wire == "bf16" else psx.F32
    esz = 2 if args.wire == "bf16" else 4
    cl = engine.TorchrunCluster(variables, ps_tasks, engine.AdamOptimizer(0.01),
                                placement=placement, stripes=args.stripes,
                                fused=(args.path == "fused"), wire=wire, device=local_rank)
    gen = torch.Generator(device="cuda").manual_seed(100 + rank)
    for t in cl.worker.grad_flat:
        t.copy_(torch.randn(t.numel(), device="cuda", generator=gen) * 1e-2)
    torch.cuda.synchronize()

    # L2 flush buffer for workloads smaller than L2
    flush = None
    if n_full * 4 <= 126e6:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def one_step(timer=None, host=False):
        if flush is not None:
            with torch.cuda.stream(cl.worker_stream):
                flush.zero_()
        if host:
            cl.round_host(mode)
        else:
            cl.round(mode, timer)

    def timed(n_steps, host=False, timer=None):
        cl.barrier()
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        launches0 = psx.launch_count()
        ev0.record(cl.worker_stream)
        for _ in range(n_steps):
            one_step(timer, host)
        ev1.record(cl.worker_stream)
        cl.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64)
        launches = torch.tensor([psx.launch_count() - launches0], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(launches, op=dist.ReduceOp.SUM)
        return ms.item() / n_steps, int(launches.item())

    for _ in range(warmup):
        one_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    timer = KernelTimer()
    ms_step, launches = timed(steps, timer=timer)
    bytes_step = world * n_full * 2 * esz       # W * N * (s_g + s_p)
    value = bytes_step / (ms_step * 1e-3) / 1e9

    # dominant kernel: the fused reduce+apply (or gather/apply/scatter) kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    dom = cl.dominant
    shard_elems = ((dom.spec.nelem + 1023) // 1024) * 1024 if dom is not None else 0
    k_ms = timer.mean_ms()
    if args.path == "fused":
        per_elem = 24 + esz * world + esz * world  # var/m/v r+w, W gradient reads, W param writes
        kname = "k_apply<ADAM,%s,SCATTER,PeerSrc<%s>>" % (args.mode, args.wire)
    else:
        per_elem = 24 + esz * world                # var/m/v r+w, W landing-slot reads
        kname = "k_apply<ADAM,%s,SlotSrc<%s>>" % (args.mode, args.wire)
    roofline = None
    if k_ms:
        # the timer brackets the launches over this rank's largest shard only
        bytes_per_launch = per_elem * shard_elems
        achieved = bytes_per_launch / (k_ms * 1e-3) / 1e9
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            traffic = tr.get("%s/%s/n%d" % (args.workload, args.path, world))
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": hbm_peak,
                    "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                    "peak_source": peak_src, "avg_launch_ms": k_ms,
                    "algorithmic_bytes_per_elem": per_elem,
                    "algorithmic_bytes_per_launch": bytes_per_launch,
                    "launches_timed": len(timer.pairs)}
        if world > 1 and args.path == "fused":
            # the one-shot kernel is bound by this GPU's NVLink port, not by HBM:
            # per direction it carries (N-1) remote gradient stripes in + the
            # other owners' parameter stripes in (and the mirror image out)
            nvl_bytes = 2 * (world - 1) * shard_elems * esz
            nvl = nvl_bytes / (k_ms * 1e-3) / 1e9
            roofline.update({"bound": "nvlink", "achieved": nvl, "peak": 770.0,
                             "frac": nvl / 770.0,
                             "peak_source": "measured peer copy per direction "
                                            "(B200_PROFILING.md)",
                             "nvlink_bytes_per_direction_per_launch": nvl_bytes,
                             "hbm_achieved": achieved, "hbm_frac": achieved / hbm_peak})
    cl.close()

    staged = None
    if args.path == "fused" and not args.no_staged:
        # the three-kernel path (push -> landing slot, reduce+apply, pull), the one
        # asynchronous / cross-process workers use; reported beside the headline
        cl = engine.TorchrunCluster(variables, ps_tasks, engine.AdamOptimizer(0.01),
                                    placement=placement, stripes=args.stripes, wire=wire,
                                    device=local_rank)
        for t in cl.worker.grad_flat:
            t.copy_(torch.randn(t.numel(), device="cuda", generator=gen) * 1e-2)
        for _ in range(warmup):
            one_step()
        t2 = KernelTimer()
        ms_staged, l_staged = timed(max(3, steps // 2), timer=t2)
        se = ((cl.dominant.spec.nelem + 1023) // 1024) * 1024
        per = 24 + esz * world
        staged = {"value": bytes_step / (ms_staged * 1e-3) / 1e9, "unit": "GB/s",
                  "ms_per_step": ms_staged, "gpu_launches": l_staged,
                  "apply_kernel": {"kernel": "k_apply<ADAM,%s,SlotSrc<f32>>" % args.mode,
                                   "avg_launch_ms": t2.mean_ms(),
                                   "algorithmic_bytes_per_elem": per,
                                   "achieved": per * se / (t2.mean_ms() * 1e-3) / 1e9,
                                   "frac": per * se / (t2.mean_ms() * 1e-3) / 1e9 / hbm_peak}}
        cl.close()

    e2e = None
    if not args.no_e2e:
        # same workload through the host-in / host-out public call; more, smaller
        # shards per bucket so H2D, the kernels and D2H pipeline across shards
        e2e_stripes = max(16, world)
        cl = engine.TorchrunCluster(variables, ps_tasks, engine.AdamOptimizer(0.01),
                                    placement=placement, stripes=e2e_stripes, wire=wire,
                                    device=local_rank)
        cl.staging = engine.HostStaging(cl.worker)
        for t in cl.staging.grad:
            t.copy_(torch.randn(t.numel(), generator=torch.Generator().manual_seed(200 + rank))
                    * 1e-2)
        for _ in range(2):
            one_step(host=True)
        ms_e2e, _ = timed(max(3, steps // 2), host=True)
        st = cl.staging
        e2e = {"value": bytes_step / (ms_e2e * 1e-3) / 1e9, "unit": "GB/s",
               "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": st.h2d_bytes(), "d2h_bytes_per_step": st.d2h_bytes(),
               "stripes_per_bucket": e2e_stripes,
               "api": "tfmesos_b200.engine.TorchrunCluster.round_host (pinned host "
                      "gradients in, host parameters out, per rank; H2D / kernels / D2H "
                      "pipelined over the shards)"}
        cl.close()

    # clocks were sampled across all timed regions above (headline, staged, e2e)
    clocks = sampler.stop() if rank == 0 else None

    mnist = softmax = None
    if not args.no_mnist:
        mnist = mnist_section(torch, engine, psx, world, rank, dist, "mlp")
        softmax = mnist_section(torch, engine, psx, world, rank, dist, "softmax")

    cpu = cpu_grpc = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)          # the CPU arm gets every host core
        cb = cpu_ps(args, steps=5, warmup=1)
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        try:
            # the same path over the transport the reference actually selects
            # (protocol='grpc'): loopback gRPC, one RPC per variable per direction
            from oracle import cpu_ps_grpc
            cpu_grpc = cpu_ps_grpc.time_round(min(n_full, 10_000_000), steps=3, warmup=1)
            cpu_grpc.update({"kind": "port", "transport": "python grpcio, loopback, raw bytes "
                             "(TensorFlow's C++ gRPC core moves tensors a few times faster; "
                             "the memcpy figure above is the upper bound for this path)"})
        except Exception as exc:
            cpu_grpc = {"unavailable": str(exc)[:200]}

    if rank == 0:
        line = {
            "metric": "ps_push_pull_GBps", "value": value, "unit": "GB/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.wire == "f32" else "f32 master / bf16 wire",
            "data": "synthetic",
            "config": workload_config(args, world, n_full),
            "steps_per_sec": 1e3 / ms_step * (world if args.mode == "async" else 1),
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "staged_path": staged,
            "e2e": e2e,
            "cpu_baseline": cpu,
            "cpu_baseline_grpc": cpu_grpc,
            "mnist_replica": mnist,
            "mnist_softmax_sgd": softmax,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
