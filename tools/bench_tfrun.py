#!/usr/bin/env python
"""Steady-state step time of the LITERAL reference-API path:

    tfrun -w W -s 1 [-Gw 1] -- python examples/mnist/mnist_replica.py --ps_hosts {ps_hosts} ...

one OS process per ps / worker task (tfmesos/scheduler.py:201-217, server.py:95-98),
free-running async Adam -- the reference's default (mnist_replica.py:198-205).  The
chief prints one line per step with a wall-clock stamp; the step time is the
median difference over the second half of the run.  Variants: exact step fetch
(the value TF's sess.run returns: an 8-byte copy the host waits for) vs --lag_step
(never blocks the host), numpy batches (as the reference feeds them) vs
--device_batches.  One JSON object on stdout."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(workers, extra, steps, gw=0, timeout=240):
    cmd = [sys.executable, os.path.join(ROOT, "script", "tfrun"), "-w", str(workers), "-s", "1"] + \
          (["-Gw", str(gw)] if gw else []) + ["--", sys.executable, os.path.join(ROOT, "examples", "mnist", "mnist_replica.py"),
           "--ps_hosts", "{ps_hosts}", "--worker_hosts", "{worker_hosts}",
           "--job_name", "{job_name}", "--worker_index", "{task_index}",
           "--train_steps", str(steps)] + extra
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""),
               PYTHONUNBUFFERED="1")
    import signal
    from types import SimpleNamespace
    p = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         start_new_session=True)
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(p.pid, signal.SIGKILL)       # exactly the process group started above
        except OSError:
            pass
        p.communicate()
        return {"error": "tfrun did not finish within %d s" % timeout}
    r = SimpleNamespace(returncode=p.returncode, stdout=out, stderr=err)
    if r.returncode != 0:
        return {"error": r.stderr[-600:]}
    stamps = [float(m.group(1)) for m in
              re.finditer(r"^([0-9.]+): Worker 0: training step \d+ done", r.stdout, re.M)]
    el = re.search(r"Training elapsed time: ([0-9.]+) s", r.stdout)
    val = re.search(r"validation cross entropy = ([-+0-9.eE]+|nan|inf)", r.stdout)
    gs = [int(m.group(1)) for m in re.finditer(r"\(global step: (\d+)\)", r.stdout)]
    out = {"chief_local_steps": len(stamps), "final_global_step": max(gs) if gs else None,
           "elapsed_s": float(el.group(1)) if el else None,
           "validation_cross_entropy": float(val.group(1)) if val else None}
    if len(stamps) > 20:
        half = stamps[len(stamps) // 2:]
        d = sorted(b - a for a, b in zip(half, half[1:]))
        out["ms_per_chief_step_median"] = 1e3 * d[len(d) // 2]
        out["ms_per_chief_step_p90"] = 1e3 * d[(9 * len(d)) // 10]
    if el and stamps:
        out["ms_per_chief_step_mean"] = 1e3 * float(el.group(1)) / len(stamps)
    return out


def main():
    steps = int(os.environ.get("TFRUN_STEPS", "3000"))
    res = {"command": "tfrun -w W -s 1 -- python examples/mnist/mnist_replica.py ... "
                      "--train_steps %d (async Adam, batch 100, 784-100-10)" % steps}
    import torch
    n_gpus = torch.cuda.device_count()
    res["gpus"] = n_gpus
    res["note"] = ("1 GPU: the ps task and the worker tasks are separate PROCESSES time-slicing "
                   "one GPU (a context switch per step each way); with -Gw 1 on >= 2 GPUs every "
                   "worker has its own GPU and the ps task shares GPU 0 with worker 0")
    full = os.environ.get("TFRUN_FULL", "0") == "1"
    for workers in (1, 2):
        gw = 1 if n_gpus >= workers and n_gpus >= 2 else 0
        variants = [("exact_step_numpy_batches", []),
                    ("lag_step_device_batches", ["--device_batches", "--lag_step"])]
        if full:
            variants.insert(1, ("exact_step_device_batches", ["--device_batches"]))
        for name, extra in variants:
            res["w%d%s/%s" % (workers, "-Gw1" if gw else "", name)] = run(
                workers, extra + ["--quiet_steps"] * 0, steps * workers, gw)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
