"""Why is the CPU arm's round time bimodal on the GPU boxes (~20 ms and ~100 ms for
the same round)?  Prints the container's CPU quota (cgroup v1 / v2) and throttling
counters, then times the CPU-PS round with different pool sizes, each in its own
process (the pool is created once per process)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def read(path):
    try:
        return open(path).read().strip()
    except OSError:
        return None


def quota():
    out = {"cpu.max(v2)": read("/sys/fs/cgroup/cpu.max"),
           "cfs_quota_us(v1)": read("/sys/fs/cgroup/cpu/cpu.cfs_quota_us"),
           "cfs_period_us(v1)": read("/sys/fs/cgroup/cpu/cpu.cfs_period_us"),
           "cpu.stat(v2)": read("/sys/fs/cgroup/cpu.stat"),
           "cpu.stat(v1)": read("/sys/fs/cgroup/cpu/cpu.stat"),
           "cpuset": read("/sys/fs/cgroup/cpuset.cpus.effective") or read("/sys/fs/cgroup/cpuset/cpuset.cpus"),
           "nproc": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)),
           "proc_cgroup": read("/proc/self/cgroup")}
    cores = None
    if out["cpu.max(v2)"] and not out["cpu.max(v2)"].startswith("max"):
        q, p = out["cpu.max(v2)"].split()
        cores = float(q) / float(p)
    elif out["cfs_quota_us(v1)"] and int(out["cfs_quota_us(v1)"]) > 0:
        cores = int(out["cfs_quota_us(v1)"]) / float(out["cfs_period_us(v1)"])
    out["quota_cores"] = cores
    return out


def child(threads):
    import time
    from oracle import ps_oracle as o
    b = o.CpuPsBaseline(50_000_000, 1, threads=threads)
    ts = []
    for _ in range(12):
        t0 = time.perf_counter()
        b.round(o.SUM)
        ts.append(round((time.perf_counter() - t0) * 1e3, 1))
    print(json.dumps({"threads": b.threads, "rounds_ms": ts}))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(int(sys.argv[1]))
    else:
        q = quota()
        print(json.dumps(q))
        sizes = [0, 64, 32, 16]
        if q["quota_cores"]:
            sizes.insert(1, max(1, int(q["quota_cores"])))
        for t in sizes:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), str(t)],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
            print(r.stdout.strip() or r.stderr[-300:])
            st = quota()
            print(json.dumps({"after_threads": t, "cpu.stat(v2)": st["cpu.stat(v2)"],
                              "cpu.stat(v1)": st["cpu.stat(v1)"]}))
