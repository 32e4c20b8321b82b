"""Empirical NVLink P2P peak between two B200s, measured with this library's own
copy kernel ("the build must measure an empirical P2P read and write
peak ... and use it as the roofline denominator").

One process, two GPUs, 1 GiB buffers, CUDA events, median of 10:
  write_uni   GPU0 kernel stores into GPU1                (push direction)
  read_uni    GPU0 kernel loads from GPU1                 (pull direction)
  write_bidir both GPUs store into each other at once
  read_bidir  both GPUs load from each other at once
  mixed_bidir each GPU loads AND stores remotely at once  (what psx_round does)
GB/s are per direction per GPU."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tfmesos_b200 import psx  # noqa: E402

N = 1 << 30


def main():
    assert torch.cuda.device_count() >= 2
    psx.init(0)
    psx.enable_peer(0, 1)
    psx.enable_peer(1, 0)
    a = [torch.empty(N, dtype=torch.uint8, device="cuda:%d" % d) for d in (0, 1)]
    b = [torch.empty(N, dtype=torch.uint8, device="cuda:%d" % d) for d in (0, 1)]
    c = [torch.empty(N, dtype=torch.uint8, device="cuda:%d" % d) for d in (0, 1)]
    for t in a + b + c:
        t.zero_()
    s = [torch.cuda.Stream(device=d) for d in (0, 1)]
    s2 = [torch.cuda.Stream(device=d) for d in (0, 1)]

    def copy(dev, dst, src, stream):
        psx.copy(dev, dst.data_ptr(), src.data_ptr(), N, stream)

    def run(ops, per_dir_bytes):
        """ops: list of (device, dst, src, stream)"""
        times = []
        for it in range(13):
            torch.cuda.synchronize(0)
            torch.cuda.synchronize(1)
            ev = []
            for d in (0, 1):
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                ev.append((e0, e1))
            used = sorted({op[0] for op in ops})
            for d in used:
                with torch.cuda.device(d):
                    ev[d][0].record(s[d])
                    s2[d].wait_stream(s[d])
            for dev, dst, src, which in ops:
                copy(dev, dst, src, s[dev] if which == 0 else s2[dev])
            for d in used:
                with torch.cuda.device(d):
                    s[d].wait_stream(s2[d])
                    ev[d][1].record(s[d])
            torch.cuda.synchronize(0)
            torch.cuda.synchronize(1)
            if it >= 3:
                times.append(max(ev[d][0].elapsed_time(ev[d][1]) for d in used))
        times.sort()
        ms = times[len(times) // 2]
        return {"ms": ms, "GBps_per_direction": per_dir_bytes / ms / 1e6}

    out = {
        "local_copy_gpu0": run([(0, b[0], a[0], 0)], N),
        "write_uni_0to1": run([(0, b[1], a[0], 0)], N),
        "read_uni_0from1": run([(0, b[0], a[1], 0)], N),
        "write_bidir": run([(0, b[1], a[0], 0), (1, b[0], a[1], 0)], N),
        "read_bidir": run([(0, b[0], a[1], 0), (1, b[1], a[0], 0)], N),
        # each GPU: one kernel loading remotely + one kernel storing remotely, concurrently
        "mixed_bidir": run([(0, b[0], a[1], 0), (0, c[1], a[0], 1),
                            (1, b[1], a[0], 0), (1, c[0], a[1], 1)], 2 * N),
    }
    out["note"] = ("GB/s per direction per GPU; local_copy counts bytes once (read+write "
                   "= 2x that); mixed_bidir carries 2 GiB per direction")
    G = torch.cuda.device_count()
    if G >= 4:
        out.update(many_gpu(G))
    print(json.dumps(out))


def many_gpu(G):
    """W -> 1 incast (W GPUs store into GPU 0 / GPU 0's kernel loads from W GPUs) and
    the all-pairs pattern of the striped round (every GPU loads 1/G from every other
    GPU and stores 1/G into every other GPU), G = all visible GPUs."""
    for a in range(G):
        for b_ in range(G):
            if a != b_:
                psx.enable_peer(a, b_)
    src = [torch.zeros(N, dtype=torch.uint8, device="cuda:%d" % d) for d in range(G)]
    dst = [torch.zeros(N, dtype=torch.uint8, device="cuda:%d" % d) for d in range(G)]
    land = [torch.zeros(N, dtype=torch.uint8, device="cuda:0") for _ in range(min(G - 1, 7))]
    st = [[torch.cuda.Stream(device=d) for _ in range(G)] for d in range(G)]

    def run(ops, used):
        times = []
        for it in range(9):
            for d in range(G):
                torch.cuda.synchronize(d)
            ev = {}
            for d in used:
                with torch.cuda.device(d):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st[d][0])
                    for k in range(1, G):
                        st[d][k].wait_stream(st[d][0])
                    ev[d] = (e0, e1)
            for dev, dptr, sptr, nbytes, k in ops:
                psx.copy(dev, dptr, sptr, nbytes, st[dev][k])
            for d in used:
                with torch.cuda.device(d):
                    for k in range(1, G):
                        st[d][0].wait_stream(st[d][k])
                    ev[d][1].record(st[d][0])
            for d in range(G):
                torch.cuda.synchronize(d)
            if it >= 3:
                times.append(max(ev[d][0].elapsed_time(ev[d][1]) for d in used))
        times.sort()
        return times[len(times) // 2]

    out = {"gpus": G}
    for W in [w for w in (1, 2, 4, 7) if w < G]:
        # W workers each STORE 1 GiB into GPU 0 (push incast)
        ops = [(w, land[w - 1].data_ptr(), src[w].data_ptr(), N, 0) for w in range(1, W + 1)]
        ms = run(ops, list(range(1, W + 1)))
        out["incast_write_%dto1" % W] = {"ms": ms, "ingress_GBps_at_gpu0": W * N / ms / 1e6}
        # GPU 0 LOADS 1 GiB from each of W workers (PS-side gather), one stream per source
        ops = [(0, land[w - 1].data_ptr(), src[w].data_ptr(), N, w) for w in range(1, W + 1)]
        ms = run(ops, [0])
        out["gather_read_1from%d" % W] = {"ms": ms, "ingress_GBps_at_gpu0": W * N / ms / 1e6}
    part = N // G // 16 * 16
    ops = []
    for a in range(G):
        for k, b_ in enumerate([x for x in range(G) if x != a]):
            ops.append((a, dst[a].data_ptr() + b_ * part, src[b_].data_ptr() + a * part, part, k))
            ops.append((a, dst[b_].data_ptr() + a * part, src[a].data_ptr() + b_ * part, part,
                        (k + 1) % (G - 1) + 0))
    ms = run(ops, list(range(G)))
    out["all_pairs_read_plus_write"] = {
        "ms": ms, "GBps_per_direction_per_gpu": 2 * (G - 1) * part / ms / 1e6,
        "note": "every GPU loads 1/G GiB from and stores 1/G GiB into every other GPU at once "
                "(the unicast striped round's traffic pattern)"}
    return out


if __name__ == "__main__":
    main()
