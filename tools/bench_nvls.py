"""NVLS primitives vs their unicast equivalents on 2 GPUs (1 GiB, CUDA events,
median of 7).  Output GB/s = payload bytes / time."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tfmesos_b200 import psx  # noqa: E402

N = 1 << 30


def main():
    psx.init(0)
    if not (psx.nvls_supported(0) and psx.nvls_supported(1)):
        print(json.dumps({"nvls": "not supported on this box"}))
        return
    try:
        mc = psx.MulticastBuffer([0, 1], N)
    except RuntimeError as exc:
        print(json.dumps({"nvls": "multicast object creation failed", "error": str(exc)[:300]}))
        return
    psx.enable_peer(0, 1)
    psx.enable_peer(1, 0)
    t = [mc.tensor(0), mc.tensor(1)]
    for x in t:
        x.fill_(1.0)
    src = [torch.ones(N // 4, device="cuda:%d" % d) for d in (0, 1)]
    dst = [torch.zeros(N // 4, device="cuda:%d" % d) for d in (0, 1)]
    s = [torch.cuda.Stream(device=d) for d in (0, 1)]

    def timed(fn, devs):
        times = []
        for it in range(10):
            for d in (0, 1):
                torch.cuda.synchronize(d)
            ev = {}
            for d in devs:
                with torch.cuda.device(d):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(s[d])
                    ev[d] = (e0, e1)
            fn()
            for d in devs:
                with torch.cuda.device(d):
                    ev[d][1].record(s[d])
            for d in (0, 1):
                torch.cuda.synchronize(d)
            if it >= 3:
                times.append(max(ev[d][0].elapsed_time(ev[d][1]) for d in devs))
        times.sort()
        return times[len(times) // 2]

    half = N // 2
    out = {}
    ms = timed(lambda: mc.broadcast(0, src[0].data_ptr(), N, 0, s[0]), [0])
    out["multicast_broadcast_from_gpu0"] = {"ms": ms, "payload_GBps": N / ms / 1e6}
    ms = timed(lambda: mc.reduce(0, dst[0].data_ptr(), N, 0, s[0]), [0])
    out["switch_reduce_into_gpu0"] = {"ms": ms, "payload_GBps": N / ms / 1e6}

    def both_reduce():
        mc.reduce(0, dst[0].data_ptr(), half, 0, s[0])
        mc.reduce(1, dst[1].data_ptr(), half, half, s[1])
    ms = timed(both_reduce, [0, 1])
    out["striped_reduce_both_gpus_half_each"] = {"ms": ms, "payload_GBps_per_gpu": half / ms / 1e6}

    def both_bcast():
        mc.broadcast(0, src[0].data_ptr(), half, 0, s[0])
        mc.broadcast(1, src[1].data_ptr(), half, half, s[1])
    ms = timed(both_bcast, [0, 1])
    out["striped_broadcast_both_gpus_half_each"] = {"ms": ms, "payload_GBps_per_gpu": half / ms / 1e6}

    def round_like():
        both_reduce()
        both_bcast()
    ms = timed(round_like, [0, 1])
    out["striped_round_like_reduce_then_broadcast"] = {
        "ms": ms, "note": "gather+scatter of a 1 GiB bucket striped over 2 GPUs through the "
                          "switch; the unicast fused kernel moves the same bucket in "
                          "~1.3 ms x (1 GiB / 0.8 GB)"}
    # unicast equivalents with the library's copy kernel
    ms = timed(lambda: psx.copy(0, t[1].data_ptr(), src[0].data_ptr(), N, s[0]), [0])
    out["unicast_write_gpu0_to_gpu1"] = {"ms": ms, "payload_GBps": N / ms / 1e6}
    print(json.dumps(out))
    mc.destroy()


if __name__ == "__main__":
    main()
