"""NVLS primitives vs their unicast equivalents on all visible GPUs of one
process (1 GiB bucket, CUDA events, median of 7).  GB/s = payload bytes / time."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tfmesos_b200 import psx  # noqa: E402

N = 1 << 30


def main():
    psx.init(0)
    G = torch.cuda.device_count()
    devs = list(range(G))
    if G < 2 or not all(psx.nvls_supported(d) for d in devs):
        print(json.dumps({"nvls": "not supported on this box"}))
        return
    try:
        mc = psx.MulticastBuffer(devs, N)
    except RuntimeError as exc:
        print(json.dumps({"nvls": "multicast object creation failed", "error": str(exc)[:300]}))
        return
    for a in devs:
        for b in devs:
            if a != b:
                psx.enable_peer(a, b)
    t = [mc.tensor(d) for d in devs]
    for x in t:
        x.fill_(1.0)
    src = [torch.ones(N // 4, device="cuda:%d" % d) for d in devs]
    dst = [torch.zeros(N // 4, device="cuda:%d" % d) for d in devs]
    s = [torch.cuda.Stream(device=d) for d in devs]

    def timed(fn, used):
        times = []
        for it in range(10):
            for d in devs:
                torch.cuda.synchronize(d)
            ev = {}
            for d in used:
                with torch.cuda.device(d):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(s[d])
                    ev[d] = (e0, e1)
            fn()
            for d in used:
                with torch.cuda.device(d):
                    ev[d][1].record(s[d])
            for d in devs:
                torch.cuda.synchronize(d)
            if it >= 3:
                times.append(max(ev[d][0].elapsed_time(ev[d][1]) for d in used))
        times.sort()
        return times[len(times) // 2]

    part = N // G // 16 * 16
    out = {"gpus": G, "bucket_bytes": N}
    ms = timed(lambda: mc.broadcast(0, src[0].data_ptr(), N, 0, s[0]), [0])
    out["multicast_broadcast_from_gpu0"] = {"ms": ms, "payload_GBps": N / ms / 1e6}
    ms = timed(lambda: mc.reduce(0, dst[0].data_ptr(), N, 0, s[0]), [0])
    out["switch_reduce_into_gpu0"] = {"ms": ms, "payload_GBps": N / ms / 1e6}

    def all_reduce_stripes():
        for d in devs:
            mc.reduce(d, dst[d].data_ptr(), part, d * part, s[d])

    def all_bcast_stripes():
        for d in devs:
            mc.broadcast(d, src[d].data_ptr(), part, d * part, s[d])

    ms = timed(all_reduce_stripes, devs)
    out["striped_reduce_every_gpu_its_stripe"] = {"ms": ms}
    ms = timed(all_bcast_stripes, devs)
    out["striped_broadcast_every_gpu_its_stripe"] = {"ms": ms}

    def round_like():
        all_reduce_stripes()
        all_bcast_stripes()
    ms = timed(round_like, devs)
    out["striped_round_like_reduce_then_broadcast"] = {
        "ms": ms, "note": "gather+scatter of the whole bucket striped over the GPUs through "
                          "the switch (no optimizer); compare with the unicast fused kernel "
                          "scaled to 1 GiB"}
    ms = timed(lambda: psx.copy(0, t[1].data_ptr(), src[0].data_ptr(), N, s[0]), [0])
    out["unicast_write_gpu0_to_gpu1"] = {"ms": ms, "payload_GBps": N / ms / 1e6}
    print(json.dumps(out))
    mc.destroy()


if __name__ == "__main__":
    main()
