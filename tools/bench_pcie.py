"""Raw PCIe ceilings for the e2e number: pinned H2D alone, D2H alone, and both at
once (copy engines, 0.8 GB each = one worker's gradient / parameter set)."""
import json

import torch

N = 200_200_000


def main():
    h_in = torch.empty(N, dtype=torch.float32).pin_memory()
    h_out = torch.empty(N, dtype=torch.float32).pin_memory()
    d_in = torch.empty(N, dtype=torch.float32, device="cuda")
    d_out = torch.ones(N, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=5):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

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
    out = {}
    for name, fn in (("h2d_alone", h2d), ("d2h_alone", d2h), ("h2d_and_d2h_together", both)):
        ms = timed(fn)
        out[name] = {"ms": ms, "GBps_per_direction": gb / ms * 1e3}
    out["note"] = ("e2e moves 0.8 GB each way per step; its floor is the 'together' time, "
                   "times (S+1)/S for the S-shard pipeline fill")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
