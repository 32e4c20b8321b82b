#!/usr/bin/env python
"""Summaries of ncu outputs for profiles/:
    python tools/ncu_summary.py launches <launches.csv>      kernel / launches / total us / share
    python tools/ncu_summary.py raw <prof_raw.csv> [regex]   the roofline-relevant metrics per kernel
"""
import collections
import csv
import re
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__waves_per_multiprocessor",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "nvlrx__bytes.sum", "nvltx__bytes.sum", "pcie__read_bytes.sum", "pcie__write_bytes.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])[:100]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-100s %8s %12s %7s" % ("kernel", "launches", "total us", "share"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-100s %8d %12.1f %6.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))


def raw(path, pattern=None):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H, units = rows[hdr], rows[hdr + 1]
    ki = H.index("Kernel Name")
    for r in rows[hdr + 2:]:
        if len(r) < len(H) or (pattern and not re.search(pattern, r[ki])):
            continue
        print("kernel: %s" % re.sub(r"\(.*", "", r[ki])[:120])
        for m in KEEP:
            if m in H:
                print("  %-80s %s %s" % (m, r[H.index(m)], units[H.index(m)]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
