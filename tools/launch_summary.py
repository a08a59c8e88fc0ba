#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total/mean time, share.

    python tools/launch_summary.py gpurun_out/launches.csv [--ours-only]
Times under ncu are cold-cache and serialised: compare SHARES with the CUDA-event stage split of bench.py, not absolutes."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    ours_only = "--ours-only" in sys.argv
    rows = []
    with open(path, newline="") as f:
        rd = csv.reader(f)
        hdr = None
        for r in rd:
            if hdr is None:
                if r and r[0] == "ID":
                    hdr = r
                continue
            if len(r) == len(hdr):
                rows.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        ours = "dcma_" in name
        if ours_only and not ours:
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        a = agg.setdefault(name, [0, 0.0, r["Grid Size"], r["Block Size"], ours])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values() if a[4])
    print(f"{'kernel':78s} {'launches':>8s} {'total_ms':>9s} {'mean_us':>9s} {'share':>6s}  grid block")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = f"{a[1] / tot:6.3f}" if a[4] and tot else "   -  "
        print(f"{k[:78]:78s} {a[0]:8d} {a[1] / 1e6:9.3f} {a[1] / a[0] / 1e3:9.1f} {share}  {a[2]} {a[3]}")
    print(f"library kernels total: {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values() if a[4])} launches")


if __name__ == "__main__":
    main()
