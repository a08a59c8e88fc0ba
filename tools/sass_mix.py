#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source sass` dump by opcode: executed warp-instructions per kernel.

    ncu -i prof.ncu-rep --page source --csv --print-source sass > src.csv ; python tools/sass_mix.py src.csv [top]
"""
import csv
import collections
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    kernel = None
    hdr = None
    mix = collections.OrderedDict()
    stall = collections.OrderedDict()
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            kernel = row[1].split("(")[0]
            mix[kernel] = collections.Counter()
            stall[kernel] = collections.Counter()
            hdr = None
            continue
        if row[0] == "Address":
            hdr = row
            continue
        if hdr is None or kernel is None:
            continue
        try:
            src = row[hdr.index("Source")]
            n = float(row[hdr.index("Instructions Executed")] or 0)
            smp = float(row[hdr.index("# Samples")] or 0)
        except (ValueError, IndexError):
            continue
        toks = src.split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.rstrip(";")
        base = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("F2F", "F2I", "I2F", "LDG", "STG", "LDS", "STS", "MUFU", "FRND") else op.split(".")[0]
        mix[kernel][base] += n
        stall[kernel][base] += smp
    for k, c in mix.items():
        tot = sum(c.values())
        ts = sum(stall[k].values()) or 1
        print(f"== {k}: {tot/1e6:.1f} M warp-instructions")
        for op, n in c.most_common(top):
            print(f"   {op:<14} {n/1e6:9.1f} M  {100*n/tot:5.1f}%   stall-samples {100*stall[k][op]/ts:5.1f}%")


if __name__ == "__main__":
    main()
