#!/usr/bin/env python3
"""Regenerate profiles/traffic.json from `ncu --set full` captures: DRAM bytes (read + write) per launch of every loop
kernel, per field mode.  bench.py republishes the smoother's figure as `roofline.traffic`.

    python tools/make_traffic_json.py fp64=gpurun_out/r2n_prof_fp64.ncu-rep fp32=gpurun_out/r2n_prof_fp32.ncu-rep
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def per_kernel(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    acc = {}
    for d in data:
        name = re.sub(r"[<(].*", "", d[ik]).replace("void ", "").split("::")[-1]
        b = float(d[ir]) * SCALE[units[ir]] + float(d[iw]) * SCALE[units[iw]]
        acc.setdefault(name, []).append(b)
    return {k: sum(v) / len(v) for k, v in acc.items()}


def main():
    res = {"_source": "tools/make_traffic_json.py: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean per launch, "
                      "512x512x256, r=4; captures: " + ", ".join(a.split("=", 1)[1] for a in sys.argv[1:])}
    for arg in sys.argv[1:]:
        mode, rep = arg.split("=", 1)
        k = per_kernel(rep)
        smooth = next((v for n, v in k.items() if n.startswith("k_smooth3d")), None)
        res[mode] = {"k_smooth3d_dram_bytes_per_launch": smooth, "per_kernel_dram_bytes_per_launch": k}
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
