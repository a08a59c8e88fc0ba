#!/usr/bin/env python3
"""Pretty-print bench.py JSON lines: python tools/show_bench.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if d.get("impl") == "reference":
            print(path, "REFERENCE value %.3g" % d["value"], d["cpu_baseline"])
            continue
        r = d["roofline"]
        print(f"{path}: mode={d['config']['field_mode']} value={d['value']:.4g} vox-upd/s  ms/iter={d['ms_per_iteration']:.3f}"
              f"  e2e={d['e2e']['value']:.4g} (step {d['e2e']['ms_per_step']:.0f} ms: h2d {d['e2e']['h2d_ms']:.0f} loop {d['e2e']['loop_ms']:.0f} d2h {d['e2e']['d2h_ms']:.0f})")
        print(f"   smooth kernel: {r['achieved']:.0f} GB/s = {r['frac']:.3f} of peak, {r['ms_per_launch']:.3f} ms/launch, share {r['share_of_step']:.2f};"
              f" whole iteration {r['whole_iteration']['achieved_GBs']:.0f} GB/s = {r['whole_iteration']['frac_of_peak']:.3f}")
        print("   stage ms/launch:", {k: round(v['ms'] / max(v['launches'], 1), 3) for k, v in r['stages'].items()})
        print("   clocks", d["clocks"], "launches", d["gpu_launches"], "mse", d["final_mse"])
        if d.get("cpu_baseline"):
            print("   cpu_baseline", d["cpu_baseline"])
