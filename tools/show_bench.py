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
        e = d["e2e"]
        c = e.get("c_abi") or e
        print(f"{path}: mode={d['config']['field_mode']} value={d['value']:.4g} vox-upd/s  ms/iter={d['ms_per_iteration']:.3f}"
              f"  e2e={e['value']:.4g} (step {e['ms_per_step']:.0f} ms; C ABI: step {c['ms_per_step']:.0f} ms = h2d "
              f"{c.get('h2d_ms', 0):.0f} + loop {c.get('loop_ms', 0):.0f} + d2h {c.get('d2h_ms', 0):.0f})")
        print(f"   dominant kernel {r.get('kernel', '')[:24]}: {r['achieved']:.0f} GB/s = {r['frac']:.3f} of peak, {r['ms_per_launch']:.3f} ms/launch, "
              f"share {r['share_of_step']:.2f}; whole iteration {r['whole_iteration'].get('achieved_GBs', 0):.0f} GB/s = "
              f"{r['whole_iteration']['frac_of_peak']:.3f}")
        for k in r.get("per_kernel", []):
            print(f"      {k['kernel'][:30]:30s} {k['ms_per_launch']:.3f} ms  frac {k['frac']:.3f}  share {k['share_of_step']:.3f}")
        print("   stage ms/launch:", {k: round(v['ms'] / max(v['launches'], 1), 3) for k, v in r.get('stages', {}).items()})
        print("   clocks", d["clocks"], "launches", d["gpu_launches"], "mse", d["final_mse"])
        if d.get("cpu_baseline"):
            print("   cpu_baseline", d["cpu_baseline"])
