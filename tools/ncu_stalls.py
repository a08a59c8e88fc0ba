#!/usr/bin/env python3
"""Per-kernel warp stall breakdown (PC sampling) + a few memory counters from an .ncu-rep:  python tools/ncu_stalls.py rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
extra = ["l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
         "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_active.avg", "smsp__warps_eligible.avg.per_cycle_active",
         "smsp__issue_active.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active",
         "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
         "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
for d in data:
    print("==", d[hdr.index("Kernel Name")].split("(")[0])
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                st.append((float(d[i].replace(",", "")), h[len("smsp__pcsamp_warps_issue_stalled_"):]))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1
    print("   stalls: " + "  ".join(f"{n}={100*v/tot:.1f}%" for v, n in sorted(st, reverse=True)[:9]))
    print("   " + "  ".join(f"{k.split('__',1)[1][:48]}={d[hdr.index(k)]}" for k in extra if k in hdr))
