#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of numbers we track per kernel.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("launch__grid_size", "grid"),
    ("launch__waves_per_multiprocessor", "waves"),
    ("launch__occupancy_limit_registers", "lim_regs"),
    ("launch__occupancy_limit_shared_mem", "lim_smem"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        name = d[hdr.index("Kernel Name")].split("(")[0]
        print("==", name)
        parts = []
        for key, short in WANT:
            if key in hdr:
                i = hdr.index(key)
                parts.append(f"{short}={d[i]}{(' ' + units[i]) if units[i] not in ('', '%') else ''}")
        print("   " + "  ".join(parts))


if __name__ == "__main__":
    main()
