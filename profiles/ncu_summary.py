#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / bench.py cite.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shfl.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.per_cycle_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active")


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("==", name[:100], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        stalls = []
        for i, h in enumerate(hdr):
            base = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h
            if h in KEYS:
                print(f"  {h:70s} {r[i]:>18s} {units[i]}")
            m = STALL.search(h)
            if m and h.endswith(".ratio"):
                try:
                    stalls.append((float(r[i]), m.group(1)))
                except ValueError:
                    pass
        for v, n in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {n:40s} {v:8.3f} warps/issue")


if __name__ == "__main__":
    main()
