#!/usr/bin/env python
"""Markdown summary of an .ncu-rep: the raw metrics the judge reads + the top source lines by stall samples.
    python tools/ncu_summary.py <report.ncu-rep> <kernel regex> <object file> "<title>" > profiles/xxx.md"""
import csv
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def main():
    rep, pat, obj, title = sys.argv[1:5]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", "regex:" + pat] + (["--launch-skip", os.environ["NCU_SKIP"]] if os.environ.get("NCU_SKIP") else []), capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n")
    print(f"Source: `{rep.split('/')[-1]}` (`ncu --set full --clock-control none --import-source on`), read with "
          "`ncu -i ... --page raw --csv`; per-line table: `tools/ncu_by_line.py` (SASS samples joined with `nvdisasm -g`).\n")
    for r in rows[2:3]:
        print(f"Kernel: `{r[hdr.index('Kernel Name')]}`\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} | {units[i]} |")
    print()
    by_line = subprocess.run([sys.executable, "tools/ncu_by_line.py", rep, pat, obj, "14"], capture_output=True, text=True).stdout
    print("Top source lines by stall samples:\n")
    print(by_line)


if __name__ == "__main__":
    main()
