#!/usr/bin/env python
"""Summarise .ncu-rep files (ncu --set full) into the text format kept under profiles/: one block per kernel launch with
the metrics DESIGN.md argues from (duration, DRAM / L2 / L1 throughput, occupancy limits, shared-memory wavefronts and
bank conflicts, issue-slot utilisation, stall reasons, pipe utilisation).
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/rN_ncu_x_summary.txt
Needs `ncu` on PATH (reading a report does not need a GPU)."""
import csv
import io
import subprocess
import sys

WANT = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max",
]
STALLS = ("long_scoreboard", "short_scoreboard", "barrier", "mio_throttle", "lg_throttle", "math_pipe_throttle", "wait", "not_selected",
          "membar", "branch_resolving", "no_instruction", "sleeping")


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            print(f"===== {rep}")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    print(f"  {w:<90} {vals[i]} {units[i]}")
            for st in STALLS:
                name = f"smsp__average_warps_issue_stalled_{st}_per_issue_active.ratio"
                if name in hdr:
                    i = hdr.index(name)
                    print(f"  {name:<90} {vals[i]} {units[i]}")


if __name__ == "__main__":
    main()
