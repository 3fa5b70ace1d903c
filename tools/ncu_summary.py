#!/usr/bin/env python
"""Key metrics of an ncu report as text (the summaries committed under profiles/).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rN_ncu_<kernel>_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    for v in vals:
        d = dict(zip(hdr, v))
        u = dict(zip(hdr, units))
        print("%-90s %s" % ("Kernel Name", d.get("Kernel Name")))
        for k in KEYS:
            if k in d:
                print("%-90s %s %s" % (k, d[k], u[k]))
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") \
                    and "not_issued" not in k:
                print("%-90s %s %s" % (k, d[k], u[k]))
        print()


if __name__ == "__main__":
    main()
