#!/usr/bin/env python
"""Short text summary of one `ncu --set full` capture (the metrics DESIGN.md quotes).
usage: python tools/ncu_summary.py X.ncu-rep [rows] > profiles/NAME_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("kernel: %s" % d.get("Kernel Name", ("", "?"))[1])
        for k in KEYS:
            if k in d:
                print("  %-82s %-16s %s" % (k, d[k][0], d[k][1]))
        for k in sorted(d):
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    if float(d[k][1]) >= 0.02:
                        print("  %-82s %-16s %s" % (k, d[k][0], d[k][1]))
                except ValueError:
                    pass


if __name__ == "__main__":
    main()
