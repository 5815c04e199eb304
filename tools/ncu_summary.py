#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (ncu --set full) into the text form kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep 'header line' > profiles/x_summary.txt"""
import csv, subprocess, sys
rep, head = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active_mem_lgds.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "idc__requests.sum"]
print("# kernel: %s" % v[h.index("Kernel Name")])
if head:
    print("# " + head)
for n in want:
    if n in h:
        i = h.index(n)
        print("%-95s %-15s %s" % (n, u[i], v[i]))
for i, n in enumerate(h):
    if "issue_stalled" in n and n.endswith("per_issue_active.ratio") and "not_issued" not in n:
        try:
            if float(v[i]) > 0.1:
                print("%-95s %-15s %s" % (n, u[i], v[i]))
        except ValueError:
            pass
