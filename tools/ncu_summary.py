"""Prints the metrics we quote in profiles/ from an .ncu-rep (read here, with no GPU): python tools/ncu_summary.py file.ncu-rep"""
import csv
import re
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print("%-72s %-10s %s" % (k, units[i], [r[i] for r in data]))
for i, h in enumerate(hdr):
    if re.search(r"issue_stalled.*per_issue_active", h):
        vals = [r[i] for r in data]
        if any(float(v or 0) > 0.15 for v in vals):
            print("%-72s %-10s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall/issue "), "", vals))
