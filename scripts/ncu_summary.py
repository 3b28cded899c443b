#!/usr/bin/env python
"""Key metrics of the first kernel in an .ncu-rep.  usage: python scripts/ncu_summary.py rep [out.txt]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_fp64.sum', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']
out = []
for i, h in enumerate(hdr):
    if h in keep:
        out.append(f"{h} [{units[i]}] = {vals[i]}")
    elif 'issue_stalled' in h and h.endswith('ratio'):
        try:
            if float(vals[i]) >= 0.05: out.append(f"{h} = {vals[i]}")
        except ValueError: pass
    elif h == "Kernel Name":
        out.append(f"kernel = {vals[i]}")
s = "\n".join(out)
print(s)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(s + "\n")
