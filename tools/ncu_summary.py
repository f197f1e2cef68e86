#!/usr/bin/env python3
"""Print the handful of ncu raw metrics we track for a kernel. usage: ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:72s} {units[i]:16s} " + " | ".join(r[i][:60] for r in rows[2:]))
tot = {}
for i, h in enumerate(hdr):
    if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
        try:
            tot[h.replace('smsp__pcsamp_warps_issue_stalled_', '')] = float(rows[2][i].replace(',', ''))
        except Exception:
            pass
s = sum(tot.values()) or 1
print("stall samples: " + ", ".join(f"{k} {100*v/s:.0f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v / s > 0.02))
