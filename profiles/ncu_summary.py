#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics we track.  Usage: ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:80])
    for w in WANT:
        if w in hdr:
            print(f'   {w:75s} {r[hdr.index(w)]} {units[hdr.index(w)]}')
    st = sorted(((float(r[hdr.index(s)] or 0), s) for s in stall), reverse=True)[:5]
    for v, s in st:
        print(f'   stall {s.split("issue_stalled_")[1].replace("_per_issue_active.ratio",""):30s} {v:.2f}')
