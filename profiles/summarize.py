#!/usr/bin/env python
"""Summarise ncu output brought back from a gpurun call.

    python profiles/summarize.py launches gpurun_out/launches_rNN.csv
    python profiles/summarize.py full gpurun_out/prof_rNN.ncu-rep

`launches`: per-kernel launch count, mean duration and share of the profiled
command (gpu__time_duration.sum pass; cold-cache, serialised: compare shares).
`full`: the raw-page metrics the roofline section of DESIGN.md quotes.
"""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            agg[r[ki]].append(float(r[vi].replace(',', '')))
        except (ValueError, IndexError):
            continue
    tot = sum(sum(v) for v in agg.values())
    print(f'{"kernel":90s} {"n":>4s} {"mean us":>10s} {"share":>6s}')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f'{k[:90]:90s} {len(v):4d} {sum(v)/len(v)/1e3:10.1f} {sum(v)/tot:6.3f}')


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print('---', r[h.index('Kernel Name')][:100])
        for w in WANT:
            if w in h:
                print(f'  {w}: {r[h.index(w)]} {units[h.index(w)]}')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
