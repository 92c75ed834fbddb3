#!/usr/bin/env python
"""Digest of `ncu --page raw --csv` and `--page source --csv` exports (development tool).

    python tools/ncu_digest.py raw  gpurun_out/raw_x.csv
    python tools/ncu_digest.py src  gpurun_out/src_x.csv [top]
"""
import collections
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.sum.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__shared_mem_per_block_dynamic',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.avg']


def raw(path):
    rows = list(csv.reader(open(path)))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        print('===', r[h.index('Kernel Name')][:100])
        for k in KEYS:
            if k in h:
                print(f'  {k}: {r[h.index(k)]} {u[h.index(k)]}')


def src(path, top=25):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    h = rows[hi]
    body = [r for r in rows[hi + 1:] if len(r) >= len(h)]
    stall = [i for i, k in enumerate(h) if k.startswith('stall_') and 'Not Issued' not in k]
    ns = h.index('# Samples')
    ie = h.index('Instructions Executed')
    tot = collections.Counter()
    mix = collections.Counter()
    n = 0
    for r in body:
        for i in stall:
            tot[h[i]] += int(r[i] or 0)
        n += int(r[ns] or 0)
        ops = [o for o in r[1].split() if not o.startswith('@')]
        mix[ops[0].split('.')[0] if ops else '?'] += int(r[ie] or 0)
    print('samples', n)
    for k, v in tot.most_common(10):
        print(f'  {k:28s} {v:8d} {v/max(n,1):.3f}')
    t = sum(mix.values())
    print('warp instructions', t)
    for k, v in mix.most_common(22):
        print(f'  {k:10s} {v:10d} {v/t:.3f}')
    print('hottest instructions')
    for r in sorted(body, key=lambda r: -int(r[ns] or 0))[:top]:
        st = {h[i]: int(r[i] or 0) for i in stall if int(r[i] or 0) > 0}
        st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f'  {int(r[ns]):6d} {r[1].strip()[:70]:70s} {st}')


if __name__ == '__main__':
    if sys.argv[1] == 'raw':
        raw(sys.argv[2])
    else:
        src(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
