#!/usr/bin/env python
"""Opcode mix and stall reasons of one kernel from an `ncu --page source --csv` export.

    python tools/opmix.py <source.csv> <units> [label]

`units` = the work items of the launch the per-unit columns are divided by (frames for the
spectrogram ring kernel: 60 000 in the bench's configs[1] launch).
"""
import collections
import csv
import sys


def main():
    path, units = sys.argv[1], float(sys.argv[2])
    label = sys.argv[3] if len(sys.argv) > 3 else 'unit'
    rows = list(csv.reader(open(path)))
    kernel = rows[0][1]
    h = rows[1]
    si, ei, ni = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    ops = collections.Counter()
    samples = collections.Counter()
    stalls = collections.Counter()
    for r in rows[2:]:
        if len(r) <= ei:
            continue
        op = r[si].split()
        if not op:
            continue
        name = op[1] if op[0].startswith('@') else op[0]
        name = name.split('.')[0]
        try:
            ops[name] += float(r[ei])
            samples[name] += float(r[ni])
        except ValueError:
            continue
        for i, c in stall_cols:
            try:
                stalls[c] += float(r[i])
            except (ValueError, IndexError):
                pass
    tot = sum(ops.values())
    fp64 = sum(v for k, v in ops.items() if k in ('DADD', 'DMUL', 'DFMA'))
    stot = sum(samples.values()) or 1.0
    print(kernel)
    print(f'executed warp instructions per {label}: {tot/units:.0f}; fp64 (DADD+DMUL+DFMA) {fp64/units:.0f}\n')
    print(f'{"opcode":10s} {"per " + label:>10s} {"share":>7s} {"stall samples":>14s}')
    for k, v in ops.most_common(26):
        print(f'{k:10s} {v/units:10.1f} {v/tot:7.3f} {samples[k]/stot:14.3f}')
    st = sum(stalls.values()) or 1.0
    print('\nwarp stall samples by reason:')
    for k, v in stalls.most_common(10):
        print(f'  {k:24s} {v/st:6.3f}')


if __name__ == '__main__':
    main()
