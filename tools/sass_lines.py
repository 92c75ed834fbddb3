#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel (development tool).

    python tools/sass_lines.py <lib.so> <cubin-stem> <kernel-substring> <ncu source csv>

Joins `nvdisasm -g` line info of the kernel with the per-instruction counters of an
`ncu --page source --csv` export (same instruction order) and prints, per CUDA source
line, executed warp instructions and stall samples, split into fp64 / other.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    lib, stem, kname, ncsv = sys.argv[1:5]
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp,
                   stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(tmp) if f.startswith(stem)][0]
    txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cub)], capture_output=True,
                         text=True).stdout.splitlines()
    inside = False
    line = 0
    seq = []      # (source line, opcode)
    for l in txt:
        if l.startswith('//---') and '.text.' in l:
            inside = kname in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = int(m.group(2))
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', l)
        if m:
            ops = [o for o in m.group(1).split() if not o.startswith('@')]
            seq.append((line, ops[0].split('.')[0] if ops else '?'))
    rows = list(csv.reader(open(ncsv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    h = rows[hi]
    body = [r for r in rows[hi + 1:] if len(r) >= len(h)]
    if len(body) != len(seq):
        print('instruction count mismatch', len(body), len(seq), file=sys.stderr)
    ie, ns = h.index('Instructions Executed'), h.index('# Samples')
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for (ln, op), r in zip(seq, body):
        a = agg[ln]
        n = int(r[ie] or 0)
        if op in ('DADD', 'DMUL', 'DFMA'):
            a[0] += n
        else:
            a[1] += n
        a[2] += int(r[ns] or 0)
    src = open('audian_b200/csrc/%s.cu' % stem.split('.')[0]).read().splitlines()
    tot = sum(a[0] + a[1] for a in agg.values())
    print(f'{"line":>5s} {"fp64":>10s} {"other":>10s} {"samples":>8s}  source')
    for ln in sorted(agg):
        a = agg[ln]
        if a[0] + a[1] < tot*0.002 and a[2] < 20:
            continue
        s = src[ln - 1].strip()[:90] if 0 < ln <= len(src) else ''
        print(f'{ln:5d} {a[0]:10d} {a[1]:10d} {a[2]:8d}  {s}')


if __name__ == '__main__':
    main()
