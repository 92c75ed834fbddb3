#!/usr/bin/env python
"""Static SASS summary of the library: per kernel, instruction count and the mnemonics that show
which hardware paths it uses (TMA bulk copies UBLKCP / bulk L2 prefetch UBLKPF, mbarrier SYNCS,
Ampere-style LDGSTS, named barriers, the fp64 pipe).

    python tools/sass_summary.py [audian_b200/libaudian_b200.so] > profiles/rNN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

KEYS = ['UTMALDG', 'UBLKCP', 'UBLKPF', 'SYNCS', 'LDGSTS', 'BAR', 'DFMA', 'DADD', 'DMUL', 'LDS', 'STS', 'SHFL', 'LDG', 'STG',
        'LDL', 'STL']


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), '..', 'audian_b200',
                                                              'libaudian_b200.so')
    txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    kern = None
    counts = collections.OrderedDict()
    for line in txt.splitlines():
        m = re.match(r'\s+Function : (\S+)', line)
        if m:
            kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r'adn::\(anonymous namespace\)::|adn::', '', kern)
            kern = re.sub(r'\(.*$', '', kern)
            counts[kern] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and kern:
            counts[kern]['total'] += 1
            counts[kern][m.group(1)] += 1
    print('SASS of', os.path.basename(lib), '(sm_100a): static instruction counts per kernel')
    print(f'{"kernel":58s} {"total":>6s} ' + ' '.join(f'{k:>6s}' for k in KEYS))
    for k, c in counts.items():
        if c['total'] < 400:
            continue
        print(f'{k[:58]:58s} {c["total"]:6d} ' + ' '.join(f'{c[x]:6d}' for x in KEYS))


if __name__ == '__main__':
    main()
