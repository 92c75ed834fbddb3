#!/usr/bin/env python
"""Per-block timing of the pipelined zero-phase kernel (development; needs the -DADN_ZP_TIMING build:
tools/build_alt.sh zt zerophase.cu -DADN_ZP_TIMING; ADN_LIB=audian_b200/libaudian_b200_zt.so)."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scipy.signal import butter
from audian_b200 import _lib, device

C, rate, seconds = 8, 48000., 80.
_lib.init(0)
n = int(rate*seconds)
x = device.synth(0, n, C, rate, 1)
sos = butter(2, 500., 'lowpass', fs=rate, output='sos')
for _ in range(3):
    y = device.envelope(sos, x)
torch.cuda.synchronize()
lib = _lib.lib()
nb = 148
out = (ctypes.c_ulonglong*(2*1024*8))()
for rep in range(3):
    y = device.envelope(sos, x)
    y = device.envelope(sos, x)
    rc = lib.adn_debug_zp_times(out, 2*1024*8)
    full = np.array(out[:], dtype=np.int64).reshape(2, 1024, 8)[:, :nb]
    # which parity came first?
    first = 0 if full[0, :, 0].min() < full[1, :, 0].min() else 1
    A, B = full[first], full[1 - first]
    t0 = A[:, 0].min()
    for name, a in (('launch k', A), ('launch k+1', B)):
        start = (a[:, 0] - t0)/1e3
        pro = (a[:, 1] - a[:, 0])/1e3
        ends = (a[:, 2:5] - t0)/1e3
        dur = ends.max(axis=1) - start
        print(name, 'start us min/max %.1f %.1f' % (start.min(), start.max()), 'prologue med %.1f max %.1f' % (np.median(pro), pro.max()),
              'end min/med/max %.1f %.1f %.1f' % (ends.max(axis=1).min(), np.median(ends.max(axis=1)), ends.max()),
              'dur med/max %.1f %.1f' % (np.median(dur), dur.max()))
        order = np.argsort(dur)
        print('  slowest (blk, smid, dur):', [(int(b), int(a[b, 6]), round(float(dur[b]), 1)) for b in order[-5:]])
        rot = int(os.environ.get('ADN_ZP_ROT', '0'))
        b0 = (nb - rot) % nb
        print('  block 0: %.1f us; block of run 0 (blk %d, sm %d): %.1f us' % (dur[0], b0, a[b0, 6], dur[b0]))
    print('  gap end(k) -> start(k+1): %.1f us; period start->start %.1f us' % ((B[:, 0].min() - A[:, 2:5].max())/1e3, (B[:, 0].min() - A[:, 0].min())/1e3))
tt = (ctypes.c_ulonglong*(4*64))()
lib.adn_debug_zp_tile_times(tt)
tt = np.array(tt[:], dtype=np.int64).reshape(4, 64)
for r in range(4):
    row = tt[r]
    ok = row > 0
    rel = (row - row[ok].min())/1e3
    print('run', r, 'tile end times (us since the first end of the run), tiles a+0..:', [round(float(v), 1) if o else None for v, o in zip(rel[:56], ok[:56])])
