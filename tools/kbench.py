#!/usr/bin/env python
"""Single-kernel micro-benchmark on device-resident data (development tool).

    python tools/kbench.py --op filter|envelope|spectrogram|minmax [--C 8] [--rate 48000]
        [--seconds 80] [--order 2] [--nfft 1024] [--hop 512] [--step 2000] [--steps 10]

Prints one JSON line: mean ms per call (CUDA events), Gsamples/s and algorithmic GB/s.
Rotates over 3 input windows so that no call finds its input in L2.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
from scipy.signal import butter

from audian_b200 import _lib, device


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--op', default='filter')
    ap.add_argument('--C', type=int, default=8)
    ap.add_argument('--rate', type=float, default=48000.)
    ap.add_argument('--seconds', type=float, default=80.)
    ap.add_argument('--order', type=int, default=2)
    ap.add_argument('--kind', default='bandpass')
    ap.add_argument('--nfft', type=int, default=1024)
    ap.add_argument('--hop', type=int, default=512)
    ap.add_argument('--step', type=int, default=2000)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--onepass', type=int, default=1, help='0: envelope as two sweeps through memory')
    ap.add_argument('--env-cutoff', type=float, default=500.0)
    a = ap.parse_args()
    _lib.init(0)
    _lib.set_option(_lib.ADN_OPT_ZERO_PHASE_ONEPASS, a.onepass)
    n = int(a.rate*a.seconds)
    C = a.C
    xs = [device.synth(i*n, n, C, a.rate) for i in range(3)]
    if a.kind == 'bandpass':
        sos = butter(a.order, (0.02*a.rate, 0.3*a.rate), 'bandpass', fs=a.rate, output='sos')
    else:
        sos = butter(a.order, 0.1*a.rate, a.kind, fs=a.rate, output='sos')
    esos = butter(a.order, a.env_cutoff*a.rate/48000., 'lowpass', fs=a.rate, output='sos')
    out = torch.empty((n, C), dtype=torch.float64, device='cuda')
    nsp = (n - (a.nfft - a.hop))//a.hop
    if a.op == 'spectrogram':
        sp = torch.empty((nsp, C, a.nfft//2 + 1), dtype=torch.float64, device='cuda')

    def call(i):
        x = xs[i % 3]
        if a.op == 'filter':
            device.sosfilt(sos, x, 0, out=out)
        elif a.op == 'envelope':
            device.envelope(esos, x, 0, True, out=out)
        elif a.op == 'spectrogram':
            device.spectrogram(x, a.rate, a.nfft, a.hop, nsp, out=sp)
        elif a.op == 'minmax':
            device.minmax(x, a.step)
        elif a.op == 'state':
            device.sosfilt(sos, x, 0, state_only=True)
        else:
            raise SystemExit('unknown op')

    for i in range(a.warmup):
        call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        call(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/a.steps
    bps = {'filter': 16.0, 'envelope': 16.0, 'minmax': 8.0, 'state': 8.0,
           'spectrogram': 8.0 + 8.0*(a.nfft//2 + 1)/a.hop}[a.op]
    print(json.dumps({'op': a.op, 'C': C, 'n': n, 'S': int(sos.shape[0]), 'nfft': a.nfft,
                      'hop': a.hop, 'ms': ms, 'gsamples_s': n*C/ms/1e6,
                      'alg_gbs': n*C*bps/ms/1e6}))


if __name__ == '__main__':
    main()
