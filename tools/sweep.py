#!/usr/bin/env python
"""BASELINE config 5: parameter sweep nfft 128..16384 x overlap 0..87.5 % x filter cut-offs on
64-channel 250 kHz array data (device-resident, one table row per combination).

    python tools/sweep.py [--C 64] [--rate 250000] [--seconds 4] [--out file.json]
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


def timed(fn, reps=5, batches=3):
    """Mean ms per call of the best of `batches` batches of `reps` calls (one warm-up call)."""
    fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(batches):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/reps
        best = ms if best is None else min(best, ms)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--C', type=int, default=64)
    ap.add_argument('--rate', type=float, default=250000.)
    ap.add_argument('--seconds', type=float, default=4.)
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    _lib.init(0)
    C, rate = a.C, a.rate
    n = int(rate*a.seconds)
    xs = [device.synth(i*n, n, C, rate, 0xA0D1A9 + 5) for i in range(2)]
    peak = 6546.2
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    rows = []
    k = [0]

    def nxt():
        k[0] += 1
        return xs[k[0] % 2]
    for nfft in (128, 256, 512, 1024, 2048, 4096, 8192, 16384):
        for div in (1, 2, 4, 8):
            hop = nfft//div
            nf = (n - (nfft - hop))//hop
            F = nfft//2 + 1
            if nf*C*F*8 > 24e9:
                continue
            out = torch.empty((nf, C, F), dtype=torch.float64, device='cuda')
            ms = timed(lambda: device.spectrogram(nxt(), rate, nfft, hop, nf, out=out), 3, 2)
            del out
            bps = 8.0 + 8.0*F/hop
            # second floor: the fp64 pipe.  Flops of a frame by the usual count for a real
            # transform (2.5 N log2 N) + window, mean and power (2 N + 3 F), at the B200's nominal
            # 64 DFMA / clock / SM x 148 SMs x 1.965 GHz x 2 flops (every operation counted as
            # half an FMA: optimistic for the kernel, i.e. a floor)
            flops = nf*C*(2.5*nfft*np.log2(nfft) + 2*nfft + 3*F)
            fp64_ms = flops/(64*148*1.965e9*2)*1e3
            hbm_ms = n*C*bps/peak/1e6
            rows.append({'op': 'spectrogram', 'nfft': nfft, 'overlap': 1 - 1/div, 'ms': ms,
                         'gsamples_s': n*C/ms/1e6, 'alg_gbs': n*C*bps/ms/1e6,
                         'frac': n*C*bps/ms/1e6/peak,
                         'hbm_floor_ms': hbm_ms, 'fp64_floor_ms': fp64_ms,
                         'bound': 'fp64' if fp64_ms > hbm_ms else 'hbm',
                         'frac_of_binding_floor': max(hbm_ms, fp64_ms)/ms})
            print(json.dumps(rows[-1]), flush=True)
    out = torch.empty((n, C), dtype=torch.float64, device='cuda')
    for hp, lp, order in ((0, 20000., 2), (1000., 60000., 2), (5000., rate/2, 2), (1000., 60000., 4)):
        if hp > 0 and lp < rate/2:
            sos = butter(order, (hp, lp), 'bandpass', fs=rate, output='sos')
        elif hp > 0:
            sos = butter(order, hp, 'highpass', fs=rate, output='sos')
        else:
            sos = butter(order, lp, 'lowpass', fs=rate, output='sos')
        ms = timed(lambda: device.sosfilt(sos, nxt(), 0, out=out))
        rows.append({'op': 'filter', 'highpass': hp, 'lowpass': lp, 'order': order, 'sections': int(sos.shape[0]),
                     'ms': ms, 'gsamples_s': n*C/ms/1e6, 'alg_gbs': n*C*16/ms/1e6, 'frac': n*C*16/ms/1e6/peak})
        print(json.dumps(rows[-1]), flush=True)
    esos = butter(2, 500., 'lowpass', fs=rate, output='sos')
    ms = timed(lambda: device.envelope(esos, nxt(), 0, True, out=out))
    rows.append({'op': 'envelope', 'cutoff': 500., 'ms': ms, 'gsamples_s': n*C/ms/1e6,
                 'alg_gbs': n*C*16/ms/1e6, 'frac': n*C*16/ms/1e6/peak})
    print(json.dumps(rows[-1]), flush=True)
    ms = timed(lambda: device.minmax(nxt(), max(1, n//2000)))
    rows.append({'op': 'minmax', 'step': max(1, n//2000), 'ms': ms, 'gsamples_s': n*C/ms/1e6,
                 'alg_gbs': n*C*8/ms/1e6, 'frac': n*C*8/ms/1e6/peak})
    print(json.dumps(rows[-1]), flush=True)
    if a.out:
        json.dump({'C': C, 'rate': rate, 'frames': n, 'peak_gbs': peak, 'rows': rows}, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
