#!/usr/bin/env python
"""Whole-file passes of BASELINE configs 3 and 4 on synthetic recordings generated on the device.

    python tools/wholefile_bench.py --config 3|4 [--seconds S] [--chunk-frames N]
    torchrun --nproc-per-node N tools/wholefile_bench.py --gpus N --config 4

config 3: 16 ch x 500 kHz x 30 min, full-file spectrogram nfft 1024 / hop 512, time-sharded;
          the frames are reduced on the device to the mean power spectrum per channel
          (spectrogramplot.py:158) instead of being stored (115 GB).
config 4:  4 ch x 96 kHz x 24 h, full-trace min/max rows (max_pixel 6000) + order-4 Butterworth
          band-pass in one pass over the data; the filtered trace is reduced to its own
          min/max rows.
--seconds shortens the recording (default: the full length).  One JSON line per run: Msamples/s
over all ranks (device timing, max over ranks) and a parity spot check of sampled windows
against the CPU oracle.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
from scipy.signal import butter


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', type=int, default=4)
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--seconds', type=float, default=None)
    ap.add_argument('--chunk-frames', type=int, default=None)
    a = ap.parse_args()
    from audian_b200 import _lib, device
    from audian_b200.wholefile import WholeFile
    from audian_b200.synth import synth
    from oracle import oracle as orc
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    _lib.init(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', rank=rank, world_size=world,
                                device_id=torch.device('cuda', local))
    if a.config == 3:
        C, rate, seconds, seed = 16, 500000., 1800., 0xA0D1A9 + 3
    else:
        C, rate, seconds, seed = 4, 96000., 86400., 0xA0D1A9 + 4
    if a.seconds:
        seconds = a.seconds
    frames = int(rate*seconds)
    chunk = a.chunk_frames or (1 << 26)//C            # 0.5 GB of float64 per chunk
    buf = torch.empty((chunk + 2048, C), dtype=torch.float64, device='cuda')

    def source(t0, n):
        return device.synth(t0, n, C, rate, seed, out=buf[:n])

    wf = WholeFile(source, frames, C, rate, None, rank, world, dist, chunk_frames=chunk)
    check = {}
    # warm-up on one chunk of rank-local data: plans, scratch and the allocator's pools
    warm = WholeFile(source, min(frames, 2*chunk), C, rate, None, 0, 1, None, chunk_frames=chunk)
    if a.config == 3:
        warm.spectrogram(1024, 512, lambda k, P: P.sum(dim=0))
    else:
        wsos = butter(4, (1000., 15000.), 'bandpass', fs=rate, output='sos')
        wstep = max(1, frames//6000)
        warm.fulltrace_and_filter(wsos, wstep, lambda t0, y: device.minmax(y, wstep))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if a.config == 3:
        nfft, hop = 1024, 512
        acc = torch.zeros((C, nfft//2 + 1), dtype=torch.float64, device='cuda')
        keep = {}

        def sink(k, P):
            acc.add_(P.sum(dim=0))
            if k == 0 and rank == 0:
                keep['first'] = P[:8].clone()
        nf = wf.spectrogram(nfft, hop, sink)
        if world > 1:
            dist.all_reduce(acc)
        result = (acc/nf)
    else:
        sos = butter(4, (1000., 15000.), 'bandpass', fs=rate, output='sos')
        step = max(1, frames//6000)
        frows = []

        def sink(t0, y):
            frows.append(device.minmax(y, step) if t0 % step == 0 else None)
        rows = wf.fulltrace_and_filter(sos, step, sink)
        result = rows
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        # parity spot checks on the host (not timed)
        if a.config == 3:
            n = 7*hop + nfft
            x = synth(0, n, C, rate, seed)
            ref = np.empty((8, C, nfft//2 + 1))
            orc.spectrogram_process(x, ref, rate, nfft, hop)
            got = keep['first'].cpu().numpy()
            check['first_frames_max_rel_err'] = float(np.max(np.abs(got - ref)/np.maximum(ref, 1e-20*ref.max())))
            bps = 8.0 + 8.0*(nfft//2 + 1)/hop
        else:
            n = min(frames, 20*step)
            x = synth(0, n, C, rate, seed)
            ref = orc.minmax_rows(x, step)
            got = result[:len(ref)].cpu().numpy()
            check['fulltrace_rows_bit_exact'] = bool(np.array_equal(got.view(np.uint64), ref.view(np.uint64)))
            m = min(frames, 400000)
            yref = np.empty((m, C))
            orc.filter_process(sos, synth(0, m, C, rate, seed), yref, 0)
            y0 = device.sosfilt(sos, device.synth(0, m, C, rate, seed), 0)
            check['filter_max_abs_err_first_rows'] = float(np.max(np.abs(y0.cpu().numpy() - yref)))
            bps = 8.0 + 16.0
        samples = frames*C
        print(json.dumps({'config': a.config, 'channels': C, 'rate_hz': rate, 'seconds': seconds,
                          'frames': frames, 'n_gpus': world, 'chunk_frames': chunk,
                          'ms': ms, 'msamples_s': samples/ms/1e3,
                          'alg_gbs': samples*bps/ms/1e6, 'bytes_per_sample': bps,
                          'includes': 'on-device generation of the input (8 B/sample written + read back)',
                          'check': check}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
