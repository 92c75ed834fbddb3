#!/usr/bin/env python
"""Whole-file passes of BASELINE configs 3 and 4 on synthetic recordings generated on the device.

    python tools/wholefile_bench.py --config 3|4 [--seconds S] [--budget-s B] [--chunk-frames N]
    torchrun --nproc-per-node N tools/wholefile_bench.py --gpus N --config 4

config 3: 16 ch x 500 kHz x 30 min, full-file spectrogram nfft 1024 / hop 512, time-sharded;
          the frames are reduced on the device to the mean power spectrum per channel
          (spectrogramplot.py:158) instead of being stored (115 GB); all-reduced over the ranks.
config 4:  4 ch x 96 kHz x 24 h, full-trace min/max rows (max_pixel 6000, gathered to rank 0) +
          order-4 Butterworth band-pass in one pass over the data; the filtered trace is reduced
          to its own min/max rows.
The recording is processed at its full length unless that would take longer than the time
budget (estimated from a probe of two chunks; every rank uses the same, all-reduced estimate):
then it is shortened to what fits and the line says so (`seconds` < `full_seconds`).  One dict per
run: Msamples/s over all ranks (device timing, max over ranks), roofline fraction of the
algorithmic bytes, parity of sampled windows (the first rows / frames of EVERY rank's range,
i.e. every seam between shards) against the CPU oracle.

`run_config()` is what `bench.py` calls for its `wholefile` key.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

CONFIGS = {
    3: dict(channels=16, rate=500000., seconds=1800., seed=0xA0D1A9 + 3, nfft=1024, hop=512,
            what='full-file spectrogram nfft1024/hop512 -> mean power spectrum',
            bytes_per_sample=8.0 + 8.0*513/512,
            # what the chunk loop really moves through HBM: the generated chunk is written (8), read by the
            # spectrogram kernel (8), the PSD chunk written (8 x 513/512) and read again by the column sums
            moved_bytes_per_sample=16.0 + 16.0*513/512),
    4: dict(channels=4, rate=96000., seconds=86400., seed=0xA0D1A9 + 4, max_pixel=6000,
            what='full-trace min/max (6000 px) + band-pass 1-15 kHz order 4 + min/max of the result',
            bytes_per_sample=8.0 + 16.0,
            # generated chunk written (8), read once by the fused filter + min/max pass (8), filtered written (8)
            moved_bytes_per_sample=24.0),
}


def run_config(config, rank=0, world=1, dist=None, seconds=None, budget_s=None, chunk_frames=None,
               peak_gbs=None):
    import torch
    from scipy.signal import butter
    from audian_b200 import _lib, device
    from audian_b200.wholefile import WholeFile
    from audian_b200.sharded import shard_bounds
    from audian_b200.synth import synth
    from oracle import oracle as orc

    cfg = CONFIGS[config]
    C, rate, seed = cfg['channels'], cfg['rate'], cfg['seed']
    full_seconds = cfg['seconds']
    if seconds is None:
        seconds = full_seconds
    chunk = chunk_frames or (1 << 26)//C            # 0.5 GB of float64 per chunk
    buf = torch.empty((chunk + 2048, C), dtype=torch.float64, device='cuda')

    def source(t0, n):
        return device.synth(t0, n, C, rate, seed, out=buf[:n])

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if config == 3:
        nfft, hop = cfg['nfft'], cfg['hop']
        F = nfft//2 + 1
        align = hop
    else:
        sos = butter(4, (1000., 15000.), 'bandpass', fs=rate, output='sos')
        align = 1

    def one_pass(frames, keep):
        """The timed pass over a recording of `frames` rows; returns the reduced result."""
        wf = WholeFile(source, frames, C, rate, None, rank, world, dist, chunk_frames=chunk)
        if config == 3:
            acc = torch.zeros((C, F), dtype=torch.float64, device='cuda')
            first = {}

            def sink(k, P):
                device.colsum(P, acc)
                if 'k' not in first:
                    first['k'] = k
                    keep['frames'] = (k, P[:8].clone())
            nf = wf.spectrogram(nfft, hop, sink)
            if world > 1:
                dist.all_reduce(acc)
            return acc/max(nf, 1)
        step = max(1, frames//cfg['max_pixel'])
        first = {}

        def sink(t0, y):
            if 't0' not in first:
                first['t0'] = t0
                keep['filtered'] = (t0, y[:200000].clone())
        # min/max of the raw rows and of the filtered rows in the filter's own pass over the chunk
        rows, frows = wf.fulltrace_filter_minmax(sos, step, sink)
        keep['step'] = step
        keep['filtered_rows'] = frows
        return rows

    # ---- probe: two chunks per rank (plans, scratch, allocator pools) -> rate estimate
    probe_frames = min(int(rate*seconds), 2*chunk*world)
    probe_frames = max(align, probe_frames//align*align)
    one_pass(probe_frames, {})
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    one_pass(probe_frames, {})
    e1.record()
    torch.cuda.synchronize()
    probe_ms = allmax(e0.elapsed_time(e1))
    frames = int(rate*seconds)
    est_s = frames/probe_frames*probe_ms*1e-3
    if budget_s is not None and est_s > budget_s:
        # shortened to what fits the budget: whole chunks for every rank
        unit = chunk*world
        frames = int(probe_frames*budget_s/(probe_ms*1e-3))
        frames = max(unit, frames//unit*unit)
    frames = min(frames, int(rate*full_seconds))
    seconds = frames/rate

    # ---- the timed pass
    keep = {}
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = ev(), ev()
    e0.record()
    result = one_pass(frames, keep)
    e1.record()
    torch.cuda.synchronize()
    ms = allmax(e0.elapsed_time(e1))
    launches = _lib.launch_count() - l0

    # ---- parity of sampled windows: the first frames / rows of this rank's range (= the seam to
    # its left neighbour), against the oracle on host-generated input; worst case over the ranks
    check = {}
    if config == 3:
        k, P = keep['frames']
        nchk = P.shape[0]
        x = synth(k*hop, (nchk - 1)*hop + nfft, C, rate, seed)
        ref = np.empty((nchk, C, F))
        orc.spectrogram_process(x, ref, rate, nfft, hop)
        got = P.cpu().numpy()
        err = float(np.max(np.abs(got - ref)/np.maximum(ref, 1e-20*ref.max())))
        check['seam_frames_max_rel_err'] = allmax(err)
        check['tolerance'] = 'spectrogram power rtol 1e-5'
        check['ok'] = check['seam_frames_max_rel_err'] <= 1e-5
        if rank == 0:
            check['mean_spectrum_finite'] = bool(torch.isfinite(result).all().item())
    else:
        step = keep['step']
        t0, y = keep['filtered']
        m = y.shape[0]
        pre = min(t0, 48000)                      # the cascade has forgotten its state long before
        xr = synth(t0 - pre, pre + m, C, rate, seed)
        yref = np.empty((pre + m, C))
        orc.filter_process(sos, xr, yref, 0)
        err = float(np.max(np.abs(y.cpu().numpy() - yref[pre:])))
        check['seam_filter_max_abs_err'] = allmax(err)
        ok_rows = True
        if rank == 0:
            # rows of the first two segments and of the two segments around the first seam
            bounds = shard_bounds(frames, world, step)
            segs = [0, 1]
            if world > 1:
                j = bounds[1][0]//step
                segs += [j - 1, j]
            got = result.cpu().numpy()
            for j in segs:
                a, b = j*step, min(frames, (j + 1)*step)
                if a >= b:
                    continue
                ref = orc.minmax_rows(synth(a, b - a, C, rate, seed), step)
                ok_rows = ok_rows and bool(np.array_equal(got[2*j:2*j + 2].view(np.uint64),
                                                          ref.view(np.uint64)))
            check['fulltrace_rows_checked'] = len(segs)
            # the rows of the filtered recording against the oracle's filter of the first segment
            frows = keep['filtered_rows'].cpu().numpy()
            m1 = min(frames, step)
            yr = np.empty((m1, C))
            orc.filter_process(sos, synth(0, m1, C, rate, seed), yr, 0)
            rref = orc.minmax_rows(yr, step)
            check['filtered_rows_max_abs_err'] = float(np.max(np.abs(frows[:2] - rref[:2])))
        check['fulltrace_rows_bit_exact'] = allmax(0.0 if ok_rows else 1.0) == 0.0
        check['tolerance'] = 'min/max bit-exact; filter max abs err 1e-6 of full scale'
        check['ok'] = bool(check['fulltrace_rows_bit_exact'] and check['seam_filter_max_abs_err'] <= 1e-6 and
                           check.get('filtered_rows_max_abs_err', 0.0) <= 1e-6)
    samples = frames*C
    bps = cfg['bytes_per_sample']
    out = {'config': config, 'workload': cfg['what'], 'channels': C, 'rate_hz': rate,
           'seconds': seconds, 'full_seconds': full_seconds, 'full_length': seconds >= full_seconds,
           'frames': frames, 'n_gpus': world, 'chunk_frames': chunk, 'ms': ms,
           'msamples_s': samples/ms/1e3, 'bytes_per_sample': bps, 'alg_gbs': samples*bps/ms/1e6,
           'gpu_launches': int(launches), 'scaling': 'strong (the recording is split over the ranks)',
           'gathered': 'mean power spectrum all-reduced' if config == 3 else 'min/max rows gathered to rank 0',
           'includes': 'on-device generation of the input inside the timed region (+16 B/sample of traffic)',
           'parity': check}
    if peak_gbs:
        out['roofline_frac'] = out['alg_gbs']/world/peak_gbs
        # the same time against the bytes the chunk loop moves, generation of the input included
        out['moved_bytes_per_sample'] = cfg['moved_bytes_per_sample']
        out['moved_frac'] = samples*cfg['moved_bytes_per_sample']/ms/1e6/world/peak_gbs
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', type=int, default=4)
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--seconds', type=float, default=None)
    ap.add_argument('--budget-s', type=float, default=None)
    ap.add_argument('--chunk-frames', type=int, default=None)
    a = ap.parse_args()
    import torch
    from audian_b200 import _lib
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
    res = run_config(a.config, rank, world, dist, a.seconds, a.budget_s, a.chunk_frames)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
