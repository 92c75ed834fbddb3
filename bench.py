#!/usr/bin/env python
"""Benchmark of the derived-trace DSP path (BASELINE.json metric:
Msamples/s spectrogram+filter+envelope; % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): synthetic 8-channel 48 kHz recording,
one step = one visible-window update of audian's 80-s buffer (3 840 000
frames x 8 channels = 30.72 M samples, 246 MB fp64):
    data -> filtered   Butterworth band-pass 1-15 kHz, order 2 (runaudian.py:4)
         -> spectrogram nfft 1024, 50 % overlap, PSD (BASELINE configs[0])
         -> envelope    (pi/2)|x| zero-phase low-pass 500 Hz (bufferedenvelope.py:15)
`value`  device-resident inputs/outputs, CUDA events, max over ranks.
`e2e`    the same step through the host-array plugin path (BufferedFilter /
         BufferedSpectrogram / BufferedEnvelope.process on pinned numpy
         buffers), host<->device copies inside the timed region.
N > 1    weak scaling: an N x 80 s recording time-sharded over N GPUs
         (audian_b200.sharded.HaloChain): every rank holds its 80-s shard plus the
         halo rows the STFT and the two cascades need (decay length to 1e-20), so the
         step needs no exchange; results equal one pass over the whole recording
         (checked at every seam, `parity`).
Every N runs the same code: eager launches, CUDA events around each op inside the
timed region (op_ms sums to the step).  One sample = one channel-sample of input.
Successive steps use different windows of the recording; every window (246 MB)
exceeds the 126 MB L2.
Besides the line's headline it carries: `minmax` (the fourth op of the path, timed on
its own), `cpu_baseline` (the oracle on the host: the chain and the full-trace pass with
the reference's own worker scheme), `parity` at every N, and `wholefile` -- BASELINE
configs 3 and 4 streamed through audian_b200.wholefile, strong-scaled over the ranks.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATE = 48000.0
CHANNELS = 8
WINDOW_S = 80.0
FRAMES = int(RATE*WINDOW_S)
HIGHPASS, LOWPASS, ORDER = 1000.0, 15000.0, 2
NFFT, HOP = 1024, 512
ENV_CUTOFF = 500.0
N_WINDOWS = 4                      # distinct windows rotated through the steps
MAX_PIXEL = 1920                   # full-trace plot of the 1-h recording: step = frames // max_pixel
RECORDING_S = 3600.0
WINDOW_STRIDE_S = 450.0            # offsets inside the 1-h recording
SEED = 0xA0D1A9 + 2

WORKLOAD = ('synthetic 8ch 48kHz 1h recording, visible-window update of the 80-s buffer: '
            'bandpass 1-15kHz o2 -> spectrogram nfft1024/hop512 + envelope 500Hz')


def algorithmic_bytes(op, frames, channels):
    """SURVEY.md 8(d) bytes per input sample x samples of one launch."""
    n = frames*channels
    if op == 'filter':
        return 16.0*n
    if op == 'spectrogram':
        return (8.0 + 8.0*(NFFT//2 + 1)/HOP)*n
    if op == 'envelope':             # one pass: every row read once, written once
        return 16.0*n
    if op == 'minmax':
        return 8.0*n
    raise KeyError(op)


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


def designs():
    from scipy.signal import butter
    sos = butter(ORDER, (HIGHPASS, LOWPASS), 'bandpass', fs=RATE, output='sos')
    esos = butter(2, ENV_CUTOFF, 'lowpass', fs=RATE, output='sos')
    return sos, esos


class ClockSampler(threading.Thread):
    """SM clock and clock-event (throttle) reasons of one GPU while the timed region runs.

    Polls NVML in this process (nvidia_ml_py) every millisecond or so: the timed region of the
    default run lasts some 10 ms, far below what an `nvidia-smi -lms` child process resolves
    (its first row arrives after the region has ended).  Falls back to one-shot `nvidia-smi`
    queries when NVML cannot be loaded."""

    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index = index
        self.uuid = uuid
        self.rows = []                  # (sm_mhz, reasons bitmask or list)
        self.mx = None
        self.power = []
        self.how = None
        self._stop_evt = threading.Event()
        self._ready = threading.Event()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        if self.uuid:
            for u in (self.uuid, self.uuid.encode()):
                try:
                    return pynvml, pynvml.nvmlDeviceGetHandleByUUID(u)
                except Exception:
                    pass
        vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
        idx = self.index
        try:
            if vis:
                idx = int(vis.split(',')[self.index])
        except Exception:
            pass
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def run(self):
        try:
            nv, h = self._nvml_handle()
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            masks = {'hw_slowdown': nv.nvmlClocksEventReasonHwSlowdown,
                     'hw_thermal_slowdown': nv.nvmlClocksEventReasonHwThermalSlowdown,
                     'sw_thermal_slowdown': nv.nvmlClocksEventReasonSwThermalSlowdown,
                     'sw_power_cap': nv.nvmlClocksEventReasonSwPowerCap}
            self.how = 'nvml'
            self._ready.set()
            while not self._stop_evt.is_set():
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                if len(self.rows) % 8 == 0:             # every NVML query takes about a millisecond
                    try:
                        self.power.append(nv.nvmlDeviceGetPowerUsage(h)/1000.0)
                    except Exception:
                        pass
                self.rows.append((time.perf_counter(), sm, [k for k, m in masks.items() if bits & m]))
                time.sleep(float(os.environ.get('BENCH_CLOCK_SLEEP', '0.0002')))
            return
        except Exception:
            pass
        # fallback: repeated one-shot nvidia-smi queries (each takes tens of milliseconds)
        self.how = 'nvidia-smi'
        self._ready.set()
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(
                    ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                     '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=10).stdout
                for line in out.strip().splitlines():
                    r = [c.strip() for c in line.split(',')]
                    self.mx = float(r[2])
                    self.rows.append((time.perf_counter(), float(r[1]),
                                      [nm for k, nm in enumerate(self.NAMES)
                                       if r[5 + k].lower().startswith('active')]))
            except Exception:
                time.sleep(0.05)

    def wait_ready(self, timeout=5.0):
        self._ready.wait(timeout)

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken in [t0, t1] (perf_counter times; all samples if None or if
        fewer than two fall inside)."""
        self._stop_evt.set()
        self.join(timeout=15)
        rows = self.rows
        if t0 is not None and t1 is not None:
            inside = [r for r in rows if t0 <= r[0] <= t1]
            if len(inside) >= 2:
                rows = inside
        sm = [r[1] for r in rows]
        reasons = set()
        for r in rows:
            reasons.update(r[2])
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': self.mx,
                'reasons': sorted(reasons), 'samples': len(sm), 'how': self.how,
                'power_w_max': max(self.power) if self.power else None}


# ------------------------------------------------------------------ reference arm

def cpu_chain(x, sos, esos):
    """The reference's CPU path for one step: the oracle = the scipy/numpy calls of
    BufferedFilter/Spectrogram/Envelope.process, single thread as in audian."""
    from oracle import oracle as orc
    n, C = x.shape
    filt = np.empty((n, C))
    orc.filter_process(sos, x, filt, 0)
    spec = np.empty((n//HOP, C, NFFT//2 + 1))
    orc.spectrogram_process(filt, spec, RATE, NFFT, HOP)
    env = np.empty((n, C))
    orc.envelope_process(esos, filt, env, 0, 0)
    return filt, spec, env


def _chain_worker(job):
    """One channel subset of the CPU chain in a worker process (top level: picklable under spawn)."""
    x, sos, esos, reps = job
    t0 = time.perf_counter()
    for _ in range(reps):
        cpu_chain(x, sos, esos)
    return time.perf_counter() - t0


def cpu_chain_parallel(x, sos, esos, reps):
    """The same oracle calls with every host thread the path can use: the channels are
    independent, so one worker process per channel (as audian's own full-trace workers are
    processes, compresseddata.py:107-122).  Returns (seconds per repetition, workers)."""
    import multiprocessing as mp
    C = x.shape[1]
    workers = max(1, min(C, os.cpu_count() or 1))
    cols = np.array_split(np.arange(C), workers)
    jobs = [(np.ascontiguousarray(x[:, c]), sos, esos, reps) for c in cols]
    ctx = mp.get_context('spawn')                # never fork a process that may hold a CUDA context
    with ctx.Pool(workers) as pool:
        pool.map(_chain_worker, [(j[0][:1000], sos, esos, 1) for j in jobs])   # imports, warm-up
        t0 = time.perf_counter()
        pool.map(_chain_worker, jobs)
        dt = time.perf_counter() - t0
    return dt/reps, workers


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from audian_b200.synth import synth
    sos, esos = designs()
    sample_s = 10.0                         # bounded sample of the 80-s window per step
    n = int(RATE*sample_s)
    x = synth(0, n, CHANNELS, RATE, SEED)
    # (1) as audian runs the path: one thread (every entry point is a Qt slot of the GUI thread)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_chain(x, sos, esos)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_chain(x, sos, esos)
    dt1 = (time.perf_counter() - t0)/args.steps
    single = n*CHANNELS/dt1/1e6
    # (2) with all the host threads the path can use: one process per channel.  This is the
    # headline of the reference arm (the harder baseline)
    dt, workers = cpu_chain_parallel(x, sos, esos, args.steps)
    value = n*CHANNELS/dt/1e6
    sample = f'{sample_s:g}-s slice of the 80-s window ({n} frames x {CHANNELS} ch) per step'
    line = {
        'impl': 'reference', 'metric': 'Msamples/s spectrogram+filter+envelope',
        'value': value, 'unit': 'Msamples/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt*1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': 'Msamples/s', 'cores': workers, 'kind': 'port',
                         'sample': sample,
                         'single_thread_value': single,
                         'note': 'oracle = the reference\'s scipy/numpy calls; value: one worker '
                                 'process per channel (all the parallelism the path admits); '
                                 'single_thread_value: one thread, as audian itself runs the path '
                                 '(GUI thread); host has %d cpus' % os.cpu_count()},
        'e2e': {'value': value, 'unit': 'Msamples/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    emit(line)


# ------------------------------------------------------------------ our arm

class gpu_near_cpus:
    """Binds this process to the CPUs NVML names as nearest to the GPU (the NUMA node its PCIe
    root hangs on) until restore(); a no-op when NVML gives no mask or the mask is every CPU."""

    def __init__(self, index):
        self.old = None
        self.cpus = None
        if os.environ.get('ADN_BENCH_NUMA', '1') == '0' or not hasattr(os, 'sched_getaffinity'):
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
            idx = int(vis.split(',')[index]) if vis else index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63)//64)
            cpus = {64*w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
            old = os.sched_getaffinity(0)
            cpus &= old
            if cpus and cpus != old:
                os.sched_setaffinity(0, cpus)
                self.old, self.cpus = old, cpus
        except Exception as exc:                       # pragma: no cover
            sys.stderr.write('gpu_near_cpus: %s\n' % (exc,))

    def describe(self):
        if self.cpus is None:
            return 'unbound (no NVML mask, or the mask is every CPU)'
        c = sorted(self.cpus)
        return '%d cpus near the GPU: %d..%d' % (len(c), c[0], c[-1])

    def restore(self):
        if self.old is not None:
            os.sched_setaffinity(0, self.old)
            self.old = None


def seam_parity(chain, sos, esos, C, filt_rows, spec, env, abs0, where):
    """Outputs of one rank's shard against the oracle at one end of the shard (`where` =
    'lo' | 'hi'): the oracle filters a host-generated window that starts 1 s before the rows
    compared (both cascades forget their state within a few thousand rows, so this equals the
    rows of one pass over the whole recording to rounding; at the ends of the recording the
    window starts / ends there and the oracle sees the true edge)."""
    from audian_b200.synth import synth
    from oracle import oracle as orc
    W, M = 48000, 20480
    lo, hi, frames = chain.lo, chain.hi, chain.frames
    if where == 'lo':
        a, b = lo, min(hi, lo + M)
    else:
        a, b = max(lo, hi - M), hi
    w0, w1 = max(0, a - W), min(frames, b + W)
    x = synth(abs0 + w0, w1 - w0, C, RATE, SEED)
    rf = np.empty_like(x)
    orc.filter_process(sos, x, rf, 0)
    re_ = np.empty_like(x)
    orc.envelope_process(esos, rf, re_, 0, 0)
    gf = filt_rows[a - lo:b - lo].cpu().numpy()
    ge = env[a - lo:b - lo].cpu().numpy()
    out = {'filter_max_abs_err': float(np.max(np.abs(gf - rf[a - w0:b - w0]))),
           'envelope_max_abs_err': float(np.max(np.abs(ge - re_[a - w0:b - w0])))}
    # frames that start in [a, b) and lie inside the window
    k0 = (a + HOP - 1)//HOP
    k1 = min(chain.k1, (b + HOP - 1)//HOP, (w1 - NFFT)//HOP + 1)
    if k1 > k0:
        seg = rf[k0*HOP - w0:(k1 - 1)*HOP + NFFT - w0]
        rs = np.empty((k1 - k0, C, NFFT//2 + 1))
        nn = orc.spectrogram_process(np.ascontiguousarray(seg), rs, RATE, NFFT, HOP)
        gs = spec[k0 - chain.k0:k0 - chain.k0 + nn].cpu().numpy()
        out['spectrogram_max_rel_err'] = float(np.max(np.abs(gs - rs[:nn])/np.maximum(rs[:nn], 1e-20*rs.max())))
    else:
        out['spectrogram_max_rel_err'] = 0.0
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from audian_b200 import _lib, device, sharded
    from audian_b200.bufferedfilter import BufferedFilter
    from audian_b200.bufferedspectrogram import BufferedSpectrogram
    from audian_b200.bufferedenvelope import BufferedEnvelope

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torch.distributed.run --nproc-per-node %d' % args.gpus)
    torch.cuda.set_device(local)
    _lib.init(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', rank=rank, world_size=world,
                                device_id=torch.device('cuda', local))
    sos, esos = designs()
    C, n = CHANNELS, FRAMES
    stride = int(RATE*WINDOW_STRIDE_S)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def allmax(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- geometry: the recording of a step is N x 80 s, rank r owns [r*80 s, (r+1)*80 s) and
    # holds the halo rows around it (none at the ends of the recording); no exchange
    frames_total = n*world
    bounds = [(r*n, (r + 1)*n) for r in range(world)]
    if not sharded.HaloChain.supported(sos, esos, bounds):
        raise SystemExit('the cascades of the bench forget fast: HaloChain must apply')
    chain = sharded.HaloChain(frames_total, RATE, C, bounds, rank, sos, esos, NFFT, HOP)
    r0, r1 = chain.raw_range()
    windows = [device.synth(w*stride + r0, r1 - r0, C, RATE, SEED) for w in range(N_WINDOWS)]
    filt_ext = torch.empty((chain.f1 - chain.f0, C), dtype=torch.float64, device='cuda')
    spec = torch.empty((chain.n_frames, C, NFFT//2 + 1), dtype=torch.float64, device='cuda')
    env = torch.empty((n, C), dtype=torch.float64, device='cuda')
    marks = []
    last = {}

    def step(i, record=False):
        x = windows[i % N_WINDOWS]
        e = [ev() for _ in range(4)] if record else None
        a = chain.lo - chain.f0
        if record:
            e[0].record()
        fext = device.sosfilt(sos, x, chain.f0 - chain.r0, out=filt_ext)
        if record:
            e[1].record()
        device.spectrogram(fext[a:], RATE, NFFT, HOP, chain.n_frames, out=spec)
        if record:
            e[2].record()
        device.zero_phase_range(esos, fext, a, n, chain.first, chain.last, True, True, out=env)
        if record:
            e[3].record()
            marks.append(e)
        last['filt'] = fext[a:a + n]

    for i in range(args.warmup):
        step(i)
    sync_all()
    try:
        uuid = 'GPU-' + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    launches0 = _lib.launch_count()
    sync_all()
    t_start, t_stop = ev(), ev()
    wall0 = time.perf_counter()
    torch.cuda.nvtx.range_push('timed')          # ncu --nvtx --nvtx-include "timed/"
    t_start.record()
    for i in range(args.steps):
        step(args.warmup + i, record=True)
    t_stop.record()
    sync_all()
    launches = _lib.launch_count() - launches0
    torch.cuda.nvtx.range_pop()
    wall1 = time.perf_counter()
    ms_step = allmax(t_start.elapsed_time(t_stop))/args.steps
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    samples_step = n*C*world
    value = samples_step/(ms_step*1e-3)/1e6

    # per-op times from the events recorded inside the timed region (this rank's)
    t_f = float(np.mean([e[0].elapsed_time(e[1]) for e in marks]))
    t_s = float(np.mean([e[1].elapsed_time(e[2]) for e in marks]))
    t_e = float(np.mean([e[2].elapsed_time(e[3]) for e in marks]))
    peak, peak_kind = measured_peak()

    # ---- the fourth op of the path, full-trace min/max of the raw window, timed on its own
    mm_step = max(1, int(RATE*RECORDING_S)//MAX_PIXEL)
    raw_own = [w[chain.lo - r0:chain.lo - r0 + n] for w in windows]
    for i in range(3):
        device.minmax(raw_own[i % N_WINDOWS], mm_step)
    sync_all()
    m0, m1 = ev(), ev()
    m0.record()
    for i in range(args.steps):
        mm_rows = device.minmax(raw_own[i % N_WINDOWS], mm_step)
    m1.record()
    sync_all()
    t_m = allmax(m0.elapsed_time(m1))/args.steps

    # ---- the same step as ONE device call (adn_chain_f64_dev), single-GPU geometry
    t_chain = None
    if world == 1:
        for i in range(3):
            device.chain(sos, windows[i % N_WINDOWS], filt_ext, RATE, 0, spec=spec, nfft=NFFT, hop=HOP,
                         esos=esos, env=env)
        sync_all()
        c0_, c1_ = ev(), ev()
        c0_.record()
        for i in range(args.steps):
            device.chain(sos, windows[i % N_WINDOWS], filt_ext, RATE, 0, spec=spec, nfft=NFFT, hop=HOP,
                         esos=esos, env=env)
        c1_.record()
        sync_all()
        t_chain = c0_.elapsed_time(c1_)/args.steps

    def roof(op, ms, frames, kernel, share_of=None):
        ach = algorithmic_bytes(op, frames, C)/(ms*1e-3)/1e9
        return {'kernel': kernel, 'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                'frac': ach/peak, 'traffic': None, 'peak_kind': peak_kind,
                'ms_per_launch': ms, 'launches_per_step': 1,
                'share_of_step': (ms/(t_f + t_s + t_e)) if share_of is None else share_of}

    scan_name = 'sos_run_kernel' if _lib.scan_run_count() > 0 else 'sos_scan_kernel'
    filt_name = ('sos_fwd_park_kernel<2,0,3>: pipelined, tile in registers' if _lib.fwd_park_count() > 0
                 else scan_name + '<S=2,FWD>')
    env_name = ('sos_zp_park_kernel<1,ENVF,3>: one pass, tile in registers' if _lib.zero_phase_count() > 0
                else scan_name + '<S=1,ENVF> + <S=1,REV>')
    roofs = {
        'filter': roof('filter', t_f, chain.f1 - chain.f0, filt_name),
        'spectrogram': roof('spectrogram', t_s, n, 'spectrogram_ring_kernel<10>'),
        'envelope': roof('envelope', t_e, n, env_name),
        'minmax': roof('minmax', t_m, n, 'minmax_split_kernel (step %d, not part of the step)' % mm_step, 0.0),
    }
    traffic_file = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.isfile(traffic_file):
        try:
            with open(traffic_file) as f:
                tr = json.load(f)
            for k in roofs:
                if k in tr:
                    roofs[k]['traffic'] = tr[k]
        except Exception:
            pass
    dominant = max(('filter', 'spectrogram', 'envelope'), key=lambda k: roofs[k]['share_of_step'])

    # ---- parity at every N: both ends of every shard (= every seam) against the oracle
    step(0)
    torch.cuda.synchronize()
    pr = {}
    for where in ('lo', 'hi'):
        one = seam_parity(chain, sos, esos, C, last['filt'], spec, env, 0*stride, where)
        for k, v in one.items():
            pr[k] = max(pr.get(k, 0.0), v)
    mm_ref = None
    from oracle import oracle as orc
    mm_got = device.minmax(raw_own[0], mm_step).cpu().numpy()
    mm_ref = orc.minmax_rows(raw_own[0].cpu().numpy(), mm_step)
    mm_ok = bool(np.array_equal(mm_got.view(np.uint64), mm_ref.view(np.uint64)))
    parity = {k: allmax(v) for k, v in sorted(pr.items())}
    parity['minmax_bit_exact'] = allmax(0.0 if mm_ok else 1.0) == 0.0
    parity['checked'] = ('first and last 20480 rows (and the frames inside them) of every rank\'s shard '
                         '= every seam and both ends of the %d x 80-s recording; min/max of the whole '
                         'window of every rank' % world)
    parity['tolerance'] = 'filter/envelope 1e-6 of full scale, spectrogram rtol 1e-5, min/max bit-exact'
    parity['ok'] = bool(parity['filter_max_abs_err'] <= 1e-6 and parity['envelope_max_abs_err'] <= 1e-6 and
                        parity['spectrogram_max_rel_err'] <= 1e-5 and parity['minmax_bit_exact'])

    # ---- end to end through the plugin path: host buffers, copies inside the timed region.
    # With N ranks: N independent replicas of the single-GPU plugin path (audian's interactive
    # updates stay on one GPU), each on its own 80-s window.
    own = slice(chain.lo - r0, chain.lo - r0 + n)
    # the host buffers of the plugin path are first touched (and page-locked) by threads on the CPUs
    # next to this rank's GPU, as a NUMA-aware host application would place them; restored below
    near = gpu_near_cpus(local)
    # page-locked arrays from the library's allocator, as the traces' allocate_buffer() makes them
    host_x = []
    for w in windows[:2]:
        a_ = _lib.pinned_empty((n, C))
        a_[:] = w[own].cpu().numpy()
        host_x.append(a_)
    nspec = n//HOP
    h_filt = _lib.pinned_empty((n, C))
    h_spec = _lib.pinned_empty((nspec, C, NFFT//2 + 1))
    h_env = _lib.pinned_empty((n, C))
    tf = BufferedFilter()
    tf.configure_standalone(RATE, C, highpass_cutoff=HIGHPASS, lowpass_cutoff=LOWPASS)
    ts = BufferedSpectrogram(nfft=NFFT, overlap_frac=0.5)
    ts.configure_standalone(RATE, C, source=tf)
    te = BufferedEnvelope(envelope_cutoff=ENV_CUTOFF)
    te.configure_standalone(RATE, C, source=tf)

    def host_step_traces(i):
        x = host_x[i % len(host_x)]
        tf.process(x, h_filt, 0)
        ts.process(h_filt, h_spec, 0)
        te.process(h_filt, h_env, 0)

    def host_step_chain(i):
        # the same three results through ONE call: what BufferedFilter.recompute_all() issues when
        # its dests are a spectrogram and an envelope (audian_b200/bufferedfilter.py)
        _lib.chain(sos, host_x[i % len(host_x)], h_filt, RATE, 0, spec=h_spec, nfft=NFFT, hop=HOP,
                   esos=esos, env=h_env, clamp_negative=True)

    def time_host(fn):
        ksteps = max(1, min(args.steps, 5))
        for i in range(2):
            fn(i)
        sync_all()
        moved0 = _lib.transfer_bytes()
        t0 = time.perf_counter()
        for i in range(ksteps):
            fn(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        moved1 = _lib.transfer_bytes()
        t = allmax(dt/ksteps)
        return {'value': samples_step/t/1e6, 'unit': 'Msamples/s',
                # counted by the library around every copy it issues
                'h2d_bytes_per_step': int((moved1[0] - moved0[0])//ksteps),
                'd2h_bytes_per_step': int((moved1[1] - moved0[1])//ksteps),
                'steps': ksteps, 'ms_per_step': t*1e3}

    e2e_traces = time_host(host_step_traces)
    e2e = time_host(host_step_chain)
    chk = np.empty_like(h_env)
    host_step_traces(0)
    chk[:] = h_env
    f_chk = h_filt.copy()
    host_step_chain(0)
    e2e['equals_separate_calls'] = bool(np.array_equal(f_chk, h_filt) and np.max(np.abs(chk - h_env)) <= 1e-12)
    e2e['api'] = ('adn_chain_f64 (audian_b200._lib.chain) on page-locked numpy buffers (adn_host_alloc, what the traces\' allocate_buffer() uses): the call '
                  'BufferedFilter.recompute_all() makes for filtered -> spectrogram + envelope; the source '
                  'goes up once, results come down while the next kernel runs'
                  + ('' if world == 1 else '; %d independent replicas, one per GPU' % world))
    e2e['separate_process_calls'] = dict(e2e_traces, api='BufferedFilter / BufferedSpectrogram / '
                                         'BufferedEnvelope.process one after the other, the filtered buffer '
                                         'handed over through the filter trace\'s device mirror')
    del chk, f_chk
    del tf, ts, te
    e2e['host_cpus'] = near.describe()
    near.restore()

    # ---- the oracle on the host's cores, bounded samples of the same workload (N = 1 only)
    cpu_baseline = None
    if rank == 0 and world == 1:
        sample_s = 10.0
        m = int(RATE*sample_s)
        xs = np.ascontiguousarray(host_x[0][:m])
        cpu_chain(xs[:m//4], sos, esos)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            cpu_chain(xs, sos, esos)
        dt = (time.perf_counter() - t0)/reps
        try:
            dtp, workers = cpu_chain_parallel(xs, sos, esos, reps)
        except Exception as exc:                       # pragma: no cover
            sys.stderr.write('parallel cpu baseline failed (%s)\n' % (exc,))
            dtp, workers = dt, 1
        cpu_baseline = {'value': m*C/dtp/1e6, 'unit': 'Msamples/s', 'cores': workers, 'kind': 'port',
                        'single_thread_value': m*C/dt/1e6,
                        'sample': f'{sample_s:g}-s slice of the 80-s window ({m} frames x {C} ch), '
                                  f'{reps} repetitions; value: one worker process per channel, '
                                  f'single_thread_value: one thread as audian runs the path; '
                                  f'host has {os.cpu_count()} cpus'}
        # full-trace min/max with the reference's own worker scheme (compresseddata.py:104-122):
        # cpu_count() - 1 processes, block-cyclic 30-s blocks, one shared array under its lock
        try:
            # the workers share the recording through /dev/shm: size it to what is free there
            free = os.statvfs('/dev/shm')
            room = free.f_bavail*free.f_frsize//3
            rows_max = int(min(4*n, room//(C*8)))
            if rows_max < int(RATE*35):
                raise RuntimeError('/dev/shm has room for %d rows only' % rows_max)
            data = np.concatenate(host_x + host_x, axis=0)[:rows_max]    # up to 320 s in memory
            mp_px = max(1, len(data)//mm_step)
            _, rows, dtm, nw = orc.fulltrace_parallel(data, mp_px, RATE)
            t0 = time.perf_counter()
            orc.fulltrace_long(data, mp_px, RATE, 1)
            dts = time.perf_counter() - t0
            cpu_baseline['minmax'] = {'value': data.size/dtm/1e6, 'unit': 'Msamples/s', 'cores': nw,
                                      'single_thread_value': data.size/dts/1e6,
                                      'sample': f'{len(data)/RATE:g} s x {C} ch in memory, step {len(data)//mp_px}: '
                                                f'the reference\'s worker scheme (os.cpu_count()-1 processes, '
                                                f'block-cyclic 30-s blocks, shared array + lock; block loops timed)'}
            del data
        except Exception as exc:                       # pragma: no cover
            sys.stderr.write('min/max cpu baseline failed (%s)\n' % (exc,))

    # ---- BASELINE configs 3 and 4: whole-file passes, strong-scaled over the ranks
    wholefile = None
    if args.wholefile:
        del windows, raw_own, filt_ext, spec, env
        last.clear()
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        import wholefile_bench
        wholefile = {}
        for cfg, key in ((3, 'c3'), (4, 'c4')):
            try:
                wholefile[key] = wholefile_bench.run_config(cfg, rank, world, dist if world > 1 else None,
                                                            None, args.wholefile_budget, None, peak)
            except Exception as exc:                   # pragma: no cover
                wholefile[key] = {'error': repr(exc)}
                if world > 1:
                    raise

    if rank == 0:
        line = {
            'metric': 'Msamples/s spectrogram+filter+envelope', 'value': value,
            'unit': 'Msamples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'channels': C, 'rate_hz': RATE,
                       'frames_per_step_per_gpu': n, 'samples_per_step': samples_step,
                       'filter': 'butter(2,(1000,15000),bandpass)', 'nfft': NFFT, 'hop': HOP,
                       'envelope_cutoff_hz': ENV_CUTOFF,
                       'l2': 'inputs larger than L2: each step reads a different 246 MB window',
                       'parallelism': 'single GPU' if world == 1 else
                                      f'time-sharded x{world}: every rank holds its shard plus halo rows '
                                      f'(filter {chain.keep_f}, envelope {chain.keep_e}, STFT {NFFT - HOP}); '
                                      f'no exchange in the step',
                       'launch': 'eager, CUDA events around each op inside the timed region'},
            'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks,
            'roofline': roofs[dominant], 'roofline_all': roofs, 'dominant': dominant,
            'op_ms': {'filter': t_f, 'spectrogram': t_s, 'envelope': t_e, 'minmax_separate': t_m,
                      'chain_one_call': t_chain},
            'minmax': {'value': n*C*world/(t_m*1e-3)/1e6, 'unit': 'Msamples/s', 'step': mm_step,
                       'ms': t_m, 'frac': roofs['minmax']['frac']},
            'cpu_baseline': cpu_baseline, 'parity': parity, 'wholefile': wholefile,
        }
        emit(line)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-wholefile', dest='wholefile', action='store_false',
                    help='skip the whole-file passes of BASELINE configs 3 and 4')
    ap.add_argument('--wholefile-budget', type=float,
                    default=float(os.environ.get('ADN_BENCH_WHOLEFILE_BUDGET_S', '70')),
                    help='seconds of run time per whole-file config (the recording is shortened to fit)')
    ap.add_argument('--no-graphs', dest='graphs', action='store_false', help='(ignored: every N runs eagerly)')
    args = ap.parse_args()
    # stdout carries the one JSON line and nothing else: whatever libraries print on file descriptor 1
    # while the bench runs (NCCL's version banner, for one) goes to stderr instead
    sys.stdout.flush()
    global _STDOUT_FD
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


_STDOUT_FD = None


def emit(line):
    """Prints the JSON line on the real stdout."""
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
