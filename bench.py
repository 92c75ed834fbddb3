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
         (audian_b200.sharded): IIR boundary states all-gathered (filter: one
         exchange, envelope: one per sweep), STFT halo exchanged over NCCL.
One sample = one channel-sample of input.  Successive steps use different
windows of the recording; every window (246 MB) exceeds the 126 MB L2.
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
    if op == 'envelope_sweep':       # one of the two scan launches of sosfiltfilt
        return 16.0*n
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
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h)/1000.0)
                except Exception:
                    pass
                self.rows.append((time.perf_counter(), sm, [k for k, m in masks.items() if bits & m]))
                time.sleep(0.001)
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
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ our arm

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

    # windows of the synthetic recording, generated on the device; with N ranks the
    # recording of a step is N x 80 s long and rank r owns [r*80 s, (r+1)*80 s)
    bounds = [(r*n, (r + 1)*n) for r in range(world)]
    windows = [device.synth(w*stride + rank*n, n, C, RATE, SEED) for w in range(N_WINDOWS)]
    nspec = n//HOP
    filt = torch.empty((n, C), dtype=torch.float64, device='cuda')
    spec = torch.empty((nspec, C, NFFT//2 + 1), dtype=torch.float64, device='cuda')
    env = torch.empty((n, C), dtype=torch.float64, device='cuda')
    ops = device.CudaOps()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = []

    def step(i, record=False):
        x = windows[i % N_WINDOWS]
        e = [ev() for _ in range(4)] if record else None
        if record:
            e[0].record()
        if world == 1:
            device.sosfilt(sos, x, 0, out=filt)
            y = filt
        else:
            rec = sharded.ShardedRecording(x, n*world, RATE, ops, rank, world, bounds)
            y = rec.sosfilt(sos, room=NFFT - HOP)
            frec = sharded.ShardedRecording(y, n*world, RATE, ops, rank, world, bounds,
                                            buffer=rec.last_buffer)
        if record:
            e[1].record()
        if world == 1:
            device.spectrogram(y, RATE, NFFT, HOP, nspec, out=spec)
        else:
            frec.spectrogram(NFFT, HOP)
        if record:
            e[2].record()
        if world == 1:
            device.envelope(esos, y, 0, True, out=env)
        else:
            frec.envelope(esos, True)
        if record:
            e[3].record()
            marks.append(e)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sync_all()
    # N > 1: one CUDA graph per input window (kernels, NCCL collectives and the small torch ops
    # of the exchange in one launch); falls back to eager launches if capture is refused
    graphs = None
    graph_launches = 0
    if world > 1 and args.graphs:
        try:
            graphs = []
            l0 = _lib.launch_count()
            for w in range(N_WINDOWS):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    step(w)
                graphs.append(g)
            graph_launches = (_lib.launch_count() - l0)//N_WINDOWS
            for g in graphs:
                g.replay()
            sync_all()
        except Exception as exc:                      # pragma: no cover
            sys.stderr.write('CUDA graph capture failed (%s): eager launches\n' % (exc,))
            graphs = None
    ok = torch.tensor([1 if graphs is not None else 0], device='cuda')
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            graphs = None
    # per-op times (and, without graphs, the timed region itself) from eager steps
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
    if graphs is None:
        t_start.record()
        for i in range(args.steps):
            step(args.warmup + i, record=True)
        t_stop.record()
        sync_all()
        launches = _lib.launch_count() - launches0
    else:
        for i in range(min(args.steps, 5)):
            step(i, record=True)
        sync_all()
        t_start.record()
        for i in range(args.steps):
            graphs[(args.warmup + i) % N_WINDOWS].replay()
        t_stop.record()
        sync_all()
        launches = graph_launches*args.steps
    torch.cuda.nvtx.range_pop()
    wall1 = time.perf_counter()
    ms_total = t_start.elapsed_time(t_stop)
    tt = torch.tensor([ms_total], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item())/args.steps
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    samples_step = n*C*world
    value = samples_step/(ms_step*1e-3)/1e6

    # per-op shares from the events recorded inside the timed region
    t_f = float(np.mean([e[0].elapsed_time(e[1]) for e in marks]))
    t_s = float(np.mean([e[1].elapsed_time(e[2]) for e in marks]))
    t_e = float(np.mean([e[2].elapsed_time(e[3]) for e in marks]))
    peak, peak_kind = measured_peak()
    edge = _lib.sosfiltfilt_edge(esos)

    def roof(op, ms, frames, launches_per_step=1, kernel=''):
        # per LAUNCH of the kernel: its algorithmic bytes, its average duration, its share of the step
        ach = algorithmic_bytes(op, frames, C)/(ms/launches_per_step*1e-3)/1e9
        return {'kernel': kernel, 'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                'frac': ach/peak, 'traffic': None, 'peak_kind': peak_kind,
                'ms_per_launch': ms/launches_per_step, 'launches_per_step': launches_per_step,
                'share_of_step': ms/launches_per_step/(t_f + t_s + t_e)}

    # which scan kernel ran: the run kernel (cascades that forget fast) or the look-back kernel
    scan_name = 'sos_run_kernel' if _lib.scan_run_count() > 0 else 'sos_scan_kernel'
    roofs = {
        'filter': roof('filter', t_f, n, 1, scan_name + '<S=2,FWD>'),
        'spectrogram': roof('spectrogram', t_s, n, 1, 'spectrogram_ring_kernel<10>'),
        'envelope': roof('envelope_sweep', t_e, n + 2*edge, 2,
                         scan_name + '<S=1,ENVF> and <S=1,REV>: the two sweeps of the envelope, each'),
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
    dominant = max(roofs, key=lambda k: roofs[k]['share_of_step'])

    # ---- end to end through the plugin path: host buffers, copies inside the timed region
    e2e = None
    cpu_baseline = None
    if world == 1 or True:
        host_x = [np.ascontiguousarray(w.cpu().numpy()) for w in windows[:2]]
        h_filt = np.empty((n, C))
        h_spec = np.empty((nspec, C, NFFT//2 + 1))
        h_env = np.empty((n, C))
        for a in host_x + [h_filt, h_spec, h_env]:
            _lib.host_register(a)
        tf = BufferedFilter()
        tf.configure_standalone(RATE, C, highpass_cutoff=HIGHPASS, lowpass_cutoff=LOWPASS)
        ts = BufferedSpectrogram(nfft=NFFT, overlap_frac=0.5)
        ts.configure_standalone(RATE, C, source=tf)
        te = BufferedEnvelope(envelope_cutoff=ENV_CUTOFF)
        te.configure_standalone(RATE, C, source=tf)

        def host_step(i):
            x = host_x[i % len(host_x)]
            tf.process(x, h_filt, 0)
            ts.process(h_filt, h_spec, 0)
            te.process(h_filt, h_env, 0)

        ksteps = max(1, min(args.steps, 5))
        for i in range(2):
            host_step(i)
        sync_all()
        moved0 = _lib.transfer_bytes()
        t0 = time.perf_counter()
        for i in range(ksteps):
            host_step(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        moved1 = _lib.transfer_bytes()
        te2e = torch.tensor([dt/ksteps], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(te2e, op=dist.ReduceOp.MAX)
        e2e = {'value': samples_step/float(te2e.item())/1e6, 'unit': 'Msamples/s',
               # counted by the library around every copy it issues
               'h2d_bytes_per_step': int((moved1[0] - moved0[0])//ksteps),
               'd2h_bytes_per_step': int((moved1[1] - moved0[1])//ksteps),
               'steps': ksteps, 'ms_per_step': float(te2e.item())*1e3,
               'api': 'BufferedFilter/BufferedSpectrogram/BufferedEnvelope.process on pinned numpy '
                      'buffers; the filtered buffer stays resident on the device for its two '
                      'consumers, copies and kernels overlap chunk by chunk'}
        for a in host_x + [h_filt, h_spec, h_env]:
            _lib.host_unregister(a)

    parity = None
    if rank == 0 and world == 1:
        # outputs of the timed path against the oracle (not timed)
        from oracle import oracle as orc
        x0 = windows[0]
        device.sosfilt(sos, x0, 0, out=filt)
        device.spectrogram(filt, RATE, NFFT, HOP, nspec, out=spec)
        device.envelope(esos, filt, 0, True, out=env)
        torch.cuda.synchronize()
        m = 400000
        hx = np.ascontiguousarray(x0[:m].cpu().numpy())
        hf = filt.cpu().numpy()
        rf = np.empty((m, C))
        orc.filter_process(sos, hx, rf, 0)
        rs = np.empty((m//HOP, C, NFFT//2 + 1))
        ns_ = orc.spectrogram_process(hf[:m], rs, RATE, NFFT, HOP)
        gs = spec[:ns_].cpu().numpy()
        re_ = np.empty((n, C))
        orc.envelope_process(esos, hf, re_, 0, 0)
        parity = {'filter_max_abs_err': float(np.max(np.abs(hf[:m] - rf))),
                  'spectrogram_max_rel_err': float(np.max(np.abs(gs - rs[:ns_])/np.maximum(rs[:ns_], 1e-20*rs.max()))),
                  'envelope_max_abs_err': float(np.max(np.abs(env.cpu().numpy() - re_))),
                  'checked': f'filter/spectrogram on the first {m} frames, envelope on the whole window'}
    if rank == 0 and world == 1:
        # the oracle on the host's cores, bounded sample of the same workload
        sample_s = 10.0
        m = int(RATE*sample_s)
        xs = np.ascontiguousarray(windows[0][:m].cpu().numpy())
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
                                      f'time-sharded x{world} (IIR state all-gather + STFT halo over NCCL)',
                       'launch': 'eager' if graphs is None else 'one CUDA graph per step; op_ms from eager steps'},
            'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks,
            'roofline': roofs[dominant], 'roofline_all': roofs, 'dominant': dominant,
            'op_ms': {'filter': t_f, 'spectrogram': t_s, 'envelope': t_e},
            'cpu_baseline': cpu_baseline, 'parity': parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if graphs is not None:
            # tearing down CUDA graphs that hold captured NCCL work together with their
            # communicator can deadlock: everything is done and flushed, leave directly
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-graphs', dest='graphs', action='store_false',
                    help='N > 1: launch every step eagerly instead of replaying CUDA graphs')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == '__main__':
    main()
