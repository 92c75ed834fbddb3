"""CPU oracle for audian's derived-trace DSP path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this module; nothing under
`audian_b200/` does.  It restates, on the same scipy/numpy calls the
reference makes, what these reference functions compute (paths relative
to /root/reference):

* `BufferedSpectrogram.process`   src/audian/bufferedspectrogram.py:45-66
* `BufferedFilter.process/update` src/audian/bufferedfilter.py:31-53
* `BufferedEnvelope.process/update` src/audian/bufferedenvelope.py:34-55
* `down_sample_worker`, `CompressedData.start` src/audian/compresseddata.py:25-53,79-122
* `TraceItem.update_plot` decimation src/audian/traceitem.py:33-67
* `BufferedData.update_step/align_buffer/load_buffer` src/audian/buffereddata.py:39-109

Parity pin: the reference ships no tests and no golden vectors
(SURVEY.md section 4).  The pin used instead is the reference's own code
run in this container: `oracle/ref_harness.py` imports the reference's
modules from /root/reference with stand-ins for the absent third-party
packages (audioio, thunderlab) and `oracle/make_golden.py` stores its
outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks this
restatement against them bit for bit.  Third-party arithmetic not under
/root/reference: `thunderlab.powerspectrum.spectrogram` / `decibel`
(thunderlab >= 1.6, unpinned, pyproject.toml:20) -- restated here from
its published behaviour: scipy.signal.spectrogram(window='hann',
detrend='constant', scaling='density', mode='psd', axis=0) followed by
a (f, ch, t) -> (f, t, ch) transpose; decibel = 10*log10(power) with
power <= 1e-20 -> -inf.  Window and detrend stay parameters.
"""

from math import floor, ceil

import os

import numpy as np
from scipy.signal import butter, sosfilt, sosfiltfilt, sosfilt_zi
from scipy.signal import spectrogram as _scipy_spectrogram


# ---------------------------------------------------------------------------
# thunderlab.powerspectrum restatement (third-party, not vendored)

def tl_spectrogram(data, rate, n_fft, n_overlap, window='hann',
                   detrend='constant'):
    """thunderlab.powerspectrum.spectrogram as called at
    bufferedspectrogram.py:51-56: returns (freqs, times, Sxx) with Sxx of
    shape (f, t) for 1-D and (f, t, ch) for 2-D (frames, channels) input."""
    freqs, time, spec = _scipy_spectrogram(data, fs=rate, window=window,
                                           nperseg=n_fft, noverlap=n_overlap,
                                           detrend=detrend, scaling='density',
                                           mode='psd', axis=0)
    if data.ndim > 1:
        # scipy returns (f, ch, t)
        spec = np.transpose(spec, (0, 2, 1))
    return freqs, time, spec


def decibel(power, ref_power=1.0, min_power=1e-20):
    """thunderlab.powerspectrum.decibel (bufferedspectrogram.py:116-117,
    specitem.py:36, spectrogramplot.py:159)."""
    if np.isscalar(power):
        if power <= min_power:
            return -np.inf
        return 10.0*np.log10(power/ref_power)
    power = np.asarray(power, dtype=np.float64)
    db = power.copy()
    db[power <= min_power] = -np.inf
    mask = power > min_power
    db[mask] = 10.0*np.log10(power[mask]/ref_power)
    return db


# ---------------------------------------------------------------------------
# process() restatements

def spectrogram_process(source, dest, rate, nfft, hop, window='hann',
                        detrend='constant'):
    """bufferedspectrogram.py:45-62.  Fills `dest` (nframes, C, nfft//2+1)
    in place; returns the number of computed frames."""
    nsource = (len(dest) - 1)*hop + nfft
    if nsource > len(source):
        nsource = len(source)
    if nsource >= nfft:
        with np.errstate(under='ignore'):
            freq, time, Sxx = tl_spectrogram(source[:nsource], rate, nfft,
                                             nfft - hop, window, detrend)
        n = Sxx.shape[1]
        dest[:n] = Sxx.transpose((1, 2, 0))
        dest[n:] = 0
        return n
    dest[:] = 0
    return 0


def spectrogram_hop_open(nfft, overlap_frac):
    """bufferedspectrogram.py:32 (open): truncating int()."""
    return int(nfft*(1 - overlap_frac))


def spectrogram_set_hop(nfft, overlap_frac):
    """bufferedspectrogram.py:69-75 (set_hop): rounding, clamped to [1, nfft]."""
    hop = int(np.round((1 - overlap_frac)*nfft))
    return min(max(hop, 1), nfft)


def estimate_noiselevels(buffer, channel):
    """bufferedspectrogram.py:109-126 (without the `init` latch)."""
    nf = max(1, buffer.shape[2]//16)
    with np.errstate(all='ignore'):
        zmin = np.percentile(decibel(buffer[:, channel, -nf:]), 95)
    zmax = np.max(decibel(buffer[:, channel, :]))
    if not np.isfinite(zmin) or not np.isfinite(zmax):
        return None, None
    zmax = zmin + 0.95*(zmax - zmin)
    if zmax - zmin < 20:
        zmax = zmin + 20
    if zmax - zmin > 80:
        zmin = zmax - 80
    return zmin, zmax


def filter_design(rate, highpass_cutoff, lowpass_cutoff, filter_order=2):
    """bufferedfilter.py:39-52: returns sos (S, 6) or None."""
    if highpass_cutoff < 0.001*rate/2 and lowpass_cutoff >= rate/2 - 1e-8:
        return None
    if highpass_cutoff < 0.001*rate/2:
        return butter(filter_order, lowpass_cutoff, 'lowpass', fs=rate,
                      output='sos')
    if lowpass_cutoff >= rate/2 - 1e-8:
        return butter(filter_order, highpass_cutoff, 'highpass', fs=rate,
                      output='sos')
    return butter(filter_order, (highpass_cutoff, lowpass_cutoff),
                  'bandpass', fs=rate, output='sos')


def filter_process(sos, source, dest, nbefore):
    """bufferedfilter.py:31-36 including the per-channel Python loop."""
    if sos is None:
        dest[:, :] = source[nbefore:, :]
    else:
        for c in range(source.shape[1]):
            dest[:, c] = sosfilt(sos, source[:, c])[nbefore:]


def envelope_design(rate, envelope_cutoff, highpass_cutoff=0, filter_order=2):
    """bufferedenvelope.py:44-54: returns sos or None (ValueError swallowed)."""
    try:
        if highpass_cutoff > 0:
            return butter(filter_order, (highpass_cutoff, envelope_cutoff),
                          'bandpass', fs=rate, output='sos')
        return butter(filter_order, envelope_cutoff, 'lowpass', fs=rate,
                      output='sos')
    except ValueError:
        return None


def envelope_process(sos, source, dest, nbefore, highpass_cutoff=0):
    """bufferedenvelope.py:34-41."""
    if sos is None:
        dest[:] = np.zeros_like(dest)
    else:
        dest[:] = sosfiltfilt(sos, (np.pi/2)*np.abs(source), axis=0)[nbefore:]
        if highpass_cutoff == 0:
            dest[dest < 0] = 0


# ---------------------------------------------------------------------------
# min/max decimation

def minmax_rows(buffer, step):
    """The reduceat idiom of compresseddata.py:49-52 / :97-100 on one block:
    returns (2*nseg, C) with row 2j = min, row 2j+1 = max of segment j."""
    segments = np.arange(0, len(buffer), step)
    out = np.zeros((2*len(segments), buffer.shape[1]))
    if len(segments) > 0:
        np.minimum.reduceat(buffer, segments, out=out[0::2])
        np.maximum.reduceat(buffer, segments, out=out[1::2])
    return out


def fulltrace_params(frames, rate, max_pixel):
    """compresseddata.py:83-89: (step, nblock, times)."""
    step = max(1, frames//max_pixel)
    nblock = max(step, int(30.0*rate//step)*step)
    times = np.arange(0, frames + step - 1, step/2)/rate
    return step, nblock, times


def fulltrace_short(buffer, max_pixel, rate):
    """compresseddata.py:90-101 (whole file in the buffer)."""
    frames = len(buffer)
    step, nblock, times = fulltrace_params(frames, rate, max_pixel)
    segments = np.arange(0, frames, step)
    datas = np.zeros((1 + 2*len(segments), buffer.shape[1]))
    np.minimum.reduceat(buffer, segments, out=datas[0:0 + 2*len(segments):2])
    np.maximum.reduceat(buffer, segments, out=datas[1:1 + 2*len(segments):2])
    return times, datas


def fulltrace_long(data, max_pixel, rate, num_proc=1):
    """compresseddata.py:104-122 + down_sample_worker :25-53 with the worker
    processes run one after the other on an in-memory (frames, C) array."""
    frames, channels = data.shape
    step, nblock0, times = fulltrace_params(frames, rate, max_pixel)
    datas = np.zeros((len(times), channels))
    for proc_idx in range(max(1, num_proc)):
        nblock = nblock0
        for index in range(proc_idx*nblock0, frames, max(1, num_proc)*nblock0):
            if frames - index < nblock:
                nblock = frames - index
            buffer = data[index:index + nblock]
            segments = np.arange(0, len(buffer), step)
            i = 2*index//step
            np.minimum.reduceat(buffer, segments,
                                out=datas[i + 0:i + 0 + 2*len(segments):2])
            np.maximum.reduceat(buffer, segments,
                                out=datas[i + 1:i + 1 + 2*len(segments):2])
    return times, datas


def traceitem_decimate(trace_len, buf_offset, buf, rate, channel, t0, t1,
                       max_pixel):
    """traceitem.py:39-67: visible-window min/max decimation of one channel
    of a buffered trace.  `buf` holds rows buf_offset.. of the trace.
    Returns (step, start, plot_data) or (1, start, raw slice)."""
    start = max(0, int(t0*rate))
    tstop = int(t1*rate + 1)
    stop = min(trace_len, tstop)
    step = max(1, (tstop - start)//max_pixel)
    if step > 1:
        start = (start//step)*step
        tstop = (stop//step + 1)*step
        stop = min(trace_len, tstop)
        while start < buf_offset:
            start += step
        while stop > buf_offset + len(buf):
            stop -= step
        segments = np.arange(0, stop - start, step)
        plot_data = np.zeros(2*len(segments))
        col = buf[start - buf_offset:stop - buf_offset, channel]
        np.minimum.reduceat(col, segments, out=plot_data[0::2])
        np.maximum.reduceat(col, segments, out=plot_data[1::2])
        return step, start, plot_data
    return 1, start, buf[start - buf_offset:stop - buf_offset, channel]


# ---------------------------------------------------------------------------
# buffer index algebra (buffereddata.py:39-109), as pure functions

def update_step(src_rate, src_frames, src_bufferframes, src_offset,
                own_bufferframes, own_rate, step):
    """buffereddata.py:39-56: returns dict(rate, frames, bufferframes, offset)."""
    tbuffer = own_bufferframes/own_rate
    if step < 1:
        step = 1
    rate = src_rate/step
    frames = (src_frames + step - 1)//step
    if src_bufferframes == src_frames:
        bufferframes = frames
    else:
        bufferframes = int(tbuffer*rate)
    offset = (src_offset + step - 1)//step
    return dict(rate=rate, frames=frames, bufferframes=bufferframes,
                offset=offset)


def align_buffer(src_offset, src_buflen, src_frames, src_rate, rate,
                 source_tbefore, source_tafter):
    """buffereddata.py:75-86: (offset, nframes) handed to move_buffer."""
    soffset = src_offset
    snframes = src_buflen
    if soffset > 0:
        n = floor(source_tbefore*src_rate)
        soffset += n
        snframes -= n
    if src_offset + src_buflen < src_frames:
        n = floor(source_tafter*src_rate)
        snframes -= n
    offset = ceil(soffset*rate/src_rate)
    nframes = floor((soffset + snframes)*rate/src_rate) - offset
    return offset, nframes


def load_buffer_slice(offset, nframes, rate, src_rate, src_offset,
                      src_buflen, source_tbefore, source_tafter):
    """buffereddata.py:94-107: (soffset, snframes, nbefore) of the source
    buffer slice given to process() -- including the divide-instead-of-
    multiply quirk on lines 96 and 99."""
    soffset = floor(offset*src_rate/rate)
    snframes = ceil((offset + nframes)*src_rate/rate) - soffset
    nbefore = floor(source_tbefore/src_rate)
    soffset -= nbefore
    snframes += nbefore
    nafter = ceil(source_tafter/src_rate)
    snframes += nafter
    soffset -= src_offset
    if soffset < 0:
        nbefore += soffset
        snframes += soffset
        soffset = 0
    if soffset + snframes > src_buflen:
        snframes = src_buflen - soffset
    return soffset, snframes, nbefore


# ---------------------------------------------------------------------------
# arithmetic restated by hand; these validate that
# the scipy calls above are what the CUDA kernels must reproduce.

def sosfilt_df2t(sos, x, zi=None):
    """scipy _sosfilt restated (direct form II transposed, a0 == 1).
    x: 1-D.  Returns (y, zf) with zf of shape (S, 2)."""
    sos = np.asarray(sos, dtype=np.float64)
    S = sos.shape[0]
    z = np.zeros((S, 2)) if zi is None else np.array(zi, dtype=np.float64)
    y = np.empty(len(x))
    for n in range(len(x)):
        xc = x[n]
        for s in range(S):
            b0, b1, b2, _, a1, a2 = sos[s]
            xn = b0*xc + z[s, 0]
            z[s, 0] = b1*xc - a1*xn + z[s, 1]
            z[s, 1] = b2*xc - a2*xn
            xc = xn
        y[n] = xc
    return y, z


def sosfiltfilt_edge(sos):
    """scipy sosfiltfilt default pad length (_signaltools.py sosfiltfilt)."""
    sos = np.asarray(sos)
    ntaps = 2*sos.shape[0] + 1
    ntaps -= min((sos[:, 2] == 0).sum(), (sos[:, 5] == 0).sum())
    return 3*ntaps


def sosfiltfilt_restated(sos, x):
    """scipy sosfiltfilt (padtype='odd', default padlen) along axis 0 of a
    2-D array, restated with explicit padding and sosfilt calls."""
    edge = sosfiltfilt_edge(sos)
    if x.shape[0] <= edge:
        raise ValueError('input too short for sosfiltfilt padding')
    left = 2*x[0:1] - x[edge:0:-1]
    right = 2*x[-1:] - x[-2:-edge - 2:-1]
    ext = np.concatenate((left, x, right), axis=0)
    zi = sosfilt_zi(sos)                           # (S, 2)
    zi = zi[:, :, None]                            # (S, 2, 1) for axis 0
    y1, _ = sosfilt(sos, ext, axis=0, zi=zi*ext[0:1][None])
    y1r = y1[::-1]
    y2, _ = sosfilt(sos, y1r, axis=0, zi=zi*y1r[0:1][None])
    y = y2[::-1]
    return y[edge:-edge]


def spectrogram_frames_restated(x, rate, nfft, hop):
    """One-sided PSD frames restated by hand: periodic Hann,
    per-frame mean removal, rfft, |X|^2/(rate*sum(w^2)), interior bins x2.
    x: (ns, C).  Returns (n, C, nfft//2+1)."""
    ns, C = x.shape
    n = (ns - (nfft - hop))//hop
    j = np.arange(nfft)
    w = 0.5 - 0.5*np.cos(2*np.pi*j/nfft)
    scale = 1.0/(rate*np.sum(w*w))
    out = np.zeros((max(n, 0), C, nfft//2 + 1))
    for k in range(n):
        seg = x[k*hop:k*hop + nfft]
        seg = seg - seg.mean(axis=0, keepdims=True)
        X = np.fft.rfft(seg*w[:, None], axis=0)
        P = (X.real**2 + X.imag**2)*scale
        if nfft % 2 == 0:
            P[1:-1] *= 2
        else:
            P[1:] *= 2
        out[k] = P.T
    return out


# ---------------------------------------------------------------- f4: play-back and ingest

def play_region(data, show_channels, rate, use_heterodyne=False, heterodyne_freq=0.0):
    """DataBrowser.play_region up to the fade (src/audian/databrowser.py:1711-1728), on the
    rows `data` (frames, channels) of the region: returns (playdata, rate)."""
    from scipy.signal import butter, sosfiltfilt
    n2 = (len(show_channels) + 1)//2
    playdata = np.zeros((len(data), min(2, len(show_channels))))
    playdata[:, 0] = np.mean(data[:, show_channels[:n2]], 1)
    if len(show_channels) > 1:
        playdata[:, 1] = np.mean(data[:, show_channels[n2:]], 1)
    if use_heterodyne:
        heterodyne = np.sin(2*np.pi*heterodyne_freq*np.arange(len(playdata))/rate)
        playdata = (playdata.T * heterodyne).T
        fcutoff = 20000.0
        sos = butter(2, 20000, 'low', output='sos', fs=rate)
        nstep = int(np.round(rate/(2*fcutoff)))
        if nstep < 1:
            nstep = 1
        playdata = sosfiltfilt(sos, playdata, 0)[::nstep]
        rate /= nstep
    return playdata, rate


def unwrap(data, thresh=-1.0, clips=False):
    """audioio.unwrap(data, thresh, clips) as the loader applies it when Data.open passes
    `unwrap` (src/audian/data.py:180).  audioio is not vendored with the reference and not
    installed here: this restates its documented behaviour [recalled] -- parity unpinned.
    In place; a step between consecutive samples of a channel below -thresh adds 2 to
    everything that follows, a step above +thresh takes 2 away; clips limits to [-1, 1]."""
    if thresh <= 0:
        return data
    d = data.reshape(len(data), -1)
    dd = np.diff(d, axis=0)
    k = np.cumsum((dd < -thresh).astype(np.int64) - (dd > thresh).astype(np.int64), axis=0)
    d[1:] += 2.0*k
    if clips:
        np.clip(d, -1.0, 1.0, out=d)
    return data


# ---------------------------------------------------------------- full-trace pass with the reference's workers

def _fulltrace_worker(proc_idx, num_proc, nblock, step, data_name, data_shape, out_name, out_shape, lock,
                      stamps):
    """down_sample_worker (src/audian/compresseddata.py:25-53) on in-memory data: the block copy
    into the worker's buffer stands in for `data.load_buffer(index, nblock, buffer)`."""
    import time
    from multiprocessing import shared_memory
    dshm = shared_memory.SharedMemory(name=data_name)
    oshm = shared_memory.SharedMemory(name=out_name)
    try:
        data = np.ndarray(data_shape, dtype=np.float64, buffer=dshm.buf)
        datas = np.ndarray(out_shape, dtype=np.float64, buffer=oshm.buf)
        frames = data_shape[0]
        buffer = np.zeros((nblock, data_shape[1]))
        segments = np.arange(0, len(buffer), step)
        stamps[2*proc_idx] = time.perf_counter()      # CLOCK_MONOTONIC: comparable across processes
        for index in range(proc_idx*nblock, frames, num_proc*nblock):
            if frames - index < nblock:
                nblock = frames - index
                buffer = buffer[:nblock, :]
                segments = np.arange(0, len(buffer), step)
            buffer[:] = data[index:index + nblock]
            i = 2*index//step
            with lock:
                np.minimum.reduceat(buffer, segments, out=datas[i + 0:i + 0 + 2*len(segments):2])
                np.maximum.reduceat(buffer, segments, out=datas[i + 1:i + 1 + 2*len(segments):2])
        stamps[2*proc_idx + 1] = time.perf_counter()
    finally:
        dshm.close()
        oshm.close()


def fulltrace_parallel(data, max_pixel, rate, num_proc=None):
    """CompressedData.start's long-file path (src/audian/compresseddata.py:104-122) with real
    worker processes: os.cpu_count() - 1 of them, block-cyclic over 30-s blocks, writing into one
    shared array under its lock.  `data`: in-memory (frames, C) recording.  Returns
    (times, datas, seconds, workers); seconds = first worker entering its block loop to
    the last one leaving it (process start-up and imports are not counted)."""
    import multiprocessing as mp
    import time
    from multiprocessing import shared_memory
    frames, channels = data.shape
    step, nblock, times = fulltrace_params(frames, rate, max_pixel)
    if num_proc is None:
        num_proc = max(1, (os.cpu_count() or 2) - 1)
    ctx = mp.get_context('spawn')               # never fork a process that may hold a CUDA context
    dshm = shared_memory.SharedMemory(create=True, size=max(8, data.nbytes))
    oshm = shared_memory.SharedMemory(create=True, size=max(8, len(times)*channels*8))
    try:
        shared = np.ndarray(data.shape, dtype=np.float64, buffer=dshm.buf)
        shared[:] = data
        datas = np.ndarray((len(times), channels), dtype=np.float64, buffer=oshm.buf)
        datas[:] = 0.0
        lock = ctx.Lock()
        stamps = ctx.Array('d', 2*num_proc)
        procs = [ctx.Process(target=_fulltrace_worker,
                             args=(i, num_proc, nblock, step, dshm.name, data.shape, oshm.name,
                                   datas.shape, lock, stamps)) for i in range(num_proc)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        st = np.array(stamps[:]).reshape(num_proc, 2)
        st = st[st[:, 1] > 0]
        dt = float(st[:, 1].max() - st[:, 0].min()) if len(st) else float('nan')
        result = datas.copy()
    finally:
        dshm.close()
        dshm.unlink()
        oshm.close()
        oshm.unlink()
    return times, result, dt, num_proc
