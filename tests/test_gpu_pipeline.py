"""Host-pointer entry points: chunked upload/kernel/download pipeline and results
kept resident on the device (ADN_OPT_RESIDENT) -- same answers as the oracle whatever
the chunking, stale device copies are never used."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from audian_b200 import _lib
from audian_b200.synth import synth
from oracle import oracle as orc


@pytest.fixture
def small_chunks():
    old = _lib.get_option(_lib.ADN_OPT_CHUNK_BYTES)
    _lib.set_option(_lib.ADN_OPT_CHUNK_BYTES, 1 << 16)       # 64 KiB: dozens of chunks
    yield
    _lib.set_option(_lib.ADN_OPT_CHUNK_BYTES, old)


@pytest.fixture
def resident():
    oldr = _lib.get_option(_lib.ADN_OPT_RESIDENT)
    oldm = _lib.get_option(_lib.ADN_OPT_RESIDENT_MIN_BYTES)
    _lib.set_option(_lib.ADN_OPT_RESIDENT, 1)
    _lib.set_option(_lib.ADN_OPT_RESIDENT_MIN_BYTES, 1 << 12)
    yield
    _lib.set_option(_lib.ADN_OPT_RESIDENT_MIN_BYTES, oldm)
    _lib.set_option(_lib.ADN_OPT_RESIDENT, oldr)


@pytest.mark.parametrize('C,nbefore', [(1, 0), (2, 777), (8, 5000), (3, 13)])
def test_sosfilt_chunked_equals_one_shot(small_chunks, C, nbefore):
    fs, n = 48000., 60000
    x = synth(0, n, C, fs, seed=C)
    sos = orc.filter_design(fs, 800., 12000., 4)
    ref = np.empty((n - nbefore, C))
    orc.filter_process(sos, x, ref, nbefore)
    got = np.full_like(ref, np.nan)
    _lib.sosfilt(sos, x, got, nbefore)
    assert np.max(np.abs(got - ref)) <= 1e-9
    # shorter destination and carried state
    zi = np.zeros((C, sos.shape[0], 2))
    h = n//3
    a = np.empty((h, C))
    b = np.empty((n - h, C))
    _lib.sosfilt(sos, x[:h], a, 0, zi=zi)
    _lib.sosfilt(sos, x[h:], b, 0, zi=zi)
    full = np.empty((n, C))
    orc.filter_process(sos, x, full, 0)
    assert np.max(np.abs(np.concatenate([a, b]) - full)) <= 1e-9


@pytest.mark.parametrize('nfft,hop', [(256, 128), (1024, 512), (1024, 128), (2048, 2048), (64, 16)])
def test_spectrogram_chunked(small_chunks, nfft, hop):
    fs, C, n = 48000., 4, 50000
    x = synth(5, n, C, fs, seed=nfft)
    n_dst = n//hop + 3
    ref = np.empty((n_dst, C, nfft//2 + 1))
    nref = orc.spectrogram_process(x, ref, fs, nfft, hop)
    got = np.full_like(ref, np.nan)
    assert _lib.spectrogram(x, fs, nfft, hop, got) == nref
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-20*ref.max())


def test_minmax_chunked_bit_exact(small_chunks):
    x = synth(0, 100000, 4, 48000., seed=3)
    x[5000:5003] = np.nan
    x[70000, 1] = -0.0
    for step in (7, 1000, 33333, 100000):
        got = _lib.minmax(x, step)
        assert np.array_equal(got.view(np.uint64), orc.minmax_rows(x, step).view(np.uint64)), step


def test_resident_chain_matches_oracle_and_skips_uploads(resident):
    fs, C, n = 48000., 8, 120000
    x = synth(0, n, C, fs, seed=11)
    sos = orc.filter_design(fs, 1000., 15000., 2)
    esos = orc.envelope_design(fs, 500.)
    filt = np.empty((n, C))
    m = _lib.Mirror()                                    # owned by whoever owns `filt`
    hits0 = _lib.resident_hits()
    _lib.sosfilt(sos, x, filt, 0, dst_mirror=m)
    rf = np.empty((n, C))
    orc.filter_process(sos, x, rf, 0)
    assert np.max(np.abs(filt - rf)) <= 1e-9
    nfft, hop = 1024, 512
    spec = np.empty((n//hop, C, nfft//2 + 1))
    ns = _lib.spectrogram(filt, fs, nfft, hop, spec, src_mirror=m)
    env = np.empty((n, C))
    _lib.envelope(esos, filt, env, 0, True, src_mirror=m)
    assert _lib.resident_hits() == hits0 + 2            # both consumers read the device copy
    rs = np.empty_like(spec)
    assert orc.spectrogram_process(filt, rs, fs, nfft, hop) == ns
    assert np.allclose(spec, rs, rtol=1e-5, atol=1e-20*rs.max())
    re = np.empty((n, C))
    orc.envelope_process(esos, filt, re, 0, 0)
    assert np.max(np.abs(env - re)) <= 1e-9
    # a slice of the kept range hits too
    part = np.empty((n - 40000, C))
    _lib.envelope(esos, filt[40000:], part, 0, True, src_mirror=m)
    assert _lib.resident_hits() == hits0 + 3
    orc.envelope_process(esos, filt[40000:], re[:n - 40000], 0, 0)
    assert np.max(np.abs(part - re[:n - 40000])) <= 1e-9
    # a range that is not inside the mirror's does not
    other = filt.copy()
    _lib.envelope(esos, other, env, 0, True, src_mirror=m)
    assert _lib.resident_hits() == hits0 + 3
    m.release()


def test_stale_device_copy_is_never_used(resident):
    """Explicit hand-over: calls that name no mirror never see device copies, so an array
    edited in place -- three samples are enough -- or freed and reallocated at the same
    address is always read from the host."""
    fs, C, n = 48000., 2, 50000
    x = synth(0, n, C, fs, seed=12)
    sos = orc.filter_design(fs, 1000., 15000., 2)
    filt = np.empty((n, C))
    m = _lib.Mirror()
    _lib.sosfilt(sos, x, filt, 0, dst_mirror=m)
    # three samples change in place; a plain call sees the new values
    filt[7, 0] = 5.0
    filt[20001, 1] = -6.0
    filt[n - 1, 0] = 7.0
    hits = _lib.resident_hits()
    got = _lib.minmax(filt, 100)
    assert _lib.resident_hits() == hits
    assert np.array_equal(got.view(np.uint64), orc.minmax_rows(filt, 100).view(np.uint64))
    assert got[1, 0] == 5.0 and got[2*200, 1] == -6.0 and got[-1, 0] == 7.0
    # the owner says so: the mirror is not used until its buffer is filled again
    m.invalidate()
    got = _lib.minmax(filt, 100, src_mirror=m)
    assert _lib.resident_hits() == hits
    assert np.array_equal(got.view(np.uint64), orc.minmax_rows(filt, 100).view(np.uint64))
    _lib.sosfilt(sos, x, filt, 0, dst_mirror=m)
    got = _lib.minmax(filt, 100, src_mirror=m)
    assert _lib.resident_hits() == hits + 1
    assert np.array_equal(got.view(np.uint64), orc.minmax_rows(filt, 100).view(np.uint64))
    # address-based invalidation still works (adn_invalidate)
    filt *= 0.5
    _lib.invalidate(filt)
    got = _lib.minmax(filt, 100, src_mirror=m)
    assert _lib.resident_hits() == hits + 1
    assert np.array_equal(got.view(np.uint64), orc.minmax_rows(filt, 100).view(np.uint64))
    # results written over a kept range replace it
    _lib.sosfilt(sos, x, filt, 0, dst_mirror=m)
    _lib.sosfilt(sos, 2.0*x, filt, 0, dst_mirror=m)
    got = _lib.minmax(filt, 100, src_mirror=m)
    assert np.array_equal(got.view(np.uint64), orc.minmax_rows(filt, 100).view(np.uint64))
    # temporaries: a freed array whose address is reused is never served from the device
    m.release()
    for k in range(3):
        tmp = np.empty((n, C))
        _lib.sosfilt(sos, (k + 1.0)*x, tmp, 0)
        ref = np.empty((n, C))
        orc.filter_process(sos, (k + 1.0)*x, ref, 0)
        got = _lib.minmax(tmp, 50)
        assert np.array_equal(got.view(np.uint64), orc.minmax_rows(tmp, 50).view(np.uint64))
        assert np.max(np.abs(tmp - ref)) <= 1e-9
        del tmp


def test_failed_call_leaves_the_mirror_invalid(resident):
    fs, C = 1000., 2
    sos = orc.envelope_design(fs, 100.)
    edge = orc.sosfiltfilt_edge(sos)
    m = _lib.Mirror()
    x = synth(0, edge, C, fs)
    out = np.empty_like(x)
    with pytest.raises(ValueError):
        _lib.envelope(sos, x, out, dst_mirror=m)
    hits = _lib.resident_hits()
    _lib.minmax(out, 3, src_mirror=m)
    assert _lib.resident_hits() == hits
    m.release()


def test_host_calls_from_two_threads(resident):
    """ctypes releases the GIL: the library serialises the host-pointer entry points."""
    import threading
    fs, C, n = 48000., 4, 200000
    sos = orc.filter_design(fs, 1000., 15000., 2)
    xs = [synth(k, n, C, fs, seed=90 + k) for k in range(4)]
    outs = [np.empty((n, C)) for _ in xs]
    errs = []

    def work(k):
        try:
            for _ in range(3):
                _lib.sosfilt(sos, xs[k], outs[k], 0)
                _lib.minmax(outs[k], 64)
        except Exception as exc:                      # pragma: no cover
            errs.append(exc)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    for k in range(4):
        ref = np.empty((n, C))
        orc.filter_process(sos, xs[k], ref, 0)
        assert np.max(np.abs(outs[k] - ref)) <= 1e-9


def test_traces_invalidate_on_buffer_moves(resident):
    """The scroll / parameter-change scenario of the golden fixture with every
    result kept resident: audioio recycles and replaces the buffers between the
    calls, the traces must invalidate the device copies in time."""
    from test_host_traces import replay, check_calls
    hits = _lib.resident_hits()
    g, log, states, filt, spect, env = replay(use_gpu=True)
    check_calls(g, log, states)
    assert _lib.resident_hits() > hits                   # the chain did use device copies
    assert np.max(np.abs(filt.buffer - g['filt_buffer'])) <= 1e-6
    assert np.max(np.abs(env.buffer - g['env_buffer'])) <= 1e-6
    ref = g['spec_buffer']
    assert np.allclose(spect.buffer, ref, rtol=1e-5, atol=1e-20*ref.max())


def test_trace_buffers_are_page_locked():
    """adn_host_alloc / _lib.pinned_empty: numpy arrays in page-locked memory, blocks reused by size;
    the traces' allocate_buffer() puts `buffer` there (the scroll replay keeps its answers)."""
    import gc
    a = _lib.pinned_empty((1000, 8))
    a[:] = 1.5
    assert a.shape == (1000, 8) and a.dtype == np.float64 and a.flags.c_contiguous
    assert _lib.is_pinned_array(a) and _lib.is_pinned_array(a[10:20])
    assert not _lib.is_pinned_array(np.empty((10, 8)))
    ptr = a.ctypes.data
    del a
    gc.collect()
    b = _lib.pinned_empty((1000, 8))
    assert b.ctypes.data == ptr                          # the released block came back from the pool
    x = synth(0, 1000, 8, 48000.)
    b[:] = x
    sos = orc.filter_design(48000., 1000., 15000., 2)
    got = _lib.pinned_empty((1000, 8))
    _lib.sosfilt(sos, b, got, 0)
    ref = np.empty((1000, 8))
    orc.filter_process(sos, x, ref, 0)
    assert np.max(np.abs(got - ref)) <= 1e-9
    from test_host_traces import replay
    g, log, states, filt, spect, env = replay(use_gpu=True)
    for tr in (filt, spect, env):
        assert _lib.is_pinned_array(tr.buffer), type(tr).__name__
    assert _lib.pinned_empty((0, 8)).shape == (0, 8)


@pytest.mark.parametrize('C,nbefore,mm', [(8, 0, 0), (2, 321, 1000), (5, 7, 0)])
def test_chain_equals_separate_calls(C, nbefore, mm):
    """adn_chain_f64: data -> filtered -> {spectrogram, envelope} (+ min/max of the raw rows) in one
    call (bufferedfilter.py:53, buffereddata.py:149-153) == the four entry points one after the
    other, bit for bit."""
    fs, n = 48000., 300000
    x = synth(1, n, C, fs, seed=70 + C)
    sos = orc.filter_design(fs, 1000., 15000., 2)
    esos = orc.envelope_design(fs, 500.)
    nfft, hop = 512, 128
    nf = n - nbefore
    filt = np.empty((nf, C))
    _lib.sosfilt(sos, x, filt, nbefore)
    s0, srows = 1024, nf - 1024 - 777
    nspec = srows//hop + 3                              # more frames than the slice can fill: zeros
    spec = np.full((nspec, C, nfft//2 + 1), np.nan)
    ncomp = _lib.spectrogram(filt[s0:s0 + srows], fs, nfft, hop, spec)
    e0, erows, enb = 5000, nf - 5000, 11
    env = np.empty((erows - enb, C))
    _lib.envelope(esos, filt[e0:e0 + erows], env, enb, True)
    f2 = np.full_like(filt, np.nan)
    s2 = np.full_like(spec, np.nan)
    e2 = np.full_like(env, np.nan)
    m2 = np.full((2*((n + mm - 1)//mm), C), np.nan) if mm else None
    got = _lib.chain(sos, x, f2, fs, nbefore, spec=s2, nfft=nfft, hop=hop, spec_first=s0, spec_rows=srows,
                     esos=esos, env=e2, env_first=e0, env_rows=erows, env_nbefore=enb,
                     clamp_negative=True, mm_step=mm, minmax_out=m2)
    assert got == ncomp
    assert np.array_equal(f2, filt)
    assert np.array_equal(s2, spec)
    assert np.max(np.abs(e2 - env)) <= 1e-13            # one launch over the slice vs the same launch
    if mm:
        assert np.array_equal(m2.view(np.uint64), orc.minmax_rows(x, mm).view(np.uint64))
    # stages can be left out; no filter = copy
    f3 = np.empty_like(filt)
    assert _lib.chain(None, x, f3, fs, nbefore) == 0
    assert np.array_equal(f3, x[nbefore:])
    with pytest.raises(ValueError):
        _lib.chain(sos, x, f2, fs, nbefore, esos=esos, env=np.empty((5, C)), env_first=0, env_rows=5)


def test_fused_recompute_equals_the_walk(monkeypatch):
    """BufferedFilter.recompute_all with a spectrogram and an envelope as dests: one chain call
    leaves every buffer exactly as the reference's walk trace by trace does."""
    import audian_b200 as ab
    from oracle.ref_harness import ArrayLoader
    fs, C = 48000., 2
    x = synth(0, 2500000, C, fs, seed=91)

    def graph():
        data = ArrayLoader(x, fs, 600000, 1500000)       # 31 s buffered: 10-s margin + 21 s of filtered
        filt, spect, env = ab.BufferedFilter(), ab.BufferedSpectrogram(nfft=1024, overlap_frac=0.75), \
            ab.BufferedEnvelope(envelope_cutoff=400.)
        filt.open(data)
        spect.open(filt)
        env.open(filt)
        for t in (filt, spect, env):
            t.need_update = True
        for t in (filt, spect, env):
            t.align_buffer()                             # Data.update_times(): buffers follow the loader's
        filt.highpass_cutoff, filt.lowpass_cutoff = 800., 12000.
        return data, filt, spect, env

    calls = []
    real = _lib.chain
    monkeypatch.setattr(_lib, 'chain', lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    d1, f1, s1, e1 = graph()
    f1.update()                                          # fused: one chain call
    assert len(calls) == 1
    d2, f2, s2, e2 = graph()
    monkeypatch.setattr(ab.BufferedFilter, '_chain_stages', lambda self: [])
    f2.update()                                          # the plain walk
    assert len(calls) == 1
    for a, b in ((f1, f2), (s1, s2), (e1, e2)):
        assert a.offset == b.offset and a.buffer.shape == b.buffer.shape and a.buffer.size > 0
        assert a.buffer_changed.all()
    assert np.array_equal(f1.buffer, f2.buffer)
    assert np.array_equal(s1.buffer, s2.buffer)
    assert np.max(np.abs(e1.buffer - e2.buffer)) <= 1e-13
    assert s1.spec_rect == s2.spec_rect and np.array_equal(s1.frequencies, s2.frequencies)


@pytest.mark.parametrize('C,order,step_tiles,n', [(4, 4, 3, 3000000), (8, 2, 5, 1000003), (2, 2, 1, 2500000),
                                                  (3, 2, 2, 1500000), (16, 4, 4, 1200000)])
def test_filter_with_fused_minmax(C, order, step_tiles, n):
    """adn_sosfilt_minmax_f64_dev: the filter and the full-trace min/max rows of the raw and of the
    filtered rows in one pass (BASELINE config 4; compresseddata.py:49-52) == the separate entry
    points, bit for bit, including NaN / inf / signed-zero rows and carried filter state."""
    import torch
    from scipy.signal import butter
    from audian_b200 import device
    fs = 96000.
    x = synth(2, n, C, fs, seed=5 + C)
    x[1000, 0] = np.nan
    x[1001, 0] = np.inf
    x[70000:70010, 1] = 0.0
    x[70003, 1] = -0.0
    x[n - 1, C - 1] = -np.inf
    sos = butter(order, (1000., 15000.), 'bandpass', fs=fs, output='sos')
    cg = 1
    while cg < C and cg < 8:
        cg *= 2
    T = (128//cg)*32
    step = step_tiles*T
    xd = torch.from_numpy(x).cuda()
    zi = torch.from_numpy(np.random.default_rng(1).standard_normal((C, sos.shape[0], 2))*1e-3).cuda()
    launches = _lib.launch_count()
    y, zf, rr, rf = device.sosfilt_minmax(sos, xd, step, zi)
    fused = _lib.launch_count() - launches
    assert fused <= 3, 'one filter launch plus the two folds (recordings long enough to fill the device)'
    yr, zfr = device.sosfilt(sos, xd, 0, zi, want_zf=True)
    ok = ~np.isnan(yr.cpu().numpy())
    assert np.array_equal(y.cpu().numpy()[ok], yr.cpu().numpy()[ok])
    assert np.array_equal(zf.cpu().numpy(), zfr.cpu().numpy(), equal_nan=True)
    ref_raw = orc.minmax_rows(x, step)
    got = rr.cpu().numpy()
    rn = np.isnan(ref_raw)
    assert np.array_equal(np.isnan(got), rn) and np.array_equal(got.view(np.uint64)[~rn], ref_raw.view(np.uint64)[~rn])
    ref_f = device.minmax(yr, step).cpu().numpy()
    gf = rf.cpu().numpy()
    fn = np.isnan(ref_f)
    assert np.array_equal(np.isnan(gf), fn) and np.array_equal(gf.view(np.uint64)[~fn], ref_f.view(np.uint64)[~fn])
    # a step that is no multiple of the tile: the separate kernels, same answers
    y2, zf2, rr2, rf2 = device.sosfilt_minmax(sos, xd, step + 7, zi)
    ref2 = orc.minmax_rows(x, step + 7)
    g2 = rr2.cpu().numpy()
    r2n = np.isnan(ref2)
    assert np.array_equal(g2.view(np.uint64)[~r2n], ref2.view(np.uint64)[~r2n])
