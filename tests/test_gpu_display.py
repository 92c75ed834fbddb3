"""Display-side reductions and raw-data ingest against numpy /
the oracle's restatement of the reference's plot-item code."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from audian_b200 import _lib, display
from audian_b200.synth import synth
from oracle import oracle as orc


class Trace(object):
    def __init__(self, buffer, offset, frames, rate, fres=1.0):
        self.buffer, self.offset, self.frames, self.rate = buffer, offset, frames, rate
        self.fresolution = fres

    def __len__(self):
        return self.frames


def make_spec(n=700, C=3, nfft=256, hop=128, fs=48000.):
    x = synth(0, n*hop + nfft, C, fs, seed=5)
    spec = np.empty((n, C, nfft//2 + 1))
    _lib.spectrogram(x, fs, nfft, hop, spec)
    return spec


@pytest.mark.parametrize('resident', [0, 1])
def test_spec_image_and_power_spectrum(resident):
    spec = make_spec()
    spec[17, 1, 5] = 0.0                       # decibel floor: -inf
    if resident:
        # the spectrogram call leaves its result in a mirror; the display calls name it
        m = _lib.Mirror()
        spec2 = np.empty_like(spec)
        x = synth(0, 700*128 + 256, 3, 48000., seed=5)
        _lib.spectrogram(x, 48000., 256, 128, spec2, dst_mirror=m)
        hits = _lib.resident_hits()
        img = _lib.spec_image_db(spec2, 2, src_mirror=m)
        assert _lib.resident_hits() == hits + 1
        assert np.allclose(img, orc.decibel(spec2[:, 2, :].T), rtol=0, atol=1e-9)
        p2 = _lib.mean_power_db(spec2, 1, 10, 60, src_mirror=m)
        assert _lib.resident_hits() == hits + 2
        ref = orc.decibel(np.mean(spec2[10:60, 1, :], axis=0))
        assert np.allclose(p2, np.maximum(ref, -200), rtol=0, atol=1e-9)
        # the owner edits the buffer and says so: the next call uploads
        spec2[3, 2, 7] *= 4.0
        m.invalidate()
        img = _lib.spec_image_db(spec2, 2, src_mirror=m)
        assert _lib.resident_hits() == hits + 2
        assert np.allclose(img, orc.decibel(spec2[:, 2, :].T), rtol=0, atol=1e-9)
        m.release()
    for ch in range(3):
        img = _lib.spec_image_db(spec, ch)
        ref = orc.decibel(spec[:, ch, :].T)
        assert img.shape == ref.shape
        assert np.array_equal(np.isneginf(img), np.isneginf(ref))
        fin = np.isfinite(ref)
        assert np.allclose(img[fin], ref[fin], rtol=0, atol=1e-9)
    tr = Trace(spec, 100, 5000, 48000./128, 48000./256)
    for t0, t1 in ((0.3, 1.0), (0.27, 0.28), (1.5, 2.1)):
        p, f = display.power_spectrum(tr, 1, t0, t1)
        i0 = int(t0*tr.rate)
        i1 = max(int(t1*tr.rate) - 1, i0 + 1)
        ref = orc.decibel(np.mean(spec[i0 - 100:i1 - 100, 1, :], axis=0))
        ref[ref < -200] = -200
        assert np.allclose(p, ref, rtol=0, atol=1e-9)
        assert np.array_equal(f, np.arange(len(p))*tr.fresolution)


def test_noise_levels_match_reference_formula():
    from audian_b200.bufferedspectrogram import BufferedSpectrogram
    spec = make_spec(n=300, C=2)
    s = BufferedSpectrogram()
    s.buffer = spec
    s.init = True
    zmin, zmax = s.estimate_noiselevels(1)
    db = orc.decibel(spec[:, 1, :])
    nf = spec.shape[2]//16
    rmin = np.percentile(db[:, -nf:], 95)
    rmax = np.max(db)
    rmax = rmin + 0.95*(rmax - rmin)
    if rmax - rmin < 20:
        rmax = rmin + 20
    if rmax - rmin > 80:
        rmin = rmax - 80
    assert abs(zmin - rmin) < 1e-8 and abs(zmax - rmax) < 1e-8


def test_trace_decimate_matches_traceitem():
    fs, C = 48000., 4
    x = synth(0, 900000, C, fs, seed=8)
    off = 50000
    tr = Trace(x, off, 5_000_000, fs)
    for t0, t1, px in ((2.0, 15.0, 1920), (1.0, 19.9, 800), (1.1, 1.12, 1920), (0.0, 30.0, 1000)):
        for ch in (0, 3):
            step, start, ref = orc.traceitem_decimate(len(tr), off, x, fs, ch, t0, t1, px)
            gstep, gt, gd = display.trace_decimate(tr, ch, t0, t1, px)
            assert gstep == step
            assert np.array_equal(gd.view(np.uint64), np.asarray(ref).view(np.uint64))
            if step > 1:
                assert np.array_equal(gt, np.arange(start, start + len(ref)*step/2, step/2)/fs)
    # with the trace on the device (its process() left a mirror) the column is gathered there
    from scipy.signal import butter
    sos = butter(2, 5000., 'lowpass', fs=fs, output='sos')
    tr2 = Trace(np.empty_like(x), off, 5_000_000, fs)
    tr2._mirror = _lib.Mirror()
    _lib.sosfilt(sos, x, tr2.buffer, 0, dst_mirror=tr2._mirror)
    hits = _lib.resident_hits()
    step, start, ref = orc.traceitem_decimate(len(tr2), off, tr2.buffer, fs, 2, 2.0, 15.0, 1920)
    gstep, gt, gd = display.trace_decimate(tr2, 2, 2.0, 15.0, 1920)
    assert _lib.resident_hits() == hits + 1
    assert gstep == step and np.array_equal(gd.view(np.uint64), np.asarray(ref).view(np.uint64))


@pytest.mark.parametrize('C,show,het', [(8, [0, 1, 2, 5], 0.0), (8, [3], 0.0), (4, [0, 1, 2, 3], 40000.),
                                        (16, list(range(16)), 35000.), (2, [1], 25000.), (3, [0, 2, 1], 0.0)])
def test_play_region_matches_reference(C, show, het):
    """DataBrowser.play_region up to the fade (databrowser.py:1702-1728): channel means,
    heterodyne, butter(2, 20 kHz) sosfiltfilt, [::nstep]."""
    fs = 250000.
    n = 300000
    x = synth(2, n, C, fs, seed=60 + C)
    ref, rrate = orc.play_region(x, show, fs, het > 0, het)
    n2 = (len(show) + 1)//2
    got, grate = _lib.play_region(x, show[:n2], show[n2:] if len(show) > 1 else [], fs, het)
    assert got.shape == ref.shape and grate == rrate
    assert np.max(np.abs(got - ref)) <= 1e-6*max(1.0, np.max(np.abs(ref)))
    if het == 0:
        assert np.array_equal(got, ref)        # numpy's pairwise mean, bit for bit


def test_play_region_through_trace_class():
    fs, C = 96000., 4
    x = synth(0, 400000, C, fs, seed=3)
    class Loaded(Trace):
        def __getitem__(self, key):
            return self.buffer[key]

    tr = Loaded(x, 0, len(x), fs)
    pd, rate, t0, t1 = display.play_region(tr, [0, 1, 3], -0.5, 2.0, True, 30000.)
    ref, rrate = orc.play_region(x[:int(np.round(2.0*fs))], [0, 1, 3], fs, True, 30000.)
    assert t0 == 0.0 and rate == rrate and pd.shape == ref.shape
    assert np.max(np.abs(pd - ref)) <= 1e-6


@pytest.mark.parametrize('C,clips', [(1, False), (4, True), (7, False)])
def test_unwrap_ingest(C, clips):
    """The loader's unwrap option (data.py:180 -> audioio unwrap): a recording that exceeded
    +-1 and wrapped around in its file comes back."""
    rng = np.random.default_rng(5 + C)
    n = 50000
    t = np.arange(n)[:, None]/48000.
    true = 1.7*np.sin(2*np.pi*(40. + 7.*np.arange(C))*t) + 0.01*rng.standard_normal((n, C))
    wrapped = (true + 1.0) % 2.0 - 1.0
    ref = orc.unwrap(wrapped.copy(), 1.5, clips)
    got = _lib.unwrap(wrapped.copy(), 1.5, clips)
    assert np.array_equal(got, ref)
    if not clips:
        assert np.max(np.abs(got - true)) < 1e-12
    same = wrapped.copy()
    assert _lib.unwrap(same, -1.0) is same and np.array_equal(same, wrapped)   # switched off



@pytest.mark.parametrize('bits', [16, 24, 32])
def test_pcm_ingest(bits):
    rng = np.random.default_rng(bits)
    n, C = 50001, 3
    lim = 1 << (bits - 1)
    v = rng.integers(-lim, lim, size=n*C, dtype=np.int64)
    v[:4] = [-lim, lim - 1, 0, -1]
    if bits == 16:
        raw = v.astype('<i2').view(np.uint8)
    elif bits == 32:
        raw = v.astype('<i4').view(np.uint8)
    else:
        raw = (v.astype('<i4').view(np.uint8).reshape(-1, 4)[:, :3]).copy().reshape(-1)
    got = _lib.pcm_to_f64(raw, bits, C, gain=2.5)
    ref = (v.astype(np.float64)/lim*2.5).reshape(n, C)
    assert got.shape == (n, C) and np.array_equal(got, ref)


def test_compresseddata_class_matches_reference_fixture():
    """audian_b200.CompressedData.start() (short and long path) against what the reference's
    CompressedData.start / down_sample_worker produced (tests/golden/fulltrace.npz)."""
    import json
    import os
    from audian_b200.compresseddata import CompressedData
    from oracle.ref_harness import ArrayLoader
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'fulltrace.npz'))
    for name in ('short_3ch', 'short_1ch_step1', 'long_4ch'):
        a = json.loads(str(g[name + '_args']))
        x = synth(0, a['frames'], a['channels'], a['rate'], a['seed'])
        if name.startswith('short'):
            data = ArrayLoader(x, a['rate'])                     # whole recording in the buffer
        else:
            data = ArrayLoader(x, a['rate'], 0, 50000)           # buffer shorter than the file
        cd = CompressedData(data)
        cd.start(a['max_pixel'], {})
        cd.wait()                                  # long recordings are reduced in the background
        assert not cd.is_busy()
        assert cd.short_data == name.startswith('short')
        assert np.array_equal(cd.times, g[name + '_times'])
        ref = g[name + '_datas']
        assert cd.datas.shape == ref.shape
        assert np.array_equal(cd.datas.view(np.uint64), ref.view(np.uint64)), name


@pytest.mark.parametrize('C,order,fc', [(2, 2, 20000.), (1, 4, 3000.), (8, 2, 500.)])
def test_sosfiltfilt_playback_filter(C, order, fc):
    """The zero-phase low-pass of play_region (databrowser.py:1702-1731): heterodyne, sosfiltfilt,
    decimate -- the filter on the device, against scipy."""
    from scipy.signal import butter, sosfiltfilt
    rate, n = 250000., 120001
    x = synth(0, n, C, rate, seed=60 + C)
    het = np.sin(2*np.pi*40000.*np.arange(n)/rate)
    play = np.ascontiguousarray((x.T*het).T)
    sos = butter(order, fc, 'low', output='sos', fs=rate)
    ref = sosfiltfilt(sos, play, 0)
    got = _lib.sosfiltfilt(sos, play)
    assert np.max(np.abs(got - ref)) <= 1e-9
    nstep = max(1, int(np.round(rate/(2*fc))))
    assert np.max(np.abs(got[::nstep] - ref[::nstep])) <= 1e-9
    with pytest.raises(ValueError):
        _lib.sosfiltfilt(sos, play[:5])
