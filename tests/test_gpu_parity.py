"""Parity of the sm_100a kernels (through the C ABI) with the CPU oracle.

Tolerances (BASELINE.json north_star): min/max bit-exact; filtered/envelope
traces max abs error <= 1e-6 of full scale; spectrogram power rtol 1e-5 (with
an absolute floor of 1e-20 x the largest bin, below which two fp64 FFTs cannot
agree to 1e-5 either).
"""

import glob
import json
import os

import numpy as np
import pytest
from scipy.signal import butter, sosfilt

from audian_b200 import _lib
from audian_b200.synth import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
FS_TOL = 1e-6          # of full scale
SPEC_RTOL = 1e-5
SPEC_ATOL_REL = 1e-20


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def assert_minmax_equal(got, ref, what=''):
    """Bit-exact where the reference is a number (signed zeros included); NaN
    wherever the reference is NaN.  NaN payloads depend on numpy's SIMD dispatch
    on the host that runs the oracle, so they are compared separately."""
    assert got.shape == ref.shape, what
    rn = np.isnan(ref)
    assert np.array_equal(np.isnan(got), rn), what
    assert np.array_equal(bits(got)[~rn], bits(ref)[~rn]), what


def assert_trace_close(got, ref, what=''):
    scale = max(1.0, float(np.max(np.abs(ref))) if ref.size else 1.0)
    err = float(np.max(np.abs(got - ref))) if ref.size else 0.0
    assert err <= FS_TOL*scale, f'{what}: max abs err {err:g} (scale {scale:g})'


def assert_spec_close(got, ref, what=''):
    atol = SPEC_ATOL_REL*float(np.max(ref)) if ref.size else 0.0
    bad = np.abs(got - ref) > atol + SPEC_RTOL*np.abs(ref)
    assert not bad.any(), f'{what}: {bad.sum()} bins off, worst rel ' \
        f'{np.max(np.abs(got - ref)[bad]/np.maximum(np.abs(ref[bad]), 1e-300)):g}'


# ---------------------------------------------------------------- min/max

@pytest.mark.parametrize('n,C,step', [
    (1, 1, 1), (7, 1, 1), (1000, 1, 7), (1000, 2, 1000), (1000, 2, 5000),
    (4099, 3, 17), (50000, 3, 71), (100003, 4, 333), (65536, 8, 2048),
    (300001, 8, 40000), (200000, 16, 100000), (30000, 64, 977), (5000, 5, 1250),
    (1 << 20, 1, 1 << 18), (1 << 20, 2, 1 << 20), (777777, 6, 111111),
])
def test_minmax_bit_exact(n, C, step):
    x = synth(5, n, C, 48000., seed=n + C)
    ref = orc.minmax_rows(x, step)
    got = _lib.minmax(x, step)
    assert got.shape == ref.shape
    assert np.array_equal(bits(got), bits(ref))


def test_minmax_special_values():
    rng = np.random.default_rng(3)
    n, C = 60000, 4
    x = rng.standard_normal((n, C))
    x[rng.integers(0, n, 4000), rng.integers(0, C, 4000)] = 0.0
    x[rng.integers(0, n, 4000), rng.integers(0, C, 4000)] = -0.0
    x[rng.integers(0, n, 30), rng.integers(0, C, 30)] = np.inf
    x[rng.integers(0, n, 30), rng.integers(0, C, 30)] = -np.inf
    nan1 = np.array([0x7ff8000000000001], dtype=np.uint64).view(np.float64)[0]
    nan2 = np.array([0xfff8000000000abc], dtype=np.uint64).view(np.float64)[0]
    x[rng.integers(0, n, 6), rng.integers(0, C, 6)] = nan1
    x[rng.integers(0, n, 6), rng.integers(0, C, 6)] = nan2
    for step in (3, 500, 9000, 60000):
        ref = orc.minmax_rows(x, step)
        got = _lib.minmax(x, step)
        assert_minmax_equal(got, ref, step)
    # signed zeros only: the later row decides the sign
    z = np.where(rng.random((40000, 2)) < 0.5, 0.0, -0.0)
    for step in (2, 64, 5000, 40000):
        assert np.array_equal(bits(_lib.minmax(z, step)), bits(orc.minmax_rows(z, step)))


@pytest.mark.parametrize('C,step', [(1, 2048), (1, 2051), (2, 1100), (4, 700), (8, 256), (8, 1921),
                                    (16, 130), (32, 64), (64, 40), (64, 333)])
def test_minmax_many_short_segments(C, step):
    """Thousands of short segments: the warp-per-segment kernel, bit-exact incl. NaN payloads,
    signed zeros and a partial last segment."""
    rng = np.random.default_rng(C*1000 + step)
    n = step*1300 + step//3
    x = synth(1, n, C, 48000., seed=C + step)
    x[rng.integers(0, n, 3000), rng.integers(0, C, 3000)] = 0.0
    x[rng.integers(0, n, 3000), rng.integers(0, C, 3000)] = -0.0
    # one NaN payload only: which of several different NaNs of a segment numpy returns depends
    # on its SIMD dispatch (the first for 2 columns, the last for >= 4 on this build); the
    # library returns the last (test_minmax_special_values pins that for 4 columns)
    x[rng.integers(0, n, 80), rng.integers(0, C, 80)] = np.array([0x7ff8000000000123], dtype=np.uint64).view(np.float64)[0]
    x[rng.integers(0, n, 40), rng.integers(0, C, 40)] = -np.inf
    # whole segments of zeros of one sign, then a late zero of the other
    x[5*step:6*step] = 0.0
    x[6*step - 1, 0] = -0.0
    got = _lib.minmax(x, step)
    ref = orc.minmax_rows(x, step)
    if C == 1:
        # numpy's 1-D SIMD path: signed-zero ties are lane-order dependent there
        same = (bits(got) == bits(ref)) | ((got == 0) & (ref == 0))
        assert same.all()
    else:
        assert np.array_equal(bits(got), bits(ref))


def test_minmax_golden_fulltrace():
    g = np.load(os.path.join(GOLDEN, 'fulltrace.npz'))
    for name in ('short_3ch', 'short_1ch_step1', 'long_4ch'):
        a = json.loads(str(g[name + '_args']))
        x = synth(0, a['frames'], a['channels'], a['rate'], a['seed'])
        step = max(1, a['frames']//a['max_pixel'])
        got = _lib.minmax(x, step)
        ref = g[name + '_datas']
        assert np.array_equal(bits(got), bits(ref[:len(got)]))
        assert not ref[len(got):].any()


# ---------------------------------------------------------------- filter

FILTERS = {
    'lp2': lambda fs: butter(2, 0.2*fs, 'lowpass', fs=fs, output='sos'),
    'hp2': lambda fs: butter(2, 0.02*fs, 'highpass', fs=fs, output='sos'),
    'bp2': lambda fs: butter(2, (0.02*fs, 0.3*fs), 'bandpass', fs=fs, output='sos'),
    'lp4': lambda fs: butter(4, 0.1*fs, 'lowpass', fs=fs, output='sos'),
    'bp4': lambda fs: butter(4, (0.02*fs, 0.3*fs), 'bandpass', fs=fs, output='sos'),
    'bp3': lambda fs: butter(3, (0.05*fs, 0.2*fs), 'bandpass', fs=fs, output='sos'),
    'lp8': lambda fs: butter(8, 0.15*fs, 'lowpass', fs=fs, output='sos'),
    'bp8': lambda fs: butter(8, (0.1*fs, 0.2*fs), 'bandpass', fs=fs, output='sos'),
    'hp_low': lambda fs: butter(2, 0.0006*fs, 'highpass', fs=fs, output='sos'),
}


@pytest.mark.parametrize('C', [1, 2, 3, 4, 8, 16, 24, 64])
@pytest.mark.parametrize('filt', ['lp2', 'bp2', 'bp4'])
def test_sosfilt_channels(C, filt):
    fs = 48000.
    n = 70001
    x = synth(0, n, C, fs, seed=C)
    sos = FILTERS[filt](fs)
    ref = np.empty((n, C))
    orc.filter_process(sos, x, ref, 0)
    got = np.empty((n, C))
    _lib.sosfilt(sos, x, got, 0)
    assert_trace_close(got, ref, f'{filt} C={C}')


@pytest.mark.parametrize('filt', sorted(FILTERS))
def test_sosfilt_filters(filt):
    fs = 96000.
    n, C = 123457, 4
    x = synth(11, n, C, fs, seed=77)
    sos = FILTERS[filt](fs)
    ref = np.empty((n, C))
    orc.filter_process(sos, x, ref, 0)
    got = np.empty((n, C))
    _lib.sosfilt(sos, x, got, 0)
    assert_trace_close(got, ref, filt)


@pytest.mark.parametrize('n', [1, 2, 31, 32, 33, 255, 1024, 8191, 8192, 8193, 20000])
def test_sosfilt_lengths_and_nbefore(n):
    fs = 20000.
    sos = FILTERS['bp2'](fs)
    for C in (1, 2, 8):
        x = synth(3, n, C, fs, seed=n)
        for nbefore in sorted({0, min(5, n - 1), n//2}):
            ref = np.empty((n - nbefore, C))
            orc.filter_process(sos, x, ref, nbefore)
            got = np.full((n - nbefore, C), np.nan)
            _lib.sosfilt(sos, x, got, nbefore)
            assert_trace_close(got, ref, f'n={n} C={C} nbefore={nbefore}')


def test_sosfilt_none_is_copy():
    x = synth(0, 5000, 3, 1000.)
    got = np.empty((4990, 3))
    _lib.sosfilt(None, x, got, 10)
    assert np.array_equal(got, x[10:])


def test_sosfilt_streaming_zi():
    fs = 48000.
    C, n = 4, 300000
    x = synth(0, n, C, fs, seed=5)
    sos = FILTERS['bp4'](fs)
    S = sos.shape[0]
    ref = np.empty((n, C))
    orc.filter_process(sos, x, ref, 0)
    zi = np.zeros((C, S, 2))
    got = np.empty((n, C))
    pos = 0
    for chunk in (1, 777, 8192, 100000, 50001, n):
        stop = min(n, pos + chunk)
        _lib.sosfilt(sos, x[pos:stop], got[pos:stop], 0, zi=zi)
        pos = stop
        if pos >= n:
            break
    assert pos == n
    assert_trace_close(got, ref, 'streamed')
    # final state equals scipy's zf
    for c in range(C):
        _, zf = sosfilt(sos, x[:, c], zi=np.zeros((S, 2)))
        assert np.max(np.abs(zi[c] - zf)) <= 1e-9*max(1.0, np.max(np.abs(zf)))


def test_sosfilt_impulse_response():
    fs = 1000.
    sos = FILTERS['lp4'](fs)
    x = np.zeros((20000, 2))
    x[0, 0] = 1.0
    x[9000, 1] = -2.0
    got = np.empty_like(x)
    _lib.sosfilt(sos, x, got)
    ref = np.empty_like(x)
    orc.filter_process(sos, x, ref, 0)
    assert np.max(np.abs(got - ref)) <= 1e-12


# ---------------------------------------------------------------- run kernel vs look-back kernel

def _both_scan_kernels(call):
    """call() once with the run kernel allowed and once with the look-back kernel only;
    returns (result with runs, result with look-back, run-kernel launches of the first call)."""
    _lib.set_option(_lib.ADN_OPT_SCAN_RUNS, 1)
    chunk = _lib.get_option(_lib.ADN_OPT_CHUNK_BYTES)
    _lib.set_option(_lib.ADN_OPT_CHUNK_BYTES, 1 << 30)      # the whole trace in one piece
    try:
        r0 = _lib.scan_run_count()
        a = call()
        nrun = _lib.scan_run_count() - r0
        _lib.set_option(_lib.ADN_OPT_SCAN_RUNS, 0)
        r1 = _lib.scan_run_count()
        b = call()
        assert _lib.scan_run_count() == r1, 'look-back only: the run kernel must not launch'
    finally:
        _lib.set_option(_lib.ADN_OPT_SCAN_RUNS, 1)
        _lib.set_option(_lib.ADN_OPT_CHUNK_BYTES, chunk)
    return a, b, nrun


@pytest.mark.parametrize('C,n,filt', [
    (8, 1200000, 'bp2'), (8, 600001, 'lp2'), (8, 1800123, 'bp4'), (64, 200000, 'bp2'),
    (3, 2000003, 'bp2'), (1, 8000000, 'lp4'), (2, 4000000, 'hp2'), (24, 500000, 'bp3'),
    (4, 3500000, 'bp4'), (16, 600000, 'lp2'),
])
def test_sosfilt_run_kernel(C, n, filt):
    """Long traces take the run kernel (a block per run of tiles, run-in from zero state): same
    result as the look-back kernel (to rounding) and as scipy (to the parity tolerance)."""
    fs = 48000.
    x = synth(5, n, C, fs, seed=C + n % 97)
    sos = FILTERS[filt](fs)
    nbefore = 37

    def call():
        got = np.empty((n - nbefore, C))
        _lib.sosfilt(sos, x, got, nbefore)
        return got

    a, b, nrun = _both_scan_kernels(call)
    assert nrun >= 1, 'this shape is meant to take the run kernel'
    assert np.max(np.abs(a - b)) <= 1e-12*max(1.0, np.max(np.abs(b)))
    ref = np.empty((n - nbefore, C))
    orc.filter_process(sos, x, ref, nbefore)
    assert_trace_close(a, ref, f'run kernel {filt} C={C}')
    assert np.max(np.abs(a - ref)) <= 1e-11*max(1.0, np.max(np.abs(ref)))


def test_sosfilt_run_kernel_streaming_state():
    """zi/zf through the run kernel: chunks long enough for it, state carried between them."""
    fs = 96000.
    C, n = 4, 3000000
    x = synth(0, n, C, fs, seed=15)
    sos = FILTERS['bp2'](fs)
    S = sos.shape[0]
    ref = np.empty((n, C))
    orc.filter_process(sos, x, ref, 0)
    zi = np.zeros((C, S, 2))
    got = np.empty((n, C))
    r0 = _lib.scan_run_count()
    pos = 0
    old = _lib.get_option(_lib.ADN_OPT_CHUNK_BYTES)
    _lib.set_option(_lib.ADN_OPT_CHUNK_BYTES, 1 << 30)
    try:
        for chunk in (1200000, 1000001, n):
            stop = min(n, pos + chunk)
            _lib.sosfilt(sos, x[pos:stop], got[pos:stop], 0, zi=zi)
            pos = stop
    finally:
        _lib.set_option(_lib.ADN_OPT_CHUNK_BYTES, old)
    assert _lib.scan_run_count() - r0 >= 1
    assert_trace_close(got, ref, 'streamed through the run kernel')
    assert np.max(np.abs(got - ref)) <= 1e-11
    for c in range(C):
        _, zf = sosfilt(sos, x[:, c], zi=np.zeros((S, 2)))
        assert np.max(np.abs(zi[c] - zf)) <= 1e-9*max(1.0, np.max(np.abs(zf)))


def test_sosfilt_slow_cascade_keeps_look_back():
    """A cascade that needs tens of tiles to forget its state is not given a run-in."""
    fs = 96000.
    n, C = 800000, 8
    x = synth(1, n, C, fs, seed=3) + 0.25            # offset: large slowly decaying state
    sos = FILTERS['hp_low'](fs)
    r0 = _lib.scan_run_count()
    got = np.empty((n, C))
    _lib.sosfilt(sos, x, got, 0)
    assert _lib.scan_run_count() == r0
    ref = np.empty((n, C))
    orc.filter_process(sos, x, ref, 0)
    assert_trace_close(got, ref, 'hp_low')


@pytest.mark.parametrize('C,n,order,hp', [(8, 1500000, 2, 0), (64, 300000, 2, 0), (2, 5000000, 4, 0),
                                          (3, 1200000, 2, 50.), (8, 700001, 4, 20.)])
def test_envelope_run_kernel(C, n, order, hp):
    fs = 48000.
    x = synth(9, n, C, fs, seed=C + 200)
    sos = orc.envelope_design(fs, 500., hp, order)
    nbefore = 11

    def call():
        got = np.empty((n - nbefore, C))
        _lib.envelope(sos, x, got, nbefore, hp == 0)
        return got

    _lib.set_option(_lib.ADN_OPT_ZERO_PHASE_ONEPASS, 0)     # the two sweeps through memory
    try:
        a, b, nrun = _both_scan_kernels(call)
    finally:
        _lib.set_option(_lib.ADN_OPT_ZERO_PHASE_ONEPASS, 1)
    if hp == 0:
        assert nrun >= 2, 'both sweeps are meant to take the run kernel'
    assert np.max(np.abs(a - b)) <= 1e-11*max(1.0, np.max(np.abs(b)))
    ref = np.empty((n - nbefore, C))
    orc.envelope_process(sos, x, ref, nbefore, hp)
    assert_trace_close(a, ref, f'envelope run kernel C={C} order={order} hp={hp}')
    assert np.max(np.abs(a - ref)) <= 1e-10*max(1.0, np.max(np.abs(ref)))


# ---------------------------------------------------------------- one-pass zero-phase kernel

@pytest.mark.parametrize('C,n,order,fc,nbefore', [
    (8, 1500000, 2, 500., 0), (8, 700001, 2, 500., 11), (64, 300000, 2, 500., 0),
    (1, 3000001, 2, 500., 5), (2, 2500000, 4, 500., 0), (3, 1200000, 2, 800., 7),
    (4, 900000, 2, 2000., 0), (13, 400000, 2, 500., 1), (16, 600000, 4, 1000., 100000),
    (8, 123457, 2, 4000., 0)])
@pytest.mark.parametrize('pipe', [1, 0])
def test_envelope_onepass_kernel(C, n, order, fc, nbefore, pipe, monkeypatch):
    """csrc/zerophase.cu (one pass, tile in registers; pipe = 1: the register pipeline where the
    cascade allows it, 0: the variant that reads every tile twice) against the two sweeps
    through memory and against scipy's sosfiltfilt (oracle); bufferedenvelope.py:34-41."""
    monkeypatch.setenv('ADN_ZP_PIPE', str(pipe))
    fs = 48000.
    x = synth(3, n, C, fs, seed=C + 300)
    sos = orc.envelope_design(fs, fc, 0, order)

    def call():
        got = np.empty((n - nbefore, C))
        _lib.envelope(sos, x, got, nbefore, True)
        return got

    z0 = _lib.zero_phase_count()
    a = call()
    assert _lib.zero_phase_count() == z0 + 1, 'the one-pass kernel is meant to take this shape'
    _lib.set_option(_lib.ADN_OPT_ZERO_PHASE_ONEPASS, 0)
    try:
        b = call()
        assert _lib.zero_phase_count() == z0 + 1
    finally:
        _lib.set_option(_lib.ADN_OPT_ZERO_PHASE_ONEPASS, 1)
    ref = np.empty((n - nbefore, C))
    orc.envelope_process(sos, x, ref, nbefore, 0)
    scale = max(1.0, np.max(np.abs(ref)))
    assert np.max(np.abs(a - b)) <= 1e-11*scale
    assert np.max(np.abs(a - ref)) <= 1e-10*scale, f'one-pass envelope C={C} n={n} order={order}'
    assert (a >= 0).all()


@pytest.mark.parametrize('C,n,fs,fc', [(16, 900000, 250000., 500.), (8, 1500000, 250000., 500.),
                                       (16, 900000, 500000., 1000.), (6, 800001, 250000., 300.),
                                       (64, 600000, 250000., 500.)])
def test_envelope_onepass_narrow_groups(C, n, fs, fc):
    """Slow cascades (the 500-Hz envelope of ultrasound recordings): the pipelined kernel takes
    narrower channel groups = longer tiles, so that the tiles it has to park still fit
    (zerophase.cu: zero_phase_regs_dev); same answers as the two sweeps and scipy."""
    x = synth(3, n, C, fs, seed=C + 700)
    sos = orc.envelope_design(fs, fc, 0, 2)

    def call():
        got = np.empty((n, C))
        _lib.envelope(sos, x, got, 0, True)
        return got

    z0 = _lib.zero_phase_count()
    a = call()
    assert _lib.zero_phase_count() == z0 + 1
    _lib.set_option(_lib.ADN_OPT_ZERO_PHASE_ONEPASS, 0)
    try:
        b = call()
    finally:
        _lib.set_option(_lib.ADN_OPT_ZERO_PHASE_ONEPASS, 1)
    ref = np.empty((n, C))
    orc.envelope_process(sos, x, ref, 0, 0)
    scale = max(1.0, np.max(np.abs(ref)))
    assert np.max(np.abs(a - b)) <= 1e-11*scale
    assert np.max(np.abs(a - ref)) <= 1e-10*scale
    assert (a >= 0).all()


@pytest.mark.parametrize('C,n,order', [(8, 800000, 2), (2, 1000003, 4), (5, 500000, 6)])
def test_sosfiltfilt_onepass_kernel(C, n, order):
    """The same kernel without the rectification == scipy.signal.sosfiltfilt (databrowser.py:1725)."""
    from scipy.signal import sosfiltfilt
    fs = 96000.
    x = synth(1, n, C, fs, seed=C + 400)
    sos = butter(order, 12000., 'lowpass', fs=fs, output='sos')
    z0 = _lib.zero_phase_count()
    got = _lib.sosfiltfilt(sos, x)
    assert _lib.zero_phase_count() == z0 + 1
    ref = sosfiltfilt(sos, x, axis=0)
    assert np.max(np.abs(got - ref)) <= 1e-10*max(1.0, np.max(np.abs(ref)))


def test_envelope_onepass_falls_back():
    """Cascades that forget slowly (band-pass envelope with a 20-Hz high-pass) and short inputs
    keep the two sweeps."""
    fs = 48000.
    x = synth(0, 400000, 4, fs, seed=77)
    z0 = _lib.zero_phase_count()
    sos = orc.envelope_design(fs, 500., 20., 2)
    got = np.empty_like(x)
    _lib.envelope(sos, x, got, 0, False)
    ref = np.empty_like(x)
    orc.envelope_process(sos, x, ref, 0, 20.)
    assert_trace_close(got, ref, 'slow cascade')
    sos = orc.envelope_design(fs, 500.)
    got = np.empty((5000, 4))
    _lib.envelope(sos, x[:5000], got, 0, True)
    ref = np.empty((5000, 4))
    orc.envelope_process(sos, x[:5000], ref, 0, 0)
    assert_trace_close(got, ref, 'short input')
    assert _lib.zero_phase_count() == z0


# ---------------------------------------------------------------- envelope

@pytest.mark.parametrize('C', [1, 2, 3, 8, 32, 64])
def test_envelope_channels(C):
    fs = 48000.
    n = 50000
    x = synth(0, n, C, fs, seed=C + 100)
    sos = orc.envelope_design(fs, 500.)
    ref = np.empty((n, C))
    orc.envelope_process(sos, x, ref, 0, 0)
    got = np.empty((n, C))
    _lib.envelope(sos, x, got, 0, True)
    assert_trace_close(got, ref, f'C={C}')
    assert (got >= 0).all()


@pytest.mark.parametrize('order,hp', [(2, 0), (4, 0), (2, 50.), (4, 20.)])
def test_envelope_orders(order, hp):
    fs = 20000.
    n, C = 90000, 2
    x = synth(7, n, C, fs, seed=9)
    sos = orc.envelope_design(fs, 300., hp, order)
    nbefore = 123
    ref = np.empty((n - nbefore, C))
    orc.envelope_process(sos, x, ref, nbefore, hp)
    got = np.empty((n - nbefore, C))
    _lib.envelope(sos, x, got, nbefore, hp == 0)
    assert_trace_close(got, ref, f'order={order} hp={hp}')


def test_envelope_short_input_raises():
    sos = orc.envelope_design(1000., 100.)
    edge = orc.sosfiltfilt_edge(sos)
    x = synth(0, edge, 2, 1000.)
    with pytest.raises(ValueError):
        _lib.envelope(sos, x, np.empty_like(x))
    x = synth(0, edge + 1, 2, 1000.)
    ref = np.empty_like(x)
    orc.envelope_process(sos, x, ref, 0, 0)
    got = np.empty_like(x)
    _lib.envelope(sos, x, got)
    assert_trace_close(got, ref, 'edge+1')


def test_envelope_none_is_zero():
    x = synth(0, 100, 2, 1000.)
    got = np.full_like(x, 7.0)
    _lib.envelope(None, x, got)
    assert not got.any()


# ---------------------------------------------------------------- spectrogram

@pytest.mark.parametrize('nfft', [8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384,
                                  32768, 131072])
def test_spectrogram_nfft(nfft):
    fs = 48000.
    C = 2
    for hop in sorted({nfft, nfft//2, nfft//8}):
        n_src = nfft*5 + hop*3 + 1
        x = synth(0, n_src, C, fs, seed=nfft)
        n_dst = (n_src + hop - 1)//hop
        ref = np.empty((n_dst, C, nfft//2 + 1))
        nref = orc.spectrogram_process(x, ref, fs, nfft, hop)
        got = np.full_like(ref, np.nan)
        ngot = _lib.spectrogram(x, fs, nfft, hop, got)
        assert ngot == nref
        assert_spec_close(got, ref, f'nfft={nfft} hop={hop}')
        assert not got[ngot:].any()


@pytest.mark.parametrize('C', [1, 2, 3, 5, 8, 16, 64])
def test_spectrogram_channels(C):
    fs = 250000.
    nfft, hop = 256, 128
    x = synth(100, 9000, C, fs, seed=C)
    n_dst = 9000//hop
    ref = np.empty((n_dst, C, nfft//2 + 1))
    orc.spectrogram_process(x, ref, fs, nfft, hop)
    got = np.empty_like(ref)
    _lib.spectrogram(x, fs, nfft, hop, got)
    assert_spec_close(got, ref, f'C={C}')


@pytest.mark.parametrize('fill', ['0', '1'])
@pytest.mark.parametrize('C,nfft,hop', [(8, 1024, 512), (8, 256, 128), (4, 128, 32), (16, 512, 512),
                                        (64, 256, 128), (6, 1024, 128), (2, 512, 64)])
def test_spectrogram_ring_fill_variants(C, nfft, hop, fill, monkeypatch):
    """The ring kernel fills its ring through registers or by cp.async (a template parameter the
    launcher picks by shape, spectrogram.cu: launch_ring) and prefetches the next rows into L2 whole
    or in shares: every variant on long runs (several steps per block), against the oracle."""
    monkeypatch.setenv('ADN_SPEC_ASYNC', fill)
    monkeypatch.setenv('ADN_SPEC_PFSPLIT', fill)
    fs = 96000.
    n_src = 300000 + 7
    x = synth(11, n_src, C, fs, seed=nfft + C + hop) + 0.25
    n_dst = (n_src - (nfft - hop))//hop
    ref = np.empty((n_dst, C, nfft//2 + 1))
    nref = orc.spectrogram_process(x, ref, fs, nfft, hop)
    got = np.full_like(ref, np.nan)
    assert _lib.spectrogram(x, fs, nfft, hop, got) == nref
    assert_spec_close(got, ref, f'C={C} nfft={nfft} hop={hop} fill={fill}')


@pytest.mark.parametrize('tma', ['0', '1'])
@pytest.mark.parametrize('C,nfft,hop', [(8, 2048, 2048), (8, 2048, 1024), (16, 2048, 512), (4, 4096, 2048),
                                        (64, 4096, 1024), (6, 4096, 4096), (8, 2048, 256), (8, 4096, 768)])
def test_spectrogram_tma_gather(C, nfft, hop, tma, monkeypatch):
    """nfft 2048 / 4096: the rows of a channel pair come in through the TMA unit (2-D tensor map over the
    source, frame-per-block kernel and ring-staged kernel for overlapping frames) or through the
    threads' own loads (ADN_SPEC_TMA=0): both against the oracle, runs of many frames per block,
    hops the tensor boxes do not divide fall back."""
    monkeypatch.setenv('ADN_SPEC_TMA', tma)
    monkeypatch.setenv('ADN_SPEC_MPRT', '1' if tma == '1' else '-1')
    fs = 250000.
    n_src = 150 * hop + nfft + 5
    x = synth(3, n_src, C, fs, seed=nfft + C + hop) - 0.5
    n_dst = (n_src - (nfft - hop))//hop
    ref = np.empty((n_dst, C, nfft//2 + 1))
    nref = orc.spectrogram_process(x, ref, fs, nfft, hop)
    got = np.full_like(ref, np.nan)
    assert _lib.spectrogram(x, fs, nfft, hop, got) == nref
    assert_spec_close(got, ref, f'C={C} nfft={nfft} hop={hop} tma={tma}')


@pytest.mark.parametrize('nfft,C', [(2048, 1), (2048, 3), (4096, 8), (8192, 5), (16384, 2), (4096, 64)])
def test_spectrogram_midsize_channels(nfft, C):
    fs = 250000.
    hop = nfft//4
    n_src = nfft*2 + hop*5 + 3
    x = synth(7, n_src, C, fs, seed=nfft + C) - 0.125
    n_dst = n_src//hop
    ref = np.empty((n_dst, C, nfft//2 + 1))
    nref = orc.spectrogram_process(x, ref, fs, nfft, hop)
    got = np.full_like(ref, np.nan)
    assert _lib.spectrogram(x, fs, nfft, hop, got) == nref
    assert_spec_close(got, ref, f'nfft={nfft} C={C}')
    db = np.empty_like(ref)
    _lib.spectrogram(x, fs, nfft, hop, db, out_db=True)
    assert np.allclose(db[:nref], orc.decibel(ref[:nref]), rtol=0, atol=1e-6)


def test_spectrogram_largest_gui_nfft():
    # 2^19 is the largest size the GUI offers (databrowser.py:516)
    fs, nfft = 500000., 1 << 19
    hop = nfft//2
    x = synth(0, nfft*2 + hop + 17, 1, fs, seed=19)
    n_dst = 4
    ref = np.empty((n_dst, 1, nfft//2 + 1))
    nref = orc.spectrogram_process(x, ref, fs, nfft, hop)
    got = np.full_like(ref, np.nan)
    assert _lib.spectrogram(x, fs, nfft, hop, got) == nref == 4
    assert_spec_close(got, ref, 'nfft=2^19')
    db = np.empty_like(ref)
    _lib.spectrogram(x, fs, nfft, hop, db, out_db=True)
    assert np.allclose(db, orc.decibel(ref), rtol=0, atol=1e-6)


def test_spectrogram_short_source_zero_fill():
    x = synth(0, 100, 2, 1000.)
    got = np.full((4, 2, 129), 5.0)
    assert _lib.spectrogram(x, 1000., 256, 128, got) == 0
    assert not got.any()


def test_spectrogram_known_answers():
    fs, N = 1000., 256
    t = np.arange(N*8)
    # sine exactly on bin k0: peak A^2 N/(3 fs), neighbours a quarter of it
    k0, A = 20, 0.7
    x = (A*np.sin(2*np.pi*k0*t/N))[:, None].copy()
    got = np.empty((x.shape[0]//N, 1, N//2 + 1))
    _lib.spectrogram(x, fs, N, N, got)
    peak = A*A*N/(3*fs)
    assert np.allclose(got[:, 0, k0], peak, rtol=1e-9)
    assert np.allclose(got[:, 0, k0 - 1], peak/4, rtol=1e-9)
    assert np.allclose(got[:, 0, k0 + 1], peak/4, rtol=1e-9)
    # DC: the mean removal leaves nothing
    x = np.full((N*4, 2), 0.37)
    got = np.empty((4, 2, N//2 + 1))
    _lib.spectrogram(x, fs, N, N, got)
    assert np.max(got) <= 1e-30
    db = np.empty_like(got)
    _lib.spectrogram(x, fs, N, N, db, out_db=True)
    assert np.all(np.isneginf(db))


def test_spectrogram_db_and_decibel():
    fs = 48000.
    x = synth(0, 20000, 2, fs)
    lin = np.empty((77, 2, 129))
    _lib.spectrogram(x, fs, 256, 128, lin)
    db = np.empty_like(lin)
    _lib.spectrogram(x, fs, 256, 128, db, out_db=True)
    ref = orc.decibel(lin)
    assert np.allclose(db, ref, rtol=0, atol=1e-9)
    assert np.allclose(_lib.decibel(lin), ref, rtol=0, atol=1e-9)
    p = np.array([1e-20, 1e-21, 0.0, 2e-20, 1.0, 100.0])
    assert np.array_equal(np.isneginf(_lib.decibel(p)), np.isneginf(orc.decibel(p)))


@pytest.mark.parametrize('nfft,hop,C', [(1000, 500, 2), (12, 5, 3), (999, 333, 1), (3072, 3072, 2),
                                        (25000, 12500, 2), (100, 67, 4), (40001, 20000, 1)])
def test_spectrogram_nfft_not_a_power_of_two(nfft, hop, C):
    # update() clamps nfft to len(source)//2 (bufferedspectrogram.py:88), open() truncates
    # hop = int(nfft*(1 - overlap)): any integer pair can arrive
    fs = 44100.
    n_src = nfft*3 + hop*2 + 5
    x = synth(0, n_src, C, fs, seed=nfft) + 0.25          # a DC offset the mean removal must take out
    n_dst = (n_src + hop - 1)//hop
    ref = np.empty((n_dst, C, nfft//2 + 1))
    nref = orc.spectrogram_process(x, ref, fs, nfft, hop)
    got = np.full_like(ref, np.nan)
    assert _lib.spectrogram(x, fs, nfft, hop, got) == nref
    assert_spec_close(got, ref, f'nfft={nfft} hop={hop}')
    assert not got[nref:].any()


def test_spectrogram_unsupported_nfft_fails_loudly():
    x = synth(0, (1 << 20) + 5000, 1, 1000.)
    with pytest.raises(_lib.AdnError):
        _lib.spectrogram(x, 1000., (1 << 20) + 2, 500, np.empty((2, 1, (1 << 19) + 2)))
    with pytest.raises(_lib.AdnError):
        _lib.spectrogram(x[:5000], 1000., 4, 2, np.empty((8, 1, 3)))


# ---------------------------------------------------------------- golden chains

@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN, 'chain_*.npz'))))
def test_golden_chain(path):
    g = np.load(path)
    a = json.loads(str(g['args']))
    x = synth(0, a['frames'], a['channels'], a['rate'], a['seed'])
    boff = a['buf_offset']
    blen = a['frames'] - boff if a['buf_frames'] is None else a['buf_frames']
    raw = x[boff:boff + blen]
    C = a['channels']
    # filtered: the reference filters the part of the raw buffer that survives
    # the align_buffer margins, from zero state
    fbuf = g['filt_buffer']
    foff = int(g['filt_offset'])
    sos = g['filt_sos'] if len(g['filt_sos']) else None
    src = raw[foff - boff:foff - boff + len(fbuf)]
    got = np.empty_like(fbuf)
    _lib.sosfilt(sos, src, got, 0)
    assert_trace_close(got, fbuf, 'filtered')
    # spectrogram of the reference's filtered buffer
    sbuf = g['spec_buffer']
    hop = int(g['spec_hop'])
    so, sn, nb = orc.load_buffer_slice(int(g['spec_offset']), len(sbuf), float(g['spec_rate']),
                                       a['rate'], foff, len(fbuf), 0, 10)
    gots = np.empty_like(sbuf)
    _lib.spectrogram(fbuf[so:so + sn], a['rate'], a['nfft'], hop, gots)
    assert_spec_close(gots, sbuf, 'spectrogram')
    # envelope of the reference's filtered buffer
    ebuf = g['env_buffer']
    eoff = int(g['env_offset'])
    esos = g['env_sos'] if len(g['env_sos']) else None
    so, sn, nb = orc.load_buffer_slice(eoff, len(ebuf), a['rate'], a['rate'], foff, len(fbuf), 1, 0)
    gote = np.empty_like(ebuf)
    _lib.envelope(esos, fbuf[so:so + sn], gote, nb, True)
    assert_trace_close(gote, ebuf, 'envelope')


def test_launch_counter_moves():
    before = _lib.launch_count()
    _lib.minmax(synth(0, 5000, 2, 1000.), 100)
    assert _lib.launch_count() > before


def test_minmax_nan_payload_rule():
    """The rule the kernel implements (numpy 2.3.5, 2-D axis-0 reduceat): the
    last NaN of a segment survives; for C == 1 the canonical quiet NaN."""
    nan1 = np.array([0x7ff8000000000001], dtype=np.uint64).view(np.float64)[0]
    nan2 = np.array([0xfff8000000000abc], dtype=np.uint64).view(np.float64)[0]
    for n in (20, 50000):
        x = np.random.default_rng(0).standard_normal((n, 4))
        x[5, :] = nan1
        x[n - 3, :] = nan2
        got = _lib.minmax(x, n)
        assert (bits(got) == 0xfff8000000000abc).all()
        x[5, :] = nan2
        x[n - 3, :] = nan1
        got = _lib.minmax(x, n)
        assert (bits(got) == 0x7ff8000000000001).all()
        got = _lib.minmax(np.ascontiguousarray(x[:, :1]), n)
        assert (bits(got) == 0x7ff8000000000000).all()
