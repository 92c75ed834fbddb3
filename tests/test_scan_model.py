"""Host-side model of the sm_100a scan kernel (audian_b200/csrc/sosfilt.cu).

Re-enacts, in numpy, exactly the decomposition the kernel uses -- pass A dot
products, Kogge-Stone over sub-chunks with A^(L 2^k), warp prefixes with A^(L GW k),
look-back over tiles with (A^T)^j weights, fix-up with A^(L gl), pass B -- with
the matrices the library's host code produces (adn_sos_state_space), and checks
it against scipy.signal.sosfilt.  Runs on CPU: it pins the algebra the kernel
relies on before any GPU time is spent.
"""

import numpy as np
import pytest
from scipy.signal import butter, sosfilt

from audian_b200 import _lib

L, NT, NW = 32, 128, 4


def pick_cg(C):
    cg = 1
    while cg < C and cg < 32:
        cg <<= 1
    return cg


def df2t_chunk(sos, x, z):
    """pass B: the exact recurrence from state z (D,), returns y and final state."""
    S = sos.shape[0]
    z = z.copy()
    y = np.empty(len(x))
    for i, xc in enumerate(x):
        for s in range(S):
            b0, b1, b2, _, a1, a2 = sos[s]
            yn = b0*xc + z[2*s]
            z[2*s] = b1*xc - a1*yn + z[2*s + 1]
            z[2*s + 1] = b2*xc - a2*yn
            xc = yn
        y[i] = xc
    return y, z


def run_model(sos, x, CG, rng, zi=None):
    S = sos.shape[0]
    D = 2*S
    n = len(x)
    GW = 32//CG
    G = NT//CG
    T = G*L
    A, B, AL = _lib.sos_state_space(sos, L)
    W = np.empty((D, L))
    v = B.copy()
    for i in range(L - 1, -1, -1):
        W[:, i] = v
        v = A @ v
    scan = [_lib.sos_state_space(sos, L*2**k)[2] for k in range(5)]
    fix = [_lib.sos_state_space(sos, L*j)[2] for j in range(32)]
    wpow = [_lib.sos_state_space(sos, L*GW*k)[2] for k in range(NW + 1)]
    Pt = [_lib.sos_state_space(sos, T*j)[2] for j in range(33)]
    ntt = (n + T - 1)//T
    xpad = np.zeros(ntt*T)
    xpad[:n] = x
    y = np.empty(ntt*T)
    s0 = np.zeros(D) if zi is None else np.asarray(zi, float).reshape(D)
    agg = np.zeros((ntt, D))
    incl = np.zeros((ntt, D))
    zf = None
    for tt in range(ntt):
        xt = xpad[tt*T:(tt + 1)*T].reshape(G, L)
        vv = xt @ W.T                                 # pass A: (G, D)
        # warp scan over gl inside each warp (NW warps x GW sub-chunks)
        vv = vv.reshape(NW, GW, D).copy()
        k = 0
        off = 1
        while off < GW:
            u = np.zeros_like(vv)
            u[:, off:] = vv[:, :-off]
            vv[:, off:] += u[:, off:] @ scan[k].T
            off *= 2
            k += 1
        ex = np.zeros_like(vv)
        ex[:, 1:] = vv[:, :-1]
        wagg = vv[:, GW - 1]                          # (NW, D)
        # every warp: its incoming state for a zero tile carry; warp 0: tile aggregate
        wpre = np.zeros((NW, D))
        for w in range(NW):
            for j in range(w):
                wpre[w] += wpow[w - 1 - j] @ wagg[j]
        acc = np.zeros(D)
        for j in range(NW):
            acc += wpow[NW - 1 - j] @ wagg[j]
        agg[tt] = acc
        # look-back: nearest inclusive at a random distance J (1..min(tt+1, 32))
        if tt == 0:
            sin = s0.copy()
        else:
            J = int(lookback_J(rng, tt))
            sin = np.zeros(D)
            for jj in range(1, J + 1):
                b = tt - jj
                if jj == J:
                    vec = s0 if b < 0 else incl[b]
                else:
                    vec = agg[b]
                sin += Pt[jj - 1] @ vec
        incl[tt] = Pt[1] @ sin + acc
        wcar = np.zeros((NW, D))
        for w in range(NW):
            wcar[w] = wpre[w] + wpow[w] @ sin
        for w in range(NW):
            for gl in range(GW):
                z = ex[w, gl] + fix[gl] @ wcar[w]
                g = w*GW + gl
                yy, zz = df2t_chunk(sos, xt[g], z)
                y[tt*T + g*L: tt*T + (g + 1)*L] = yy
                lo = tt*T + g*L
                if lo <= n - 1 < lo + L:
                    _, zf = df2t_chunk(sos, xt[g][:n - lo], z)
    return y[:n], zf


def lookback_J(rng, tt):
    return rng.integers(1, min(tt + 1, 32) + 1)


@pytest.mark.parametrize('CG', [1, 4, 8, 32])
@pytest.mark.parametrize('kind', ['bp2', 'hp2', 'lp4', 'bp4', 'slow'])
def test_scan_model_matches_sosfilt(CG, kind):
    rng = np.random.default_rng(42)
    fs = 48000.
    sos = {'bp2': butter(2, (1000., 15000.), 'bandpass', fs=fs, output='sos'),
           'hp2': butter(2, 1000., 'highpass', fs=fs, output='sos'),
           'lp4': butter(4, 6000., 'lowpass', fs=fs, output='sos'),
           'bp4': butter(4, (1000., 15000.), 'bandpass', fs=fs, output='sos'),
           # pole radius ~ 1 - 1e-4: no decay inside the look-back window
           'slow': butter(2, 1., 'lowpass', fs=fs, output='sos')}[kind]
    T = (NT//CG)*L
    n = 3*T + 77 if CG > 1 else 2*T + 77
    x = rng.standard_normal(n)
    zi = rng.standard_normal((sos.shape[0], 2))*0.1
    y, zf = run_model(sos, x, CG, rng, zi=zi)
    yref, zref = sosfilt(sos, x, zi=zi)
    scale = max(1.0, np.max(np.abs(yref)))
    # 'slow' has a near-double pole at z = 1: any two fp64 evaluation orders
    # (scipy's serial one included) differ by ~1e5 * eps there
    tol = 1e-7 if kind == 'slow' else 1e-10
    assert np.max(np.abs(y - yref)) <= tol*scale
    assert np.max(np.abs(zf.reshape(-1, 2) - zref)) <= tol*max(1.0, np.max(np.abs(zref)))


def test_state_space_matches_recurrence():
    sos = butter(4, (1000., 15000.), 'bandpass', fs=48000., output='sos')
    A, B, A5 = _lib.sos_state_space(sos, 5)
    D = A.shape[0]
    assert np.allclose(np.linalg.matrix_power(A, 5), A5, rtol=1e-12, atol=1e-15)
    # block lower triangular: section k only sees sections <= k
    for r in range(D):
        assert np.all(A[r, (r | 1) + 1:] == 0)
    rng = np.random.default_rng(0)
    z = rng.standard_normal(D)
    _, z1 = df2t_chunk(sos, np.array([0.7]), z)
    assert np.allclose(A @ z + B*0.7, z1, rtol=1e-13, atol=1e-15)


def test_sosfiltfilt_edge():
    from oracle.oracle import sosfiltfilt_edge
    for order, wn, kind in [(2, 500., 'lowpass'), (4, 500., 'lowpass'),
                            (2, (100., 500.), 'bandpass'), (3, 2000., 'highpass')]:
        sos = butter(order, wn, kind, fs=48000., output='sos')
        assert _lib.sosfiltfilt_edge(sos) == sosfiltfilt_edge(sos)


# ---------------------------------------------------------------- run kernel: run-in from zero state

def run_in_tiles(sos, T, tol=1e-20, look=32):
    """Plan::jpre of sosfilt.cu: tiles after which max|A^(T j)| < tol, or None."""
    for j in range(1, look + 1):
        if np.max(np.abs(_lib.sos_state_space(sos, T*j)[2])) < tol:
            return j
    return None


@pytest.mark.parametrize('design', [
    lambda fs: butter(2, (1000., 15000.), 'bandpass', fs=fs, output='sos'),
    lambda fs: butter(4, (1000., 15000.), 'bandpass', fs=fs, output='sos'),
    lambda fs: butter(2, 500., 'lowpass', fs=fs, output='sos'),
    lambda fs: butter(4, 500., 'lowpass', fs=fs, output='sos'),
    lambda fs: butter(2, 0.02*fs, 'highpass', fs=fs, output='sos'),
])
@pytest.mark.parametrize('CG', [1, 8])
def test_run_in_from_zero_state_is_exact_to_rounding(design, CG):
    """sos_run_kernel starts a run `jpre` tiles early from ZERO state instead of taking the true
    state from its predecessors: with the kernel's criterion (max|A^(T jpre)| < 1e-20) what it
    stores equals sosfilt over the whole trace to rounding, even with a DC offset that loads the
    states (worst case for a high-pass)."""
    fs = 48000.
    sos = design(fs)
    T = (NT//CG)*L
    jpre = run_in_tiles(sos, T)
    assert jpre is not None and jpre <= 8, 'the audio-rate designs of the bench take the run kernel'
    rng = np.random.default_rng(CG)
    run_tiles = 4*jpre
    n = (jpre + run_tiles + 3)*T
    x = rng.standard_normal(n)*0.1 + 0.7
    ref = sosfilt(sos, x)
    start = 3*T + jpre*T                    # first stored row of the run
    y, _ = df2t_chunk(sos, x[start - jpre*T:start + run_tiles*T], np.zeros(2*sos.shape[0]))
    got = y[jpre*T:]
    # two rounding trajectories of a narrow low-pass differ by ~1e-14 before they lock in; a
    # run-in that is too short would leave 1e-20^(fraction) of an O(1) state, orders above this
    assert np.max(np.abs(got - ref[start:start + run_tiles*T])) <= 1e-12*max(1.0, np.max(np.abs(ref)))


def test_slow_cascade_gets_no_run_in():
    """A 30-Hz high-pass at 96 kHz needs far more than the look-back window to forget: the host
    logic must keep the look-back kernel for it (Plan::jpre > SOS_LOOK)."""
    sos = butter(2, 0.0006*96000., 'highpass', fs=96000., output='sos')
    assert run_in_tiles(sos, (NT//8)*L) is None
