"""Host-side model of the FFT arithmetic in audian_b200/csrc/spectrogram.cu:
half-size complex FFT of the real frame, radix-2^2 decimation in frequency (two
stages per pass, optional last radix-2 stage), bit-reversed read-out, split
step, one-sided density scaling.  Checked against scipy.signal.spectrogram."""

import numpy as np
import pytest
from scipy.signal import spectrogram


def bitrev(k, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (k & 1)
        k >>= 1
    return r


def model_frame(x, rate):
    N = len(x)
    M = N//2
    logM = M.bit_length() - 1
    j = np.arange(N)
    w = 0.5 - 0.5*np.cos(2*np.pi*j/N)
    tw = np.exp(-2j*np.pi*np.arange(N//2)/N)
    xw = (x - x.sum()/N)*w
    z = xw[0::2] + 1j*xw[1::2]
    lg = logM
    while lg >= 2:
        n = 1 << lg
        q4 = n >> 2
        tstep = N >> lg
        for r in range(M >> 2):
            blk, k = r >> (lg - 2), r & (q4 - 1)
            b = (blk << lg) + k
            a0, a1, a2, a3 = z[b], z[b + q4], z[b + 2*q4], z[b + 3*q4]
            w1, w2 = tw[k*tstep], tw[2*k*tstep]
            b0 = a0 + a2
            b2 = (a0 - a2)*w1
            b1 = a1 + a3
            d = a1 - a3
            b3 = complex(d.imag, -d.real)*w1
            z[b] = b0 + b1
            z[b + q4] = (b0 - b1)*w2
            z[b + 2*q4] = b2 + b3
            z[b + 3*q4] = (b2 - b3)*w2
        lg -= 2
    if lg == 1:
        for r in range(M >> 1):
            u, v = z[2*r], z[2*r + 1]
            z[2*r], z[2*r + 1] = u + v, u - v
    scale = 1.0/(rate*np.sum(w*w))
    P = np.empty(M + 1)
    for k in range(M + 1):
        if k == 0 or k == M:
            xr = z[0].real + z[0].imag if k == 0 else z[0].real - z[0].imag
            P[k] = xr*xr*scale
        else:
            zk, zm = z[bitrev(k, logM)], z[bitrev(M - k, logM)]
            er, ei = 0.5*(zk.real + zm.real), 0.5*(zk.imag - zm.imag)
            orr, oi = 0.5*(zk.imag + zm.imag), -0.5*(zk.real - zm.real)
            wv = tw[k]
            xr = er + (orr*wv.real - oi*wv.imag)
            xi = ei + (orr*wv.imag + oi*wv.real)
            P[k] = (xr*xr + xi*xi)*scale*2.0
    return P


@pytest.mark.parametrize('N', [8, 16, 32, 64, 256, 1024])
def test_fft_model(N):
    rng = np.random.default_rng(N)
    x = rng.standard_normal(N) + 0.3
    f, t, S = spectrogram(x, fs=1000., window='hann', nperseg=N, noverlap=0,
                          detrend='constant', scaling='density', mode='psd')
    P = model_frame(x, 1000.)
    assert np.allclose(P, S[:, 0], rtol=1e-9, atol=1e-22*np.max(S))


# ---- the exchange-buffer layouts of the ring kernel's second pass (spectrogram.cu, T < 32)

def _groups(addrs):
    """16-byte bank groups (a 128-byte wavefront has eight) of complex-element addresses."""
    return [a % 8 for a in addrs]


def test_exchange_rows_of_nfft_256_are_conflict_free():
    """T = 8 (nfft 256): rows of 8 padded to 9, lane t owns rows t and t + 8.  A 128-bit access is
    served a quarter warp (= the 8 lanes of one frame) at a time: the eight lanes must hit eight
    different bank groups in the first-pass stores and in the second-pass loads.  The layout it
    replaced (rows 2 t + r at stride 17, results written back in natural order) had two-way
    conflicts in the natural-order stores."""
    for k1 in range(16):                                  # first pass: element (k1, t) of the frame
        assert sorted(_groups([k1*9 + t for t in range(8)])) == list(range(8))
    for i in range(16):                                   # second pass: a[i] of lane t
        assert sorted(_groups([(t + 8*(i >> 3))*9 + (i & 7) for t in range(8)])) == list(range(8))
    for r in range(2):                                    # the old natural-order stores, for the record
        for k2 in range(8):
            old = _groups([2*t + r + 16*k2 for t in range(8)])
            assert len(set(old)) == 4


def test_exchange_rows_of_nfft_128_are_conflict_free():
    """T = 4 (nfft 128): rows of 4 rotated by the row index, lane t owns rows t + 4 r; a quarter warp
    holds the lanes of two frames whose buffers sit FS = 68 (== 4 mod 8) elements apart."""
    FS = 68
    for k1 in range(16):
        a = [s*FS + k1*4 + ((t + k1) & 3) for s in range(2) for t in range(4)]
        assert sorted(_groups(a)) == list(range(8))
    for i in range(16):
        a = [s*FS + (t + 4*(i >> 2))*4 + ((t + i) & 3) for s in range(2) for t in range(4)]
        assert sorted(_groups(a)) == list(range(8))
    # every element of a frame has exactly one place
    places = sorted(k1*4 + ((t + k1) & 3) for k1 in range(16) for t in range(4))
    assert places == list(range(64))
    # what lane t loads as a[4 r + j] is element j of row t + 4 r
    for t in range(4):
        for i in range(16):
            r, j = i >> 2, i & 3
            row = t + 4*r
            assert (t + 4*(i >> 2))*4 + ((t + i) & 3) == row*4 + ((j + row) & 3)


def test_register_split_step_partners():
    """Split step in registers (all ring sizes): lane t of a frame holds Z[t + T m], m < 16, in register
    RG(m); the partner Z[M - k] of k = t + T m (m < 8) sits in lane (T - t) % T at m' = 15 - m, for
    lane 0 in its own register m' = 16 - m; the twiddle W_N^(t + T m) is W_N^t W_32^m."""
    def RG(T, m):
        return m if T >= 16 else (8*(m & 1) + (m >> 1) if T == 8 else 4*(m & 3) + (m >> 2))
    for T in (4, 8, 16, 32):
        M, N = 16*T, 32*T
        Q = max(1, 16//T)                              # rows per lane in the second pass
        # register i of lane t = output k2 of the DFT of row t + T r:  Z[(t + T r) + 16 k2]
        for t in range(T):
            held = {}
            for i in range(16):
                r, k2 = (i // (16//Q), i % (16//Q)) if T < 32 else (0, i)
                k = (t + T*r) + 16*k2 if T < 32 else t + 32*i
                held[i] = k
            assert sorted(held.values()) == [t + T*m for m in range(16)]
            for m in range(16):
                assert held[RG(T, m)] == t + T*m
        for t in range(T):
            for m in range(8):
                k = t + T*m
                if t > 0:
                    assert ((T - t) % T) + T*(15 - m) == M - k
                elif m > 0:
                    assert T*(16 - m) == M - k
                w = np.exp(-2j*np.pi*k/N)
                assert abs(w - np.exp(-2j*np.pi*t/N)*np.exp(-2j*np.pi*m/32)) < 1e-15
        assert T*8 == M//2                              # the self-paired bin is lane 0, m = 8
