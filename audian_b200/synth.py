"""Deterministic synthetic recordings.

Integer-only and counter based, so the host generator here (numpy uint64)
and the device generator (`adn_synth_f64_dev`, csrc/synth.cu) produce
identical bytes for any (t, c) without keeping state: a gated triangular
carrier per channel ("cricket chirps", 20 pulses/s, 50 % duty) plus
broadband noise at about -30 dBFS, on the 2**-15 grid of int16 WAV data.

    noise  n(t,c) = int16(top 16 bits of splitmix64(seed ^ (t*C + c)))
    phase  p(t,c) = (t * inc_c) mod 2**32,  inc_c = round(2**32*(0.05 + 0.005*(c % 64)))
    tri16(p)      = q < 32768 ? 2q - 32767 : 98303 - 2q,   q = p >> 16
    gate   g(t)   = (t mod round(rate/20)) < round(rate/40)
    v = clip(((13107*tri16*g) >> 15) + (n >> 5), -32768, 32767);  x = v/32768
"""

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30)))*np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27)))*np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def carrier_increment(c):
    """Phase increment of channel c; carriers at (0.05 + 0.005*(c%64))*rate."""
    return int(round(2**32*(0.05 + 0.005*(c % 64)))) & 0xFFFFFFFF


def gate_periods(rate):
    period = max(2, int(round(rate/20.0)))
    on = max(1, int(round(rate/40.0)))
    return period, on


def synth(t0, nframes, channels, rate, seed=0xA0D1A9):
    """Rows t0 .. t0+nframes of the synthetic recording: (nframes, channels)
    float64, C-contiguous, channels interleaved."""
    with np.errstate(over='ignore'):
        t = (np.arange(nframes, dtype=np.uint64) + np.uint64(t0))[:, None]
        c = np.arange(channels, dtype=np.uint64)[None, :]
        key = np.uint64(seed) ^ (t*np.uint64(channels) + c)
        r = splitmix64(key)
        noise = (r >> np.uint64(48)).astype(np.uint16).view(np.int16).astype(np.int64)
        inc = np.array([carrier_increment(k) for k in range(channels)],
                       dtype=np.uint64)[None, :]
        phase = (t*inc) & np.uint64(0xFFFFFFFF)
        q = (phase >> np.uint64(16)).astype(np.int64)
        tri = np.where(q < 32768, 2*q - 32767, 98303 - 2*q)
        period, on = gate_periods(rate)
        g = ((t % np.uint64(period)) < np.uint64(on)).astype(np.int64)
        v = ((13107*tri*g) >> 15) + (noise >> 5)
        v = np.clip(v, -32768, 32767)
    return np.ascontiguousarray(v.astype(np.float64)/32768.0)
