"""Device-pointer operators: the *_dev entry points of the C ABI applied to
torch CUDA tensors (torch is used for device memory and streams only).

All functions enqueue work on torch's current stream and return without
synchronising.  Traces are (frames, C) float64 contiguous CUDA tensors,
spectrograms (frames, C, nfft//2+1).
"""

import ctypes as C

import numpy as np

from . import _lib


def _torch():
    import torch
    return torch


def _stream():
    torch = _torch()
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_trace(t, name):
    torch = _torch()
    if not t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous():
        raise TypeError(f'{name} must be a contiguous float64 CUDA tensor')
    return t


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else None


class CudaOps(object):
    """The compute back end of audian_b200.sharded: sm_100a kernels."""

    name = 'cuda'

    def empty(self, shape, like=None):
        torch = _torch()
        return torch.empty(shape, dtype=torch.float64, device='cuda')

    def zeros(self, shape, like=None):
        torch = _torch()
        return torch.zeros(shape, dtype=torch.float64, device='cuda')

    def minmax(self, src, step):
        return minmax(src, step)

    def sosfilt(self, sos, src, nbefore=0, zi=None, want_zf=False, out=None,
                state_only=False, zf_out=None):
        return sosfilt(sos, src, nbefore, zi, want_zf, out, state_only, zf_out)

    def spectrogram(self, src, rate, nfft, hop, n_dst, out_db=False, out=None):
        return spectrogram(src, rate, nfft, hop, n_dst, out_db, out)

    def sosfilt_minmax(self, sos, src, step, zi=None, want_raw=True, want_filt=True, out=None):
        return sosfilt_minmax(sos, src, step, zi, want_raw, want_filt, out)

    def envelope(self, sos, src, nbefore=0, clamp_negative=True):
        return envelope(sos, src, nbefore, clamp_negative)

    def zero_phase_range(self, sos, src, first, n_dst, edge_left=False, edge_right=False,
                         rectify=True, clamp_negative=True, out=None):
        return zero_phase_range(sos, src, first, n_dst, edge_left, edge_right, rectify,
                                clamp_negative, out)

    def env_forward(self, sos, src, edge_left=0, edge_right=0, zi=None, state_only=False,
                    zf_out=None):
        return env_forward(sos, src, edge_left, edge_right, zi, state_only, zf_out)

    def sosfilt_rev(self, sos, src, zi=None, first=0, n_dst=None, clamp_negative=False,
                    state_only=False, zf_out=None):
        return sosfilt_rev(sos, src, zi, first, n_dst, clamp_negative, state_only, zf_out)

    def env_state0(self, sos, src, edge, which, out):
        return env_state0(sos, src, edge, which, out)

    def fold_states(self, packs, mats, rank, backward=False):
        return fold_states(packs, mats, rank, backward)


def minmax(src, step, out=None):
    _check_trace(src, 'src')
    n, ch = src.shape
    nseg = (n + step - 1)//step if n > 0 else 0
    if out is None:
        out = _torch().empty((2*nseg, ch), dtype=src.dtype, device=src.device)
    _lib.check(_lib.lib().adn_minmax_f64_dev(_p(src), n, ch, int(step), _p(out), _stream()))
    return out


def sosfilt(sos, src, nbefore=0, zi=None, want_zf=False, out=None, state_only=False,
            zf_out=None):
    """Returns out, or (out, zf) if want_zf; state_only skips the output.
    zf_out: (C, S, 2)-sized contiguous tensor that receives the final state."""
    torch = _torch()
    _check_trace(src, 'src')
    sos, S = _lib.sos_array(sos)
    n, ch = src.shape
    zf = zf_out
    if zf is None and (want_zf or state_only) and S > 0:
        zf = torch.empty((ch, S, 2), dtype=src.dtype, device=src.device)
    if zi is not None:
        _check_trace(zi, 'zi')
    if state_only:
        out = None
        n_dst = 0
    else:
        if out is None:
            out = torch.empty((n - nbefore, ch), dtype=src.dtype, device=src.device)
        n_dst = out.shape[0]
    _lib.check(_lib.lib().adn_sosfilt_f64_dev(
        None if sos is None else sos.ctypes.data, S, _p(src), n, ch, int(nbefore),
        _p(out), n_dst, _p(zi), _p(zf), _stream()))
    if state_only:
        return zf
    return (out, zf) if want_zf else out


def sosfilt_minmax(sos, src, step, zi=None, want_raw=True, want_filt=True, out=None):
    """sosfilt of src (state zi carried) plus the full-trace min/max rows of the raw and of the
    filtered rows in the same pass (compresseddata.py:49-52; BASELINE config 4).
    Returns (filtered, zf, rows_raw or None, rows_filt or None)."""
    torch = _torch()
    _check_trace(src, 'src')
    sos, S = _lib.sos_array(sos)
    n, ch = src.shape
    nseg = (n + step - 1)//step
    if out is None:
        out = torch.empty((n, ch), dtype=src.dtype, device=src.device)
    zf = torch.empty((ch, S, 2), dtype=src.dtype, device=src.device) if S > 0 else None
    rr = torch.empty((2*nseg, ch), dtype=src.dtype, device=src.device) if want_raw else None
    rf = torch.empty((2*nseg, ch), dtype=src.dtype, device=src.device) if want_filt else None
    if zi is not None:
        _check_trace(zi, 'zi')
    _lib.check(_lib.lib().adn_sosfilt_minmax_f64_dev(
        None if sos is None else sos.ctypes.data, S, _p(src), n, ch, 0, _p(out), n, _p(zi), _p(zf),
        int(step), _p(rr), _p(rf), _stream()))
    return out, zf, rr, rf


def envelope(sos, src, nbefore=0, clamp_negative=True, out=None):
    torch = _torch()
    _check_trace(src, 'src')
    sos, S = _lib.sos_array(sos)
    n, ch = src.shape
    if out is None:
        out = torch.empty((n - nbefore, ch), dtype=src.dtype, device=src.device)
    _lib.check(_lib.lib().adn_envelope_f64_dev(
        None if sos is None else sos.ctypes.data, S, _p(src), n, ch, int(nbefore),
        _p(out), out.shape[0], 1 if clamp_negative else 0, _stream()))
    return out


def zero_phase_range(sos, src, first, n_dst, edge_left=False, edge_right=False, rectify=True,
                     clamp_negative=True, out=None):
    """sosfiltfilt (rectify: of (pi/2)|src|) over the rows `src` of a longer recording, rows
    first..first+n_dst of the result; scipy's edge handling only at the ends flagged as ends of
    the recording, zero state at the others (the caller brings halo rows)."""
    torch = _torch()
    _check_trace(src, 'src')
    sos, S = _lib.sos_array(sos)
    n, ch = src.shape
    if out is None:
        out = torch.empty((n_dst, ch), dtype=src.dtype, device=src.device)
    _lib.check(_lib.lib().adn_zero_phase_range_f64_dev(
        sos.ctypes.data, S, _p(src), n, ch, 1 if rectify else 0, 1 if edge_left else 0,
        1 if edge_right else 0, int(first), _p(out), int(n_dst), 1 if clamp_negative else 0, _stream()))
    return out


def chain(sos, src, filtered, rate, nbefore=0, spec=None, nfft=0, hop=0, spec_first=0, spec_rows=None,
          esos=None, env=None, env_first=0, env_rows=None, env_nbefore=0, clamp_negative=True,
          mm_step=0, minmax_out=None, out_db=False):
    """adn_chain_f64_dev on torch tensors: filtered (and spec / env / minmax_out if given) are
    filled; returns the number of spectrogram frames computed."""
    _check_trace(src, 'src')
    _check_trace(filtered, 'filtered')
    n_filt = filtered.shape[0]
    if spec_rows is None:
        spec_rows = n_filt - spec_first
    if env_rows is None:
        env_rows = n_filt - env_first
    cs, keep = _lib.chain_spec(sos, nbefore, nfft, hop, spec_first, spec_rows,
                               0 if spec is None else spec.shape[0], out_db, esos, env_first, env_rows,
                               env_nbefore, 0 if env is None else env.shape[0], clamp_negative, mm_step)
    n = _lib._i64(0)
    _lib.check(_lib.lib().adn_chain_f64_dev(C.byref(cs), _p(src), src.shape[0], src.shape[1], float(rate),
                                            _p(filtered), n_filt, _p(spec), _p(env), _p(minmax_out),
                                            C.byref(n), _stream()))
    del keep
    return n.value


def env_forward(sos, src, edge_left=0, edge_right=0, zi=None, state_only=False, zf_out=None):
    """Forward sweep of the envelope over one time shard: sosfilt of (pi/2)|src|
    with scipy's odd extension at the ends that are ends of the recording.
    Returns (y1, zf), y1 (edge_left + n + edge_right, C) or None if state_only."""
    torch = _torch()
    _check_trace(src, 'src')
    sos, S = _lib.sos_array(sos)
    n, ch = src.shape
    zf = zf_out if zf_out is not None else torch.empty((ch, S, 2), dtype=src.dtype, device=src.device)
    out = None
    if not state_only:
        out = torch.empty((n + edge_left + edge_right, ch), dtype=src.dtype, device=src.device)
    if zi is not None:
        _check_trace(zi, 'zi')
    _lib.check(_lib.lib().adn_envelope_forward_f64_dev(
        sos.ctypes.data, S, _p(src), n, ch, int(edge_left), int(edge_right), _p(zi), _p(out),
        _p(zf), _stream()))
    return out, zf


def sosfilt_rev(sos, src, zi=None, first=0, n_dst=None, clamp_negative=False, state_only=False,
                zf_out=None):
    """sosfilt over the rows of src in reversed order from state zi.
    Returns (rows first..first+n_dst of the result or None, zf)."""
    torch = _torch()
    _check_trace(src, 'src')
    sos, S = _lib.sos_array(sos)
    n, ch = src.shape
    zf = zf_out if zf_out is not None else torch.empty((ch, S, 2), dtype=src.dtype, device=src.device)
    out = None
    if not state_only:
        if n_dst is None:
            n_dst = n - first
        out = torch.empty((n_dst, ch), dtype=src.dtype, device=src.device)
    if zi is not None:
        _check_trace(zi, 'zi')
    _lib.check(_lib.lib().adn_sosfilt_reverse_f64_dev(
        sos.ctypes.data, S, _p(src), n, ch, _p(zi), _p(out), int(first),
        0 if out is None else out.shape[0], 1 if clamp_negative else 0, _p(zf), _stream()))
    return out, zf


def env_state0(sos, src, edge, which, out):
    """out (C, S, 2)-sized = sosfilt_zi(sos) * x0: which 0: x0 = 2 r[0] - r[edge] of
    the rows at src, r = (pi/2)|src|; which 1: x0 = src[0]."""
    sos, S = _lib.sos_array(sos)
    _lib.check(_lib.lib().adn_envelope_state0_f64_dev(
        sos.ctypes.data, S, _p(src), src.shape[-1], int(edge), int(which), _p(out), _stream()))
    return out


def fold_states(packs, mats, rank, backward=False):
    """State entering shard `rank`: packs (W, 2, C, D), mats (W, D, D) -> (C, D)."""
    torch = _torch()
    W, _, ch, D = packs.shape
    out = torch.empty((ch, D), dtype=packs.dtype, device=packs.device)
    _lib.check(_lib.lib().adn_fold_states_f64_dev(_p(packs), _p(mats), W, ch, D, int(rank),
                                                  1 if backward else 0, _p(out), _stream()))
    return out


def spectrogram(src, rate, nfft, hop, n_dst, out_db=False, out=None):
    """Returns (dst, n_computed)."""
    torch = _torch()
    _check_trace(src, 'src')
    n, ch = src.shape
    if out is None:
        out = torch.empty((n_dst, ch, nfft//2 + 1), dtype=src.dtype, device=src.device)
    ncomp = _lib._i64(0)
    _lib.check(_lib.lib().adn_spectrogram_f64_dev(
        _p(src), n, ch, float(rate), int(nfft), int(hop), _lib.ADN_WINDOW_HANN,
        _lib.ADN_DETREND_CONSTANT, _p(out), out.shape[0], 1 if out_db else 0,
        C.byref(ncomp), _stream()))
    return out, ncomp.value


def colsum(spec, acc):
    """acc (C, F) += sum over the frames of spec (n, C, F); returns acc."""
    n = spec.shape[0]
    _lib.check(_lib.lib().adn_colsum_f64_dev(_p(spec), n, spec.numel()//max(n, 1), _p(acc), _stream()))
    return acc


def decibel(power, ref_power=1.0, min_power=1e-20, out=None):
    torch = _torch()
    if out is None:
        out = torch.empty_like(power)
    _lib.check(_lib.lib().adn_decibel_f64_dev(_p(power), power.numel(), float(ref_power),
                                              float(min_power), _p(out), _stream()))
    return out


def synth(t0, nframes, channels, rate, seed=0xA0D1A9, out=None):
    """Rows t0..t0+nframes of the synthetic recording, generated on the device
    (bit-identical to audian_b200.synth.synth)."""
    torch = _torch()
    if out is None:
        out = torch.empty((nframes, channels), dtype=torch.float64, device='cuda')
    _lib.check(_lib.lib().adn_synth_f64_dev(_p(out), int(t0), int(nframes), int(channels),
                                            float(rate), C.c_uint64(seed), _stream()))
    return out
