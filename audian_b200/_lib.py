"""ctypes binding of libaudian_b200.so (C ABI: include/audian_b200.h).

Thin by design: argument checks that need numpy (dtype, contiguity), pointer
extraction, status -> exception.  There is no CPU implementation behind these
functions; if the shared library is missing or no CUDA device is present the
calls raise.
"""

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ADN_LIB: a differently built copy of the library (kernel experiments), default the in-tree one
LIB_PATH = os.environ.get('ADN_LIB') or os.path.join(_HERE, 'libaudian_b200.so')

ADN_OK = 0
ADN_ERR_INVALID = 1
ADN_ERR_CUDA = 2
ADN_ERR_UNSUPPORTED = 3
ADN_ERR_SHORT = 4
ADN_MAX_SECTIONS = 8
ADN_MIN_NFFT = 8
ADN_MAX_NFFT = 1 << 20
ADN_OPT_RESIDENT = 0
ADN_OPT_VERIFY = 1
ADN_OPT_CHUNK_BYTES = 2
ADN_OPT_RESIDENT_MIN_BYTES = 3
ADN_OPT_RESIDENT_CAP_BYTES = 4
ADN_OPT_ENVELOPE_CHUNK_BYTES = 5
ADN_OPT_SCAN_RUNS = 6
ADN_OPT_ZERO_PHASE_ONEPASS = 7
ADN_WINDOW_HANN = 0
ADN_DETREND_NONE = 0
ADN_DETREND_CONSTANT = 1

_dp = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f64 = C.c_double

class ChainSpec(C.Structure):
    """adn_chain_t of include/audian_b200.h."""
    _fields_ = [('sos', C.c_void_p), ('S', C.c_int32), ('out_db', C.c_int32), ('nbefore', C.c_int64),
                ('nfft', C.c_int32), ('hop', C.c_int32),
                ('spec_first', C.c_int64), ('spec_rows', C.c_int64), ('n_spec', C.c_int64),
                ('esos', C.c_void_p), ('ES', C.c_int32), ('clamp_negative', C.c_int32),
                ('env_first', C.c_int64), ('env_rows', C.c_int64), ('env_nbefore', C.c_int64),
                ('n_env', C.c_int64), ('mm_step', C.c_int64)]


# name: (restype, argtypes) -- every symbol include/audian_b200.h declares
SIGNATURES = {
    'adn_init': (_i32, [_i32]),
    'adn_shutdown': (_i32, []),
    'adn_last_error': (C.c_char_p, []),
    'adn_version': (_i32, []),
    'adn_launch_count': (_i64, []),
    'adn_scan_run_count': (_i64, []),
    'adn_fwd_park_count': (_i64, []),
    'adn_zero_phase_count': (_i64, []),
    'adn_synchronize': (_i32, []),
    'adn_host_register': (_i32, [_dp, _i64]),
    'adn_host_unregister': (_i32, [_dp]),
    'adn_host_alloc': (_i32, [_i64, C.POINTER(C.c_void_p)]),
    'adn_host_free': (_i32, [C.c_void_p]),
    'adn_set_option': (_i32, [_i32, _i64]),
    'adn_get_option': (_i64, [_i32]),
    'adn_mirror_create': (_i32, [C.POINTER(_i64)]),
    'adn_mirror_release': (_i32, [_i64]),
    'adn_mirror_invalidate': (_i32, [_i64]),
    'adn_invalidate': (_i32, [_dp, _i64]),
    'adn_resident_hits': (_i64, []),
    'adn_transfer_bytes': (_i32, [C.POINTER(_i64), C.POINTER(_i64)]),
    'adn_minmax_f64': (_i32, [_dp, _i64, _i32, _i64, _dp]),
    'adn_sosfilt_f64': (_i32, [_dp, _i32, _dp, _i64, _i32, _i64, _dp, _i64, _dp]),
    'adn_envelope_f64': (_i32, [_dp, _i32, _dp, _i64, _i32, _i64, _dp, _i64, _i32]),
    'adn_spectrogram_f64': (_i32, [_dp, _i64, _i32, _f64, _i32, _i32, _i32, _i32,
                                   _dp, _i64, _i32, C.POINTER(_i64)]),
    'adn_decibel_f64': (_i32, [_dp, _i64, _f64, _f64, _dp]),
    'adn_sosfiltfilt_f64': (_i32, [_dp, _i32, _dp, _i64, _i32, _dp, _i64]),
    'adn_minmax_f64_m': (_i32, [_dp, _i64, _i32, _i64, _dp, _i64]),
    'adn_chain_f64': (_i32, [C.POINTER(ChainSpec), _dp, _i64, _i32, _f64, _dp, _i64, _dp, _dp, _dp,
                             C.POINTER(_i64), _i64, _i64]),
    'adn_chain_f64_dev': (_i32, [C.POINTER(ChainSpec), _dp, _i64, _i32, _f64, _dp, _i64, _dp, _dp, _dp,
                                 C.POINTER(_i64), _dp]),
    'adn_sosfilt_f64_m': (_i32, [_dp, _i32, _dp, _i64, _i32, _i64, _dp, _i64, _dp, _i64, _i64]),
    'adn_envelope_f64_m': (_i32, [_dp, _i32, _dp, _i64, _i32, _i64, _dp, _i64, _i32, _i64, _i64]),
    'adn_spectrogram_f64_m': (_i32, [_dp, _i64, _i32, _f64, _i32, _i32, _i32, _i32,
                                     _dp, _i64, _i32, C.POINTER(_i64), _i64, _i64]),
    'adn_spec_image_db_f64_m': (_i32, [_dp, _i64, _i32, _i32, _i32, _dp, _i64]),
    'adn_minmax_channel_f64_m': (_i32, [_dp, _i64, _i32, _i32, _i64, _dp, _i64]),
    'adn_unwrap_f64': (_i32, [_dp, _i64, _i32, _f64, _i32]),
    'adn_unwrap_f64_dev': (_i32, [_dp, _i64, _i32, _f64, _i32, _dp, _dp]),
    'adn_play_region_f64_m': (_i32, [_dp, _i64, _i32, _dp, _i32, _dp, _i32, _f64, _f64, _dp, _i32, _i64,
                                     _dp, _i64]),
    'adn_play_region_f64_dev': (_i32, [_dp, _i64, _i32, _dp, _i32, _dp, _i32, _f64, _f64, _dp, _i32, _i64,
                                       _dp, _dp]),
    'adn_mean_power_db_f64_m': (_i32, [_dp, _i64, _i32, _i32, _i32, _i64, _i64, _f64, _dp, _i64]),
    'adn_sosfiltfilt_f64_dev': (_i32, [_dp, _i32, _dp, _i64, _i32, _dp, _i64, _dp]),
    'adn_spec_image_db_f64': (_i32, [_dp, _i64, _i32, _i32, _i32, _dp]),
    'adn_mean_power_db_f64': (_i32, [_dp, _i64, _i32, _i32, _i32, _i64, _i64, _f64, _dp]),
    'adn_pcm_to_f64': (_i32, [_dp, _i64, _i32, _f64, _dp]),
    'adn_spec_image_db_f64_dev': (_i32, [_dp, _i64, _i32, _i32, _i32, _dp, _dp]),
    'adn_mean_power_db_f64_dev': (_i32, [_dp, _i32, _i32, _i32, _i64, _i64, _f64, _dp, _dp]),
    'adn_colsum_f64_dev': (_i32, [_dp, _i64, _i64, _dp, _dp]),
    'adn_pcm_to_f64_dev': (_i32, [_dp, _i64, _i32, _f64, _dp, _dp]),
    'adn_minmax_f64_dev': (_i32, [_dp, _i64, _i32, _i64, _dp, _dp]),
    'adn_sosfilt_f64_dev': (_i32, [_dp, _i32, _dp, _i64, _i32, _i64, _dp, _i64,
                                   _dp, _dp, _dp]),
    'adn_sosfilt_minmax_f64_dev': (_i32, [_dp, _i32, _dp, _i64, _i32, _i64, _dp, _i64, _dp, _dp, _i64, _dp, _dp,
                                          _dp]),
    'adn_envelope_f64_dev': (_i32, [_dp, _i32, _dp, _i64, _i32, _i64, _dp, _i64,
                                    _i32, _dp]),
    'adn_zero_phase_range_f64_dev': (_i32, [_dp, _i32, _dp, _i64, _i32, _i32, _i32, _i32, _i64, _dp, _i64,
                                            _i32, _dp]),
    'adn_envelope_forward_f64_dev': (_i32, [_dp, _i32, _dp, _i64, _i32, _i32, _i32, _dp, _dp,
                                            _dp, _dp]),
    'adn_envelope_state0_f64_dev': (_i32, [_dp, _i32, _dp, _i32, _i32, _i32, _dp, _dp]),
    'adn_fold_states_f64_dev': (_i32, [_dp, _dp, _i32, _i32, _i32, _i32, _i32, _dp, _dp]),
    'adn_sosfilt_reverse_f64_dev': (_i32, [_dp, _i32, _dp, _i64, _i32, _dp, _dp, _i64, _i64,
                                           _i32, _dp, _dp]),
    'adn_spectrogram_f64_dev': (_i32, [_dp, _i64, _i32, _f64, _i32, _i32, _i32,
                                       _i32, _dp, _i64, _i32, C.POINTER(_i64), _dp]),
    'adn_decibel_f64_dev': (_i32, [_dp, _i64, _f64, _f64, _dp, _dp]),
    'adn_synth_f64_dev': (_i32, [_dp, _i64, _i64, _i32, _f64, C.c_uint64, _dp]),
    'adn_sos_state_space': (_i32, [_dp, _i32, _dp, _dp, _i64, _dp]),
    'adn_sos_decay_length': (_i64, [_dp, _i32, _f64]),
    'adn_sosfiltfilt_edge': (_i32, [_dp, _i32]),
}


class AdnError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f'libaudian_b200 error {code}: {message}')
        self.code = code


_lib = None


def lib():
    """The loaded shared library (loaded on first use; raises if absent)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} is missing: build it with `python -m audian_b200.build` '
                '(nvcc, sm_100a). audian_b200 has no CPU fallback.')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status):
    if status != ADN_OK:
        msg = lib().adn_last_error().decode('utf-8', 'replace')
        if status == ADN_ERR_SHORT:
            # scipy raises ValueError from sosfiltfilt for too short inputs
            raise ValueError(msg)
        raise AdnError(status, msg)


def _f64_array(a, name):
    if not isinstance(a, np.ndarray) or a.dtype != np.float64:
        raise TypeError(f'{name} must be a float64 ndarray')
    if not a.flags.c_contiguous:
        raise ValueError(f'{name} must be C-contiguous')
    return a


def ptr(a):
    return a.ctypes.data if a.size > 0 else None


def sos_array(sos):
    """(S, 6) float64 C-contiguous, or None -> (None, 0)."""
    if sos is None:
        return None, 0
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    if sos.ndim != 2 or sos.shape[1] != 6:
        raise ValueError('sos must have shape (n_sections, 6)')
    if sos.shape[0] > ADN_MAX_SECTIONS:
        raise AdnError(ADN_ERR_UNSUPPORTED,
                       f'{sos.shape[0]} sections (at most {ADN_MAX_SECTIONS})')
    if not np.all(sos[:, 3] == 1.0):
        sos = sos/sos[:, 3:4]
    return sos, sos.shape[0]


# ---------------------------------------------------------------- mirrors

class Mirror(object):
    """Device copy of (part of) one host buffer, handed explicitly from the call that
    fills the buffer (`dst_mirror=`) to the calls that read it (`src_mirror=`).  The
    owner invalidates it whenever the host buffer changes by other means."""

    def __init__(self):
        h = _i64(0)
        check(lib().adn_mirror_create(C.byref(h)))
        self.handle = h.value

    def invalidate(self):
        if self.handle:
            lib().adn_mirror_invalidate(self.handle)

    def release(self):
        if self.handle and _lib is not None:
            _lib.adn_mirror_release(self.handle)
        self.handle = 0

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def _h(mirror):
    return 0 if mirror is None else int(mirror.handle)


# ---------------------------------------------------------------- host arrays

def init(device=-1):
    check(lib().adn_init(device))


def launch_count():
    return int(lib().adn_launch_count())


def scan_run_count():
    """Launches of the SOS run kernel so far (the look-back kernel is the other scan kernel)."""
    return int(lib().adn_scan_run_count())


def fwd_park_count():
    """Launches of the pipelined forward kernel so far (csrc/sosfwd.cu)."""
    return int(lib().adn_fwd_park_count())


def zero_phase_count():
    """Launches of the one-pass zero-phase kernel so far (csrc/zerophase.cu)."""
    return int(lib().adn_zero_phase_count())


def minmax(src, step, dst=None, src_mirror=None):
    """dst (2*ceil(n/step), C): rows 2j / 2j+1 = min / max of source rows
    [j*step, (j+1)*step).  compresseddata.py:49-52,97-100; traceitem.py:58-61."""
    src = _f64_array(src, 'src')
    if src.ndim != 2:
        raise ValueError('src must be (frames, channels)')
    n, ch = src.shape
    nseg = (n + step - 1)//step if n > 0 else 0
    if dst is None:
        dst = np.empty((2*nseg, ch))
    dst = _f64_array(dst, 'dst')
    if dst.shape != (2*nseg, ch):
        raise ValueError(f'dst must have shape {(2*nseg, ch)}')
    check(lib().adn_minmax_f64_m(ptr(src), n, ch, int(step), ptr(dst), _h(src_mirror)))
    return dst


def minmax_channel(src, channel, step, src_mirror=None):
    """Interleaved min / max of `step` rows of one column of a (frames, channels) trace:
    what TraceItem.update_plot draws (traceitem.py:58-61).  Returns (2*ceil(n/step),)."""
    src = _f64_array(src, 'src')
    if src.ndim != 2:
        raise ValueError('src must be (frames, channels)')
    n, ch = src.shape
    nseg = (n + step - 1)//step if n > 0 else 0
    dst = np.empty(2*nseg)
    check(lib().adn_minmax_channel_f64_m(ptr(src), n, ch, int(channel), int(step), ptr(dst),
                                         _h(src_mirror)))
    return dst


def unwrap(data, thresh, clips=False):
    """audioio's unwrap(data, thresh, clips) in place on a (frames, channels) array
    (the loader option of Data.open, data.py:180)."""
    data = _f64_array(data, 'data')
    if data.ndim == 1:
        n, ch = data.shape[0], 1
    else:
        n, ch = data.shape
    check(lib().adn_unwrap_f64(ptr(data), n, ch, float(thresh), 1 if clips else 0))
    return data


def play_region(src, left, right, rate, het_freq=0.0, cutoff=20000.0, src_mirror=None):
    """(playdata, rate) as DataBrowser.play_region computes them before the fade
    (databrowser.py:1711-1728): means over the channel groups `left` / `right` (right may be
    empty: one column), heterodyne + butter(2, cutoff) sosfiltfilt + [::nstep] if het_freq > 0."""
    from scipy.signal import butter
    src = _f64_array(src, 'src')
    n, ch = src.shape
    left = np.ascontiguousarray(left, dtype=np.int32)
    right = np.ascontiguousarray(right, dtype=np.int32)
    ncols = 2 if right.size > 0 else 1
    sos, S, nstep = None, 0, 1
    if het_freq > 0:
        sos = np.ascontiguousarray(butter(2, cutoff, 'low', output='sos', fs=rate))
        S = sos.shape[0]
        nstep = max(1, int(np.round(rate/(2*cutoff))))
    out = np.empty(((n + nstep - 1)//nstep, ncols))
    check(lib().adn_play_region_f64_m(ptr(src), n, ch, left.ctypes.data, left.size,
                                      right.ctypes.data if right.size else None, right.size,
                                      float(rate), float(het_freq),
                                      None if sos is None else sos.ctypes.data, S, nstep,
                                      ptr(out), _h(src_mirror)))
    return out, rate/nstep


def chain_spec(sos, nbefore=0, nfft=0, hop=0, spec_first=0, spec_rows=0, n_spec=0, out_db=False,
               esos=None, env_first=0, env_rows=0, env_nbefore=0, n_env=0, clamp_negative=True,
               mm_step=0):
    """(ChainSpec, keep-alive references) for adn_chain_f64[_dev]."""
    sos, S = sos_array(sos)
    esos, ES = sos_array(esos)
    cs = ChainSpec(None if sos is None else sos.ctypes.data, S, 1 if out_db else 0, int(nbefore),
                   int(nfft), int(hop), int(spec_first), int(spec_rows), int(n_spec),
                   None if esos is None else esos.ctypes.data, ES, 1 if clamp_negative else 0,
                   int(env_first), int(env_rows), int(env_nbefore), int(n_env), int(mm_step))
    return cs, (sos, esos)


def chain(sos, src, filtered, rate, nbefore=0, spec=None, nfft=0, hop=0, spec_first=0, spec_rows=None,
          out_db=False, esos=None, env=None, env_first=0, env_rows=None, env_nbefore=0,
          clamp_negative=True, mm_step=0, minmax_out=None, src_mirror=None, filt_mirror=None):
    """data -> filtered -> {spectrogram, envelope} (+ min/max of the raw rows) in one call
    (bufferedfilter.py:53 / buffereddata.py:149-153: a parameter change recomputes filtered,
    then its dests).  Each stage equals sosfilt() / spectrogram() / envelope() / minmax() on
    the same arrays.  Returns the number of spectrogram frames computed."""
    src = _f64_array(src, 'src')
    filtered = _f64_array(filtered, 'filtered')
    n_filt = filtered.shape[0]
    if spec is not None:
        spec = _f64_array(spec, 'spec')
    if env is not None:
        env = _f64_array(env, 'env')
    if minmax_out is not None:
        minmax_out = _f64_array(minmax_out, 'minmax_out')
    if spec_rows is None:
        spec_rows = n_filt - spec_first
    if env_rows is None:
        env_rows = n_filt - env_first
    cs, keep = chain_spec(sos, nbefore, nfft, hop, spec_first, spec_rows,
                          0 if spec is None else spec.shape[0], out_db, esos, env_first, env_rows,
                          env_nbefore, 0 if env is None else env.shape[0], clamp_negative, mm_step)
    n = _i64(0)
    check(lib().adn_chain_f64(C.byref(cs), ptr(src), src.shape[0], src.shape[1], float(rate),
                              ptr(filtered), n_filt, None if spec is None else ptr(spec),
                              None if env is None else ptr(env),
                              None if minmax_out is None else ptr(minmax_out), C.byref(n),
                              _h(src_mirror), _h(filt_mirror)))
    del keep
    return n.value


def sosfilt(sos, src, dst, nbefore=0, zi=None, src_mirror=None, dst_mirror=None):
    """dst[i, c] = scipy.signal.sosfilt(sos, src[:, c])[nbefore + i].
    zi: None or (C, S, 2) float64, updated in place.  bufferedfilter.py:31-36."""
    src = _f64_array(src, 'src')
    dst = _f64_array(dst, 'dst')
    sos, S = sos_array(sos)
    if src.ndim != 2 or dst.ndim != 2 or src.shape[1] != dst.shape[1]:
        raise ValueError('src and dst must be (frames, channels) with equal channels')
    if zi is not None:
        zi = _f64_array(zi, 'zi')
        if zi.shape != (src.shape[1], S, 2):
            raise ValueError('zi must have shape (channels, sections, 2)')
    check(lib().adn_sosfilt_f64_m(None if sos is None else sos.ctypes.data, S,
                                  ptr(src), src.shape[0], src.shape[1], int(nbefore),
                                  ptr(dst), dst.shape[0],
                                  None if zi is None else zi.ctypes.data,
                                  _h(src_mirror), _h(dst_mirror)))
    return dst


def envelope(sos, src, dst, nbefore=0, clamp_negative=True, src_mirror=None, dst_mirror=None):
    """dst = sosfiltfilt(sos, (pi/2)|src|, axis=0)[nbefore:], negatives clamped.
    bufferedenvelope.py:34-41."""
    src = _f64_array(src, 'src')
    dst = _f64_array(dst, 'dst')
    sos, S = sos_array(sos)
    if src.ndim != 2 or dst.ndim != 2 or src.shape[1] != dst.shape[1]:
        raise ValueError('src and dst must be (frames, channels) with equal channels')
    check(lib().adn_envelope_f64_m(None if sos is None else sos.ctypes.data, S,
                                   ptr(src), src.shape[0], src.shape[1], int(nbefore),
                                   ptr(dst), dst.shape[0], 1 if clamp_negative else 0,
                                   _h(src_mirror), _h(dst_mirror)))
    return dst


def sosfiltfilt(sos, src, dst=None):
    """scipy.signal.sosfiltfilt(sos, src, axis=0) (databrowser.py:1725, play-back path)."""
    src = _f64_array(src, 'src')
    sos, S = sos_array(sos)
    if dst is None:
        dst = np.empty_like(src)
    dst = _f64_array(dst, 'dst')
    check(lib().adn_sosfiltfilt_f64(sos.ctypes.data, S, ptr(src), src.shape[0], src.shape[1],
                                    ptr(dst), dst.shape[0]))
    return dst


def spectrogram(src, rate, nfft, hop, dst, window=ADN_WINDOW_HANN,
                detrend=ADN_DETREND_CONSTANT, out_db=False, src_mirror=None, dst_mirror=None):
    """Fills dst (n_dst, C, nfft//2+1) like BufferedSpectrogram.process
    (bufferedspectrogram.py:45-62); returns the number of computed frames."""
    src = _f64_array(src, 'src')
    dst = _f64_array(dst, 'dst')
    if src.ndim != 2 or dst.ndim != 3 or dst.shape[1] != src.shape[1] or \
       dst.shape[2] != nfft//2 + 1:
        raise ValueError('src must be (frames, C) and dst (n, C, nfft//2+1)')
    n = _i64(0)
    check(lib().adn_spectrogram_f64_m(ptr(src), src.shape[0], src.shape[1], float(rate),
                                      int(nfft), int(hop), window, detrend, ptr(dst),
                                      dst.shape[0], 1 if out_db else 0, C.byref(n),
                                      _h(src_mirror), _h(dst_mirror)))
    return n.value


def decibel(power, ref_power=1.0, min_power=1e-20):
    """thunderlab.powerspectrum.decibel on the GPU (specitem.py:36)."""
    power = np.ascontiguousarray(power, dtype=np.float64)
    out = np.empty_like(power)
    check(lib().adn_decibel_f64(ptr(power), power.size, float(ref_power),
                                float(min_power), ptr(out)))
    return out


def spec_image_db(spec, channel, src_mirror=None):
    """decibel(spec[:, channel, :].T) as a new (F, n) array (specitem.py:33-39)."""
    spec = _f64_array(spec, 'spec')
    if spec.ndim != 3:
        raise ValueError('spec must be (frames, channels, bins)')
    n, ch, F = spec.shape
    out = np.empty((F, n))
    check(lib().adn_spec_image_db_f64_m(ptr(spec), n, ch, F, int(channel), ptr(out), _h(src_mirror)))
    return out


def mean_power_db(spec, channel, i0, i1, floor_db=-200.0, src_mirror=None):
    """max(decibel(mean(spec[i0:i1, channel, :], axis=0)), floor_db) (spectrogramplot.py:158-160)."""
    spec = _f64_array(spec, 'spec')
    n, ch, F = spec.shape
    out = np.empty(F)
    check(lib().adn_mean_power_db_f64_m(ptr(spec), n, ch, F, int(channel), int(i0), int(i1),
                                        float(floor_db), ptr(out), _h(src_mirror)))
    return out


def pcm_to_f64(pcm, bits, channels, gain=1.0, dst=None):
    """(frames, channels) float64 = PCM / 2**(bits-1) * gain from a bytes-like /
    integer array of interleaved little-endian samples (int16, packed int24, int32)."""
    raw = np.ascontiguousarray(pcm).view(np.uint8).reshape(-1)
    n = raw.size//(bits//8)
    if dst is None:
        dst = np.empty((n//channels, channels))
    dst = _f64_array(dst, 'dst')
    if dst.size != n:
        raise ValueError('dst does not match the number of samples')
    check(lib().adn_pcm_to_f64(raw.ctypes.data if n else None, n, int(bits), float(gain), ptr(dst)))
    return dst


def sos_state_space(sos, power=1):
    """Host-side plan inspection: (A, B, A**power) of the cascade."""
    sos, S = sos_array(sos)
    D = 2*S
    A = np.empty((D, D))
    B = np.empty(D)
    P = np.empty((D, D))
    check(lib().adn_sos_state_space(sos.ctypes.data, S, A.ctypes.data, B.ctypes.data,
                                    int(power), P.ctypes.data))
    return A, B, P


def sos_decay_length(sos, tol=1e-30):
    """Samples after which the cascade has forgotten its state to within tol
    (a power of two), or -1."""
    sos, S = sos_array(sos)
    return int(lib().adn_sos_decay_length(sos.ctypes.data, S, float(tol)))


def sosfiltfilt_edge(sos):
    sos, S = sos_array(sos)
    return int(lib().adn_sosfiltfilt_edge(sos.ctypes.data, S))


def set_option(option, value):
    check(lib().adn_set_option(int(option), int(value)))


def get_option(option):
    return int(lib().adn_get_option(int(option)))


def enable_resident(on=True):
    """Whether mirrors are honoured at all (default on).  Needs no GPU: only flips an option."""
    lib().adn_set_option(ADN_OPT_RESIDENT, 1 if on else 0)


def invalidate(a):
    """Invalidate every mirror that overlaps the host array `a` (it was changed by
    other means than a call into the library)."""
    if a is not None and a.size > 0:
        lib().adn_invalidate(a.ctypes.data, a.nbytes)


def resident_hits():
    return int(lib().adn_resident_hits())


def transfer_bytes():
    """(host->device, device->host) bytes copied by the host-array calls so far."""
    a, b = _i64(0), _i64(0)
    lib().adn_transfer_bytes(C.byref(a), C.byref(b))
    return a.value, b.value


def host_register(a):
    check(lib().adn_host_register(a.ctypes.data, a.nbytes))


def host_unregister(a):
    check(lib().adn_host_unregister(a.ctypes.data))


_PINNED_POOL = {}                   # nbytes -> [pointers]: blocks of released arrays, reused by size
_PINNED_POOL_LIMIT = 4 << 30        # bytes kept for reuse (page-locking a few hundred MB takes ~0.1 s)
_pinned_pooled = [0]


class _PinnedBlock(object):
    """Owner of one adn_host_alloc() allocation; released with the last array that views it
    (back into a small pool keyed by size: a scrolling trace asks for the same size again)."""

    def __init__(self, nbytes):
        self.nbytes = nbytes
        free = _PINNED_POOL.get(nbytes)
        if free:
            self.ptr = free.pop()
            _pinned_pooled[0] -= nbytes
            return
        p = C.c_void_p()
        check(lib().adn_host_alloc(nbytes, C.byref(p)))
        self.ptr = p.value

    def __del__(self):
        try:
            if self.ptr:
                if _pinned_pooled[0] + self.nbytes <= _PINNED_POOL_LIMIT:
                    _PINNED_POOL.setdefault(self.nbytes, []).append(self.ptr)
                    _pinned_pooled[0] += self.nbytes
                else:
                    lib().adn_host_free(C.c_void_p(self.ptr))
        except Exception:                               # interpreter shutdown
            pass
        self.ptr = None


def pinned_pool_clear():
    """Gives the pooled page-locked blocks back to the driver."""
    for free in _PINNED_POOL.values():
        while free:
            lib().adn_host_free(C.c_void_p(free.pop()))
    _pinned_pooled[0] = 0


def pinned_empty(shape, dtype=np.float64):
    """np.empty(shape, dtype) in page-locked memory from the driver's allocator (adn_host_alloc):
    copies to and from such arrays run at full PCIe rate and overlap with the kernels.  Raises
    RuntimeError without a CUDA device (callers that may run without one catch it and use np.empty)."""
    shape = (shape,) if np.isscalar(shape) else tuple(int(v) for v in shape)
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64))*dt.itemsize
    if nbytes == 0:
        return np.empty(shape, dt)
    blk = _PinnedBlock(nbytes)
    buf = (C.c_char*nbytes).from_address(blk.ptr)
    buf._adn_block = blk                                # the array's base keeps the block alive
    return np.frombuffer(buf, dtype=dt).reshape(shape)


def buffer_empty(shape, dtype=np.float64):
    """The allocation behind a trace's `buffer`: page-locked when a CUDA device is there, plain
    np.empty otherwise (host-only tests of the index algebra; computing needs the device anyway)."""
    global _pinned_ok
    if _pinned_ok is not False:
        try:
            a = pinned_empty(shape, dtype)
            _pinned_ok = True
            return a
        except (RuntimeError, OSError):
            if not _pinned_ok:                          # never worked: no device, stop trying
                _pinned_ok = False
            # worked before: this size cannot be page-locked right now; pageable memory works too
    return np.empty(shape, dtype)


_pinned_ok = None


def is_pinned_array(a):
    """True for arrays made by pinned_empty() (and views of them)."""
    b = a
    while isinstance(b, np.ndarray) and b.base is not None:
        b = b.base
    while b is not None and not hasattr(b, '_adn_block'):
        b = getattr(b, 'obj', None) if isinstance(b, memoryview) else None
    return b is not None
