"""The C-ABI library loads on a machine without a GPU and exports every symbol
include/audian_b200.h declares; compute entry points fail loudly (no CPU path)."""

import ctypes
import os
import re

import numpy as np
import pytest

from audian_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'audian_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(adn_[a-z0-9_]+)\s*\(', text)))


def test_every_declared_symbol_is_exported_and_bound():
    syms = header_symbols()
    assert len(syms) >= 20
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), s
    assert sorted(_lib.SIGNATURES) == syms


def test_version_and_error_string():
    lib = _lib.lib()
    assert lib.adn_version() >= 100
    assert isinstance(lib.adn_last_error(), bytes)


def test_argument_errors_without_touching_the_gpu():
    x = np.zeros((10, 2))
    with pytest.raises(_lib.AdnError) as e:
        _lib.check(_lib.lib().adn_minmax_f64(x.ctypes.data, 10, 2, 0, x.ctypes.data))
    assert e.value.code == _lib.ADN_ERR_INVALID
    with pytest.raises(_lib.AdnError):
        _lib.check(_lib.lib().adn_sosfilt_f64(None, 9, x.ctypes.data, 10, 2, 0, x.ctypes.data, 10, None))
    with pytest.raises(TypeError):
        _lib.minmax(np.zeros((4, 2), dtype=np.float32), 2)
    with pytest.raises(ValueError):
        _lib.minmax(np.zeros((4, 4))[:, ::2], 2)


def test_no_cpu_fallback():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip('a GPU is present')
    except ImportError:
        pass
    x = np.zeros((100, 2))
    with pytest.raises(_lib.AdnError) as e:
        _lib.minmax(x, 10)
    assert e.value.code == _lib.ADN_ERR_CUDA


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'audian_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f


def test_options_and_invalidate_need_no_gpu():
    assert _lib.get_option(_lib.ADN_OPT_RESIDENT) == 1
    assert _lib.get_option(_lib.ADN_OPT_ZERO_PHASE_ONEPASS) == 1
    m = _lib.Mirror()                            # mirrors are host-side bookkeeping until used
    assert m.handle > 0
    m.invalidate()
    m.release()
    old = _lib.get_option(_lib.ADN_OPT_CHUNK_BYTES)
    _lib.set_option(_lib.ADN_OPT_CHUNK_BYTES, 12345)
    assert _lib.get_option(_lib.ADN_OPT_CHUNK_BYTES) == 12345
    _lib.set_option(_lib.ADN_OPT_CHUNK_BYTES, old)
    with pytest.raises(_lib.AdnError):
        _lib.set_option(99, 1)
    _lib.invalidate(np.zeros((10, 2)))           # nothing kept: a no-op
    assert _lib.resident_hits() >= 0
    # the copy / zero branches of the reference (sos is None) are host-side and need no device
    x = np.arange(20.).reshape(10, 2)
    y = np.empty((7, 2))
    _lib.sosfilt(None, x, y, 3)
    assert np.array_equal(y, x[3:])
    _lib.envelope(None, x, y, 3)
    assert not y.any()
