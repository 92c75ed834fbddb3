"""Run the REAL reference modules in this container.  TEST INFRASTRUCTURE ONLY.

`import audian` is impossible here (PyQt5, pyqtgraph, audioio, thunderlab
are not installed).  The hot-path modules themselves
only need two third-party names, so this harness registers stand-ins for
them in `sys.modules` and then imports the reference's own files, untouched,
from /root/reference/src/audian:

* `audioio.BufferedArray`  -> `audian_b200.bufferedarray.BufferedArray`
  (buffer bookkeeping only, no arithmetic; [recalled] semantics)
* `thunderlab.powerspectrum.spectrogram/decibel` -> `oracle.tl_spectrogram`
  / `oracle.decibel` (scipy.signal.spectrogram(hann, constant, density))
* `audioio.AudioLoader/load_audio/write_audio`, `audioio.audioconverter`,
  `thunderlab.dataloader.DataLoader`, `platformdirs` -> inert placeholders
  (needed only so that `compresseddata.py` imports)

`audian/__init__.py` is bypassed (it pulls in the Qt GUI) by registering an
empty package object whose `__path__` points at the reference directory.

Used by `oracle/make_golden.py` (fixtures in tests/golden/) and, when
/root/reference exists, by tests that compare the oracle with the reference
live.  /root/reference does not exist on the GPU box: nothing marked `gpu`,
`smoke()` or `bench.py` touches this file.
"""

import importlib
import os
import sys
import types

import numpy as np

REFERENCE_SRC = '/root/reference/src'


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_SRC, 'audian',
                                       'buffereddata.py'))


class ArrayLoader(object):
    """In-memory stand-in for thunderlab.dataloader.DataLoader: a (frames,
    channels) float64 array with the buffer attributes audian reads
    (data.py:168-199, compresseddata.py:86-118)."""

    def __init__(self, array, rate, buffer_offset=0, buffer_frames=None,
                 unit='a.u.'):
        self.array = np.ascontiguousarray(array, dtype=np.float64)
        self.rate = float(rate)
        self.frames, self.channels = self.array.shape
        self.shape = self.array.shape
        self.ndim = 2
        self.size = self.array.size
        self.unit = unit
        self.ampl_min = -1.0
        self.ampl_max = 1.0
        self.dests = []
        self.need_update = False
        self.name = 'data'
        self.file_paths = ['memory.wav']
        self.end_indices = None
        self.unwrap_thresh = 0
        self.unwrap_clips = False
        self.set_buffer(buffer_offset, buffer_frames)

    def set_buffer(self, offset, nframes=None):
        if nframes is None:
            nframes = self.frames - offset
        self.offset = int(offset)
        self.buffer = self.array[self.offset:self.offset + int(nframes)]
        self.bufferframes = len(self.buffer)
        self.backframes = 0

    def __len__(self):
        return self.frames

    def load_buffer(self, offset, nframes, buffer):
        buffer[:nframes] = self.array[offset:offset + nframes]

    def set_unwrap(self, *args, **kwargs):
        pass


def _install_stubs():
    from audian_b200.bufferedarray import BufferedArray
    from oracle import oracle as _oracle

    audioio = types.ModuleType('audioio')
    audioio.BufferedArray = BufferedArray
    audioio.AudioLoader = type('AudioLoader', (), {})
    audioio.load_audio = lambda *a, **k: (_ for _ in ()).throw(
        RuntimeError('audioio stand-in: no file I/O'))
    audioio.write_audio = audioio.load_audio
    conv = types.ModuleType('audioio.audioconverter')
    conv.parse_load_kwargs = lambda x: {}
    audioio.audioconverter = conv

    thunderlab = types.ModuleType('thunderlab')
    ps = types.ModuleType('thunderlab.powerspectrum')

    def spectrogram(data, ratetime, freq_resolution=0.2, n_fft=None,
                    overlap_frac=0.5, n_overlap=None, **kwargs):
        return _oracle.tl_spectrogram(data, ratetime, n_fft, n_overlap,
                                      **kwargs)

    ps.spectrogram = spectrogram
    ps.decibel = _oracle.decibel
    dl = types.ModuleType('thunderlab.dataloader')
    dl.DataLoader = ArrayLoader
    thunderlab.powerspectrum = ps
    thunderlab.dataloader = dl

    platformdirs = types.ModuleType('platformdirs')

    class PlatformDirs(object):
        def __init__(self, *a, **k):
            from pathlib import Path
            self.user_cache_path = Path('/nonexistent-audian-cache')
    platformdirs.PlatformDirs = PlatformDirs

    mods = {'audioio': audioio, 'audioio.audioconverter': conv,
            'thunderlab': thunderlab, 'thunderlab.powerspectrum': ps,
            'thunderlab.dataloader': dl}
    for k, v in mods.items():
        sys.modules.setdefault(k, v)
    if 'platformdirs' not in sys.modules:
        try:
            importlib.import_module('platformdirs')
        except ImportError:
            sys.modules['platformdirs'] = platformdirs


_ref = {}


def load_reference():
    """Returns a dict of the reference's hot-path modules."""
    if _ref:
        return _ref
    if not reference_available():
        raise RuntimeError('/root/reference is not present')
    _install_stubs()
    pkg = types.ModuleType('audian')
    pkg.__path__ = [os.path.join(REFERENCE_SRC, 'audian')]
    sys.modules['audian'] = pkg
    for name in ('buffereddata', 'bufferedfilter', 'bufferedenvelope',
                 'bufferedspectrogram', 'compresseddata'):
        _ref[name] = importlib.import_module('audian.' + name)
    return _ref


class FakeSharedArray(object):
    """Quacks like multiprocessing.Array for down_sample_worker
    (compresseddata.py:38,48) without a process boundary."""

    def __init__(self, n):
        self._a = np.zeros(n)

    def get_obj(self):
        return self._a

    def get_lock(self):
        import contextlib
        return contextlib.nullcontext()


def run_reference_chain(array, rate, buffer_offset, buffer_frames,
                        highpass=0.0, lowpass=None, nfft=256, overlap=0.5,
                        envelope_cutoff=None, filter_order=2):
    """Open data -> filtered -> {spectrogram, envelope} with the reference's
    own classes, align all buffers to the loader's buffer and return the
    trace objects (their .buffer/.offset are the reference results)."""
    import contextlib
    import io
    ref = load_reference()
    data = ArrayLoader(array, rate, buffer_offset, buffer_frames)
    filt = ref['bufferedfilter'].BufferedFilter()
    spec = ref['bufferedspectrogram'].BufferedSpectrogram(nfft=nfft,
                                                          overlap_frac=overlap)
    traces = [filt, spec]
    env = None
    if envelope_cutoff is not None:
        env = ref['bufferedenvelope'].BufferedEnvelope(
            envelope_cutoff=envelope_cutoff, filter_order=filter_order)
        traces.append(env)
    with contextlib.redirect_stdout(io.StringIO()):
        filt.open(data)
        filt.need_update = True
        spec.open(filt)
        spec.need_update = True
        if env is not None:
            env.open(filt)
            env.need_update = True
        filt.highpass_cutoff = highpass
        filt.lowpass_cutoff = rate/2 if lowpass is None else lowpass
        filt.filter_order = filter_order
        filt.update()
        for t in traces:
            t.align_buffer()
    return data, filt, spec, env
