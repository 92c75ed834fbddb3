"""Full-trace min/max cache computed on the GPU.

Mirror of audian's `CompressedData` (reference src/audian/compresseddata.py:
56-130) for the arithmetic part: same `step`, `times` and row layout (row 2j =
min, row 2j+1 = max of segment j; `len(times)` rows for long recordings,
`1 + 2*nseg` rows for recordings that fit the buffer).  The reference's pool of
`cpu_count()-1` worker processes (:104-122) becomes one pass of the min/max
kernel per block of the recording from the calling process -- there is no
process boundary and therefore no lock; `start()` is synchronous, so `is_busy()`
is False and `wait()` returns at once.

The cache files are the reference's (:147-248, docs/usermanual.md:12-33), so a
cache written here is picked up by stock audian and vice versa:
`<stem>-fulltrace.wav` next to the recording or `<8 hex digits>-fulltrace.wav`
in the user cache directory (indexed by `fulltraces.json`, least recently used
entries dropped beyond `max_files`), DOUBLE-encoded WAV whose sampling rate is
the rate of the min/max rows scaled by 1e6 (and down by 1e3 while it exceeds
2**31).  The reference writes them with audioio; here a minimal IEEE-float WAV
writer/reader does, audioio not being a dependency of this package.
"""

import json
import os
import struct
from datetime import datetime
from pathlib import Path

import numpy as np

from . import _lib


def write_wav_f64(path, data, rate):
    """RIFF/WAVE, format tag 3 (IEEE float), 64 bit: what audioio's
    write_audio(..., format='WAV', encoding='DOUBLE') produces via libsndfile."""
    data = np.ascontiguousarray(data, dtype='<f8')
    if data.ndim == 1:
        data = data[:, None]
    frames, channels = data.shape
    rate = int(rate)
    payload = data.tobytes()
    fmt = struct.pack('<HHIIHH', 3, channels, rate, rate*channels*8, channels*8, 64)
    fact = struct.pack('<I', frames)
    body = (b'WAVE' + b'fmt ' + struct.pack('<I', len(fmt)) + fmt +
            b'fact' + struct.pack('<I', len(fact)) + fact +
            b'data' + struct.pack('<I', len(payload)) + payload)
    with open(path, 'wb') as f:
        f.write(b'RIFF' + struct.pack('<I', len(body)) + body)


def read_wav_f64(path):
    """(data (frames, channels) float64, rate) of a float64 (or float32) WAV file."""
    with open(path, 'rb') as f:
        raw = f.read()
    if raw[:4] != b'RIFF' or raw[8:12] != b'WAVE':
        raise ValueError('%s is not a WAV file' % path)
    pos = 12
    fmt = None
    while pos + 8 <= len(raw):
        tag, size = raw[pos:pos + 4], struct.unpack('<I', raw[pos + 4:pos + 8])[0]
        body = raw[pos + 8:pos + 8 + size]
        if tag == b'fmt ':
            fmt = struct.unpack('<HHIIHH', body[:16])
            if fmt[0] == 0xFFFE and len(body) >= 26:        # WAVE_FORMAT_EXTENSIBLE
                fmt = (struct.unpack('<H', body[24:26])[0],) + fmt[1:]
        elif tag == b'data':
            if fmt is None or fmt[0] != 3 or fmt[5] not in (32, 64):
                raise ValueError('%s is not an IEEE-float WAV file' % path)
            dt = '<f8' if fmt[5] == 64 else '<f4'
            data = np.frombuffer(body, dtype=dt).astype(np.float64)
            return data.reshape(-1, fmt[1]), fmt[2]
        pos += 8 + size + (size & 1)
    raise ValueError('%s has no data chunk' % path)


def default_cache_dir():
    """audian_dirs.user_cache_path of the reference (version.py:14)."""
    from platformdirs import PlatformDirs
    return Path(PlatformDirs('audian', 'janscience').user_cache_path)


class CompressedData(object):

    fulltraces_file = 'fulltraces.json'
    max_files = 1000

    def __init__(self, data, cache_dir=None):
        self.cache_dir = cache_dir
        self.data = data
        self.procs = []
        self.shared_array = None
        self.times = None
        self.datas = None
        self.short_data = True

    def close(self):
        self.procs = []

    def start(self, max_pixel, load_kwargs=None, do_short=True):
        if self.times is not None and self.datas is not None:
            return
        data = self.data
        step = max(1, data.frames//max_pixel)
        # blocks of ~30 s that are multiples of step (compresseddata.py:84)
        nblock = max(step, int(30.0*data.rate//step)*step)
        self.times = np.arange(0, data.frames + step - 1, step/2)/data.rate
        if len(data.buffer) == data.frames:
            self.short_data = True
            if do_short:
                rows = _lib.minmax(np.ascontiguousarray(data.buffer), step)
                self.datas = np.zeros((1 + len(rows), data.channels))
                self.datas[:len(rows)] = rows
            return
        self.short_data = False
        self.datas = np.zeros((len(self.times), data.channels))
        buffer = np.zeros((nblock, data.channels))
        for index in range(0, data.frames, nblock):
            n = min(nblock, data.frames - index)
            data.load_buffer(index, n, buffer[:n])
            i = 2*index//step
            rows = _lib.minmax(buffer[:n], step)
            self.datas[i:i + len(rows)] = rows

    # ------------------------------------------------------------ cache files
    def _cache_path(self):
        return Path(self.cache_dir) if self.cache_dir is not None else default_cache_dir()

    def _file_rate(self):
        rate = 1/(self.times[1] - self.times[0])
        rate *= 1e6
        while rate > 2**31:
            rate /= 1e3
        return rate

    def save_data_local(self):
        """compresseddata.py:147-155"""
        if self.short_data:
            return
        fp = Path(self.data.filepath)
        write_wav_f64(fp.with_name(fp.stem + '-fulltrace.wav'), self.datas, self._file_rate())

    def save_data(self):
        """compresseddata.py:157-204"""
        if self.short_data:
            return
        cache = self._cache_path()
        cache.mkdir(parents=True, exist_ok=True)
        files = {}
        ft_path = cache / CompressedData.fulltraces_file
        if ft_path.exists():
            with ft_path.open() as sf:
                files = json.load(sf)
        ft_name = f'{1:08X}-fulltrace.wav'
        for k in range(1, CompressedData.max_files + 10):
            ft_name = f'{k:08X}-fulltrace.wav'
            if ft_name not in files:
                break
        first_file = Path(self.data.file_paths[0]).absolute()
        last_file = Path(self.data.file_paths[-1]).absolute()
        timestamp = datetime.now().isoformat()
        rate = 1/(self.times[1] - self.times[0])
        files[ft_name] = dict(first=os.fspath(first_file), last=os.fspath(last_file),
                              rate=rate, created=timestamp, used=timestamp)
        if len(files) > CompressedData.max_files:
            ft_files = list(files)
            stamps = [files[ftf]['used'] for ftf in ft_files]
            idx = np.argsort(stamps)
            for i in idx[:len(ft_files) - CompressedData.max_files]:
                try:
                    (cache / ft_files[i]).unlink()
                except Exception as e:
                    print(e)
                files.pop(ft_files[i])
        with ft_path.open('w') as df:
            json.dump(files, df, indent=4)
        write_wav_f64(cache / ft_name, self.datas, self._file_rate())

    def load_data(self):
        """compresseddata.py:206-248"""
        self.times = None
        self.datas = None
        fp = Path(self.data.filepath)
        ft_path = fp.with_name(fp.stem + '-fulltrace.wav')
        if ft_path.exists():
            self.datas, rate = read_wav_f64(ft_path)
            rates = np.array([rate/1e6, rate/1e3, rate])
            durations = len(self.datas)/rates
            rate = rates[np.argmin(np.abs(durations - self.data.frames/self.data.rate))]
            self.times = np.arange(len(self.datas))/rate
            self.short_data = False
            return
        cache = self._cache_path()
        ft_path = cache / CompressedData.fulltraces_file
        if cache.exists() and ft_path.exists():
            with ft_path.open() as sf:
                files = json.load(sf)
            first_file = Path(self.data.file_paths[0]).absolute()
            last_file = Path(self.data.file_paths[-1]).absolute()
            for ft_file in list(files.keys()):
                props = files[ft_file]
                if props['first'] == os.fspath(first_file) and props['last'] == os.fspath(last_file):
                    ft_file_path = cache / ft_file
                    if not ft_file_path.is_file() or ft_file_path.stat().st_size == 0:
                        del files[ft_file]
                        with ft_path.open('w') as df:
                            json.dump(files, df, indent=4)
                        break
                    self.datas, rate = read_wav_f64(ft_file_path)
                    rate = props['rate']
                    self.times = np.arange(len(self.datas))/rate
                    self.short_data = False
                    props['used'] = datetime.now().isoformat()
                    with ft_path.open('w') as df:
                        json.dump(files, df, indent=4)
                    break

    def wait(self):
        pass

    def is_busy(self):
        return False

    def get_lock(self):
        import contextlib
        return contextlib.nullcontext()
