"""Full-trace min/max cache computed on the GPU.

Mirror of audian's `CompressedData` (reference src/audian/compresseddata.py:
56-145) for the arithmetic part: same `step`, `times` and row layout (row 2j =
min, row 2j+1 = max of segment j; `len(times)` rows for long recordings,
`1 + 2*nseg` rows for recordings that fit the buffer).  The reference's pool of
`cpu_count()-1` worker processes (:104-122) becomes ONE background thread that
owns a private loader: it decodes block k+1 of the recording into one of two
page-locked host buffers while the min/max kernel reduces block k from the
other, and writes the rows into `datas` under a lock.  `start()` returns at
once, `is_busy()` / `get_lock()` / `wait()` / `close()` behave as the
reference's (fulltraceplot.py:166-190 polls them every 500 ms and draws the
partial result), so the drop-in keeps audian's window responsive while a
long recording is read.

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
import threading
from concurrent.futures import ThreadPoolExecutor
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


class FulltraceIndex(object):
    """`fulltraces.json` of the user cache directory: which cached full trace belongs to
    which recording.  {file name: {first, last, rate, created, used}} (the reference's
    format, compresseddata.py:178-186), file names '%08X-fulltrace.wav'."""

    def __init__(self, folder, name='fulltraces.json'):
        self.folder = Path(folder)
        self.path = self.folder / name
        self.entries = {}
        if self.path.exists():
            with self.path.open() as f:
                self.entries = json.load(f)

    @staticmethod
    def key_of(data):
        first = os.fspath(Path(data.file_paths[0]).absolute())
        last = os.fspath(Path(data.file_paths[-1]).absolute())
        return first, last

    def store(self):
        self.folder.mkdir(parents=True, exist_ok=True)
        with self.path.open('w') as f:
            json.dump(self.entries, f, indent=4)

    def free_name(self, limit):
        k = 1
        while k < limit and f'{k:08X}-fulltrace.wav' in self.entries:
            k += 1
        return f'{k:08X}-fulltrace.wav'

    def add(self, name, key, rate):
        now = datetime.now().isoformat()
        self.entries[name] = {'first': key[0], 'last': key[1], 'rate': rate,
                              'created': now, 'used': now}

    def find(self, key):
        for name, e in self.entries.items():
            if (e['first'], e['last']) == key:
                return name
        return None

    def touch(self, name):
        self.entries[name]['used'] = datetime.now().isoformat()

    def drop(self, name, unlink=True):
        self.entries.pop(name, None)
        if unlink:
            try:
                (self.folder / name).unlink()
            except OSError as exc:
                print(exc)

    def trim(self, keep):
        """Least recently used entries (and their files) go until `keep` are left."""
        by_age = sorted(self.entries, key=lambda nm: self.entries[nm]['used'])
        for name in by_age[:max(0, len(by_age) - keep)]:
            self.drop(name)


def file_rate(row_rate):
    """Sampling rate written into the header of a cache file: the rate of the min/max rows
    times 1e6, scaled down by 1e3 while it exceeds 2**31 (compresseddata.py:151-153)."""
    rate = row_rate*1e6
    while rate > 2**31:
        rate /= 1e3
    return rate


class _Lock(object):
    """threading.Lock behind the acquire(block=...) signature of the multiprocessing lock the
    reference hands out (fulltraceplot.py:184: lock.acquire(block=False))."""

    def __init__(self):
        self._lock = threading.Lock()

    def acquire(self, block=True, timeout=None):
        if not block:
            return self._lock.acquire(False)
        return self._lock.acquire(True, -1 if timeout is None else timeout)

    def release(self):
        self._lock.release()

    def __enter__(self):
        self._lock.acquire()
        return self

    def __exit__(self, *exc):
        self._lock.release()


class _Worker(threading.Thread):
    """What the reference's `procs` entries offer (is_alive / join / terminate / close)."""

    def __init__(self, body):
        super().__init__(daemon=True)
        self.body = body                      # body(worker)
        self.cancelled = threading.Event()
        self.error = None

    def run(self):
        self.body(self)

    def terminate(self):
        self.cancelled.set()

    def close(self):
        pass


class CompressedData(object):

    fulltraces_file = 'fulltraces.json'
    max_files = 1000

    def __init__(self, data, cache_dir=None, loader_factory=None):
        """loader_factory: callable returning a private loader (load_buffer(index, n, buffer))
        for the background pass; default: a thunderlab DataLoader opened like the reference's
        workers open theirs (compresseddata.py:29-37), else `data` itself."""
        self.cache_dir = cache_dir
        self.loader_factory = loader_factory
        self.data = data
        self.procs = []
        self.shared_array = None
        self.times = None
        self.datas = None
        self.short_data = True
        self._lock = _Lock()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        """Stops a running pass (compresseddata.py:72-77)."""
        for proc in self.procs:
            proc.terminate()
        for proc in self.procs:
            proc.join()
            proc.close()
        self.procs = []

    # ------------------------------------------------------------ the pass
    def _private_loader(self, nblock, load_kwargs):
        if self.loader_factory is not None:
            return self.loader_factory()
        data = self.data
        try:
            from thunderlab.dataloader import DataLoader
        except ImportError:
            return data
        kwargs = dict(load_kwargs or {})
        tbuffer = nblock/data.rate + 0.1
        if len(data.file_paths) > 1:
            loader = DataLoader(data.file_paths, tbuffer, 0, verbose=0, rate=data.rate,
                                channels=data.channels, unit=data.unit, amax=data.ampl_max,
                                end_indices=data.end_indices, **kwargs)
        else:
            loader = DataLoader(data.file_paths, tbuffer, 0, verbose=0, **kwargs)
        loader.set_unwrap(data.unwrap_thresh, data.unwrap_clips, False, loader.unit)
        return loader

    def _reduce_blocks(self, worker, step, nblock, load_kwargs):
        data = self.data
        frames, C = data.frames, data.channels
        loader = None
        blocks = [np.zeros((nblock, C)) for _ in range(2)]
        pinned = []
        pool = ThreadPoolExecutor(max_workers=1)
        try:
            loader = self._private_loader(nblock, load_kwargs)
            for b in blocks:
                try:
                    _lib.host_register(b)
                    pinned.append(b)
                except Exception:
                    pass                               # pageable memory works too
            pending = None                             # (future, first row) of the block in flight
            k = 0
            for index in range(0, frames, nblock):
                if worker.cancelled.is_set():
                    break
                n = min(nblock, frames - index)
                buf = blocks[k & 1][:n]
                loader.load_buffer(index, n, buf)      # decode block k while block k-1 is reduced
                if pending is not None:
                    self._store_rows(pending)
                pending = (pool.submit(_lib.minmax, buf, step), 2*index//step)
                k += 1
            if pending is not None:
                self._store_rows(pending)
        except Exception as exc:                       # pragma: no cover - surfaced by wait()
            worker.error = exc
        finally:
            pool.shutdown(wait=True)
            for b in pinned:
                try:
                    _lib.host_unregister(b)
                except Exception:
                    pass
            if loader is not None and loader is not data and hasattr(loader, 'close'):
                try:
                    loader.close()
                except Exception:
                    pass

    def _store_rows(self, pending):
        fut, i = pending
        rows = fut.result()
        with self._lock:
            self.datas[i:i + len(rows)] = rows

    def start(self, max_pixel, load_kwargs=None, do_short=True):
        """compresseddata.py:79-122.  Returns at once for long recordings; the rows appear in
        `datas` block by block."""
        if self.times is not None and self.datas is not None:
            return
        self.close()
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
        worker = _Worker(lambda w: self._reduce_blocks(w, step, nblock, load_kwargs))
        self.procs = [worker]
        worker.start()

    def wait(self):
        """Blocks until the pass is done; raises what the pass raised (compresseddata.py:124-129)."""
        errors = [p.error for p in self.procs]
        for p in self.procs:
            p.join()
            errors.append(p.error)
            p.close()
        self.procs = []
        for e in errors:
            if e is not None:
                raise e

    def is_busy(self):
        busy = any(p.is_alive() for p in self.procs)
        if not busy:
            self.procs = []
        return busy

    def get_lock(self):
        return self._lock

    # ------------------------------------------------------------ cache files
    def _cache_path(self):
        return Path(self.cache_dir) if self.cache_dir is not None else default_cache_dir()

    def _row_rate(self):
        return 1/(self.times[1] - self.times[0])

    def _sidecar(self):
        fp = Path(self.data.filepath)
        return fp.with_name(fp.stem + '-fulltrace.wav')

    def save_data_local(self):
        """<stem>-fulltrace.wav next to the recording (compresseddata.py:147-155)."""
        if self.short_data:
            return
        write_wav_f64(self._sidecar(), self.datas, file_rate(self._row_rate()))

    def save_data(self):
        """Into the user cache directory, registered in its index (compresseddata.py:157-202)."""
        if self.short_data:
            return
        index = FulltraceIndex(self._cache_path(), CompressedData.fulltraces_file)
        name = index.free_name(CompressedData.max_files + 10)
        index.add(name, FulltraceIndex.key_of(self.data), self._row_rate())
        index.trim(CompressedData.max_files)
        index.store()
        write_wav_f64(index.folder / name, self.datas, file_rate(self._row_rate()))

    def load_data(self):
        """The sidecar file if there is one, else the user cache (compresseddata.py:204-248)."""
        self.times = None
        self.datas = None
        sidecar = self._sidecar()
        if sidecar.exists():
            rows, header_rate = read_wav_f64(sidecar)
            # which scaling did the writer apply?  the one that gives the recording's duration
            duration = self.data.frames/self.data.rate
            rate = min((header_rate/f for f in (1e6, 1e3, 1.0)),
                       key=lambda r: abs(len(rows)/r - duration))
            self._adopt(rows, rate)
            return
        folder = self._cache_path()
        if not (folder / CompressedData.fulltraces_file).exists():
            return
        index = FulltraceIndex(folder, CompressedData.fulltraces_file)
        name = index.find(FulltraceIndex.key_of(self.data))
        if name is None:
            return
        path = folder / name
        if not path.is_file() or path.stat().st_size == 0:
            index.drop(name, unlink=False)              # the file vanished: forget the entry
            index.store()
            return
        rows, _ = read_wav_f64(path)
        self._adopt(rows, index.entries[name]['rate'])
        index.touch(name)
        index.store()

    def _adopt(self, rows, rate):
        self.datas = rows
        self.times = np.arange(len(rows))/rate
        self.short_data = False
