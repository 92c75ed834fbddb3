"""Full-trace cache files in the reference's format (compresseddata.py:147-248): DOUBLE WAV with
the scaled rate next to the recording or in the cache directory with the fulltraces.json index."""

import json
import wave

import numpy as np

from audian_b200.compresseddata import CompressedData, write_wav_f64, read_wav_f64


class FakeData(object):
    def __init__(self, path, frames, rate, channels):
        self.filepath = path
        self.file_paths = [str(path)]
        self.frames, self.rate, self.channels = frames, rate, channels
        self.buffer = np.zeros((0, channels))


def test_wav_f64_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1234, 3))
    x[5, 1] = 0.0
    p = tmp_path / 'a.wav'
    write_wav_f64(p, x, 96000)
    y, rate = read_wav_f64(p)
    assert rate == 96000 and np.array_equal(x.view(np.uint64), y.view(np.uint64))
    raw = p.read_bytes()
    assert raw[:4] == b'RIFF' and raw[8:12] == b'WAVE' and raw[20:22] == bytes([3, 0])   # IEEE float


def test_wav_f64_against_an_independent_implementation(tmp_path):
    """libsndfile (which audioio uses for encoding='DOUBLE') is not installed here; scipy's
    WAV reader / writer is an independent implementation of the same IEEE-float WAV layout."""
    from scipy.io import wavfile
    rng = np.random.default_rng(3)
    x = rng.standard_normal((777, 4))
    p = tmp_path / 'ours.wav'
    write_wav_f64(p, x, 144144)
    rate, y = wavfile.read(p)
    assert rate == 144144 and y.dtype == np.float64 and np.array_equal(x, y)
    q = tmp_path / 'theirs.wav'
    wavfile.write(q, 48000, x)
    z, rate = read_wav_f64(q)
    assert rate == 48000 and np.array_equal(x, z)


def make(tmp_path, name='rec.wav'):
    frames, rate, C, step = 4_000_000, 48000., 2, 666
    d = FakeData(tmp_path / name, frames, rate, C)
    cd = CompressedData(d, cache_dir=tmp_path / 'cache')
    cd.short_data = False
    cd.times = np.arange(0, frames + step - 1, step/2)/rate
    cd.datas = np.random.default_rng(1).standard_normal((len(cd.times), C))
    return d, cd


def test_local_cache_file(tmp_path):
    d, cd = make(tmp_path)
    cd.save_data_local()
    assert (tmp_path / 'rec-fulltrace.wav').exists()
    cd2 = CompressedData(d, cache_dir=tmp_path / 'cache')
    cd2.load_data()
    assert np.array_equal(cd2.datas, cd.datas)
    # the rate survives the 1e6 scaling and integer truncation of the header to ~1e-6
    assert np.allclose(cd2.times, cd.times, rtol=1e-5)


def test_user_cache_index_and_lru(tmp_path):
    d, cd = make(tmp_path)
    cd.save_data()
    idx = json.load(open(tmp_path / 'cache' / 'fulltraces.json'))
    assert list(idx) == ['00000001-fulltrace.wav']
    e = idx['00000001-fulltrace.wav']
    assert e['first'] == e['last'] == str((tmp_path / 'rec.wav').absolute())
    assert abs(e['rate'] - 2*48000./666) < 1e-9
    cd2 = CompressedData(d, cache_dir=tmp_path / 'cache')
    cd2.load_data()
    assert np.array_equal(cd2.datas, cd.datas) and np.allclose(cd2.times, cd.times, rtol=1e-12)
    # a second recording gets the next free name; beyond max_files the oldest goes
    d3, cd3 = make(tmp_path, 'other.wav')
    old = CompressedData.max_files
    CompressedData.max_files = 1
    try:
        cd3.save_data()
    finally:
        CompressedData.max_files = old
    idx = json.load(open(tmp_path / 'cache' / 'fulltraces.json'))
    assert list(idx) == ['00000002-fulltrace.wav']
    assert not (tmp_path / 'cache' / '00000001-fulltrace.wav').exists()
    # an index entry whose file vanished is dropped on load
    (tmp_path / 'cache' / '00000002-fulltrace.wav').unlink()
    cd4 = CompressedData(d3, cache_dir=tmp_path / 'cache')
    cd4.load_data()
    assert cd4.datas is None
    assert json.load(open(tmp_path / 'cache' / 'fulltraces.json')) == {}


def test_start_is_non_blocking_and_cancellable(monkeypatch):
    """CompressedData.start returns at once for long recordings (compresseddata.py:104-122): the
    pass runs in a background thread with its own loader, rows appear under get_lock(), is_busy()
    is true until the last block, close() stops it (fulltraceplot.py:166-190 polls these).  The
    kernel is replaced by the oracle's reduceat here -- what is under test is the threading."""
    import threading
    import time
    from audian_b200 import compresseddata
    from oracle import oracle as orc
    from audian_b200.synth import synth

    monkeypatch.setattr(compresseddata._lib, 'minmax', lambda buf, step: orc.minmax_rows(buf, step))
    monkeypatch.setattr(compresseddata._lib, 'host_register', lambda a: None)
    monkeypatch.setattr(compresseddata._lib, 'host_unregister', lambda a: None)
    rate, C = 8000., 2
    frames = int(rate*200)                          # seven 30-s blocks
    x = synth(0, frames, C, rate, seed=5)
    gate = threading.Semaphore(0)
    loads = []

    class Loader(FakeData):
        def load_buffer(self, index, n, buffer):
            gate.acquire()                          # the test releases one block at a time
            loads.append((threading.get_ident(), index))
            buffer[:] = x[index:index + n]

    d = Loader('rec.wav', frames, rate, C)
    private = Loader('rec.wav', frames, rate, C)
    cd = compresseddata.CompressedData(d, loader_factory=lambda: private)
    t0 = time.perf_counter()
    cd.start(1000, {})
    assert time.perf_counter() - t0 < 1.0 and cd.is_busy() and not cd.short_data
    lock = cd.get_lock()
    assert lock.acquire(block=False)                # the reference's polling idiom
    assert not cd.datas.any()
    lock.release()
    gate.release(); gate.release(); gate.release()
    deadline = time.time() + 20
    step = max(1, frames//1000)
    nblock = max(step, int(30.0*rate//step)*step)
    while time.time() < deadline:
        with cd.get_lock():
            done = cd.datas[:2*(nblock//step)].any()
        if done:
            break
        time.sleep(0.01)
    assert done and cd.is_busy()                    # partial rows while the pass is running
    for _ in range(10):
        gate.release()
    cd.wait()
    assert not cd.is_busy()
    _, ref = orc.fulltrace_long(x, 1000, rate, 1)
    assert np.array_equal(cd.datas.view(np.uint64), ref.view(np.uint64))
    assert all(tid != threading.get_ident() for tid, _ in loads)      # never on the caller's thread
    # close() cancels a running pass
    cd2 = compresseddata.CompressedData(d, loader_factory=lambda: private)
    cd2.start(1000, {})
    gate.release()
    cd2.procs[0].terminate()
    for _ in range(10):
        gate.release()
    cd2.close()
    assert not cd2.is_busy() and cd2.procs == []
    assert not cd2.datas[-4:].any()                 # it stopped before the end of the recording
