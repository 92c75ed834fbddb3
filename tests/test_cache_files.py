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
