"""World-size-2 (and 3) CPU tests of the time-sharded whole-recording drivers
(audian_b200/sharded.py) over the gloo backend.  The arithmetic is injected
from the oracle (scipy on CPU tensors); what is under test is the shard
geometry, the STFT halo exchange, the IIR boundary-state fold and the gather of
the min/max rows -- shard-count invariance against one pass over the whole
recording."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy.signal import butter, sosfilt

from audian_b200 import sharded
from audian_b200.synth import synth
from oracle import oracle as orc


class OracleOps(object):
    """CPU stand-in for audian_b200.device.CudaOps (test infrastructure)."""

    name = 'oracle'

    def minmax(self, src, step):
        return torch.from_numpy(orc.minmax_rows(src.numpy(), step))

    def sosfilt(self, sos, src, nbefore=0, zi=None, want_zf=False, out=None,
                state_only=False, zf_out=None):
        x = src.numpy()
        S = sos.shape[0]
        y = np.empty_like(x)
        zf = np.empty((x.shape[1], S, 2))
        for c in range(x.shape[1]):
            z0 = np.zeros((S, 2)) if zi is None else zi.numpy().reshape(-1, S, 2)[c]
            y[:, c], zf[c] = sosfilt(sos, x[:, c], zi=z0)
        if zf_out is not None:
            zf_out.copy_(torch.from_numpy(zf).reshape(zf_out.shape))
        if state_only:
            return torch.from_numpy(zf)
        y = torch.from_numpy(y[nbefore:].copy())
        if out is not None:
            out.copy_(y)
            y = out
        return (y, torch.from_numpy(zf)) if want_zf else y

    def empty(self, shape):
        return torch.empty(shape, dtype=torch.float64)

    def zeros(self, shape):
        return torch.zeros(shape, dtype=torch.float64)

    def env_state0(self, sos, src, edge, which, out):
        from scipy.signal import sosfilt_zi
        zi = sosfilt_zi(sos).reshape(1, -1)
        x = src.numpy()
        x0 = (np.pi/2)*(2*np.abs(x[0]) - np.abs(x[edge])) if which == 0 else x[0]
        out.copy_(torch.from_numpy(x0.reshape(-1, 1)*zi).reshape(out.shape))
        return out

    def fold_states(self, packs, mats, rank, backward=False):
        P, M = packs.numpy(), mats.numpy()
        W = P.shape[0]
        s = P[W - 1 if backward else 0, 1].copy()
        order = range(W - 1, rank, -1) if backward else range(rank)
        for i in order:
            s = s @ M[i].T + P[i, 0]
        return torch.from_numpy(s)

    def _filt(self, sos, seq, zi, zf_out=None):
        S = sos.shape[0]
        y = np.empty_like(seq)
        zf = np.empty((seq.shape[1], S, 2))
        for c in range(seq.shape[1]):
            z0 = np.zeros((S, 2)) if zi is None else zi.numpy().reshape(-1, S, 2)[c]
            y[:, c], zf[c] = sosfilt(sos, seq[:, c], zi=z0)
        if zf_out is not None:
            zf_out.copy_(torch.from_numpy(zf).reshape(zf_out.shape))
        return y, zf

    def env_forward(self, sos, src, edge_left=0, edge_right=0, zi=None, state_only=False,
                    zf_out=None):
        r = (np.pi/2)*np.abs(src.numpy())
        parts = []
        if edge_left:
            parts.append(2*r[0] - r[edge_left:0:-1])
        parts.append(r)
        if edge_right:
            parts.append(2*r[-1] - r[-2:-edge_right - 2:-1])
        y, zf = self._filt(sos, np.concatenate(parts), zi, zf_out)
        return (None if state_only else torch.from_numpy(y)), torch.from_numpy(zf)

    def sosfilt_rev(self, sos, src, zi=None, first=0, n_dst=None, clamp_negative=False,
                    state_only=False, zf_out=None):
        y, zf = self._filt(sos, src.numpy()[::-1].copy(), zi, zf_out)
        if state_only:
            return None, torch.from_numpy(zf)
        y = y[::-1]
        if n_dst is None:
            n_dst = len(y) - first
        y = y[first:first + n_dst].copy()
        if clamp_negative:
            y[y < 0] = 0
        return torch.from_numpy(y), torch.from_numpy(zf)

    def zero_phase_range(self, sos, src, first, n_dst, edge_left=False, edge_right=False,
                         rectify=True, clamp_negative=True, out=None):
        from scipy.signal import sosfilt_zi
        x = src.numpy()
        r = (np.pi/2)*np.abs(x) if rectify else x
        edge = orc.sosfiltfilt_edge(sos)
        parts = [2*r[0] - r[edge:0:-1]] if edge_left else []
        parts.append(r)
        if edge_right:
            parts.append(2*r[-1] - r[-2:-edge - 2:-1])
        ext = np.concatenate(parts)
        zi = sosfilt_zi(sos)
        S = sos.shape[0]
        y = np.empty_like(ext)
        for c in range(ext.shape[1]):
            z0 = zi*ext[0, c] if edge_left else np.zeros((S, 2))
            y1, _ = sosfilt(sos, ext[:, c], zi=z0)
            z1 = zi*y1[-1] if edge_right else np.zeros((S, 2))
            y2, _ = sosfilt(sos, y1[::-1], zi=z1)
            y[:, c] = y2[::-1]
        el = edge if edge_left else 0
        res = y[el + first:el + first + n_dst].copy()
        if clamp_negative:
            res[res < 0] = 0
        res = torch.from_numpy(res)
        if out is not None:
            out.copy_(res)
            return out
        return res

    def spectrogram(self, src, rate, nfft, hop, n_dst, out_db=False, out=None):
        dst = np.empty((n_dst, src.shape[1], nfft//2 + 1))
        n = orc.spectrogram_process(src.numpy(), dst, rate, nfft, hop)
        if out is not None:
            out.copy_(torch.from_numpy(dst))
            return out, n
        return torch.from_numpy(dst), n


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, frames, C, rate, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ops = OracleOps()
        out = {}
        # ---- min/max
        step = 777
        b = sharded.shard_bounds(frames, world, step)
        lo, hi = b[rank]
        x = torch.from_numpy(synth(lo, hi - lo, C, rate, seed=99))
        rec = sharded.ShardedRecording(x, frames, rate, ops, bounds=b)
        rows = rec.minmax(step)
        if rank == 0:
            out['minmax'] = rows.numpy()
        # ---- filter -> spectrogram chain, hop-aligned shards
        nfft, hop = 256, 64
        b = sharded.shard_bounds(frames, world, hop)
        lo, hi = b[rank]
        x = torch.from_numpy(synth(lo, hi - lo, C, rate, seed=99))
        rec = sharded.ShardedRecording(x, frames, rate, ops, bounds=b)
        sos = butter(4, (0.01*rate, 0.2*rate), 'bandpass', fs=rate, output='sos')
        y, spec, k0, nf = rec.filter_chain(sos, nfft, hop)
        parts = [None]*world
        dist.all_gather_object(parts, (lo, y.numpy(), k0, spec.numpy(), nf))
        if rank == 0:
            out['chain'] = parts
        # ---- envelope of the whole recording, slow (long memory) and fast cut-off
        for name, fc in (('env_slow', 0.0005*rate), ('env_fast', 0.05*rate)):
            esos = butter(2, fc, 'lowpass', fs=rate, output='sos')
            e = rec.envelope(esos, True)
            parts = [None]*world
            dist.all_gather_object(parts, (lo, e.numpy()))
            if rank == 0:
                out[name] = parts
        # ---- the same chain from halo rows, no exchange (fast-forgetting cascades)
        sosf = butter(2, (0.05*rate, 0.3*rate), 'bandpass', fs=rate, output='sos')
        esosf = butter(2, 0.05*rate, 'lowpass', fs=rate, output='sos')
        assert sharded.HaloChain.supported(sosf, esosf, b)
        hc = sharded.HaloChain(frames, rate, C, b, rank, sosf, esosf, nfft, hop, ops)
        r0, r1 = hc.raw_range()
        raw = torch.from_numpy(synth(r0, r1 - r0, C, rate, seed=99))
        fy, fs_, fe, k0 = hc.run(raw)
        parts = [None]*world
        dist.all_gather_object(parts, (lo, fy.numpy(), k0, fs_.numpy(), fe.numpy(), (r0, r1)))
        if rank == 0:
            out['halo'] = parts
        if rank == 0:
            q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_matches_single_pass(world):
    frames, C, rate = 20011, 3, 8000.
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, frames, C, rate, q))
             for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x = synth(0, frames, C, rate, seed=99)
    # min/max: bit-exact regardless of the shard count
    ref = orc.minmax_rows(x, 777)
    assert np.array_equal(out['minmax'].view(np.uint64), ref.view(np.uint64))
    # filter: equal to one sosfilt over the whole recording
    sos = butter(4, (0.01*rate, 0.2*rate), 'bandpass', fs=rate, output='sos')
    yref = np.empty_like(x)
    orc.filter_process(sos, x, yref, 0)
    y = np.concatenate([p[1] for p in sorted(out['chain'], key=lambda p: p[0])])
    assert y.shape == yref.shape
    assert np.max(np.abs(y - yref)) <= 1e-9
    # spectrogram of the filtered recording: same frames as one pass
    nfft, hop = 256, 64
    nf = (frames - (nfft - hop))//hop
    sref = np.empty((nf, C, nfft//2 + 1))
    assert orc.spectrogram_process(yref, sref, rate, nfft, hop) == nf
    spec = np.concatenate([p[3] for p in sorted(out['chain'], key=lambda p: p[2])])
    assert out['chain'][0][4] == nf
    assert spec.shape == sref.shape
    assert np.allclose(spec, sref, rtol=1e-7, atol=1e-22*sref.max())


    # envelope: equal to one sosfiltfilt over the whole recording
    for name, fc in (('env_slow', 0.0005*rate), ('env_fast', 0.05*rate)):
        esos = butter(2, fc, 'lowpass', fs=rate, output='sos')
        eref = np.empty_like(x)
        orc.envelope_process(esos, x, eref, 0, 0)
        e = np.concatenate([p[1] for p in sorted(out[name], key=lambda p: p[0])])
        assert e.shape == eref.shape
        assert np.max(np.abs(e - eref)) <= 1e-9, name


    # the halo chain: every shard equals the rows of one pass over the whole recording
    sosf = butter(2, (0.05*rate, 0.3*rate), 'bandpass', fs=rate, output='sos')
    esosf = butter(2, 0.05*rate, 'lowpass', fs=rate, output='sos')
    yref = np.empty_like(x)
    orc.filter_process(sosf, x, yref, 0)
    parts = sorted(out['halo'], key=lambda p: p[0])
    y = np.concatenate([p[1] for p in parts])
    assert y.shape == yref.shape and np.max(np.abs(y - yref)) <= 1e-12
    sref = np.empty((nf, C, nfft//2 + 1))
    assert orc.spectrogram_process(yref, sref, rate, nfft, hop) == nf
    spec = np.concatenate([p[3] for p in parts])
    assert spec.shape == sref.shape and np.allclose(spec, sref, rtol=1e-7, atol=1e-22*sref.max())
    eref = np.empty_like(x)
    orc.envelope_process(esosf, yref, eref, 0, 0)
    e = np.concatenate([p[4] for p in parts])
    assert e.shape == eref.shape and np.max(np.abs(e - eref)) <= 1e-12
    # halos are a few hundred rows, not shards
    for p in parts:
        assert (p[5][1] - p[5][0]) <= frames//world + 64 + 2*2048


def test_shard_bounds():
    for frames, world, align in [(100, 4, 1), (1000, 3, 64), (5, 8, 2), (777777, 8, 1382)]:
        b = sharded.shard_bounds(frames, world, align)
        assert b[0][0] == 0 and b[-1][1] == frames
        for (l0, h0), (l1, h1) in zip(b[:-1], b[1:]):
            assert h0 == l1 and l1 % align == 0 or l1 == frames


# ---------------------------------------------------------------- streamed whole-file passes

def _wf_worker(rank, world, port, frames, C, rate, q):
    from audian_b200.wholefile import WholeFile
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        def source(t0, n):
            return torch.from_numpy(synth(t0, n, C, rate, seed=7))
        wf = WholeFile(source, frames, C, rate, OracleOps(), rank, world, dist, chunk_frames=3000)
        out = {}
        rows = wf.minmax(250)
        got = {}
        for name, sos in (('fast', butter(2, 0.2*rate, 'lowpass', fs=rate, output='sos')),
                          ('slow', butter(2, 1e-5*rate, 'highpass', fs=rate, output='sos'))):
            parts = []
            wf.sosfilt(sos, lambda t0, y: parts.append((t0, y.numpy().copy())))
            got[name] = parts
        both = []
        rows2 = wf.fulltrace_and_filter(butter(4, 0.1*rate, 'lowpass', fs=rate, output='sos'), 250,
                                        lambda t0, y: both.append((t0, y.numpy().copy())))
        frames_out = []
        nf = wf.spectrogram(128, 32, lambda k, P: frames_out.append((k, P.numpy().copy())))
        allp = [None]*world
        dist.all_gather_object(allp, (got, both, frames_out, nf))
        if rank == 0:
            q.put((rows.numpy(), rows2.numpy(), allp))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [1, 2])
def test_wholefile_streaming_matches_single_pass(world):
    frames, C, rate = 20011, 2, 8000.
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_wf_worker, args=(r, world, port, frames, C, rate, q))
             for r in range(world)]
    for p in procs:
        p.start()
    rows, rows2, allp = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x = synth(0, frames, C, rate, seed=7)
    ref = orc.minmax_rows(x, 250)
    assert np.array_equal(rows.view(np.uint64), ref.view(np.uint64))
    assert np.array_equal(rows2.view(np.uint64), ref.view(np.uint64))

    def join(parts):
        parts = sorted((p for rk in parts for p in rk), key=lambda p: p[0])
        return np.concatenate([p[1] for p in parts])
    for name, sos in (('fast', butter(2, 0.2*rate, 'lowpass', fs=rate, output='sos')),
                      ('slow', butter(2, 1e-5*rate, 'highpass', fs=rate, output='sos'))):
        yref = np.empty_like(x)
        orc.filter_process(sos, x, yref, 0)
        y = join([a[0][name] for a in allp])
        assert y.shape == yref.shape and np.max(np.abs(y - yref)) <= 1e-9, name
    yref = np.empty_like(x)
    orc.filter_process(butter(4, 0.1*rate, 'lowpass', fs=rate, output='sos'), x, yref, 0)
    assert np.max(np.abs(join([a[1] for a in allp]) - yref)) <= 1e-9
    nf = (frames - 96)//32
    sref = np.empty((nf, C, 65))
    orc.spectrogram_process(x, sref, rate, 128, 32)
    spec = join([a[2] for a in allp])
    assert allp[0][3] == nf and spec.shape == sref.shape
    assert np.allclose(spec, sref, rtol=1e-7, atol=1e-22*sref.max())
