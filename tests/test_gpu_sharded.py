"""The time-sharded drivers on the real kernels.

* fake cluster: N ranks as N host threads sharing one GPU, collectives done in
  process (host barriers only -- no kernel waits on another rank), so the seam
  logic runs on the sm_100a kernels even on a single-GPU box;
* NCCL: world size 2 over two GPUs, when the box has them.

Shard-count invariance: min/max bit-exact, filter and spectrogram equal to the
single-pass oracle within the stated tolerances."""

import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from audian_b200 import _lib, sharded
from audian_b200.synth import synth
from oracle import oracle as orc


class LockedOps(object):
    """CudaOps for rank threads that share one process: the library is specified for one
    host thread at a time (its scratch memory is shared), so every call runs under a lock
    and is complete before the next thread enters."""

    _lock = threading.Lock()

    def __init__(self):
        from audian_b200.device import CudaOps
        self._ops = CudaOps()
        self.name = self._ops.name

    def __getattr__(self, name):
        fn = getattr(self._ops, name)

        def call(*args, **kwargs):
            import torch
            with LockedOps._lock:
                out = fn(*args, **kwargs)
                torch.cuda.synchronize()
                return out
        return call


class FakeDist(object):
    """In-process stand-in for torch.distributed for `world` threads."""

    class P2POp(object):
        def __init__(self, op, tensor, peer):
            self.op, self.tensor, self.peer = op, tensor, peer

    isend, irecv = 'isend', 'irecv'

    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = {}
        self.mail = {}
        self.local = threading.local()

    def bind(self, rank):
        self.local.rank = rank

    def get_rank(self):
        return self.local.rank

    def get_world_size(self):
        return self.world

    def all_gather(self, out_list, tensor):
        import torch
        torch.cuda.synchronize()
        self.slots[self.local.rank] = tensor
        self.barrier.wait()
        for r in range(self.world):
            out_list[r].copy_(self.slots[r])
        torch.cuda.synchronize()
        self.barrier.wait()

    def all_gather_into_tensor(self, out, tensor):
        import torch
        torch.cuda.synchronize()
        self.slots[self.local.rank] = tensor
        self.barrier.wait()
        for r in range(self.world):
            out[r].copy_(self.slots[r])
        torch.cuda.synchronize()
        self.barrier.wait()

    def batch_isend_irecv(self, ops):
        import torch
        torch.cuda.synchronize()
        for op in ops:
            if op.op == 'isend':
                self.mail[(self.local.rank, op.peer)] = op.tensor
        self.barrier.wait()
        for op in ops:
            if op.op == 'irecv':
                op.tensor.copy_(self.mail[(op.peer, self.local.rank)])
        torch.cuda.synchronize()
        self.barrier.wait()

        class Done(object):
            def wait(self):
                pass
        return [Done() for _ in ops]


@pytest.mark.parametrize('world', [2, 4, 8])
def test_fake_cluster_matches_single_pass(world):
    import torch
    from audian_b200.device import CudaOps
    frames, C, rate = 400003, 4, 96000.
    x = synth(0, frames, C, rate, seed=31)
    sos = orc.filter_design(rate, 1000., 15000., 4)
    nfft, hop, step = 1024, 512, 1382
    esos = orc.envelope_design(rate, 500.)
    esos_slow = orc.envelope_design(rate, 2.0)            # remembers ~1e5 samples: crosses shards
    fd = FakeDist(world)
    res = [None]*world
    err = []

    def run(rank):
        try:
            fd.bind(rank)
            torch.cuda.set_device(0)
            ops = LockedOps()
            b = sharded.shard_bounds(frames, world, step)
            lo, hi = b[rank]
            rec = sharded.ShardedRecording(torch.from_numpy(x[lo:hi]).cuda(), frames, rate, ops,
                                           rank, world, b, fd)
            rows = rec.minmax(step)
            b = sharded.shard_bounds(frames, world, hop)
            lo, hi = b[rank]
            rec = sharded.ShardedRecording(torch.from_numpy(x[lo:hi]).cuda(), frames, rate, ops,
                                           rank, world, b, fd)
            y, spec, k0, nf = rec.filter_chain(sos, nfft, hop)
            frec = sharded.ShardedRecording(y, frames, rate, ops, rank, world, b, fd)
            env = frec.envelope(esos, True)
            slow = frec.envelope(esos_slow, True)
            torch.cuda.synchronize()
            res[rank] = (rows.cpu().numpy() if rows is not None else None, lo, y.cpu().numpy(),
                         k0, spec.cpu().numpy(), nf, env.cpu().numpy(), slow.cpu().numpy())
        except Exception as e:          # pragma: no cover
            err.append(e)
            fd.barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not err, err
    ref = orc.minmax_rows(x, step)
    assert np.array_equal(res[0][0].view(np.uint64), ref.view(np.uint64))
    yref = np.empty_like(x)
    orc.filter_process(sos, x, yref, 0)
    y = np.concatenate([r[2] for r in res])
    assert np.max(np.abs(y - yref)) <= 1e-6
    nf = (frames - (nfft - hop))//hop
    sref = np.empty((nf, C, nfft//2 + 1))
    orc.spectrogram_process(yref, sref, rate, nfft, hop)
    spec = np.concatenate([r[4] for r in res])
    assert spec.shape == sref.shape and res[0][5] == nf
    assert np.allclose(spec, sref, rtol=1e-5, atol=1e-20*sref.max())
    # envelope of the filtered recording == one sosfiltfilt over all of it
    for k, es in ((6, esos), (7, esos_slow)):
        eref = np.empty_like(x)
        orc.envelope_process(es, yref, eref, 0, 0)
        env = np.concatenate([r[k] for r in res])
        assert env.shape == eref.shape
        assert np.max(np.abs(env - eref)) <= 1e-6


@pytest.mark.parametrize('world,C', [(1, 8), (2, 8), (4, 3), (8, 2)])
def test_halo_chain_matches_single_pass(world, C):
    """HaloChain: every rank computes its shard of filtered / spectrogram / envelope from its raw
    rows plus halo rows, no exchange; equal to one pass over the whole recording (oracle)."""
    import torch
    from audian_b200 import device
    fs = 48000.
    nfft, hop = 1024, 512
    frames = 1200000//hop*hop + 137
    x = synth(0, frames, C, fs, seed=31)
    sos = orc.filter_design(fs, 1000., 15000., 2)
    esos = orc.envelope_design(fs, 500.)
    rf = np.empty_like(x)
    orc.filter_process(sos, x, rf, 0)
    nf = (frames - (nfft - hop))//hop
    rs = np.empty((nf, C, nfft//2 + 1))
    assert orc.spectrogram_process(rf, rs, fs, nfft, hop) == nf
    re = np.empty_like(x)
    orc.envelope_process(esos, rf, re, 0, 0)
    bounds = sharded.shard_bounds(frames, world, hop)
    assert sharded.HaloChain.supported(sos, esos, bounds)
    filt, spec, env = [], [], []
    for r in range(world):
        hc = sharded.HaloChain(frames, fs, C, bounds, r, sos, esos, nfft, hop)
        r0, r1 = hc.raw_range()
        assert r1 - r0 <= bounds[r][1] - bounds[r][0] + 3*4096
        raw = torch.from_numpy(x[r0:r1]).cuda()
        f, s, e, k0 = hc.run(raw)
        assert k0 == bounds[r][0]//hop
        filt.append(f.cpu().numpy()); spec.append(s.cpu().numpy()); env.append(e.cpu().numpy())
    f = np.concatenate(filt); s = np.concatenate(spec); e = np.concatenate(env)
    assert f.shape == rf.shape and s.shape == rs.shape and e.shape == re.shape
    assert np.max(np.abs(f - rf)) <= 1e-10
    assert np.max(np.abs(e - re)) <= 1e-10
    assert np.allclose(s, rs, rtol=1e-5, atol=1e-20*rs.max())


@pytest.mark.parametrize('el,er,rect', [(1, 1, 1), (0, 1, 1), (1, 0, 0), (0, 0, 1)])
def test_zero_phase_range_edges(el, er, rect):
    """adn_zero_phase_range_f64_dev: scipy's edge handling only at the flagged ends, zero state at
    the others; both the one-pass kernel (long input) and the two sweeps (short input)."""
    import torch
    from scipy.signal import sosfilt, sosfilt_zi
    from audian_b200 import device
    fs, C = 48000., 4
    sos = orc.envelope_design(fs, 500.)
    edge = orc.sosfiltfilt_edge(sos)
    for n in (3000, 400000):
        x = synth(5, n, C, fs, seed=n)
        r = (np.pi/2)*np.abs(x) if rect else x
        parts = [2*r[0] - r[edge:0:-1]] if el else []
        parts.append(r)
        if er:
            parts.append(2*r[-1] - r[-2:-edge - 2:-1])
        ext = np.concatenate(parts)
        zi = sosfilt_zi(sos)
        ref = np.empty_like(ext)
        for c in range(C):
            y1, _ = sosfilt(sos, ext[:, c], zi=zi*ext[0, c] if el else np.zeros_like(zi))
            y2, _ = sosfilt(sos, y1[::-1], zi=zi*y1[-1] if er else np.zeros_like(zi))
            ref[:, c] = y2[::-1]
        off = edge if el else 0
        first, n_dst = 100, n - 300
        z0 = _lib.zero_phase_count()
        got = device.zero_phase_range(sos, torch.from_numpy(x).cuda(), first, n_dst, bool(el), bool(er),
                                      bool(rect), False).cpu().numpy()
        assert (_lib.zero_phase_count() > z0) == (n > 100000)
        assert np.max(np.abs(got - ref[off + first:off + first + n_dst])) <= 1e-10


def test_envelope_sweeps_match_scipy_pieces():
    """The two sweeps exposed for the sharded driver, against scipy on the same pieces."""
    import torch
    from scipy.signal import sosfilt
    from audian_b200 import device
    fs, C, n = 48000., 3, 30011
    hx = synth(0, n, C, fs, seed=41)
    x = torch.from_numpy(hx).cuda()
    sos = orc.envelope_design(fs, 300.)
    S = sos.shape[0]
    edge = _lib.sosfiltfilt_edge(sos)
    r = (np.pi/2)*np.abs(hx)
    zi = np.random.default_rng(1).standard_normal((C, S, 2))
    for el, er in ((0, 0), (edge, 0), (0, edge), (edge, edge)):
        seq = np.concatenate(([2*r[0] - r[el:0:-1]] if el else []) + [r] +
                             ([2*r[-1] - r[-2:-er - 2:-1]] if er else []))
        ref = np.empty_like(seq)
        zref = np.empty((C, S, 2))
        for c in range(C):
            ref[:, c], zref[c] = sosfilt(sos, seq[:, c], zi=zi[c])
        y1, zf = device.env_forward(sos, x, el, er, torch.from_numpy(zi).cuda())
        assert np.max(np.abs(y1.cpu().numpy() - ref)) <= 1e-9
        assert np.max(np.abs(zf.cpu().numpy() - zref)) <= 1e-9
        _, z2 = device.env_forward(sos, x, el, er, torch.from_numpy(zi).cuda(), state_only=True)
        assert torch.equal(z2, zf)
    rev = np.empty_like(hx)
    zref = np.empty((C, S, 2))
    for c in range(C):
        rev[:, c], zref[c] = sosfilt(sos, hx[::-1, c], zi=zi[c])
    out, zf = device.sosfilt_rev(sos, x, torch.from_numpy(zi).cuda(), 100, n - 300, False)
    assert np.max(np.abs(out.cpu().numpy() - rev[::-1][100:n - 200])) <= 1e-9
    assert np.max(np.abs(zf.cpu().numpy() - zref)) <= 1e-9


def _nccl_worker(rank, world, port, frames, C, rate, q):
    import os
    import torch
    import torch.distributed as dist
    from audian_b200 import _lib
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    _lib.init(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        nfft, hop = 512, 256
        b = sharded.shard_bounds(frames, world, hop)
        lo, hi = b[rank]
        x = torch.from_numpy(synth(lo, hi - lo, C, rate, seed=17)).cuda()
        rec = sharded.ShardedRecording(x, frames, rate, bounds=b)
        sos = orc.filter_design(rate, 500., 9000., 2)
        y, spec, k0, nf = rec.filter_chain(sos, nfft, hop)
        frec = sharded.ShardedRecording(y, frames, rate, bounds=b)
        env = frec.envelope(orc.envelope_design(rate, 20.), True)
        torch.cuda.synchronize()
        parts = [None]*world
        dist.all_gather_object(parts, (lo, y.cpu().numpy(), k0, spec.cpu().numpy(), nf,
                                       env.cpu().numpy()))
        if rank == 0:
            q.put(parts)
    finally:
        dist.destroy_process_group()


def test_nccl_two_gpus():
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    frames, C, rate = 300007, 8, 48000.
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, frames, C, rate, q))
             for r in range(2)]
    for p in procs:
        p.start()
    parts = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x = synth(0, frames, C, rate, seed=17)
    sos = orc.filter_design(rate, 500., 9000., 2)
    yref = np.empty_like(x)
    orc.filter_process(sos, x, yref, 0)
    y = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])])
    assert np.max(np.abs(y - yref)) <= 1e-6
    nf = (frames - 256)//256
    sref = np.empty((nf, C, 257))
    orc.spectrogram_process(yref, sref, rate, 512, 256)
    spec = np.concatenate([p[3] for p in sorted(parts, key=lambda p: p[2])])
    assert np.allclose(spec, sref, rtol=1e-5, atol=1e-20*sref.max())
    eref = np.empty_like(x)
    orc.envelope_process(orc.envelope_design(rate, 20.), yref, eref, 0, 0)
    env = np.concatenate([p[5] for p in sorted(parts, key=lambda p: p[0])])
    assert np.max(np.abs(env - eref)) <= 1e-6
