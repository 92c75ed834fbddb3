"""Whole-recording passes, time-sharded over the GPUs of one box.

New functionality relative to the reference (its only parallel code is the
block-cyclic worker pool of the full-trace cache, compresseddata.py:104-122):
one process per GPU (torch.distributed, NCCL over NVLink), the recording is
split into contiguous time ranges, and only what the arithmetic needs crosses
ranks:

* min/max      -- shard boundaries are multiples of `step`; no exchange; the
                  reduced rows are gathered to rank 0.
* spectrogram  -- shard boundaries are multiples of `hop`; each rank receives the
                  first nfft-hop rows of its right neighbour (halo); frames are
                  indexed globally, results stay sharded.
* filter       -- each rank computes the end state of its shard from zero state
                  (state-only pass of the scan kernel), the (C, 2S) vectors are
                  all-gathered, every rank folds its predecessors' states with the
                  shard transition matrices A^len and filters its shard from that
                  incoming state: equal to one sosfilt over the whole recording.

The arithmetic is delegated to an `ops` object (default: the sm_100a kernels,
audian_b200.device.CudaOps); the CPU tests of the exchange logic inject their
own oracle-backed ops with the gloo backend.
"""

import numpy as np

from . import _lib


def _dist():
    import torch.distributed as dist
    return dist


def shard_bounds(frames, world, align=1):
    """Contiguous ranges [lo, hi) per rank, boundaries multiples of `align`
    (except the end of the recording)."""
    units = (frames + align - 1)//align
    per, extra = divmod(units, world)
    bounds = []
    lo = 0
    for r in range(world):
        n = per + (1 if r < extra else 0)
        hi = min(frames, lo + n*align)
        bounds.append((lo, hi))
        lo = hi
    return bounds


class ShardedRecording(object):
    """One rank's time range [lo, hi) of a (frames, C) recording."""

    _matrix_cache = {}

    def __init__(self, local, frames, rate, ops=None, rank=None, world=None,
                 bounds=None, dist=None):
        # `dist`: torch.distributed (default) or an object with the same all_gather /
        # P2POp / isend / irecv / batch_isend_irecv surface (single-device test clusters)
        self.dist = _dist() if dist is None else dist
        dist = self.dist
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.frames = int(frames)
        self.rate = float(rate)
        self.local = local
        self.channels = local.shape[1]
        if ops is None:
            from .device import CudaOps
            ops = CudaOps()
        self.ops = ops
        self.bounds = bounds
        if bounds is not None:
            self.lo, self.hi = bounds[self.rank]
            if self.hi - self.lo != local.shape[0]:
                raise ValueError('local shard does not match its bounds')

    # ------------------------------------------------------------ min/max
    @staticmethod
    def minmax_bounds(frames, world, step):
        return shard_bounds(frames, world, step)

    def minmax(self, step, dst_rank=0):
        """Full-trace min/max rows (2*ceil(frames/step), C) on `dst_rank`
        (None elsewhere).  Shards must come from minmax_bounds()."""
        import torch
        dist = self.dist
        if self.lo % step != 0:
            raise ValueError('shard boundary is not a multiple of step')
        rows = self.ops.minmax(self.local, step)
        nseg_total = (self.frames + step - 1)//step
        counts = [2*((hi - lo + step - 1)//step) for lo, hi in self.bounds]
        width = max(counts)
        padded = rows
        if rows.shape[0] < width:
            padded = torch.zeros((width, self.channels), dtype=rows.dtype, device=rows.device)
            padded[:rows.shape[0]] = rows
        gathered = [torch.empty_like(padded) for _ in range(self.world)]
        dist.all_gather(gathered, padded.contiguous())
        if dst_rank is not None and self.rank != dst_rank:
            return None
        out = torch.cat([g[:c] for g, c in zip(gathered, counts)], dim=0)
        assert out.shape[0] == 2*nseg_total
        return out

    # ------------------------------------------------------------ spectrogram
    @staticmethod
    def spectrogram_bounds(frames, world, hop):
        return shard_bounds(frames, world, hop)

    def spectrogram(self, nfft, hop, out_db=False):
        """This rank's frames [lo/hop, ...) of the PSD spectrogram of the whole
        recording, (n_local, C, nfft//2+1), plus the global index of its first
        frame and the global frame count."""
        import torch
        dist = self.dist
        if self.lo % hop != 0:
            raise ValueError('shard boundary is not a multiple of hop')
        halo = nfft - hop
        nf_total = (self.frames - halo)//hop if self.frames >= nfft else 0
        k0 = self.lo//hop
        k1 = min(nf_total, self.hi//hop if self.rank + 1 < self.world else nf_total)
        src = self.local
        recv = None
        if self.world > 1 and halo > 0:
            if any(hi - lo < halo for lo, hi in self.bounds):
                raise ValueError('shards shorter than the STFT halo')
            recv = torch.empty((halo, self.channels), dtype=src.dtype, device=src.device)
            head = src[:halo].contiguous()
            ops = []
            if self.rank > 0:
                ops.append(dist.P2POp(dist.isend, head, self.rank - 1))
            if self.rank + 1 < self.world:
                ops.append(dist.P2POp(dist.irecv, recv, self.rank + 1))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            if self.rank + 1 >= self.world:
                recv = None
        n_local = max(0, k1 - k0)
        if recv is None:
            out, ncomp = self.ops.spectrogram(src, self.rate, nfft, hop, n_local, out_db)
            assert ncomp == n_local, (ncomp, n_local)
            return out, k0, nf_total
        # frames that lie inside the shard come straight from it; only the last few, which
        # reach into the neighbour, are computed from a short tail + halo buffer
        n_main = min(n_local, max(0, (src.shape[0] - nfft)//hop + 1))
        out = self.ops.empty((n_local, self.channels, nfft//2 + 1))
        if n_main > 0:
            _, ncomp = self.ops.spectrogram(src, self.rate, nfft, hop, n_main, out_db,
                                            out=out[:n_main])
            assert ncomp == n_main, (ncomp, n_main)
        if n_local > n_main:
            tail = torch.cat([src[n_main*hop:], recv], dim=0)
            _, ncomp = self.ops.spectrogram(tail, self.rate, nfft, hop, n_local - n_main,
                                            out_db, out=out[n_main:])
            assert ncomp == n_local - n_main, (ncomp, n_local - n_main)
        return out, k0, nf_total

    # ------------------------------------------------------------ filter
    def shard_matrices(self, sos):
        """A^len for every shard: the homogeneous map of the cascade across it."""
        return [_lib.sos_state_space(sos, hi - lo)[2] for lo, hi in self.bounds]

    def sosfilt(self, sos, zi=None):
        """This rank's part of sosfilt(sos, recording, axis=0) (zero initial
        state, or `zi` (C, S, 2) applied at frame 0)."""
        import torch
        dist = self.dist
        sos_a, S = _lib.sos_array(sos)
        if S == 0:
            return self.local.clone()
        D = 2*S
        C = self.channels
        x = self.local
        if self.world == 1:
            return self.ops.sosfilt(sos_a, x, 0, zi)
        # 1. end state of this shard from zero state (aggregate); a cascade that forgets its
        # state within `keep` samples (|A^keep| < 1e-30) only needs the tail of the shard
        keep = _lib.sos_decay_length(sos_a, 1e-30)
        xs = x[-keep:] if 0 < keep < x.shape[0] else x
        v = self.ops.sosfilt(sos_a, xs, 0, None, state_only=True).reshape(C, D)
        # 2. exchange
        gathered = [torch.empty_like(v) for _ in range(self.world)]
        dist.all_gather(gathered, v.contiguous())
        # 3. fold the predecessors: s_{r+1} = A^len_r s_r + v_r
        key = (sos_a.tobytes(), tuple(self.bounds), str(v.device))
        mats = ShardedRecording._matrix_cache.get(key)
        if mats is None:
            if len(ShardedRecording._matrix_cache) > 32:
                ShardedRecording._matrix_cache.clear()
            mats = [torch.as_tensor(m.T.copy(), dtype=v.dtype, device=v.device)
                    for m in self.shard_matrices(sos_a)]
            ShardedRecording._matrix_cache[key] = mats
        s = torch.zeros((C, D), dtype=v.dtype, device=v.device)
        if zi is not None:
            s = zi.reshape(C, D).clone()
        for r in range(self.rank):
            s = s @ mats[r] + gathered[r]
        # 4. filter the shard from its true incoming state
        return self.ops.sosfilt(sos_a, x, 0, s.reshape(C, S, 2).contiguous())

    # ------------------------------------------------------------ envelope
    def envelope(self, sos, clamp_negative=True):
        """This rank's part of the envelope of the whole recording:
        sosfiltfilt(sos, (pi/2)|recording|, axis=0) with scipy's odd padding at
        the two ends of the recording (bufferedenvelope.py:34-41 over the whole
        file), negatives clamped.  Two exchange steps: the forward boundary
        states travel rank r -> r+1, the backward ones r+1 -> r."""
        import torch
        from scipy.signal import sosfilt_zi
        dist = self.dist
        sos_a, S = _lib.sos_array(sos)
        x = self.local
        C = self.channels
        if S == 0:
            return torch.zeros_like(x)
        D = 2*S
        edge = _lib.sosfiltfilt_edge(sos_a)
        if self.frames <= edge:
            raise ValueError('The length of the input vector x must be greater than '
                             'padlen, which is %d.' % edge)
        r, W = self.rank, self.world
        first, last = r == 0, r == W - 1
        if (self.bounds[0][1] - self.bounds[0][0] <= edge or
                self.bounds[-1][1] - self.bounds[-1][0] <= edge):
            raise ValueError('the first and last shard must be longer than the pad length')
        el = edge if first else 0
        er = edge if last else 0
        lens = [hi - lo + (edge if i == 0 else 0) + (edge if i == W - 1 else 0)
                for i, (lo, hi) in enumerate(self.bounds)]
        zi = torch.as_tensor(sosfilt_zi(sos_a).reshape(1, D), dtype=x.dtype, device=x.device)
        keep = _lib.sos_decay_length(sos_a, 1e-30)
        key = (sos_a.tobytes(), tuple(lens), str(x.device))
        mats = ShardedRecording._matrix_cache.get(key)
        if mats is None:
            if len(ShardedRecording._matrix_cache) > 32:
                ShardedRecording._matrix_cache.clear()
            mats = [torch.as_tensor(_lib.sos_state_space(sos_a, n)[2].T.copy(), dtype=x.dtype,
                                    device=x.device) for n in lens]
            ShardedRecording._matrix_cache[key] = mats
        # ---- forward sweep
        if W == 1:
            v = None
        elif 0 < keep < x.shape[0] - edge - 1:
            _, v = self.ops.env_forward(sos_a, x[-keep:], 0, er, None, state_only=True)
        else:
            _, v = self.ops.env_forward(sos_a, x, el, er, None, state_only=True)
        z0 = torch.zeros((C, D), dtype=x.dtype, device=x.device)
        if first:
            # scipy: zi * ext[0], ext[0] = 2 r[0] - r[edge], r = (pi/2)|x|
            x0 = (np.pi/2)*(2.0*x[0].abs() - x[edge].abs())
            z0 = x0.reshape(C, 1)*zi
        if W > 1:
            pack = torch.cat([v.reshape(C, D), z0], dim=1).contiguous()
            got = [torch.empty_like(pack) for _ in range(W)]
            dist.all_gather(got, pack)
            s = got[0][:, D:].clone()
            for i in range(r):
                s = s @ mats[i] + got[i][:, :D]
        else:
            s = z0
        y1, _ = self.ops.env_forward(sos_a, x, el, er, s.reshape(C, S, 2).contiguous())
        # ---- backward sweep (time reversed: the last rank comes first)
        q0 = torch.zeros((C, D), dtype=x.dtype, device=x.device)
        if last:
            q0 = y1[-1].reshape(C, 1)*zi
        if W > 1:
            if 0 < keep < y1.shape[0]:
                _, w = self.ops.sosfilt_rev(sos_a, y1[:keep], None, state_only=True)
            else:
                _, w = self.ops.sosfilt_rev(sos_a, y1, None, state_only=True)
            pack = torch.cat([w.reshape(C, D), q0], dim=1).contiguous()
            got = [torch.empty_like(pack) for _ in range(W)]
            dist.all_gather(got, pack)
            q = got[W - 1][:, D:].clone()
            for i in range(W - 1, r, -1):
                q = q @ mats[i] + got[i][:, :D]
        else:
            q = q0
        out, _ = self.ops.sosfilt_rev(sos_a, y1, q.reshape(C, S, 2).contiguous(), el,
                                      x.shape[0], clamp_negative)
        return out

    def filter_chain(self, sos, nfft, hop):
        """filtered -> spectrogram of the filtered trace, all sharded."""
        y = self.sosfilt(sos)
        f = ShardedRecording(y, self.frames, self.rate, self.ops, self.rank,
                             self.world, self.bounds, self.dist)
        spec, k0, nf = f.spectrogram(nfft, hop)
        return y, spec, k0, nf
