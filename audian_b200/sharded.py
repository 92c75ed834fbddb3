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

Cascades that forget their state quickly (every cut-off audian offers at audio
rates: |A^n| < 1e-20 within a few thousand samples) need no exchange at all: a rank
that holds `decay length` extra raw rows on each side of its shard (it reads or
generates them itself, like the STFT halo) computes exactly what one pass over the
whole recording computes, to below the rounding of the states -- `HaloChain`.  The
boundary-state exchange of `ShardedRecording` remains for cascades with long memory.

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


def decay_rows(sos, tol=1e-20):
    """Rows after which the cascade has forgotten its state to within `tol` (a power of
    two), 0 for no filter, -1 if it practically never does."""
    sos_a, S = _lib.sos_array(sos)
    if S == 0:
        return 0
    return _lib.sos_decay_length(sos_a, tol)


class HaloChain(object):
    """data -> filtered -> {spectrogram, envelope} (the dependency walk of the reference,
    src/audian/buffereddata.py:149-153) for ONE time shard [lo, hi) of a recording, from the
    shard's raw rows plus halo rows on both sides -- no exchange between ranks:

        raw rows       [r0, r1) = [f0 - keep_f, f1)      (clipped to the recording)
        filtered rows  [f0, f1) = [lo - keep_e, hi + max(keep_e, nfft - hop))
        spectrogram    frames lo/hop .. hi/hop (global frame index), halo nfft - hop
        envelope       rows [lo, hi); scipy's odd padding / sosfilt_zi only at the ends of the
                       recording, zero state at the ends of the halo

    keep_f / keep_e = decay lengths of the filter and of the envelope low-pass (to `tol`):
    behind them the zero state a shard starts from is forgotten to below the rounding of the
    state itself, so every shard equals the corresponding rows of one pass over the recording."""

    def __init__(self, frames, rate, channels, bounds, rank, sos, esos, nfft, hop, ops=None,
                 tol=1e-20):
        self.frames, self.rate, self.channels = int(frames), float(rate), int(channels)
        self.bounds, self.rank = bounds, rank
        self.lo, self.hi = bounds[rank]
        self.sos, self.S = _lib.sos_array(sos)
        self.esos, self.ES = _lib.sos_array(esos)
        self.nfft, self.hop = int(nfft), int(hop)
        if ops is None:
            from .device import CudaOps
            ops = CudaOps()
        self.ops = ops
        self.keep_f = decay_rows(self.sos, tol)
        self.keep_e = decay_rows(self.esos, tol)
        if self.keep_f < 0 or self.keep_e < 0:
            raise ValueError('a cascade never forgets its state: use ShardedRecording')
        if self.lo % self.hop:
            raise ValueError('shard boundary is not a multiple of hop')
        halo = self.nfft - self.hop
        self.first, self.last = self.lo == 0, self.hi == self.frames
        self.f0 = max(0, self.lo - self.keep_e)
        self.f1 = min(self.frames, self.hi + max(self.keep_e, halo))
        self.r0 = max(0, self.f0 - self.keep_f)
        self.r1 = self.f1
        nf_total = (self.frames - halo)//self.hop if self.frames >= self.nfft else 0
        self.nf_total = nf_total
        self.k0 = self.lo//self.hop
        self.k1 = nf_total if self.last else min(nf_total, self.hi//self.hop)
        self.n_frames = max(0, self.k1 - self.k0)

    @staticmethod
    def supported(sos, esos, bounds, tol=1e-20, max_fraction=0.25):
        """True if both cascades forget within a fraction of the shortest shard."""
        shortest = min(hi - lo for lo, hi in bounds)
        for q in (sos, esos):
            k = decay_rows(q, tol)
            if k < 0 or k > max_fraction*shortest:
                return False
        return True

    def raw_range(self):
        """Rows [r0, r1) of the recording this rank has to hold."""
        return self.r0, self.r1

    def run(self, raw, filt=None, spec=None, env=None, clamp_negative=True, want_env=True):
        """raw: rows r0..r1 of the recording.  Returns (filtered rows lo..hi, spectrogram
        frames k0..k1, envelope rows lo..hi, k0).  filt: optional (f1 - f0, C) buffer."""
        ops = self.ops
        if raw.shape[0] != self.r1 - self.r0:
            raise ValueError('raw rows do not match raw_range()')
        if self.S > 0:
            fext = ops.sosfilt(self.sos, raw, self.f0 - self.r0, out=filt)
        else:
            fext = raw[self.f0 - self.r0:]
        a = self.lo - self.f0
        n = self.hi - self.lo
        spec, ncomp = ops.spectrogram(fext[a:], self.rate, self.nfft, self.hop, self.n_frames, out=spec)
        if ncomp != self.n_frames:
            raise RuntimeError('spectrogram frames: %d computed, %d expected' % (ncomp, self.n_frames))
        if want_env and self.ES > 0:
            env = ops.zero_phase_range(self.esos, fext, a, n, self.first, self.last, True,
                                       clamp_negative, out=env)
        return fext[a:a + n], spec, env, self.k0


class ShardedRecording(object):
    """One rank's time range [lo, hi) of a (frames, C) recording.

    `buffer`: optionally the tensor `local` is the head of -- (hi-lo + room, C)
    with spare rows behind the shard; the STFT halo is then received in place
    and the whole shard is transformed by one kernel launch."""

    _matrix_cache = {}

    def __init__(self, local, frames, rate, ops=None, rank=None, world=None,
                 bounds=None, dist=None, buffer=None):
        # `dist`: torch.distributed (default) or an object with the same all_gather /
        # P2POp / isend / irecv / batch_isend_irecv surface (single-device test clusters)
        self.dist = _dist() if dist is None else dist
        dist = self.dist
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.frames = int(frames)
        self.rate = float(rate)
        self.local = local
        self.buffer = buffer
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

    # ------------------------------------------------------------ plumbing
    def _gather(self, pack):
        """(W,) + pack.shape tensor of every rank's `pack` (one collective, no copies)."""
        import torch
        dist = self.dist
        out = torch.empty((self.world,) + tuple(pack.shape), dtype=pack.dtype, device=pack.device)
        fn = getattr(dist, 'all_gather_into_tensor', None)
        if fn is not None:
            try:
                fn(out, pack)
                return out
            except (RuntimeError, NotImplementedError):
                pass
        parts = [out[i] for i in range(self.world)]
        dist.all_gather(parts, pack)
        return out

    def _matrices(self, sos_a, lens, like):
        """(W, D, D) tensor of A^len_i: the homogeneous map of the cascade across shard i."""
        import torch
        key = (sos_a.tobytes(), tuple(lens), str(like.device))
        mats = ShardedRecording._matrix_cache.get(key)
        if mats is None:
            if len(ShardedRecording._matrix_cache) > 32:
                ShardedRecording._matrix_cache.clear()
            m = np.stack([_lib.sos_state_space(sos_a, n)[2] for n in lens])
            mats = torch.as_tensor(m, dtype=like.dtype, device=like.device).contiguous()
            ShardedRecording._matrix_cache[key] = mats
        return mats

    # ------------------------------------------------------------ min/max
    @staticmethod
    def minmax_bounds(frames, world, step):
        return shard_bounds(frames, world, step)

    def minmax(self, step, dst_rank=0):
        """Full-trace min/max rows (2*ceil(frames/step), C) on `dst_rank`
        (None elsewhere).  Shards must come from minmax_bounds()."""
        import torch
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
        gathered = self._gather(padded.contiguous())
        if dst_rank is not None and self.rank != dst_rank:
            return None
        out = torch.cat([gathered[i, :c] for i, c in enumerate(counts)], dim=0)
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
        n = src.shape[0]
        recv = None
        inplace = False
        if self.world > 1 and halo > 0:
            if any(hi - lo < halo for lo, hi in self.bounds):
                raise ValueError('shards shorter than the STFT halo')
            buf = self.buffer
            inplace = (buf is not None and buf.shape[0] >= n + halo and
                       buf.data_ptr() == src.data_ptr() and buf.is_contiguous())
            recv = buf[n:n + halo] if inplace else \
                torch.empty((halo, self.channels), dtype=src.dtype, device=src.device)
            ops = []
            if self.rank > 0:
                ops.append(dist.P2POp(dist.isend, src[:halo], self.rank - 1))
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
        if inplace:
            out, ncomp = self.ops.spectrogram(self.buffer[:n + halo], self.rate, nfft, hop,
                                              n_local, out_db)
            assert ncomp == n_local, (ncomp, n_local)
            return out, k0, nf_total
        # frames that lie inside the shard come straight from it; only the last few, which
        # reach into the neighbour, are computed from a short tail + halo buffer
        n_main = min(n_local, max(0, (n - nfft)//hop + 1))
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

    def sosfilt(self, sos, zi=None, room=0):
        """This rank's part of sosfilt(sos, recording, axis=0) (zero initial
        state, or `zi` (C, S, 2) applied at frame 0).  room: spare rows to
        allocate behind the result (the returned tensor is the head of
        `self.last_buffer`), for a following in-place halo exchange."""
        import torch
        sos_a, S = _lib.sos_array(sos)
        x = self.local
        n, C = x.shape
        self.last_buffer = None
        if S == 0:
            return x.clone()
        D = 2*S
        ybuf = self.ops.empty((n + room, C))
        y = ybuf[:n]
        self.last_buffer = ybuf
        if self.world == 1:
            self.ops.sosfilt(sos_a, x, 0, zi, out=y)
            return y
        # 1. end state of this shard from zero state (aggregate); a cascade that forgets its
        # state within `keep` samples (|A^keep| < 1e-30) only needs the tail of the shard
        keep = _lib.sos_decay_length(sos_a, 1e-30)
        xs = x[-keep:] if 0 < keep < n else x
        pack = self.ops.zeros((2, C, D)) if zi is None or self.rank > 0 else None
        if pack is None:
            pack = self.ops.empty((2, C, D))
            pack[1].copy_(zi.reshape(C, D))
        self.ops.sosfilt(sos_a, xs, 0, None, state_only=True, zf_out=pack[0])
        # 2. exchange, 3. fold the predecessors: s_{r+1} = A^len_r s_r + v_r
        packs = self._gather(pack)
        mats = self._matrices(sos_a, [hi - lo for lo, hi in self.bounds], x)
        s = self.ops.fold_states(packs, mats, self.rank, False)
        # 4. filter the shard from its true incoming state
        self.ops.sosfilt(sos_a, x, 0, s.reshape(C, S, 2), out=y)
        return y

    # ------------------------------------------------------------ envelope
    def envelope(self, sos, clamp_negative=True):
        """This rank's part of the envelope of the whole recording:
        sosfiltfilt(sos, (pi/2)|recording|, axis=0) with scipy's odd padding at
        the two ends of the recording (bufferedenvelope.py:34-41 over the whole
        file), negatives clamped.  Two exchange steps: the forward boundary
        states travel rank r -> r+1, the backward ones r+1 -> r."""
        import torch
        sos_a, S = _lib.sos_array(sos)
        x = self.local
        n, C = x.shape
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
        keep = _lib.sos_decay_length(sos_a, 1e-30)
        mats = self._matrices(sos_a, lens, x)
        # ---- forward sweep: pack[0] = end state from zero state, pack[1] = scipy's initial
        # state zi * ext[0] (rank 0 only)
        pack = self.ops.zeros((2, C, D))
        if W > 1:
            if 0 < keep < n - edge - 1:
                self.ops.env_forward(sos_a, x[-keep:], 0, er, None, state_only=True, zf_out=pack[0])
            else:
                self.ops.env_forward(sos_a, x, el, er, None, state_only=True, zf_out=pack[0])
        if first:
            self.ops.env_state0(sos_a, x, edge, 0, pack[1])
        packs = self._gather(pack) if W > 1 else pack.reshape(1, 2, C, D)
        s = self.ops.fold_states(packs, mats, r, False)
        y1, _ = self.ops.env_forward(sos_a, x, el, er, s.reshape(C, S, 2))
        # ---- backward sweep (time reversed: the last rank comes first)
        pack = self.ops.zeros((2, C, D))
        if W > 1:
            ys = y1[:keep] if 0 < keep < y1.shape[0] else y1
            self.ops.sosfilt_rev(sos_a, ys, None, state_only=True, zf_out=pack[0])
        if last:
            self.ops.env_state0(sos_a, y1[-1:], 0, 1, pack[1])
        packs = self._gather(pack) if W > 1 else pack.reshape(1, 2, C, D)
        q = self.ops.fold_states(packs, mats, r, True)
        out, _ = self.ops.sosfilt_rev(sos_a, y1, q.reshape(C, S, 2), el, n, clamp_negative)
        return out

    def filter_chain(self, sos, nfft, hop):
        """filtered -> spectrogram of the filtered trace, all sharded."""
        y = self.sosfilt(sos, room=max(0, nfft - hop))
        f = ShardedRecording(y, self.frames, self.rate, self.ops, self.rank,
                             self.world, self.bounds, self.dist, buffer=self.last_buffer)
        spec, k0, nf = f.spectrogram(nfft, hop)
        return y, spec, k0, nf
