"""Whole-recording passes over recordings that do not fit device memory.

The time-sharded drivers of `audian_b200.sharded` hold a rank's whole shard on
the device.  A 24-h recording (BASELINE config 4: 4 ch x 96 kHz x 24 h = 265 GB
of float64) does not fit, so here every rank walks through its time range in
chunks: a `source(t0, n)` callable delivers rows t0..t0+n of the recording as a
(n, C) device tensor (generated on the device, uploaded from a file, ...), the
kernels run on the chunk, and a `sink` consumes the result -- nothing of the
size of the recording is ever resident.

* min/max      chunks are multiples of `step`; rows are collected per rank and
               gathered (compresseddata.py:79-122 over the whole file).
* filter       the IIR state is carried from chunk to chunk (streamed ==
               one-shot sosfilt); with several ranks the state entering a rank
               comes from the end states of its predecessors, which only need
               the last `decay length` samples of each shard unless the cascade
               never forgets (then the shard is streamed twice).
* spectrogram  chunks of whole frames, each fetched with its nfft-hop rows of
               halo; frames are indexed globally, so the result does not depend
               on the chunking or on the number of ranks.

New functionality relative to the reference, whose only whole-file pass is the
full-trace min/max cache.
"""

import numpy as np

from . import _lib
from .sharded import shard_bounds, ShardedRecording


class WholeFile(object):
    """One rank's view of a (frames, C) recording delivered by `source`."""

    def __init__(self, source, frames, channels, rate, ops=None, rank=0, world=1,
                 dist=None, chunk_frames=1 << 22):
        self.source = source
        self.frames = int(frames)
        self.channels = int(channels)
        self.rate = float(rate)
        if ops is None:
            from .device import CudaOps
            ops = CudaOps()
        self.ops = ops
        self.rank = rank
        self.world = world
        self.dist = dist
        self.chunk_frames = int(chunk_frames)

    # ------------------------------------------------------------ helpers
    def _shard(self, align):
        bounds = shard_bounds(self.frames, self.world, align)
        return bounds, bounds[self.rank]

    def _gatherer(self, like):
        """A ShardedRecording used only for its collective helpers."""
        return ShardedRecording(like, self.frames, self.rate, self.ops, self.rank, self.world,
                                None, self.dist)

    # ------------------------------------------------------------ min/max
    def minmax(self, step, dst_rank=0):
        """(2*ceil(frames/step), C) rows on `dst_rank`: row 2j = min, 2j+1 = max of
        segment j (compresseddata.py:49-52)."""
        import torch
        bounds, (lo, hi) = self._shard(step)
        chunk = max(step, self.chunk_frames//step*step)
        rows = []
        for t0 in range(lo, hi, chunk):
            n = min(chunk, hi - t0)
            rows.append(self.ops.minmax(self.source(t0, n), step))
        local = torch.cat(rows, dim=0) if rows else self.ops.empty((0, self.channels))
        if self.world == 1:
            return local
        counts = [2*((h - l + step - 1)//step) for l, h in bounds]
        width = max(counts)
        padded = self.ops.zeros((width, self.channels))
        padded[:local.shape[0]] = local
        g = self._gatherer(local)._gather(padded)
        if dst_rank is not None and self.rank != dst_rank:
            return None
        return torch.cat([g[i, :c] for i, c in enumerate(counts)], dim=0)

    # ------------------------------------------------------------ filter
    def _incoming_state(self, sos_a, S, bounds):
        """State entering this rank's range: fold of the predecessors' end states."""
        import torch
        C, D = self.channels, 2*S
        lo, hi = bounds[self.rank]
        if self.world == 1:
            return None
        keep = _lib.sos_decay_length(sos_a, 1e-30)
        pack = self.ops.zeros((2, C, D))
        if 0 < keep < hi - lo:
            self.ops.sosfilt(sos_a, self.source(hi - keep, keep), 0, None, state_only=True,
                             zf_out=pack[0])
        else:
            z = None
            for t0 in range(lo, hi, self.chunk_frames):
                n = min(self.chunk_frames, hi - t0)
                z = self.ops.sosfilt(sos_a, self.source(t0, n), 0, z, state_only=True)
            if z is not None:                   # an empty range leaves the pack zero
                pack[0].copy_(z.reshape(C, D))
        g = self._gatherer(pack)
        packs = g._gather(pack)
        mats = g._matrices(sos_a, [h - l for l, h in bounds], pack)
        return self.ops.fold_states(packs, mats, self.rank, False).reshape(C, S, 2)

    def sosfilt(self, sos, sink, also_raw=None):
        """sosfilt(sos, recording, axis=0), streamed: sink(t0, y) receives the filtered
        rows t0.. of this rank's range chunk by chunk; also_raw(t0, x), if given, sees
        the raw chunk while it is resident (e.g. the full-trace min/max of the same pass)."""
        sos_a, S = _lib.sos_array(sos)
        bounds, (lo, hi) = self._shard(1)
        z = self._incoming_state(sos_a, S, bounds) if S > 0 else None
        for t0 in range(lo, hi, self.chunk_frames):
            n = min(self.chunk_frames, hi - t0)
            x = self.source(t0, n)
            if also_raw is not None:
                also_raw(t0, x)
            if S == 0:
                sink(t0, x)
                continue
            y, z = self.ops.sosfilt(sos_a, x, 0, z, want_zf=True)
            sink(t0, y)

    def fulltrace_filter_minmax(self, sos, step, sink=None):
        """BASELINE config 4 with the reductions fused into the filter's pass over the data
        (adn_sosfilt_minmax_f64_dev): returns (rows of the raw recording, rows of the filtered
        recording), each (2*ceil(frames/step), C) gathered to rank 0 (None elsewhere);
        sink(t0, y), if given, also sees every filtered chunk."""
        import torch
        bounds, (lo, hi) = self._shard(step)
        chunk = max(step, self.chunk_frames//step*step)
        sos_a, S = _lib.sos_array(sos)
        saved = self.chunk_frames
        self.chunk_frames = chunk
        try:
            z = self._incoming_state(sos_a, S, bounds) if S > 0 else None
            raw, filt = [], []
            for t0 in range(lo, hi, chunk):
                n = min(chunk, hi - t0)
                x = self.source(t0, n)
                y, z, rr, rf = self.ops.sosfilt_minmax(sos_a, x, step, z)
                raw.append(rr)
                filt.append(rf)
                if sink is not None:
                    sink(t0, y)
        finally:
            self.chunk_frames = saved
        counts = [2*((h - l + step - 1)//step) for l, h in bounds]
        out = []
        for rows in (raw, filt):
            local = torch.cat(rows, dim=0) if rows else self.ops.empty((0, self.channels))
            if self.world == 1:
                out.append(local)
                continue
            padded = self.ops.zeros((max(counts), self.channels))
            padded[:local.shape[0]] = local
            g = self._gatherer(local)._gather(padded)
            out.append(torch.cat([g[i, :c] for i, c in enumerate(counts)], dim=0) if self.rank == 0 else None)
        return out[0], out[1]

    def fulltrace_and_filter(self, sos, step, sink):
        """BASELINE config 4 in one pass over the data: the full-trace min/max rows of
        the raw recording (gathered to rank 0) and the filtered recording (to `sink`)."""
        import torch
        bounds, (lo, hi) = self._shard(step)
        chunk = max(step, self.chunk_frames//step*step)
        sos_a, S = _lib.sos_array(sos)
        saved = self.chunk_frames
        self.chunk_frames = chunk
        try:
            z = self._incoming_state(sos_a, S, bounds) if S > 0 else None
            rows = []
            for t0 in range(lo, hi, chunk):
                n = min(chunk, hi - t0)
                x = self.source(t0, n)
                rows.append(self.ops.minmax(x, step))
                if S == 0:
                    sink(t0, x)
                    continue
                y, z = self.ops.sosfilt(sos_a, x, 0, z, want_zf=True)
                sink(t0, y)
        finally:
            self.chunk_frames = saved
        local = torch.cat(rows, dim=0) if rows else self.ops.empty((0, self.channels))
        if self.world == 1:
            return local
        counts = [2*((h - l + step - 1)//step) for l, h in bounds]
        padded = self.ops.zeros((max(counts), self.channels))
        padded[:local.shape[0]] = local
        g = self._gatherer(local)._gather(padded)
        if self.rank != 0:
            return None
        return torch.cat([g[i, :c] for i, c in enumerate(counts)], dim=0)

    # ------------------------------------------------------------ spectrogram
    def spectrogram(self, nfft, hop, sink, out_db=False):
        """PSD frames of the whole recording, streamed: sink(k0, P) receives frames
        k0.. (global frame index) of this rank's range, P (n, C, nfft//2+1).
        Returns the global frame count."""
        halo = nfft - hop
        nf_total = (self.frames - halo)//hop if self.frames >= nfft else 0
        bounds, (lo, hi) = self._shard(hop)
        k0 = lo//hop
        k1 = min(nf_total, hi//hop if self.rank + 1 < self.world else nf_total)
        per = max(1, self.chunk_frames//hop)
        for k in range(k0, k1, per):
            nk = min(per, k1 - k)
            x = self.source(k*hop, (nk - 1)*hop + nfft)
            P, ncomp = self.ops.spectrogram(x, self.rate, nfft, hop, nk, out_db)
            assert ncomp == nk, (ncomp, nk)
            sink(k, P)
        return nf_total
