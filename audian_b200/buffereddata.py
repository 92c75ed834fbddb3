"""Base class of the GPU-backed derived traces.

Host-side mirror of audian's `BufferedData` (reference
src/audian/buffereddata.py:10-153): same constructor, attributes and methods,
same buffer index algebra -- including its quirks -- so
that `data.py`, `databrowser.py` and the plot items of stock audian can use
these traces unchanged.  Only `process()` differs in the subclasses: it hands
the numpy views to the sm_100a kernels through the C ABI.

Index algebra, for reference (all integer frames; `src` = the source trace):

    update_step   rate = src.rate/step; frames = ceil(src.frames/step);
                  offset = ceil(src.offset/step)                     (:39-56)
    align_buffer  derived extent = source buffer extent minus
                  floor(source_tbefore*src.rate) frames at the start (if the
                  source buffer does not start at 0) and
                  floor(source_tafter*src.rate) at the end (if it does not reach
                  the end of the recording), mapped with ceil / floor to
                  derived frames; then move_buffer()                 (:75-88)
    load_buffer   source slice = [floor(offset*src.rate/rate),
                  ceil((offset+nframes)*src.rate/rate)) widened by
                  nbefore = floor(source_tbefore/src.rate) and
                  nafter = ceil(source_tafter/src.rate) frames -- divisions,
                  as in the reference -- clipped to the source buffer (:91-109)
"""

from math import ceil, floor

import numpy as np

from . import _lib
from .bufferedarray import BufferedArray


class BufferedData(BufferedArray):

    def __init__(self, name, source_name, tbefore=0, tafter=0,
                 panel='none', panel_type='trace',
                 color='#00ee00', lw_thin=1.1, lw_thick=2):
        super().__init__(verbose=0)
        self.name = name
        self.source_name = source_name
        # margins accumulated from the traces further down the chain
        # (expand_times); the constructor arguments are the margins this trace
        # needs from ITS source (buffereddata.py:18-19,27-28)
        self.tbefore = 0
        self.tafter = 0
        self.source_tbefore = tbefore
        self.source_tafter = tafter
        self.panel = panel
        self.panel_type = panel_type
        self.plot_items = []
        self.color = color
        self.lw_thin = lw_thin
        self.lw_thick = lw_thick
        self.source = None
        self.dests = []
        self.need_update = False
        self.verbose = 0
        self._mirror = None

    # ------------------------------------------------------------ device copies
    # process() of a subclass names `self.mirror()` as the destination mirror of the
    # library call that fills (part of) self.buffer, and `self.source_mirror()` as the
    # source mirror: the consumers of a trace (reference buffereddata.py:149-153 walks
    # filtered -> spectrogram, envelope) read the device copy their source left behind
    # instead of uploading the buffer again.  Explicit hand-over: only buffers owned by
    # a trace are ever kept, and the owner invalidates its mirror whenever audioio moves,
    # replaces or reloads the buffer (move_buffer / allocate_buffer / reload_buffer).
    def mirror(self):
        if self._mirror is None:
            self._mirror = _lib.Mirror()
        return self._mirror

    def source_mirror(self):
        src = self.source
        return getattr(src, '_mirror', None) if src is not None else None

    def invalidate_device(self):
        """The host buffer was changed by other means than process(): forget its device copy."""
        if self._mirror is not None:
            self._mirror.invalidate()

    # ------------------------------------------------------------ wiring
    def expand_times(self, tbefore, tafter):
        self.tbefore += tbefore
        self.tafter += tafter
        return self.source_tbefore + tbefore, self.source_tafter + tafter

    def update_step(self, step=1, more_shape=None):
        src = self.source
        seconds = self.bufferframes/self.rate
        step = max(1, step)
        self.rate = src.rate/step
        self.frames = (src.frames + step - 1)//step
        self.shape = (self.frames, self.channels)
        if more_shape is not None:
            self.shape = self.shape + tuple(more_shape)
        self.ndim = len(self.shape)
        self.size = self.frames*self.channels
        if src.bufferframes == src.frames:
            self.bufferframes = self.frames
        else:
            self.bufferframes = int(seconds*self.rate)
        self.offset = (src.offset + step - 1)//step
        self.follow = 0

    def open(self, source, step=1, more_shape=None):
        self.source = source
        source.dests.append(self)
        self.ampl_min = source.ampl_min
        self.ampl_max = source.ampl_max
        self.unit = source.unit
        self.channels = source.channels
        self.rate = source.rate
        self.bufferframes = 0
        self.backframes = 0
        self.buffer_changed = np.zeros(self.channels, dtype=bool)
        self.buffer = np.zeros((0, self.channels))
        self.plot_items = [None]*self.channels
        self.update_step(step, more_shape)

    # ------------------------------------------------------------ buffers
    # audioio moves the contents of `self.buffer` around (move_buffer recycles the
    # overlap), replaces the array (allocate_buffer) and refills it (reload_buffer):
    # the device copy is invalid from that moment on; the process() calls those methods
    # trigger make it valid again for the rows they fill.
    def move_buffer(self, offset, nframes):
        self.invalidate_device()
        super().move_buffer(offset, nframes)

    def allocate_buffer(self, *args, **kwargs):
        self.invalidate_device()
        super().allocate_buffer(*args, **kwargs)
        # the trace's buffer lives in page-locked memory from the driver's allocator (adn_host_alloc;
        # blocks are pooled by size): copies run at full PCIe rate and overlap with the kernels,
        # 6-7 % faster end to end than a registered pageable array.  The base class has just
        # allocated (and not yet filled) the array that is replaced here.
        buf = self.buffer
        if buf.size > 0 and not _lib.is_pinned_array(buf):
            self.buffer = _lib.buffer_empty(buf.shape, buf.dtype)

    def reload_buffer(self):
        self.invalidate_device()
        super().reload_buffer()

    def align_buffer(self):
        src = self.source
        first = src.offset
        count = len(src.buffer)
        if first > 0:
            margin = floor(self.source_tbefore*src.rate)
            first += margin
            count -= margin
        if src.offset + len(src.buffer) < src.frames:
            count -= floor(self.source_tafter*src.rate)
        # the reference's expression order, kept for its floating-point rounding:
        # (frames*rate)/source_rate, not frames*(rate/source_rate)
        offset = ceil(first*self.rate/src.rate)
        nframes = floor((first + count)*self.rate/src.rate) - offset
        self.move_buffer(offset, nframes)
        self.bufferframes = len(self.buffer)

    def source_slice(self, offset, nframes):
        """(start, count, nbefore) of the slice of the source buffer that
        load_buffer() hands to process() for derived frames offset..+nframes."""
        src = self.source
        start = floor(offset*src.rate/self.rate)               # same expression order as the
        count = ceil((offset + nframes)*src.rate/self.rate) - start     # reference (rounding)
        nbefore = floor(self.source_tbefore/src.rate)      # sic (8-Q1)
        nafter = ceil(self.source_tafter/src.rate)         # sic
        start -= nbefore
        count += nbefore + nafter
        start -= src.offset
        if start < 0:
            nbefore += start
            count += start
            start = 0
        count = min(count, len(src.buffer) - start)
        return start, count, nbefore

    def load_buffer(self, offset, nframes, buffer):
        if self.verbose > 0:
            print(f'load {self.name} {offset/self.rate:.3f} - '
                  f'{(offset + nframes)/self.rate:.3f}')
        start, count, nbefore = self.source_slice(offset, nframes)
        self.process(self.source.buffer[start:start + count], buffer, nbefore)

    def process(self, source, dest, nbefore):
        raise NotImplementedError

    def recompute(self):
        if len(self.source.buffer) > 0:
            self.allocate_buffer()
        self.reload_buffer()

    def recompute_all(self):
        if self.need_update:
            self.recompute()
            for d in self.dests:
                d.recompute_all()

    # ------------------------------------------------------------ visibility
    def is_visible(self):
        return any(pi is not None and pi.isVisible() for pi in self.plot_items)

    def set_visible(self, show):
        for pi in self.plot_items:
            if pi is not None:
                pi.setVisible(show)

    def set_need_update(self):
        self.need_update = self.is_visible()
        for d in self.dests:
            d.set_need_update()
        if len(self.dests) == 0:
            # end of a dependency chain: everything upstream of a needed trace
            # is needed too
            trace = self
            while hasattr(trace, 'source'):
                up = trace.source
                if up is None:
                    break
                up.need_update = trace.need_update or up.need_update
                trace = up

    # ------------------------------------------------------------ stand-alone use
    def configure_standalone(self, rate, channels, source=None, **params):
        """Use `process()` directly on arrays without audian's data graph
        (bench / scripts): sets the attributes `open()` would take from a
        source, applies `params` and designs filters via `update()`.  `source`: the
        trace whose process() fills the arrays this one reads (its device copy is
        then used instead of an upload)."""
        self.rate = float(rate)
        self.channels = int(channels)
        self.source = source if source is not None else _Standalone(rate, channels)
        for k, v in params.items():
            setattr(self, k, v)
        self._standalone_update()

    def _standalone_update(self):
        pass


class _Standalone(object):
    """Minimal source for configure_standalone()."""

    def __init__(self, rate, channels):
        self.rate = float(rate)
        self.channels = int(channels)
        self.frames = 0
        self.offset = 0
        self.bufferframes = 0
        self.buffer = np.zeros((0, channels))
        self.dests = []
        self.need_update = False
