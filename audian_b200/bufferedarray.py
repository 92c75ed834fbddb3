"""Buffer management base class for derived traces.

audian's `BufferedData` derives from `audioio.BufferedArray`
(reference `src/audian/buffereddata.py:7,10`).  audioio is a third-party
package that is neither vendored under the reference tree nor installed
in this image, so this module provides the members audian relies on,
reconstructed from audian's call sites:

* ``move_buffer(offset, nframes)``   (`buffereddata.py:87`)
* ``allocate_buffer()``              (`buffereddata.py:114`)
* ``reload_buffer()``                (`buffereddata.py:115`)
* overridable ``load_buffer(offset, nframes, buffer)`` (`buffereddata.py:91`)
* ``__len__`` / ``__getitem__`` with slices and tuples that pull the
  requested range into the buffer (`data.py:109-112`, `traceitem.py:58,94`)
* attributes ``rate frames channels shape ndim size offset buffer
  bufferframes backframes follow ampl_min ampl_max unit buffer_changed``
  (`buffereddata.py:40-72`)

If audioio is importable its `BufferedArray` is used instead, so the GPU
traces sit on the very same base class as stock audian.

Behaviour of the stand-in ([recalled] from audioio, stated here so it
can be checked): `move_buffer` keeps the rows of the old buffer that
overlap the new extent, allocates a buffer of the new length and calls
`load_buffer` only for the missing range at the front or at the back;
`reload_buffer` is `load_buffer(self.offset, len(self.buffer),
self.buffer)`; both set `buffer_changed[:] = True`.
"""

import numpy as np

try:                                       # pragma: no cover - not in this image
    from audioio import BufferedArray      # noqa: F401
    HAVE_AUDIOIO = True
except ImportError:
    HAVE_AUDIOIO = False

    class BufferedArray(object):

        def __init__(self, verbose=0):
            self.rate = 0.0
            self.channels = 0
            self.frames = 0
            self.shape = (0, 0)
            self.ndim = 2
            self.size = 0
            self.unit = ''
            self.ampl_min = -1.0
            self.ampl_max = +1.0
            self.offset = 0
            self.buffer = np.zeros((0, 0))
            self.bufferframes = 0
            self.backframes = 0
            self.follow = 0
            self.buffer_changed = np.zeros(0, dtype=bool)
            self.verbose = verbose

        def __len__(self):
            return self.frames

        def __iter__(self):
            for i in range(self.frames):
                yield self[i]

        def __getitem__(self, key):
            index = key[0] if type(key) is tuple else key
            if isinstance(index, slice):
                start = 0 if index.start is None else int(index.start)
                if start < 0:
                    start += len(self)
                stop = len(self) if index.stop is None else int(index.stop)
                if stop < 0:
                    stop += len(self)
                if stop > self.frames:
                    stop = self.frames
                step = 1 if index.step is None else int(index.step)
                self.update_buffer(start, stop)
                newindex = slice(start - self.offset, stop - self.offset, step)
            elif hasattr(index, '__len__'):
                index = [i if i >= 0 else i + len(self) for i in index]
                self.update_buffer(min(index), max(index) + 1)
                newindex = [i - self.offset for i in index]
            else:
                index = int(index)
                if index < 0:
                    index += len(self)
                if index < 0 or index >= self.frames:
                    raise IndexError('index out of range')
                self.update_buffer(index, index + 1)
                newindex = index - self.offset
            if type(key) is tuple:
                return self.buffer[(newindex,) + key[1:]]
            return self.buffer[newindex]

        def update_buffer(self, start, stop):
            """Make sure frames start..stop are in the buffer."""
            if start < self.offset or stop > self.offset + len(self.buffer):
                nframes = max(self.bufferframes, stop - start)
                offset = start - self.backframes
                if offset < 0:
                    offset = 0
                if offset + nframes > self.frames:
                    offset = max(0, self.frames - nframes)
                    nframes = self.frames - offset
                self.move_buffer(offset, nframes)

        def allocate_buffer(self, nframes=None, force=False):
            if nframes is None:
                nframes = self.bufferframes
            if nframes > self.frames:
                nframes = self.frames
            if force or nframes != len(self.buffer) or \
               self.buffer.shape[1:] != tuple(self.shape[1:]):
                shape = list(self.shape)
                shape[0] = nframes
                self.buffer = np.empty(shape)

        def reload_buffer(self):
            if len(self.buffer) > 0:
                self.load_buffer(self.offset, len(self.buffer), self.buffer)
                self.buffer_changed[:] = True

        def move_buffer(self, offset, nframes):
            """Move and resize the buffer, loading only the missing rows."""
            if offset < 0:
                offset = 0
            if offset + nframes > self.frames:
                nframes = self.frames - offset
            if nframes < 0:
                nframes = 0
            if offset == self.offset and nframes == len(self.buffer):
                return
            r_offset, r_nframes = self._recycle_buffer(offset, nframes)
            self.offset = offset
            if r_nframes > 0:
                i = r_offset - self.offset
                self.load_buffer(r_offset, r_nframes,
                                 self.buffer[i:i + r_nframes])
            self.buffer_changed[:] = True

        def _recycle_buffer(self, offset, nframes):
            old = self.buffer
            old_offset = self.offset
            old_n = len(old)
            same_tail = old.shape[1:] == tuple(self.shape[1:])
            r_offset = offset
            r_nframes = nframes
            if same_tail and old_n > 0 and \
               offset >= old_offset and offset < old_offset + old_n:
                # new extent starts inside the old buffer: keep its tail
                i = offset - old_offset
                n = min(old_n - i, nframes)
                keep = old[i:i + n]
                self.allocate_buffer(nframes, force=True)
                self.buffer[:n] = keep
                r_offset = offset + n
                r_nframes = nframes - n
            elif same_tail and old_n > 0 and \
                 offset + nframes > old_offset and \
                 offset + nframes <= old_offset + old_n:
                # new extent ends inside the old buffer: keep its head
                n = offset + nframes - old_offset
                keep = old[:n]
                self.allocate_buffer(nframes, force=True)
                self.buffer[nframes - n:] = keep
                r_nframes = nframes - n
            else:
                self.allocate_buffer(nframes, force=True)
            return r_offset, r_nframes

        def load_buffer(self, offset, nframes, buffer):
            raise NotImplementedError
