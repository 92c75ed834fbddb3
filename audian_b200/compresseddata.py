"""Full-trace min/max cache computed on the GPU.

Mirror of audian's `CompressedData` (reference src/audian/compresseddata.py:
56-130) for the arithmetic part: same `step`, `times` and row layout (row 2j =
min, row 2j+1 = max of segment j; `len(times)` rows for long recordings,
`1 + 2*nseg` rows for recordings that fit the buffer).  The reference's pool of
`cpu_count()-1` worker processes (:104-122) becomes one pass of the min/max
kernel per block of the recording from the calling process -- there is no
process boundary and therefore no lock; `start()` is synchronous, so `is_busy()`
is False and `wait()` returns at once.  The cache files (:147-248) are out of
scope of this round.
"""

import numpy as np

from . import _lib


class CompressedData(object):

    def __init__(self, data):
        self.data = data
        self.procs = []
        self.shared_array = None
        self.times = None
        self.datas = None
        self.short_data = True

    def close(self):
        self.procs = []

    def start(self, max_pixel, load_kwargs=None, do_short=True):
        if self.times is not None and self.datas is not None:
            return
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
        buffer = np.zeros((nblock, data.channels))
        for index in range(0, data.frames, nblock):
            n = min(nblock, data.frames - index)
            data.load_buffer(index, n, buffer[:n])
            i = 2*index//step
            rows = _lib.minmax(buffer[:n], step)
            self.datas[i:i + len(rows)] = rows

    def wait(self):
        pass

    def is_busy(self):
        return False

    def get_lock(self):
        import contextlib
        return contextlib.nullcontext()
