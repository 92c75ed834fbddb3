"""Power-spectral-density spectrogram trace, computed on the GPU.

Drop-in for audian's `BufferedSpectrogram` (reference
src/audian/bufferedspectrogram.py:12-127): same parameters (`nfft`,
`overlap_frac`, `hop`), same parameter algebra in `open()`, `set_hop()` and
`update()` (incl. truncation in open() vs rounding in set_hop(), SURVEY.md
8-Q2), same derived attributes (`frequencies`, `fresolution`, `tresolution`,
`spec_rect`), buffer of linear PSD values (frames, channels, nfft//2+1).
`process()` = fused mean removal, Hann window, real FFT, |X|^2 scaling on the
device (adn_spectrogram_f64); frames the source slice cannot fill are zero.
`estimate_noiselevels()` takes its decibel values from the device too.
"""

import numpy as np

from . import _lib
from .buffereddata import BufferedData


class BufferedSpectrogram(BufferedData):

    def __init__(self, name='spectrogram', source='filtered',
                 panel='spectrogram', nfft=256, overlap_frac=0.5):
        super().__init__(name, source, tafter=10, panel=panel,
                         panel_type='spectrogram')
        self.nfft = nfft
        self.hop = 0
        self.overlap_frac = overlap_frac
        self.set_hop()
        self.frequencies = np.zeros(0)
        self.fresolution = 1
        self.tresolution = 1
        self.spec_rect = []
        self.use_spec = True
        self.init = True

    def open(self, source):
        self.hop = int(self.nfft*(1 - self.overlap_frac))
        self.fresolution = source.rate/self.nfft
        self.frequencies = np.arange(0, source.rate/2 + self.fresolution/2,
                                     self.fresolution)
        self.tresolution = self.hop/source.rate
        self.spec_rect = []
        self.use_spec = True
        super().open(source, self.hop, more_shape=(self.nfft//2 + 1,))
        self.unit = f'{self.unit}^2/Hz'
        self.ampl_min = 0
        self.ampl_max = self.source.rate/2

    def set_hop(self):
        hop = int(np.round((1 - self.overlap_frac)*self.nfft))
        hop = min(max(hop, 1), self.nfft)
        if hop == self.hop:
            return False
        self.hop = hop
        self.overlap_frac = 1 - self.hop/self.nfft
        return True

    def update(self, nfft=None, overlap_frac=None):
        changed = False
        if nfft is not None:
            nfft = max(nfft, 8)
            nfft = min(nfft, len(self.source)//2, 2**30)
            if nfft != self.nfft:
                self.nfft = nfft
                changed = True
        if overlap_frac is not None:
            self.overlap_frac = min(max(overlap_frac, 0.0), 0.99999)
        if self.set_hop():
            changed = True
        if changed:
            self.tresolution = self.hop/self.source.rate
            self.fresolution = self.source.rate/self.nfft
            self.update_step(self.hop, more_shape=(self.nfft//2 + 1,))
            self.recompute_all()

    def _standalone_update(self):
        self.hop = int(self.nfft*(1 - self.overlap_frac))

    chain_kind = 'spectrogram'

    def chain_stage(self, start, count, nbefore):
        """This trace's stage of BufferedFilter's fused recompute (adn_chain_f64)."""
        return dict(spec=self.buffer, nfft=self.nfft, hop=self.hop, spec_first=start, spec_rows=count)

    def chain_done(self, n):
        self._after_process(n)

    def process(self, source, dest, nbefore):
        n = _lib.spectrogram(source, self.source.rate, self.nfft, self.hop, dest,
                             src_mirror=self.source_mirror(), dst_mirror=self.mirror())
        self._after_process(n)

    def _after_process(self, n):
        if n > 0:
            # what scipy returns as `freq` (bufferedspectrogram.py:60)
            self.frequencies = np.fft.rfftfreq(self.nfft, 1/self.source.rate)
        self.spec_rect = [self.offset/self.rate, 0,
                          len(self.buffer)/self.rate,
                          self.source.rate/2 + self.fresolution]

    def estimate_noiselevels(self, channel):
        if not self.init or len(self.buffer) == 0 or len(self.buffer.shape) < 3:
            return None, None
        nf = max(1, self.buffer.shape[2]//16)
        # (bins, frames) decibel image of the channel, from the device copy of the buffer
        db = _lib.spec_image_db(self.buffer, channel, src_mirror=self._mirror)
        with np.errstate(all='ignore'):
            zmin = np.percentile(db[-nf:, :], 95)
        zmax = np.max(db)
        if not np.isfinite(zmin) or not np.isfinite(zmax):
            return None, None
        self.init = False
        zmax = zmin + 0.95*(zmax - zmin)
        if zmax - zmin < 20:
            zmax = zmin + 20
        if zmax - zmin > 80:
            zmin = zmax - 80
        return zmin, zmax


# the process() the fused recompute of BufferedFilter stands in for: subclasses that override
# process() are recomputed trace by trace
BufferedSpectrogram.chain_process = BufferedSpectrogram.process
