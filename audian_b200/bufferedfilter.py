"""Butterworth high-/low-/band-pass filtered trace, computed on the GPU.

Drop-in for audian's `BufferedFilter` (reference src/audian/bufferedfilter.py:
9-54): same name/source/margins, same attributes (`highpass_cutoff`,
`lowpass_cutoff`, `filter_order`, `sos`) and the same mode selection and
scipy.signal.butter design in `update()`; `process()` runs the SOS cascade on
the device (adn_sosfilt_f64) from zero initial state over the whole source
slice, as the reference does for every (partial) load.
"""

from scipy.signal import butter

from . import _lib
from .buffereddata import BufferedData


class BufferedFilter(BufferedData):

    def __init__(self, name='filtered', source='data', panel='trace',
                 color='#00ee00', lw_thin=1.1, lw_thick=2):
        super().__init__(name, source, tbefore=10, panel=panel,
                         panel_type='trace', color=color,
                         lw_thin=lw_thin, lw_thick=lw_thick)
        self.highpass_cutoff = 0
        self.lowpass_cutoff = 1
        self.filter_order = 2
        self.sos = None

    def open(self, source):
        super().open(source)
        self.highpass_cutoff = 0
        self.lowpass_cutoff = self.rate/2
        self.filter_order = 2
        self.sos = None
        self.update()

    def design(self):
        """Mode thresholds of bufferedfilter.py:40-52."""
        # the reference's expressions, term for term (floating-point rounding at the thresholds)
        no_highpass = self.highpass_cutoff < 0.001*self.rate/2
        no_lowpass = self.lowpass_cutoff >= self.rate/2 - 1e-8
        if no_highpass and no_lowpass:
            return None
        if no_highpass:
            return butter(self.filter_order, self.lowpass_cutoff, 'lowpass',
                          fs=self.rate, output='sos')
        if no_lowpass:
            return butter(self.filter_order, self.highpass_cutoff, 'highpass',
                          fs=self.rate, output='sos')
        return butter(self.filter_order,
                      (self.highpass_cutoff, self.lowpass_cutoff), 'bandpass',
                      fs=self.rate, output='sos')

    def update(self):
        self.sos = self.design()
        self.recompute_all()

    def _standalone_update(self):
        self.sos = self.design()

    def process(self, source, dest, nbefore):
        # sos None -> the library copies source[nbefore:] (bufferedfilter.py:32-33)
        _lib.sosfilt(self.sos, source, dest, nbefore, src_mirror=self.source_mirror(),
                     dst_mirror=self.mirror())
