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

    def recompute_all(self):
        """buffereddata.py:149-153 for the filter: recompute filtered, then its dests.  When the
        dests that need an update are a spectrogram and / or an envelope of this package, the
        whole walk is ONE library call (adn_chain_f64): the source is uploaded once, the filtered
        buffer is consumed on the device, all results come down overlapped.  The buffers end up
        exactly as the walk trace by trace leaves them."""
        if not self.need_update:
            return
        stages = self._chain_stages()
        if not stages:
            return super().recompute_all()
        # what recompute() does for this trace: allocate_buffer() + reload_buffer()
        if len(self.source.buffer) > 0:
            self.allocate_buffer()
        self.invalidate_device()
        if len(self.buffer) == 0:
            return super().recompute_all()
        start, count, nbefore = self.source_slice(self.offset, len(self.buffer))
        src = self.source.buffer[start:start + count]
        kw = {}
        for d in stages:
            d.allocate_buffer()
            d.invalidate_device()
            s0, cnt, nb = d.source_slice(d.offset, len(d.buffer))
            if len(d.buffer) == 0 or cnt <= 0:
                return super().recompute_all()
            kw.update(d.chain_stage(s0, cnt, nb))
        try:
            n = _lib.chain(self.sos, src, self.buffer, self.source.rate, nbefore,
                           src_mirror=self.source_mirror(), filt_mirror=self.mirror(), **kw)
        except ValueError:
            # an envelope slice shorter than its pad: the plain walk raises where the reference does
            return super().recompute_all()
        self.buffer_changed[:] = True
        for d in stages:
            d.buffer_changed[:] = True
            d.chain_done(n)
        for d in self.dests:
            if d in stages:
                for dd in d.dests:
                    dd.recompute_all()
            else:
                d.recompute_all()

    def _chain_stages(self):
        """The dests the chain call can fill: at most one spectrogram and one envelope."""
        kinds = {}
        if type(self).process is not BufferedFilter.process:
            return []                   # a subclass computes its own way: keep the plain walk
        for d in self.dests:
            kind = getattr(d, 'chain_kind', None)
            if kind is None or not d.need_update:
                continue
            if kind in kinds or (kind == 'envelope' and d.sos is None) or \
               type(d).process is not getattr(type(d), 'chain_process', None):
                return []
            kinds[kind] = d
        return list(kinds.values())

    def process(self, source, dest, nbefore):
        # sos None -> the library copies source[nbefore:] (bufferedfilter.py:32-33)
        _lib.sosfilt(self.sos, source, dest, nbefore, src_mirror=self.source_mirror(),
                     dst_mirror=self.mirror())
