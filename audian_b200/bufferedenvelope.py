"""Envelope trace: zero-phase low-pass of the rectified source, on the GPU.

Drop-in for audian's `BufferedEnvelope` (reference
src/audian/bufferedenvelope.py:11-56): dest = sosfiltfilt(sos, (pi/2)|source|,
axis=0)[nbefore:], negative values clamped to 0 when there is no high-pass;
`sos is None` (butter raised ValueError) gives zeros.  A source slice not
longer than scipy's pad length raises ValueError exactly like scipy does
(SURVEY.md 8-Q6).
"""

from scipy.signal import butter

from . import _lib
from .buffereddata import BufferedData


class BufferedEnvelope(BufferedData):

    def __init__(self, name='envelope', source='filtered',
                 panel='trace', color='#ff8800',
                 lw_thin=2.5, lw_thick=4, envelope_cutoff=500,
                 filter_order=2, highpass_cutoff=0):
        super().__init__(name, source, tbefore=1, panel=panel,
                         panel_type='trace', color=color,
                         lw_thin=lw_thin, lw_thick=lw_thick)
        self.envelope_cutoff = envelope_cutoff
        self.highpass_cutoff = highpass_cutoff
        self.filter_order = filter_order
        self.sos = None

    def open(self, source):
        super().open(source)
        self.sos = None
        self.update()

    def design(self):
        try:
            if self.highpass_cutoff > 0:
                return butter(self.filter_order,
                              (self.highpass_cutoff, self.envelope_cutoff),
                              'bandpass', fs=self.rate, output='sos')
            return butter(self.filter_order, self.envelope_cutoff, 'lowpass',
                          fs=self.rate, output='sos')
        except ValueError:
            return None

    def update(self):
        self.sos = self.design()
        self.recompute_all()

    def _standalone_update(self):
        self.sos = self.design()

    chain_kind = 'envelope'

    def chain_stage(self, start, count, nbefore):
        """This trace's stage of BufferedFilter's fused recompute (adn_chain_f64)."""
        return dict(esos=self.sos, env=self.buffer, env_first=start, env_rows=count,
                    env_nbefore=nbefore, clamp_negative=(self.highpass_cutoff == 0))

    def chain_done(self, n):
        pass

    def process(self, source, dest, nbefore):
        _lib.envelope(self.sos, source, dest, nbefore,
                      clamp_negative=(self.highpass_cutoff == 0),
                      src_mirror=self.source_mirror(), dst_mirror=self.mirror())


# the process() the fused recompute of BufferedFilter stands in for: subclasses that override
# process() are recomputed trace by trace
BufferedEnvelope.chain_process = BufferedEnvelope.process
