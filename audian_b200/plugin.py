"""Trace factory that installs the GPU-backed traces in audian.

audian discovers plugins by importing every `audian*.py` in the current
directory and registering callables named `audian_*traces`
(reference src/audian/plugins.py:45-62); a factory receives the DataBrowser and
adds traces with `browser.add_trace()` (databrowser.py:210).  Put a file
`audian_b200_plugin.py` next to the recordings containing

    from audian_b200.plugin import audian_b200_traces

and start audian as usual: 'filtered', 'spectrogram' (and 'envelope') are then
computed on the B200.
"""

from .bufferedenvelope import BufferedEnvelope
from .bufferedfilter import BufferedFilter
from .bufferedspectrogram import BufferedSpectrogram


def audian_b200_traces(browser, envelope=True):
    # source names must be the interned literals 'data' / 'filtered'
    # (data.py:131-132 compares them by identity); the class
    # defaults are.
    browser.clear_traces()
    browser.add_trace(BufferedFilter())
    browser.add_trace(BufferedSpectrogram())
    if envelope:
        browser.add_trace(BufferedEnvelope())
