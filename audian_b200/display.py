"""Display-side reductions of the traces, computed on the GPU.

What audian's plot items compute from the trace buffers on every update
(SURVEY.md section 8f, rows f1 and f2):

* `spec_image()`      SpecItem.update_plot        src/audian/specitem.py:33-39
* `power_spectrum()`  SpectrogramPlot.update_plot src/audian/spectrogramplot.py:144-164
* `trace_decimate()`  TraceItem.update_plot       src/audian/traceitem.py:33-67

The index algebra is the reference's; the arithmetic runs in the library.  The
functions name the trace's mirror (the device copy its process() left behind), so
usually nothing is uploaded and only the reduced data comes back.
"""

import numpy as np

from . import _lib


def _mirror(trace):
    return getattr(trace, '_mirror', None)


def spec_image(spec_trace, channel):
    """decibel(buffer[:, channel, :].T): the image SpecItem hands to setImage()."""
    return _lib.spec_image_db(spec_trace.buffer, channel, src_mirror=_mirror(spec_trace))


def power_spectrum(spec_trace, channel, t0, t1):
    """(power_db, freqs) of the frames visible in [t0, t1]: mean over time, decibel,
    floored at -200 dB.  Frame indices are relative to the whole trace; the frames must
    lie in the loaded buffer."""
    i0 = int(t0*spec_trace.rate)
    if i0 < 0:
        i0 = 0
    i1 = max(int(t1*spec_trace.rate) - 1, i0 + 1)
    if i1 > len(spec_trace):
        i1 = len(spec_trace)
        if i1 == i0:
            i0 = max(0, i1 - 1)
    b0 = i0 - spec_trace.offset
    b1 = i1 - spec_trace.offset
    if b0 < 0 or b1 > len(spec_trace.buffer) or b1 <= b0:
        raise IndexError('frames %d..%d are not in the loaded buffer' % (i0, i1))
    power = _lib.mean_power_db(spec_trace.buffer, channel, b0, b1, -200.0,
                               src_mirror=_mirror(spec_trace))
    freqs = np.arange(len(power))*spec_trace.fresolution
    return power, freqs


def trace_decimate(trace, channel, t0, t1, max_pixel):
    """(step, plot_time, plot_data) as TraceItem.update_plot draws them: for step > 1
    interleaved min/max of `step` frames aligned to multiples of step and clipped to the
    loaded buffer; for step == 1 the raw samples."""
    rate = trace.rate
    start = max(0, int(t0*rate))
    tstop = int(t1*rate + 1)
    stop = min(len(trace), tstop)
    step = max(1, (tstop - start)//max_pixel)
    off = trace.offset
    if step > 1:
        start = (start//step)*step
        tstop = (stop//step + 1)*step
        stop = min(len(trace), tstop)
        while start < off:
            start += step
        while stop > off + len(trace.buffer):
            stop -= step
        if stop <= start:
            return step, np.zeros(0), np.zeros(0)
        # one column only: gathered on the device from the trace's mirror, else uploaded alone
        plot_data = _lib.minmax_channel(trace.buffer[start - off:stop - off], channel, step,
                                        src_mirror=_mirror(trace))
        step2 = step/2
        plot_time = np.arange(start, start + len(plot_data)*step2, step2)/rate
        return step, plot_time, plot_data
    plot_data = trace.buffer[start - off:stop - off, channel]
    return 1, np.arange(start, stop)/rate, plot_data


def play_region(trace, show_channels, t0, t1, use_heterodyne=False, heterodyne_freq=0.0):
    """(playdata, rate, t0, t1) of DataBrowser.play_region (databrowser.py:1702-1728) up to
    the fade: index clipping as in the reference, the arithmetic on the device."""
    rate = trace.rate
    i0 = int(np.round(t0*rate))
    i1 = int(np.round(t1*rate))
    if i0 < 0:
        i0 = 0
        t0 = 0.0
    if i1 > len(trace):
        i1 = len(trace)
        t1 = i1/rate
    n2 = (len(show_channels) + 1)//2
    left = list(show_channels[:n2])
    right = list(show_channels[n2:]) if len(show_channels) > 1 else []
    seg = trace[i0:i1, :]                       # pulls the range into the buffer, like the reference
    seg = np.ascontiguousarray(seg)
    playdata, prate = _lib.play_region(seg, left, right, rate,
                                       heterodyne_freq if use_heterodyne else 0.0)
    return playdata, prate, t0, t1
