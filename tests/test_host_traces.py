"""The drop-in trace classes (audian_b200.BufferedFilter/Spectrogram/Envelope on
audian_b200.BufferedData) replay the scroll / parameter-change scenario of the
golden fixture and must issue the same load_buffer -> process calls and end in
the same buffer extents as the reference's classes did.  On CPU the arithmetic
of process() is injected from the oracle (the product has no CPU path); the
`gpu` variant runs the real kernels and compares the buffer contents."""

import json
import os

import numpy as np
import pytest

import audian_b200 as ab
from audian_b200.synth import synth
from oracle import oracle as orc
from oracle.ref_harness import ArrayLoader

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def make_traces(log, use_gpu):
    def logged(cls, oracle_process):
        class Logged(cls):
            def process(self, source, dest, nbefore):
                base = self.source.buffer
                so = (source.__array_interface__['data'][0] -
                      base.__array_interface__['data'][0])//base.strides[0] \
                    if len(base) > 0 and len(source) > 0 else 0
                do = (dest.__array_interface__['data'][0] -
                      self.buffer.__array_interface__['data'][0])//self.buffer.strides[0] \
                    if len(dest) > 0 else 0
                log.append((self.name, int(so), len(source), int(do), len(dest),
                            int(nbefore), int(self.offset)))
                if use_gpu:
                    super().process(source, dest, nbefore)
                else:
                    oracle_process(self, source, dest, nbefore)
        return Logged

    def p_filt(self, s, d, nb):
        orc.filter_process(self.sos, s, d, nb)

    def p_spec(self, s, d, nb):
        orc.spectrogram_process(s, d, self.source.rate, self.nfft, self.hop)

    def p_env(self, s, d, nb):
        orc.envelope_process(self.sos, s, d, nb, self.highpass_cutoff)

    return (logged(ab.BufferedFilter, p_filt), logged(ab.BufferedSpectrogram, p_spec),
            logged(ab.BufferedEnvelope, p_env))


def replay(use_gpu):
    g = np.load(os.path.join(GOLDEN, 'scroll_2ch.npz'))
    a = json.loads(str(g['args']))
    x = synth(0, a['frames'], a['channels'], a['rate'], a['seed'])
    log = []
    F, S, E = make_traces(log, use_gpu)
    data = ArrayLoader(x, a['rate'], 0, a['buflen'])
    filt, spect, env = F(), S(nfft=a['nfft'], overlap_frac=a['overlap']), \
        E(envelope_cutoff=a['envelope_cutoff'])
    filt.open(data)
    spect.open(filt)
    env.open(filt)
    for t in (filt, spect, env):
        t.need_update = True
    filt.highpass_cutoff = a['highpass']
    filt.lowpass_cutoff = a['lowpass']
    filt.update()
    states = []
    for boff in a['offsets']:
        data.set_buffer(boff, a['buflen'])
        for t in (filt, spect, env):
            t.align_buffer()
        states.append([boff, filt.offset, len(filt.buffer), spect.offset, len(spect.buffer),
                       env.offset, len(env.buffer)])
    spect.update(nfft=a['then_nfft'], overlap_frac=a['then_overlap'])
    states.append([a['offsets'][-1], filt.offset, len(filt.buffer), spect.offset,
                   len(spect.buffer), env.offset, len(env.buffer)])
    return g, log, states, filt, spect, env


def check_calls(g, log, states):
    assert [l[0] for l in log] == [str(s) for s in g['log_names']]
    assert np.array_equal(np.array([l[1:] for l in log], dtype=np.int64), g['log'])
    assert np.array_equal(np.array(states, dtype=np.int64), g['states'])


def test_scroll_replay_matches_reference_calls_and_buffers():
    g, log, states, filt, spect, env = replay(use_gpu=False)
    check_calls(g, log, states)
    # with the oracle's arithmetic injected the buffers are the reference's, bit for bit
    assert np.array_equal(filt.buffer, g['filt_buffer'])
    assert np.array_equal(spect.buffer, g['spec_buffer'])
    assert np.array_equal(env.buffer, g['env_buffer'])
    assert spect.hop == int(g['spec_hop']) and spect.rate == float(g['spec_rate'])
    assert filt.buffer_changed.all() and spect.buffer_changed.all()


@pytest.mark.gpu
def test_scroll_replay_on_gpu():
    g, log, states, filt, spect, env = replay(use_gpu=True)
    check_calls(g, log, states)
    assert np.max(np.abs(filt.buffer - g['filt_buffer'])) <= 1e-6
    assert np.max(np.abs(env.buffer - g['env_buffer'])) <= 1e-6
    ref = g['spec_buffer']
    assert np.allclose(spect.buffer, ref, rtol=1e-5, atol=1e-20*ref.max())


def test_trace_attributes_match_reference_contract():
    f, s, e = ab.BufferedFilter(), ab.BufferedSpectrogram(), ab.BufferedEnvelope()
    assert (f.name, f.source_name, f.source_tbefore, f.source_tafter) == ('filtered', 'data', 10, 0)
    assert (s.name, s.source_name, s.source_tbefore, s.source_tafter) == ('spectrogram', 'filtered', 0, 10)
    assert (e.name, e.source_name, e.source_tbefore, e.source_tafter) == ('envelope', 'filtered', 1, 0)
    assert f.source_name is 'data' and s.source_name is 'filtered'        # noqa: F632 (8-Q9)
    assert (s.nfft, s.overlap_frac, s.hop, s.panel_type) == (256, 0.5, 128, 'spectrogram')
    assert (e.envelope_cutoff, e.filter_order, e.highpass_cutoff) == (500, 2, 0)
    assert f.expand_times(1, 2) == (11, 2) and (f.tbefore, f.tafter) == (1, 2)


def test_filter_modes_and_spectrogram_parameter_algebra():
    x = synth(0, 50000, 2, 10000.)
    data = ArrayLoader(x, 10000.)
    f = ab.BufferedFilter()
    f.open(data)
    assert f.sos is None and f.lowpass_cutoff == 5000.
    for hp, lp, S in ((0., 5000., 0), (4.9, 5000., 0), (5.0, 5000., 1), (0., 1000., 1),
                      (100., 1000., 2)):
        f.highpass_cutoff, f.lowpass_cutoff = hp, lp
        f.update()                       # need_update False: designs, does not compute
        ref = orc.filter_design(10000., hp, lp, 2)
        assert (f.sos is None) == (ref is None)
        if ref is not None:
            assert np.array_equal(f.sos, ref) and len(ref) == S
    s = ab.BufferedSpectrogram(nfft=100, overlap_frac=0.33)
    s.open(f)
    assert s.hop == int(100*(1 - 0.33)) == 67                    # open(): truncation
    assert s.rate == 10000./67 and s.frames == (50000 + 66)//67
    assert s.shape == (s.frames, 2, 51)
    s.update(nfft=4, overlap_frac=2.0)
    assert s.nfft == 8 and s.hop == 1 and abs(s.overlap_frac - 0.875) < 1e-12
    s.update(nfft=10**9)
    assert s.nfft == 25000                                       # len(source)//2 (8-Q8)
    e = ab.BufferedEnvelope(envelope_cutoff=6000.)
    e.open(f)
    assert e.sos is None                                          # butter ValueError swallowed


def test_set_need_update_propagates_upstream():
    class Item(object):
        def __init__(self, vis):
            self.vis = vis

        def isVisible(self):
            return self.vis
    x = synth(0, 1000, 1, 1000.)
    data = ArrayLoader(x, 1000.)
    f, s = ab.BufferedFilter(), ab.BufferedSpectrogram(nfft=16)
    f.open(data)
    s.open(f)
    s.plot_items = [Item(True)]
    f.plot_items = [Item(False)]
    f.set_need_update()
    assert s.need_update and f.need_update and data.need_update
    s.plot_items = [Item(False)]
    data.need_update = False
    f.set_need_update()
    assert not s.need_update and not f.need_update and not data.need_update
