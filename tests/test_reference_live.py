"""Live comparison with the reference's own classes (only where /root/reference exists, i.e.
in the build container; skipped on the GPU box): randomized scroll / parameter-change sessions
must drive the drop-in traces through exactly the same load_buffer -> process calls and leave
the same buffers as the reference's BufferedFilter / BufferedSpectrogram / BufferedEnvelope."""

import contextlib
import io

import numpy as np
import pytest

import audian_b200 as ab
from audian_b200.synth import synth
from oracle import oracle as orc
from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason='/root/reference not present')


def instrument(cls, log, process=None):
    class Logged(cls):
        def process(self, source, dest, nbefore):
            base = self.source.buffer
            so = (source.__array_interface__['data'][0] -
                  base.__array_interface__['data'][0])//base.strides[0] \
                if len(base) > 0 and len(source) > 0 else 0
            log.append((self.name, int(so), len(source), len(dest), int(nbefore), int(self.offset)))
            if process is None:
                super().process(source, dest, nbefore)
            else:
                process(self, source, dest, nbefore)
    return Logged


def ours(log):
    F = instrument(ab.BufferedFilter, log, lambda s, a, b, nb: orc.filter_process(s.sos, a, b, nb))
    S = instrument(ab.BufferedSpectrogram, log,
                   lambda s, a, b, nb: orc.spectrogram_process(a, b, s.source.rate, s.nfft, s.hop))
    E = instrument(ab.BufferedEnvelope, log,
                   lambda s, a, b, nb: orc.envelope_process(s.sos, a, b, nb, s.highpass_cutoff))
    return F, S, E


def theirs(log):
    ref = rh.load_reference()
    return (instrument(ref['bufferedfilter'].BufferedFilter, log),
            instrument(ref['bufferedspectrogram'].BufferedSpectrogram, log),
            instrument(ref['bufferedenvelope'].BufferedEnvelope, log))


def session(classes, x, rate, buflen, actions, nfft, overlap):
    F, S, E = classes
    data = rh.ArrayLoader(x, rate, 0, buflen)
    f, s, e = F(), S(nfft=nfft, overlap_frac=overlap), E(envelope_cutoff=rate/40)
    with contextlib.redirect_stdout(io.StringIO()):
        f.open(data)
        s.open(f)
        e.open(f)
        for t in (f, s, e):
            t.need_update = True
        states = []
        for act in actions:
            kind = act[0]
            if kind == 'scroll':
                data.set_buffer(act[1], buflen)
                for t in (f, s, e):
                    t.align_buffer()
            elif kind == 'filter':
                f.highpass_cutoff, f.lowpass_cutoff, f.filter_order = act[1], act[2], act[3]
                f.update()
            elif kind == 'spec':
                s.update(nfft=act[1], overlap_frac=act[2])
            elif kind == 'env':
                e.envelope_cutoff = act[1]
                e.update()
            states.append((f.offset, len(f.buffer), s.offset, len(s.buffer), s.nfft, s.hop,
                           s.rate, s.frames, e.offset, len(e.buffer)))
    return states, f, s, e


@pytest.mark.parametrize('seed', [1, 2, 3, 4, 5, 6])
def test_random_sessions_match_the_reference(seed):
    rng = np.random.default_rng(seed)
    rate = float(rng.choice([1000., 8000., 22050.]))
    C = int(rng.integers(1, 4))
    frames = int(rng.integers(40000, 90000))
    buflen = int(rng.integers(12000, 30000))
    x = synth(0, frames, C, rate, seed=100 + seed)
    nfft0 = int(rng.choice([64, 256, 100]))
    ov0 = float(rng.choice([0.5, 0.75, 0.0]))
    actions = [('filter', rate/50, rate/4, 2)]
    off = 0
    for _ in range(14):
        r = rng.random()
        if r < 0.6:
            off = int(np.clip(off + rng.integers(-buflen, buflen), 0, frames - buflen))
            actions.append(('scroll', off))
        elif r < 0.75:
            hp = float(rng.choice([0., rate/100, rate/20]))
            lp = float(rng.choice([rate/2, rate/3, rate/8]))
            actions.append(('filter', hp, lp, int(rng.choice([2, 4]))))
        elif r < 0.9:
            actions.append(('spec', int(rng.choice([32, 128, 512, 300])), float(rng.choice([0.0, 0.5, 0.875]))))
        else:
            actions.append(('env', float(rng.choice([rate/100, rate/30]))))
    log_a, log_b = [], []
    sa, fa, spa, ea = session(ours(log_a), x, rate, buflen, actions, nfft0, ov0)
    sb, fb, spb, eb = session(theirs(log_b), x, rate, buflen, actions, nfft0, ov0)
    assert log_a == log_b
    assert sa == sb
    assert np.array_equal(fa.buffer, fb.buffer)
    assert np.array_equal(ea.buffer, eb.buffer)
    assert spa.buffer.shape == spb.buffer.shape and np.array_equal(spa.buffer, spb.buffer)
    assert np.array_equal(spa.frequencies, spb.frequencies) and spa.spec_rect == spb.spec_rect
