"""Generate tests/golden/*.npz from the REFERENCE's own code.  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Every fixture is produced by the reference's classes
(`BufferedFilter/BufferedSpectrogram/BufferedEnvelope.process`,
`BufferedData.align_buffer/load_buffer`, `CompressedData.start`,
`down_sample_worker`) imported through `oracle/ref_harness.py`, on inputs
from the deterministic generator `audian_b200.synth` (inputs are not
stored; fixtures record generator arguments and a checksum).  Assumption
recorded in every file: thunderlab's spectrogram = scipy hann/constant.
"""

import contextlib
import hashlib
import io
import json
import os

import numpy as np
import scipy

from audian_b200.synth import synth
from oracle import ref_harness as rh

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                      'tests', 'golden')

META = dict(numpy=np.__version__, scipy=scipy.__version__,
            window='hann', detrend='constant',
            note='outputs of /root/reference/src/audian classes run via oracle/ref_harness.py')

# name: (rate, frames, channels, seed, buf_offset, buf_frames, hp, lp, nfft, overlap, env_cutoff, order)
CHAIN_CASES = {
    'chain_mono_44k1': (44100., 30000, 1, 0xA0D1A9 + 1, 0, None, 1000., 15000., 1024, 0.5, 500., 2),
    'chain_2ch_48k': (48000., 24000, 2, 0xA0D1A9 + 2, 0, None, 1000., 15000., 256, 0.5, 500., 2),
    'chain_8ch_lp_o4': (48000., 6000, 8, 0xA0D1A9 + 3, 0, None, 0., 6000., 128, 0.75, 300., 4),
    'chain_3ch_hp': (20000., 9000, 3, 0xA0D1A9 + 4, 0, None, 800., None, 512, 0.875, 200., 2),
    'chain_4ch_bp_o4_nofilt_env': (96000., 12000, 4, 0xA0D1A9 + 5, 0, None, 1000., 15000., 2048, 0.5, 1000., 4),
    'chain_2ch_nofilter': (8000., 5000, 2, 0xA0D1A9 + 6, 0, None, 0., None, 64, 0.0, 100., 2),
    # buffer inside a longer recording (offset > 0, not reaching the end):
    # exercises the tbefore/tafter margins of align_buffer and the quirks
    # of load_buffer
    'chain_2ch_window': (1000., 60000, 2, 0xA0D1A9 + 7, 15000, 30000, 20., 300., 64, 0.5, 10., 2),
}


def checksum(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_chain(name, spec):
    rate, frames, C, seed, boff, bfr, hp, lp, nfft, ov, envc, order = spec
    x = synth(0, frames, C, rate, seed)
    data, filt, spect, env = rh.run_reference_chain(
        x, rate, boff, bfr, highpass=hp, lowpass=lp, nfft=nfft, overlap=ov,
        envelope_cutoff=envc, filter_order=order)
    out = dict(
        args=json.dumps(dict(rate=rate, frames=frames, channels=C, seed=seed,
                             buf_offset=boff, buf_frames=bfr, highpass=hp,
                             lowpass=lp, nfft=nfft, overlap=ov,
                             envelope_cutoff=envc, order=order)),
        meta=json.dumps(META), input_sha256=checksum(x),
        filt_sos=np.zeros((0, 6)) if filt.sos is None else filt.sos,
        filt_offset=filt.offset, filt_buffer=filt.buffer,
        spec_offset=spect.offset, spec_buffer=spect.buffer,
        spec_hop=spect.hop, spec_frequencies=spect.frequencies,
        spec_rate=spect.rate, spec_frames=spect.frames,
        env_sos=np.zeros((0, 6)) if env.sos is None else env.sos,
        env_offset=env.offset, env_buffer=env.buffer)
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)
    print(name, filt.buffer.shape, spect.buffer.shape, env.buffer.shape)


def make_scroll():
    """Scroll the loader's buffer and log every load_buffer -> process call of
    the reference (partial updates): the golden for the index algebra."""
    ref = rh.load_reference()
    rate, frames, C = 1000., 120000, 2
    x = synth(0, frames, C, rate, 0xA0D1A9 + 8)
    log = []

    def logged(cls):
        class Logged(cls):
            def process(self, source, dest, nbefore):
                base = self.source.buffer
                so = (source.__array_interface__['data'][0] -
                      base.__array_interface__['data'][0])//base.strides[0] \
                    if len(base) > 0 and len(source) > 0 else 0
                do = (dest.__array_interface__['data'][0] -
                      self.buffer.__array_interface__['data'][0])//self.buffer.strides[0] \
                    if len(dest) > 0 else 0
                log.append((self.name, int(so), len(source), int(do),
                            len(dest), int(nbefore), int(self.offset)))
                super().process(source, dest, nbefore)
        return Logged

    data = rh.ArrayLoader(x, rate, 0, 40000)
    filt = logged(ref['bufferedfilter'].BufferedFilter)()
    spect = logged(ref['bufferedspectrogram'].BufferedSpectrogram)(nfft=128, overlap_frac=0.5)
    env = logged(ref['bufferedenvelope'].BufferedEnvelope)(envelope_cutoff=10.)
    states = []
    with contextlib.redirect_stdout(io.StringIO()):
        filt.open(data)
        spect.open(filt)
        env.open(filt)
        for t in (filt, spect, env):
            t.need_update = True
        filt.highpass_cutoff = 20.
        filt.lowpass_cutoff = 300.
        filt.update()
        for boff in (0, 2000, 10000, 9000, 50000, 80000):
            data.set_buffer(boff, 40000)
            for t in (filt, spect, env):
                t.align_buffer()
            states.append([boff, filt.offset, len(filt.buffer), spect.offset,
                           len(spect.buffer), env.offset, len(env.buffer)])
        # parameter change -> recompute_all
        spect.update(nfft=256, overlap_frac=0.75)
        states.append([80000, filt.offset, len(filt.buffer), spect.offset,
                       len(spect.buffer), env.offset, len(env.buffer)])
    np.savez_compressed(
        os.path.join(GOLDEN, 'scroll_2ch.npz'),
        args=json.dumps(dict(rate=rate, frames=frames, channels=C,
                             seed=0xA0D1A9 + 8, buflen=40000,
                             offsets=[0, 2000, 10000, 9000, 50000, 80000],
                             highpass=20., lowpass=300., nfft=128, overlap=0.5,
                             envelope_cutoff=10., then_nfft=256,
                             then_overlap=0.75)),
        meta=json.dumps(META), input_sha256=checksum(x),
        log_names=np.array([l[0] for l in log]),
        log=np.array([l[1:] for l in log], dtype=np.int64),
        states=np.array(states, dtype=np.int64),
        filt_buffer=filt.buffer, spec_buffer=spect.buffer,
        env_buffer=env.buffer, spec_hop=spect.hop, spec_rate=spect.rate)
    print('scroll', len(log), 'process calls')


def make_fulltrace():
    ref = rh.load_reference()
    cd = ref['compresseddata']
    out = dict(meta=json.dumps(META))
    cases = {'short_3ch': (8000., 50000, 3, 0xA0D1A9 + 9, 700),
             'short_1ch_step1': (1000., 300, 1, 0xA0D1A9 + 10, 6000),
             'long_4ch': (96000., 400000, 4, 0xA0D1A9 + 11, 333)}
    for name, (rate, frames, C, seed, max_pixel) in cases.items():
        x = synth(0, frames, C, rate, seed)
        if name.startswith('short'):
            data = rh.ArrayLoader(x, rate)
            comp = cd.CompressedData(data)
            comp.start(max_pixel, {})
            times, datas = comp.times, comp.datas
        else:
            # long path: the reference's worker, called in-process for each
            # of 3 "processes" (block-cyclic over 30-s blocks)
            step = max(1, frames//max_pixel)
            nblock = max(step, int(30.0*rate//step)*step)
            nblock = min(nblock, 20*step)      # several blocks in a short file
            times = np.arange(0, frames + step - 1, step/2)/rate
            arr = rh.FakeSharedArray(len(times)*C)
            import sys
            sys.modules['thunderlab.dataloader'].DataLoader = \
                lambda fp, tb, tback, verbose=0, **kw: rh.ArrayLoader(x, rate)
            cd.DataLoader = sys.modules['thunderlab.dataloader'].DataLoader
            for p in range(3):
                cd.down_sample_worker(p, 3, nblock, step, arr, ['m.wav'], 1.0,
                                      rate, C, 'a.u.', 1.0, None, 0, False, {})
            datas = arr.get_obj().reshape((-1, C)).copy()
            out[name + '_nblock'] = nblock
        out[name + '_args'] = json.dumps(dict(rate=rate, frames=frames,
                                              channels=C, seed=seed,
                                              max_pixel=max_pixel))
        out[name + '_sha256'] = checksum(x)
        out[name + '_times'] = times
        out[name + '_datas'] = datas
        print(name, datas.shape, len(times))
    np.savez_compressed(os.path.join(GOLDEN, 'fulltrace.npz'), **out)


def main():
    if not rh.reference_available():
        raise SystemExit('needs /root/reference')
    os.makedirs(GOLDEN, exist_ok=True)
    for name, spec in CHAIN_CASES.items():
        make_chain(name, spec)
    make_scroll()
    make_fulltrace()


if __name__ == '__main__':
    main()
