"""Device-pointer entry points on torch CUDA tensors (audian_b200.device) and the
single-rank path of the sharded drivers."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from audian_b200 import _lib
from audian_b200.synth import synth
from oracle import oracle as orc


def test_device_synth_is_bit_identical_to_host():
    import torch
    from audian_b200 import device
    for t0, n, C, rate, seed in [(0, 5000, 1, 44100., 1), (123456789, 4097, 8, 48000., 0xA0D1A9),
                                 (2**33 + 5, 3000, 64, 250000., 7), (10, 1000, 3, 500000., 99)]:
        d = device.synth(t0, n, C, rate, seed)
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy(), synth(t0, n, C, rate, seed))


def test_device_ops_match_oracle():
    import torch
    from audian_b200 import device
    fs, C, n = 96000., 4, 200000
    hx = synth(0, n, C, fs, 5)
    x = torch.from_numpy(hx).cuda()
    sos = orc.filter_design(fs, 1000., 15000., 4)
    y, zf = device.sosfilt(sos, x, 0, want_zf=True)
    ref = np.empty((n, C))
    orc.filter_process(sos, hx, ref, 0)
    assert np.max(np.abs(y.cpu().numpy() - ref)) <= 1e-6
    # state-only pass == zf of the full pass
    z2 = device.sosfilt(sos, x, 0, state_only=True)
    assert torch.equal(zf, z2)
    # carried state: second half from zf of the first half
    h = n//2
    y1, z1 = device.sosfilt(sos, x[:h].contiguous(), 0, want_zf=True)
    y2 = device.sosfilt(sos, x[h:].contiguous(), 0, zi=z1)
    assert np.max(np.abs(torch.cat([y1, y2]).cpu().numpy() - ref)) <= 1e-6
    sp, ns = device.spectrogram(y, fs, 512, 128, (n - 384)//128)
    rs = np.empty((ns, C, 257))
    assert orc.spectrogram_process(ref, rs, fs, 512, 128) == ns
    assert np.allclose(sp.cpu().numpy(), rs, rtol=1e-5, atol=1e-20*rs.max())
    esos = orc.envelope_design(fs, 300.)
    e = device.envelope(esos, y, 0, True)
    re = np.empty((n, C))
    orc.envelope_process(esos, ref, re, 0, 0)
    assert np.max(np.abs(e.cpu().numpy() - re)) <= 1e-6
    mm = device.minmax(x, 1000)
    assert np.array_equal(mm.cpu().numpy().view(np.uint64), orc.minmax_rows(hx, 1000).view(np.uint64))
    db = device.decibel(sp)
    assert np.allclose(db.cpu().numpy(), orc.decibel(sp.cpu().numpy()), atol=1e-9)


def test_side_stream_ordering():
    import torch
    from audian_b200 import device
    fs, C, n = 48000., 8, 500000
    x = device.synth(0, n, C, fs)
    sos = orc.filter_design(fs, 1000., 15000., 2)
    ref = device.sosfilt(sos, x)
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        y = device.sosfilt(sos, x)
        z = y*2.0                      # torch op on the same stream sees the kernel's output
    st.synchronize()
    assert torch.equal(z, ref*2.0)


def test_wholefile_streaming_on_device():
    """Chunked whole-file passes (audian_b200.wholefile) with the recording generated on the
    device chunk by chunk: independent of the chunking, equal to one pass."""
    import torch
    from audian_b200 import device
    from audian_b200.wholefile import WholeFile
    frames, C, rate, seed = 300017, 4, 96000., 21
    hx = synth(0, frames, C, rate, seed)

    def source(t0, n):
        return device.synth(t0, n, C, rate, seed)
    wf = WholeFile(source, frames, C, rate, chunk_frames=37000)
    step = 50
    sos = orc.filter_design(rate, 1000., 15000., 4)
    parts = []
    rows = wf.fulltrace_and_filter(sos, step, lambda t0, y: parts.append(y.cpu().numpy()))
    assert np.array_equal(rows.cpu().numpy().view(np.uint64), orc.minmax_rows(hx, step).view(np.uint64))
    yref = np.empty_like(hx)
    orc.filter_process(sos, hx, yref, 0)
    assert np.max(np.abs(np.concatenate(parts) - yref)) <= 1e-6
    frames_out = []
    nf = wf.spectrogram(1024, 512, lambda k, P: frames_out.append(P.cpu().numpy()))
    sref = np.empty((nf, C, 513))
    assert orc.spectrogram_process(hx, sref, rate, 1024, 512) == nf
    assert np.allclose(np.concatenate(frames_out), sref, rtol=1e-5, atol=1e-20*sref.max())
