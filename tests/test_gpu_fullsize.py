"""Size-independent properties at the full size of BASELINE configs[1] (8 ch x 48 kHz, the
80-s buffer: 3 840 000 frames, 246 MB) and at config-5 width (64 ch x 250 kHz), device resident:
streamed == one-shot filtering, chunk-invariant spectrogram frames, min/max of min/max,
idempotence of the clamp, linearity of the filter."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from audian_b200.synth import synth
from oracle import oracle as orc


@pytest.fixture(scope='module')
def big():
    import torch
    from audian_b200 import device
    fs, C, n = 48000., 8, 3840000
    x = device.synth(0, n, C, fs, 0xA0D1A9 + 2)
    torch.cuda.synchronize()
    return fs, C, n, x


def test_filter_streamed_equals_one_shot_and_is_linear(big):
    import torch
    from audian_b200 import device
    fs, C, n, x = big
    sos = orc.filter_design(fs, 1000., 15000., 2)
    y = device.sosfilt(sos, x, 0)
    parts = []
    z = None
    edges = [0, 1, 4097, 500000, 1234567, 2000000, 3839999, n]
    for a, b in zip(edges[:-1], edges[1:]):
        p, z = device.sosfilt(sos, x[a:b], 0, zi=z, want_zf=True)
        parts.append(p)
    ys = torch.cat(parts)
    assert float((ys - y).abs().max()) <= 1e-12
    # linearity: filt(2x + x_shifted) == 2 filt(x) + filt(x_shifted)
    x2 = torch.roll(x, 12345, 0)
    lhs = device.sosfilt(sos, 2.0*x + x2, 0)
    rhs = 2.0*y + device.sosfilt(sos, x2, 0)
    assert float((lhs - rhs).abs().max()) <= 1e-12
    # a window in the middle against the CPU oracle (the prefix only matters through its state)
    a, m = 3000000, 200000
    hx = x[a - 50000:a + m].cpu().numpy()
    ref = np.empty_like(hx)
    orc.filter_process(sos, hx, ref, 0)
    assert np.max(np.abs(y[a:a + m].cpu().numpy() - ref[50000:])) <= 1e-9     # decayed prefix


def test_spectrogram_frames_do_not_depend_on_the_window(big):
    import torch
    from audian_b200 import device
    fs, C, n, x = big
    nfft, hop = 1024, 512
    nf = (n - (nfft - hop))//hop
    full, got = device.spectrogram(x, fs, nfft, hop, nf)
    assert got == nf == 7499
    for k0, k1 in ((0, 13), (3000, 3517), (7400, 7499)):
        sub, g2 = device.spectrogram(x[k0*hop:(k1 - 1)*hop + nfft], fs, nfft, hop, k1 - k0)
        assert g2 == k1 - k0
        assert torch.equal(sub, full[k0:k1])             # frame-local arithmetic: bit identical
    ref = np.empty((40, C, nfft//2 + 1))
    k0 = 5000
    orc.spectrogram_process(x[k0*hop:(k0 + 39)*hop + nfft].cpu().numpy(), ref, fs, nfft, hop)
    assert np.allclose(full[k0:k0 + 40].cpu().numpy(), ref, rtol=1e-5, atol=1e-20*ref.max())
    # Parseval: sum of the PSD over bins * fs / nfft == mean square of the windowed frame
    # (checked on one frame as a scale test)
    fr = x[k0*hop:k0*hop + nfft, 3].cpu().numpy()
    w = 0.5 - 0.5*np.cos(2*np.pi*np.arange(nfft)/nfft)
    seg = (fr - fr.mean())*w
    assert np.isclose(full[k0, 3].cpu().numpy().sum()*fs/nfft, np.sum(seg**2)/np.sum(w**2), rtol=1e-9)


def test_minmax_of_minmax_and_envelope_clamp(big):
    import torch
    from audian_b200 import device
    fs, C, n, x = big
    fine = device.minmax(x, 1920)                         # 2000 segments
    coarse = device.minmax(x, 3840)
    mn = torch.minimum(fine[0::4], fine[2::4])
    mx = torch.maximum(fine[1::4], fine[3::4])
    assert torch.equal(coarse[0::2], mn) and torch.equal(coarse[1::2], mx)
    whole = device.minmax(x, n)
    assert torch.equal(whole[0], x.min(dim=0).values) and torch.equal(whole[1], x.max(dim=0).values)
    esos = orc.envelope_design(fs, 500.)
    env = device.envelope(esos, x, 0, True)
    raw = device.envelope(esos, x, 0, False)
    assert float(env.min()) >= 0.0
    # (two runs of the scan agree to a few ulp, not bit for bit: which predecessor record a
    # tile's look-back finds published -- aggregate or inclusive state -- depends on timing and
    # changes the association of the sum)
    assert float((env - torch.clamp(raw, min=0.0)).abs().max()) <= 1e-13
    # the envelope of |x| is the envelope of x (rectification), and scales linearly
    assert float((device.envelope(esos, x.abs(), 0, True) - env).abs().max()) <= 1e-13
    assert float((device.envelope(esos, 3.0*x, 0, True) - 3.0*env).abs().max()) <= 1e-12


def test_array_width_64_channels():
    import torch
    from audian_b200 import device
    fs, C, n = 250000., 64, 2000000                       # 8 s of config 5: 1 GB
    x = device.synth(0, n, C, fs, 0xA0D1A9 + 5)
    sos = orc.filter_design(fs, 5000., 60000., 2)
    y = device.sosfilt(sos, x, 0)
    m = 60000
    hx = x[:m].cpu().numpy()
    ref = np.empty_like(hx)
    orc.filter_process(sos, hx, ref, 0)
    assert np.max(np.abs(y[:m].cpu().numpy() - ref)) <= 1e-9
    for nfft, hop in ((256, 64), (1024, 512), (4096, 4096), (16384, 2048)):
        nf = 24
        sp, got = device.spectrogram(y[:(nf - 1)*hop + nfft], fs, nfft, hop, nf)
        assert got == nf
        rs = np.empty((nf, C, nfft//2 + 1))
        orc.spectrogram_process(y[:(nf - 1)*hop + nfft].cpu().numpy(), rs, fs, nfft, hop)
        assert np.allclose(sp.cpu().numpy(), rs, rtol=1e-5, atol=1e-20*rs.max()), (nfft, hop)
    rows = device.minmax(x, 25000)
    assert np.array_equal(rows[:8].cpu().numpy().view(np.uint64),
                          orc.minmax_rows(x[:100000].cpu().numpy(), 25000).view(np.uint64))
