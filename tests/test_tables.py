"""Host tables: consistent with the oracle's float64 formulas everywhere, and bit-identical to the
reference's own buffers whenever /root/reference is importable (build container)."""

import numpy as np
import pytest
import torch

from diffsptk_b200 import tables as TB
from oracle import np_oracle as O
from oracle.ref_shim import load_reference


def test_window_tables_vs_oracle():
    for kind in ("blackman", "hamming", "hanning", "bartlett", "trapezoidal", "rectangular", "nuttall", "povey",
                 "sine", "vorbis", "kbd"):
        for norm in ("none", "power", "magnitude"):
            for sym in (True, False):
                if kind == "kbd" and not sym:
                    continue
                w = TB.make_window(400, kind, norm, sym, dtype=torch.float64).numpy()
                np.testing.assert_allclose(w, O.window_table(400, kind, norm, sym), rtol=1e-9, atol=1e-12)


def test_matrices_vs_oracle():
    np.testing.assert_allclose(TB.make_freqt_matrix(256, 24, 0.42, dtype=torch.float64).numpy(),
                               O.freqt_matrix(256, 24, 0.42), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(TB.make_coef_freqt_matrix(256, 48, 0.42, dtype=torch.float64).numpy(),
                               O.coef_freqt_matrix(256, 48, 0.42), rtol=1e-12, atol=1e-15)
    for sc in ("htk", "mel", "bark", "linear"):
        for erb in (None, 1.0):
            np.testing.assert_allclose(
                TB.make_fbank_matrix(512, 40, 16000, scale=sc, erb_factor=erb, dtype=torch.float64).numpy(),
                O.fbank_matrix(512, 40, 16000, scale=sc, erb_factor=erb), rtol=1e-12, atol=1e-15)
    for t in (1, 2, 3, 4):
        np.testing.assert_allclose(TB.make_dct_matrix(40, t, dtype=torch.float64).numpy(), O.dct_matrix(40, t),
                                   rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(TB.make_lifter(13, 22, dtype=torch.float64).numpy(), O.lifter_vector(13, 22),
                               rtol=1e-14)


def test_column_support():
    H = TB.make_fbank_matrix(512, 40, 16000)
    cb, ce = TB.column_support(H)
    assert int((H != 0).sum()) == 493 == int((ce - cb).sum())
    Hn = H.numpy()
    for c in range(40):
        nz = np.nonzero(Hn[:, c])[0]
        assert cb[c] == nz.min() and ce[c] == nz.max() + 1
    z = torch.zeros(9, 3)
    cb, ce = TB.column_support(z)
    assert torch.equal(cb, ce)


def test_fused_mcep_tables_reproduce_the_reference_algorithm():
    """Model of the CUDA mcep kernel in float64 numpy (folded matrices, unpivoted elimination)
    against the oracle, which follows mcep.py:189-224 step by step with FFTs and LAPACK."""
    P0, G, Hm = (t.numpy() for t in TB.make_mcep_tables(512, 24, 0.42, dtype=torch.float64))
    rng = np.random.default_rng(0)
    x = O.stft(rng.standard_normal((2, 2000)))
    av = (-0.42) ** np.arange(25)
    logx = np.log(x)
    mc = logx @ P0
    i25 = np.arange(25)
    for _ in range(10):
        e = np.exp(logx - 2 * (mc @ G))
        rt = e @ Hm
        A = rt[..., np.abs(i25[:, None] - i25[None, :])] + rt[..., i25[:, None] + i25[None, :]]
        b = rt[..., :25] - av
        for p in range(24):
            f = A[..., p + 1:, p] / A[..., p:p + 1, p]
            A[..., p + 1:, :] -= f[..., None] * A[..., p:p + 1, :]
            b[..., p + 1:] -= f * b[..., p:p + 1]
        g = np.zeros_like(b)
        for i in range(24, -1, -1):
            g[..., i] = (b[..., i] - (A[..., i, i + 1:] * g[..., i + 1:]).sum(-1)) / A[..., i, i]
        mc = mc + g
    np.testing.assert_allclose(mc, O.mcep(x, 24, 0.42, 10), rtol=1e-9, atol=1e-11)


def test_levinson_recursion_equals_regularised_toeplitz_solve():
    """Model of the CUDA Levinson kernel (r0 + eps) against the oracle's dense solve (levdur.py:113-127)."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((16, 400)) * np.blackman(400)
    r = np.stack([np.correlate(v, v, "full")[399:424] for v in x])
    eps = 1e-5
    out = np.zeros_like(r)
    for n, rr in enumerate(r):
        a = np.zeros(25)
        E = rr[0] + eps
        for i in range(1, 25):
            k = -(rr[i] + np.dot(a[1:i], rr[i - 1:0:-1])) / E
            a[1:i] = a[1:i] + k * a[i - 1:0:-1]
            a[i] = k
            E *= 1 - k * k
        out[n, 0] = np.sqrt(rr[0] + np.dot(rr[1:], a[1:]))
        out[n, 1:] = a[1:]
    np.testing.assert_allclose(out, O.levdur(r, eps), rtol=1e-9, atol=1e-12)


@pytest.mark.skipif(load_reference() is None, reason="reference not present (GPU box)")
def test_tables_bit_identical_to_reference():
    import diffsptk_b200 as B
    D = load_reference()
    for dt in (torch.float32, torch.float64):
        for w in ("blackman", "hamming", "hanning", "bartlett", "trapezoidal", "rectangular", "nuttall", "povey",
                  "sine", "vorbis", 0, 3, 6):
            for norm in ("none", "power", "magnitude"):
                for sym in (True, False):
                    for L in (5, 400):
                        assert torch.equal(D.Window(L, window=w, norm=norm, symmetric=sym, dtype=dt).window,
                                           B.Window(L, window=w, norm=norm, symmetric=sym, dtype=dt).window)
        r = D.MelCepstralAnalysis(fft_length=512, cep_order=24, alpha=0.42, n_iter=1, dtype=dt)
        m = B.MelCepstralAnalysis(fft_length=512, cep_order=24, alpha=0.42, n_iter=1, dtype=dt)
        for n in ("freqt.A", "ifreqt.A", "rfreqt.A", "alpha_vector"):
            assert torch.equal(dict(r.named_buffers())[n], dict(m.named_buffers())[n]), n
        r = D.MFCC(fft_length=512, mfcc_order=13, n_channel=40, sample_rate=16000, lifter=22, dtype=dt)
        m = B.MFCC(fft_length=512, mfcc_order=13, n_channel=40, sample_rate=16000, lifter=22, dtype=dt)
        for n in ("liftering_vector", "fbank.H", "dct.W"):
            assert torch.equal(dict(r.named_buffers())[n], dict(m.named_buffers())[n]), n
        assert torch.equal(D.LevinsonDurbin(24, dtype=dt).eye, B.LevinsonDurbin(24, dtype=dt).eye)
