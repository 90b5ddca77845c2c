"""Generate golden input/output vectors by running the REAL reference.

Run in the build container only (``/root/reference`` must exist)::

    python tests/golden/make_golden.py

Writes ``tests/golden/golden_f32.npz``, ``golden_f64.npz`` and
``manifest.json``.  Every case is ``diffsptk.functional.<op>(*inputs, **params)``
executed by the unmodified reference on CPU; the shapes mirror the reference's
own test table (tests/test_frame.py:23-49,
test_window.py:23-78, test_stft.py:24-68, test_fftr.py:24-70, test_spec.py:23-91,
test_acorr.py:23-46, test_levdur.py:23-45, test_lpc.py:23-44, test_freqt.py:23-50,
test_mcep.py:23-54, test_fbank.py:25-106, test_mfcc.py:23-71, test_dct.py:24-58)
plus the BASELINE.json parameters (fl=400, fp=80, n_fft=512, M=24, alpha=0.42,
40 mel / 13 cep) at a few hundred frames.

The reference cannot travel to the GPU box, which is why these vectors are
committed.  ``cases()`` is also imported by the tests to rebuild the inputs.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def _rng(seed):
    return np.random.default_rng(seed)


def speechlike(T, seed=7, sr=16000):
    """A voiced/unvoiced/silence test signal (synthetic; no reference asset is copied).

    Harmonic source through two resonances, a noise burst, and exact-zero
    silence -- this produces ill-conditioned autocorrelation matrices and
    frames that hit the eps / floor paths, like real speech does.
    """
    rng = _rng(seed)
    t = np.arange(T) / sr
    f0 = 120 + 30 * np.sin(2 * np.pi * 2.0 * t)
    phase = 2 * np.pi * np.cumsum(f0) / sr
    src = sum(np.sin(k * phase) / k for k in range(1, 20))
    y = np.zeros(T)
    a1, a2 = 2 * 0.97 * np.cos(2 * np.pi * 700 / sr), -0.97 ** 2
    for n in range(T):
        y[n] = src[n] + (a1 * y[n - 1] if n > 0 else 0) + (a2 * y[n - 2] if n > 1 else 0)
    y = 0.1 * y / np.max(np.abs(y))
    seg = T // 4
    y[:seg // 2] = 0.0                                   # leading silence (exact zeros)
    y[2 * seg:3 * seg] = 0.02 * rng.standard_normal(seg)  # unvoiced burst
    y[3 * seg + seg // 2:] = 0.0                         # trailing silence
    return y


def cases():
    """Yield (name, op, params, inputs_float64).  Deterministic."""
    r = _rng(1234)
    ramp20 = np.arange(20, dtype=np.float64)
    # ---- frame (tests/test_frame.py:23-49) -------------------------------------------
    for fl in (1, 2, 3, 4, 5):
        for fp in (1, 2, 3, 5):
            for center in (True, False):
                yield (f"frame_ramp_l{fl}_p{fp}_c{int(center)}", "frame",
                       dict(frame_length=fl, frame_period=fp, center=center), [ramp20])
    x_small = r.standard_normal((3, 1000))
    for mode in ("constant", "reflect", "replicate", "circular"):
        for center in (True, False):
            yield (f"frame_400_80_{mode}_c{int(center)}", "frame",
                   dict(frame_length=400, frame_period=80, center=center, mode=mode), [x_small])
    yield ("frame_zmean", "frame", dict(frame_length=400, frame_period=80, zmean=True), [x_small])
    yield ("frame_2048_441", "frame", dict(frame_length=2048, frame_period=441), [r.standard_normal((2, 5000))])
    for T in (1, 79, 80, 81, 399, 400, 401):
        yield (f"frame_T{T}", "frame", dict(frame_length=400, frame_period=80), [r.standard_normal((2, T))])
    # ---- window (tests/test_window.py:23-78) -----------------------------------------
    step10 = np.ones((2, 10))
    for w in (0, 1, 2, 3, 4, 5, 6, "povey", "sine", "vorbis", "kbd"):
        for norm in (0, 1, 2):
            for sym in (True, False):
                if w == "kbd" and not sym:
                    continue
                for L1, L2 in ((8, 10), (10, 10), (10, None)):
                    yield (f"window_{w}_n{norm}_s{int(sym)}_{L1}_{L2}", "window",
                           dict(out_length=L2, window=w, norm=norm, symmetric=sym), [step10[:, :L1]])
    fr = r.standard_normal((4, 7, 400))
    yield ("window_400_512", "window", dict(out_length=512), [fr])
    yield ("window_400_none", "window", dict(out_length=None, window="hamming", norm="none"), [fr])
    # ---- fftr (tests/test_fftr.py:24-70) ---------------------------------------------
    x13 = r.standard_normal((3, 13))
    for of in ("complex", "real", "imaginary", "amplitude", "power"):
        yield (f"fftr_16_{of}", "fftr", dict(fft_length=16, out_format=of), [x13])
    yield ("fftr_512", "fftr", dict(fft_length=512), [r.standard_normal((5, 512))])
    yield ("fftr_none_len24", "fftr", dict(fft_length=None, out_format="power"), [r.standard_normal((2, 24))])
    yield ("fftr_trunc", "fftr", dict(fft_length=8, out_format="real"), [r.standard_normal((2, 13))])
    yield ("fftr_424", "fftr", dict(fft_length=424), [r.standard_normal((2, 400))])
    # ---- spec (tests/test_spec.py:23-91) ---------------------------------------------
    b = r.standard_normal((3, 5)); a = r.standard_normal((3, 4)); a[:, 0] = np.abs(a[:, 0]) + 0.5
    for of in ("db", "log-magnitude", "magnitude", "power"):
        for rf in (None, -40):
            yield (f"spec_{of}_rf{rf}", "spec",
                   dict(fft_length=16, eps=0.01, relative_floor=rf, out_format=of), [b, a])
    yield ("spec_b_only", "spec", dict(fft_length=16, eps=0.01), [b, None])
    yield ("spec_a_only", "spec", dict(fft_length=16, eps=0.01), [None, a])
    yield ("spec_512", "spec", dict(fft_length=512, eps=1e-9), [r.standard_normal((6, 400)), None])
    # ---- stft (tests/test_stft.py:24-68) ---------------------------------------------
    x100 = r.standard_normal((2, 100))
    for of in ("power", "complex", "db", "log-magnitude", "magnitude"):
        yield (f"stft_small_{of}", "stft",
               dict(frame_length=12, frame_period=10, fft_length=16, window="hamming", norm="power",
                    eps=1e-6, out_format=of), [x100])
    x_b = r.standard_normal((3, 8000))
    yield ("stft_baseline", "stft", dict(), [x_b])
    yield ("stft_baseline_complex", "stft", dict(out_format="complex"), [x_b[:1, :3000]])
    yield ("stft_hann_db", "stft", dict(window="hanning", norm="none", out_format="db"), [x_b[:1, :3000]])
    yield ("stft_rf", "stft", dict(relative_floor=-30.0), [x_b[:2, :2000]])
    yield ("stft_zmean_reflect", "stft", dict(zmean=True, mode="reflect"), [x_b[:2, :2000]])
    yield ("stft_nocenter", "stft", dict(center=False, window="rectangular", eps=0.0, out_format="complex"),
           [x_b[:2, :2000]])
    yield ("stft_1024", "stft", dict(frame_length=800, frame_period=160, fft_length=1024), [x_b[:2, :4000]])
    yield ("stft_256", "stft", dict(frame_length=200, frame_period=40, fft_length=256), [x_b[:2, :2000]])
    yield ("stft_speech", "stft", dict(), [speechlike(8000)])
    yield ("stft_silence", "stft", dict(), [np.zeros((1, 800))])
    # ---- acorr / levdur / lpc (tests/test_acorr.py:23-46, test_levdur.py, test_lpc.py)
    x14 = r.standard_normal((4, 14))
    for M in (12, 13):
        for of in ("naive", "normalized", "biased", "unbiased"):
            yield (f"acorr_14_m{M}_{of}", "acorr", dict(acr_order=M, out_format=of), [x14])
    wfr = fr * np.blackman(400)
    yield ("acorr_400_24", "acorr", dict(acr_order=24), [wfr])
    r52 = r.standard_normal((5, 52))
    # levdur needs a valid autocorrelation: take it from noise (computed in float64 here)
    ac30 = np.stack([np.correlate(v, v, "full")[51:51 + 31] for v in r52])
    yield ("levdur_30", "levdur", dict(), [ac30])
    yield ("levdur_30_eps", "levdur", dict(eps=1e-3), [ac30])
    ac24 = np.stack([np.correlate(v, v, "full")[399:399 + 25] for v in wfr.reshape(-1, 400)[:8]])
    yield ("levdur_24", "levdur", dict(), [ac24])
    yield ("lpc_30_14", "lpc", dict(lpc_order=14), [r.standard_normal((3, 30))])
    yield ("lpc_400_24", "lpc", dict(lpc_order=24), [wfr])
    sp = speechlike(8000)
    spf = np.stack([sp[i * 80:i * 80 + 400] for i in range(20, 80)]) * np.blackman(400)
    yield ("lpc_speech_24_eps", "lpc", dict(lpc_order=24, eps=1e-5), [spf])
    # ---- freqt (tests/test_freqt.py:23-50) -------------------------------------------
    c20 = r.standard_normal((4, 20))
    yield ("freqt_19_29", "freqt", dict(out_order=29, alpha=0.1), [c20])
    yield ("freqt_24_24_042", "freqt", dict(out_order=24, alpha=0.42), [r.standard_normal((6, 25))])
    yield ("freqt_256_24", "freqt", dict(out_order=24, alpha=0.42), [r.standard_normal((3, 257)) * 0.1])
    yield ("freqt_0_0", "freqt", dict(out_order=0, alpha=0.3), [r.standard_normal((3, 1))])
    # ---- mcep (tests/test_mcep.py:23-54) ---------------------------------------------
    P32 = np.square(np.abs(np.fft.rfft(r.standard_normal((4, 32)), axis=-1))) + 1e-3
    for M in (0, 7, 8):
        for it in (0, 3):
            yield (f"mcep_32_m{M}_i{it}", "mcep", dict(cep_order=M, alpha=0.1, n_iter=it), [P32])
    import oracle.np_oracle as O  # inputs only (power spectra), not outputs
    Pb = O.stft(x_b[:2, :4000])
    yield ("mcep_baseline", "mcep", dict(cep_order=24, alpha=0.42, n_iter=10), [Pb])
    yield ("mcep_speech", "mcep", dict(cep_order=24, alpha=0.42, n_iter=10), [O.stft(sp)])
    # ---- fbank / mfcc / dct ----------------------------------------------------------
    for of in ("y", "yE"):
        yield (f"fbank_32_{of}", "fbank",
               dict(n_channel=10, sample_rate=8000, f_min=300, f_max=3400, floor=1.0, out_format=of), [P32 * 50])
    for sc in ("htk", "mel", "bark", "linear"):
        yield (f"fbank_512_{sc}", "fbank", dict(n_channel=40, sample_rate=16000, scale=sc, out_format="yE"), [Pb])
    yield ("fbank_512_erb", "fbank", dict(n_channel=40, sample_rate=16000, erb_factor=1.0), [Pb])
    yield ("fbank_512_gamma_pow", "fbank", dict(n_channel=40, sample_rate=16000, gamma=-0.5, use_power=True), [Pb])
    for of in ("y", "yE", "yc", "ycE"):
        yield (f"mfcc_32_{of}", "mfcc",
               dict(mfcc_order=4, n_channel=10, sample_rate=8000, lifter=20, f_min=300, f_max=3400,
                    floor=1.0, out_format=of), [P32 * 50])
    yield ("mfcc_baseline", "mfcc", dict(mfcc_order=13, n_channel=40, sample_rate=16000), [Pb])
    yield ("mfcc_baseline_ycE_l22", "mfcc",
           dict(mfcc_order=13, n_channel=40, sample_rate=16000, lifter=22, out_format="ycE"), [Pb])
    x8 = r.standard_normal((3, 8))
    for t in (1, 2, 3, 4):
        yield (f"dct_8_t{t}", "dct", dict(dct_type=t), [x8])
    yield ("dct_40", "dct", dict(), [r.standard_normal((5, 40))])
    # ---- inverse path, SURVEY.md section 8(f) rank 2 (tests/test_ifftr.py, test_unframe.py, test_istft.py) ----
    ri = _rng(4321)

    def cplx(*shape):
        return ri.standard_normal(shape) + 1j * ri.standard_normal(shape)
    y9 = cplx(3, 9)
    for ol in (None, 16, 5):
        yield (f"ifftr_16_o{ol}", "ifftr", dict(out_length=ol), [y9])
    yield ("ifftr_512_400", "ifftr", dict(out_length=400), [cplx(2, 3, 257)])
    yield ("ifftr_24", "ifftr", dict(out_length=None), [cplx(2, 13)])
    yield ("ifftr_2", "ifftr", dict(out_length=1), [cplx(2, 2)])
    ramp_frames = np.array([[0, 0, 1, 2, 3], [1, 2, 3, 4, 5], [3, 4, 5, 6, 7], [5, 6, 7, 8, 9], [7, 8, 9, 0, 0]],
                           dtype=np.float64)
    yield ("unframe_ramp", "unframe", dict(out_length=9, frame_period=2), [ramp_frames])
    fr12 = ri.standard_normal((3, 7, 12))
    for ol in (None, 20):
        yield (f"unframe_12_5_o{ol}", "unframe", dict(out_length=ol, frame_period=5), [fr12])
        yield (f"unframe_12_3_nocenter_hamming_o{ol}", "unframe",
               dict(out_length=ol, frame_period=3, center=False, window="hamming", norm="power"), [fr12])
    yield ("unframe_12_12", "unframe", dict(out_length=None, frame_period=12), [fr12])
    yield ("unframe_400_80_blackman", "unframe", dict(out_length=2000, frame_period=80, window="blackman",
                                                      norm="power"), [ri.standard_normal((2, 2, 26, 400))])
    Yb = cplx(2, 13, 257)
    for ol in (None, 1000, 777):
        yield (f"istft_baseline_o{ol}", "istft", dict(out_length=ol), [Yb])
    yield ("istft_small", "istft", dict(out_length=30, frame_length=12, frame_period=5, fft_length=16), [cplx(3, 7, 9)])
    yield ("istft_nonpow2", "istft", dict(out_length=None, frame_length=30, frame_period=7, fft_length=48,
                                          window="hanning", norm="magnitude"), [cplx(2, 9, 25)])
    yield ("istft_nocenter", "istft", dict(out_length=None, frame_length=40, frame_period=10, fft_length=64,
                                           center=False, window="hamming", norm="none", symmetric=False),
           [cplx(2, 3, 12, 33)])
    yield ("istft_one_frame", "istft", dict(out_length=None), [cplx(1, 1, 257)])
    # ---- fftcep, section 8(f) rank 3 (tests/test_fftcep.py) --------------------------------------------------
    P9 = ri.standard_normal((2, 9)) ** 2 + 0.1
    for M in (3, 8):
        for it in (0, 3):
            yield (f"fftcep_16_m{M}_i{it}", "fftcep", dict(cep_order=M, accel=0.5 if it else 0.0, n_iter=it), [P9])
    P257 = ri.standard_normal((2, 5, 257)) ** 2 + 1e-3
    yield ("fftcep_512_m24", "fftcep", dict(cep_order=24, accel=0.0, n_iter=0), [P257])
    # ---- delta, section 8(f) rank 4 (tests/test_delta.py) ---------------------------------------------------
    xd = ri.standard_normal((2, 9, 3))
    for i, (seed, so) in enumerate(([[[-0.5, 0, 0.5]], True], [[[-0.5, 0, 0.5], [1, -2, 1]], True],
                                    [[[1, -1]], False], [[2], True], [[3, 2], False], [[1, 1], True])):
        yield (f"delta_seed{i}", "delta", dict(seed=seed, static_out=so), [xd])
    yield ("delta_2d", "delta", dict(seed=[[-0.5, 0, 0.5], [1, -2, 1]], static_out=True), [ri.standard_normal((7, 4))])
    yield ("delta_one_frame", "delta", dict(seed=[2, 2], static_out=True), [ri.standard_normal((3, 1, 5))])
    yield ("delta_mfcc_like", "delta", dict(seed=[2, 2], static_out=True), [ri.standard_normal((4, 200, 13))])
    yield ("fftcep_512_m24_i5", "fftcep", dict(cep_order=24, accel=1.0, n_iter=5), [P257])
    # ---- per-row converters, section 8(f) rank 4 (tests/test_lpc2par.py, test_par2lpc.py, test_gnorm.py,
    #      test_ignorm.py, test_norm0.py, test_mc2b.py, test_b2mc.py) ------------------------------------------
    rc = _rng(99)
    # stable LPC rows: step-up recursion from random PARCOR coefficients in (-0.9, 0.9), gain > 0
    def lpc_rows(shape, M):
        k = rc.uniform(-0.9, 0.9, (*shape, M))
        a = np.zeros((*shape, M))
        for m in range(M):
            km = k[..., m:m + 1]
            prev = a[..., :m].copy()
            a[..., :m] = prev + km * prev[..., ::-1]
            a[..., m] = km[..., 0]
        return np.concatenate([rc.uniform(0.5, 2.0, (*shape, 1)), a], axis=-1)
    for M in (0, 1, 2, 7, 24):
        yield (f"lpc2par_m{M}", "lpc2par", dict(gamma=1, c=None), [lpc_rows((3, 5), M)])
        yield (f"par2lpc_m{M}", "par2lpc", dict(gamma=1, c=None),
               [np.concatenate([rc.uniform(0.5, 2.0, (3, 5, 1)), rc.uniform(-0.9, 0.9, (3, 5, M))], axis=-1)])
    yield ("lpc2par_g05", "lpc2par", dict(gamma=0.5, c=None), [lpc_rows((40,), 12)])
    yield ("lpc2par_c2", "lpc2par", dict(gamma=1, c=2), [lpc_rows((40,), 12)])
    yield ("par2lpc_g05", "par2lpc", dict(gamma=0.5, c=None),
           [np.concatenate([rc.uniform(0.5, 2.0, (40, 1)), rc.uniform(-0.9, 0.9, (40, 12))], axis=-1)])
    yield ("par2lpc_c3", "par2lpc", dict(gamma=1, c=3),
           [np.concatenate([rc.uniform(0.5, 2.0, (40, 1)), rc.uniform(-0.9, 0.9, (40, 12))], axis=-1)])
    yield ("lpc2par_many", "lpc2par", dict(gamma=1, c=None), [lpc_rows((700,), 24)])
    cep = 0.3 * rc.standard_normal((3, 50, 25))
    cep[..., 0] = rc.uniform(0.1, 1.5, (3, 50))
    for g, c in ((0, None), (0.5, None), (-0.5, None), (-1, None), (1, None), (0, 2), (0, 4)):
        tag = f"g{g}_c{c}".replace("-", "m").replace(".", "")
        yield (f"gnorm_{tag}", "gnorm", dict(gamma=g, c=c), [cep])
        gain = np.concatenate([rc.uniform(0.3, 3.0, (3, 50, 1)), cep[..., 1:]], axis=-1)
        yield (f"ignorm_{tag}", "ignorm", dict(gamma=g, c=c), [gain])
    yield ("gnorm_m0", "gnorm", dict(gamma=-0.5, c=None), [rc.uniform(0.1, 1.0, (7, 1))])
    yield ("norm0_m24", "norm0", dict(), [lpc_rows((3, 50), 24)])
    yield ("norm0_m0", "norm0", dict(), [rc.uniform(0.5, 2.0, (9, 1))])
    for M in (0, 1, 4, 24):
        for alpha in (0.0, 0.42, -0.3):
            tag = f"m{M}_a{alpha}".replace("-", "m").replace(".", "")
            yield (f"mc2b_{tag}", "mc2b", dict(alpha=alpha), [rc.standard_normal((2, 6, M + 1))])
            yield (f"b2mc_{tag}", "b2mc", dict(alpha=alpha), [rc.standard_normal((2, 6, M + 1))])
    # ---- mgc2mgc / mgc2sp / plp, section 8(f) rank 3 (tests/test_mgc2mgc.py:23-100, test_mgc2sp.py:23-51,
    #      test_plp.py:23-70) -- inputs are valid (normalised / multiplied) cepstra made from a small random one
    rm = _rng(2024)
    base = 0.2 * rm.standard_normal((2, 6, 5))
    base[..., 0] = rm.uniform(0.2, 1.0, (2, 6))

    def shaped(c, g, norm, mul):
        """A gamma-g cepstrum in the (normalised, multiplied) form the flags announce."""
        c = c.copy()
        if norm:
            z = 1 + g * c[..., :1]
            c = np.concatenate([z ** (1 / g), c[..., 1:] / z], axis=-1)
        if mul:
            c = np.concatenate([c[..., :1], c[..., 1:] * g], axis=-1)
            if not norm:
                c = np.concatenate([c[..., :1] * g + 1, c[..., 1:]], axis=-1)
        return c
    i = 0
    for (M, A, G) in ((4, 0, 0.1), (4, 0, -0.1), (2, 0.1, 0.1), (6, 0.1, 0.2)):
        for in_norm in (False, True):
            for out_norm in (False, True):
                for in_mul in (False, True):
                    for out_mul in (False, True):
                        i += 1
                        yield (f"mgc2mgc_{i:02d}", "mgc2mgc",
                               dict(out_order=M, in_alpha=0, out_alpha=A, in_gamma=0.1, out_gamma=G, in_norm=in_norm,
                                    out_norm=out_norm, in_mul=in_mul, out_mul=out_mul, n_fft=256),
                               [shaped(base, 0.1, in_norm, in_mul)])
    mc25 = 0.1 * rm.standard_normal((3, 40, 25))
    mc25[..., 0] = rm.uniform(-2.0, 1.0, (3, 40))
    yield ("mgc2mgc_mc2c", "mgc2mgc", dict(out_order=40, in_alpha=0.42, out_alpha=0, n_fft=512), [mc25])
    yield ("mgc2mgc_to_lsp_domain", "mgc2mgc", dict(out_order=24, in_alpha=0.42, out_alpha=0.42, in_gamma=0,
                                                      out_gamma=-0.5, n_fft=512), [mc25])
    yield ("mgc2mgc_lpc2c", "mgc2mgc", dict(out_order=12, in_gamma=-1, out_gamma=0, in_norm=True, in_mul=True,
                                              n_fft=512), [shaped(0.5 * base, -1.0, True, True)])
    c8 = 0.3 * rm.standard_normal((2, 8))
    for fmt in (0, 1, 2, 3, 4, 5, 6, "complex"):
        yield (f"mgc2sp_16_o{fmt}", "mgc2sp", dict(fft_length=16, out_format=fmt), [c8])
    yield ("mgc2sp_mcep_512", "mgc2sp", dict(fft_length=512, alpha=0.42), [mc25])
    yield ("mgc2sp_mgc_512", "mgc2sp", dict(fft_length=512, alpha=0.42, gamma=-0.5, n_fft=1024),
           [np.concatenate([mc25[..., :1] * 0.2, mc25[..., 1:]], axis=-1)])
    yield ("mgc2sp_lpc_like", "mgc2sp", dict(fft_length=64, gamma=-1, norm=True, mul=True, n_fft=256, out_format=1),
           [shaped(0.5 * base, -1.0, True, True)])
    P17 = rm.standard_normal((2, 4, 32))
    P17 = np.abs(np.fft.rfft(P17, axis=-1)) ** 2
    for fmt in (0, 1, 2, 3):
        yield (f"plp_32_o{fmt}", "plp", dict(plp_order=4, n_channel=10, sample_rate=8000, compression_factor=0.3,
                                              lifter=20, floor=1, out_format=fmt), [P17])
    Pw = np.abs(np.fft.rfft(rm.standard_normal((2, 30, 400)) * np.hanning(400), n=512, axis=-1)) ** 2 + 1e-9
    yield ("plp_512_m12", "plp", dict(plp_order=12, n_channel=40, sample_rate=16000), [Pw])
    yield ("plp_512_bark", "plp", dict(plp_order=8, n_channel=20, sample_rate=16000, f_min=100, f_max=7000,
                                        scale="bark", lifter=22, out_format="ycE"), [Pw])
    # ---- mgcep, section 8(f) rank 3 (tests/test_mgcep.py:23-48); MODULE-ONLY in the reference: no functional ----
    P17g = np.abs(np.fft.rfft(rm.standard_normal((2, 3, 32)), axis=-1)) ** 2 + 1e-6
    for it in (0, 3):
        for gm in (0, -0.5, -1):
            tag = f"g{gm}_i{it}".replace("-", "m").replace(".", "")
            yield (f"mgcep_32_{tag}", "mgcep", dict(fft_length=32, cep_order=8, alpha=0.1, gamma=gm, n_iter=it),
                   [P17g])
    yield ("mgcep_512_m24", "mgcep", dict(fft_length=512, cep_order=24, alpha=0.42, gamma=-0.5, n_iter=5), [Pw[:, :8]])
    yield ("mgcep_512_c3", "mgcep", dict(fft_length=512, cep_order=12, alpha=0.42, gamma=0, c=3, n_iter=2), [Pw[:, :8]])
    # ---- lpc2lsp, section 8(f) rank 4 (tests/test_lpc2lsp.py:23-55): LPC rows obtained from windowed noise and
    #      from the speech-like signal by the oracle-independent normal equations below --------------------------------
    def lpc_of(frames, M):
        L = frames.shape[-1]
        r = np.stack([(frames[..., : L - k] * frames[..., k:]).sum(-1) for k in range(M + 1)], axis=-1)
        if M == 0:
            return np.sqrt(r)
        idx = np.abs(np.arange(M)[:, None] - np.arange(M)[None, :])
        a = np.linalg.solve(r[..., :-1][..., idx] + 1e-9 * np.eye(M), -r[..., 1:, None])[..., 0]
        return np.concatenate([np.sqrt(r[..., :1] + (r[..., 1:] * a).sum(-1, keepdims=True)), a], axis=-1)
    noise = rm.standard_normal((2, 6, 32))
    for M in (0, 1, 7, 8):
        for fmt in (0, 1, 2, 3):
            yield (f"lpc2lsp_m{M}_o{fmt}", "lpc2lsp", dict(log_gain=True, sample_rate=8000, out_format=fmt),
                   [lpc_of(noise, M)])
    sp = speechlike(16000)
    spf = np.lib.stride_tricks.sliding_window_view(sp, 400)[4000:12000:80] * np.hanning(400)
    spf = spf[np.abs(spf).max(-1) > 1e-4]
    yield ("lpc2lsp_speech_m24", "lpc2lsp", dict(), [lpc_of(spf, 24)])
    yield ("lpc2lsp_speech_m12", "lpc2lsp", dict(log_gain=True, sample_rate=16000, out_format="khz"), [lpc_of(spf, 12)])
    yield ("lpc2lsp_noise_m24", "lpc2lsp", dict(), [lpc_of(rm.standard_normal((50, 400)) * np.hanning(400), 24)])


MODULE_ONLY = {"mgcep": "MelGeneralizedCepstralAnalysis"}   # ops the reference exposes as nn.Module only


def main():
    sys.path.insert(0, ROOT)
    import torch
    from oracle.ref_shim import load_reference

    D = load_reference()
    if D is None:
        raise SystemExit("reference not importable: golden vectors can only be made in the build container")
    F = D.functional
    manifest = {}
    for dt_name, tdt, ndt in (("f32", torch.float32, np.float32), ("f64", torch.float64, np.float64)):
        store = {}
        for name, op, params, inputs in cases():
            cdt = np.complex64 if ndt is np.float32 else np.complex128
            cast = lambda v: v.astype(cdt if np.iscomplexobj(v) else ndt)  # noqa: E731
            tin = [None if v is None else torch.from_numpy(np.ascontiguousarray(cast(v))) for v in inputs]
            with torch.no_grad():
                if op in MODULE_ONLY:
                    out = getattr(D, MODULE_ONLY[op])(**params, dtype=tdt)(*tin)
                else:
                    out = getattr(F, op)(*tin, **params)
            outs = out if isinstance(out, tuple) else (out,)
            for i, v in enumerate(inputs):
                if v is not None:
                    store[f"{name}/in{i}"] = cast(v)
            for i, o in enumerate(outs):
                store[f"{name}/out{i}"] = o.numpy()
            manifest[name] = dict(op=op, params=params, n_in=len(inputs),
                                  none_in=[i for i, v in enumerate(inputs) if v is None], n_out=len(outs))
        np.savez_compressed(os.path.join(HERE, f"golden_{dt_name}.npz"), **store)
        print(dt_name, len(store), "arrays")
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("cases:", len(manifest))


if __name__ == "__main__":
    main()
