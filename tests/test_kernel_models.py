"""The numpy lane-level models of the specialised kernels agree with the oracle (CPU test)."""

import os

import numpy as np
import pytest

import kernel_models as KM
from oracle import np_oracle as O


def test_fft16_model():
    rng = np.random.default_rng(0)
    a = rng.standard_normal(16) + 1j * rng.standard_normal(16)
    np.testing.assert_allclose(KM.fft16(list(a)), np.fft.fft(a), rtol=1e-12, atol=1e-12)


def test_fft16_fma_form_model():
    """Round 2: twiddles folded into the butterflies' multiply-adds (fft16.cuh layer2) -- same transform."""
    rng = np.random.default_rng(5)
    for _ in range(4):
        a = rng.standard_normal(16) + 1j * rng.standard_normal(16)
        np.testing.assert_allclose(KM.fft16_fma(list(a)), np.fft.fft(a), rtol=1e-12, atol=1e-12)
    pruned = np.concatenate([a[:13], np.zeros(3)])            # the first pass sees structural zeros for j >= 13
    np.testing.assert_allclose(KM.fft16_fma(list(pruned)), np.fft.fft(pruned), rtol=1e-12, atol=1e-12)


def test_stftn_shared_memory_fft_model_matches_oracle():
    """Round 2: fft_length 1024 / 2048 (stftn.cu) -- pass structure, padded positions, digit-reversed split."""
    rng = np.random.default_rng(6)
    for n, L in ((1024, 1024), (1024, 400), (2048, 2048)):
        x = rng.standard_normal(L + 7 * 160)
        fr = O.window(O.frame(x, L, 160), n)
        want = O.fftr(fr, n)
        got = np.stack([KM.stftn_frame_model(f) for f in fr[:3]])
        np.testing.assert_allclose(got, want[:3], rtol=1e-10, atol=1e-10)


def test_stft512_lane_model_matches_oracle():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(1200)
    fr = O.window(O.frame(x, 400, 80), 512)
    want = O.fftr(fr, 512)
    got = np.stack([KM.stft512_frame_model(f) for f in fr])
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-11)


def test_lpc_wave_model_matches_oracle():
    rng = np.random.default_rng(2)
    fr = O.window(O.frame(rng.standard_normal(2000), 400, 80), None)
    want = O.lpc(fr, 24, eps=1e-5)
    got = np.stack([KM.lpc_wave_model(f, 24, 1e-5) for f in fr])
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("L,P,n", [(2048, 441, 2048), (1500, 300, 2048), (2048, 512, 2048), (700, 1201, 2048)])
def test_stftn_prefetched_span_reaches_both_frames(L, P, n):
    """Every sample of frame A (span[j]) and of frame B (span[P + j]) is written exactly once; the rest is padding."""
    A, B = KM.stftn_prefetch_scatter_model(L, P, n)
    assert (A[:L] == np.arange(L)).all() and (A[L:] == -1).all()
    assert (B[:L] == np.arange(L) + P).all() and (B[L:] == -1).all()


def test_sweep_tool_spec_parser():
    from tools.sweep_knobs import parse_spec
    wl, settings = parse_spec("lpc:LPC_V=0,7+LPC_W2=12,16")
    assert wl == "lpc" and len(settings) == 4 and settings[-1] == {"LPC_V": 7, "LPC_W2": 16}


def test_lpc_lagpair_lag_sums_match_the_direct_sums():
    """Two accumulator sets on aligned sample pairs (no odd-aligned pair, scalar lag 0 for odd samples) give every
    lag exactly once."""
    rng = np.random.default_rng(31)
    fr = O.window(O.frame(rng.standard_normal(1000), 400, 80), None)
    for f in fr[2:6]:
        want = np.array([np.dot(f[:400 - k], f[k:]) for k in range(25)])
        np.testing.assert_allclose(KM.lpc_lagpair_autocorr_model(f), want, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("M", [1, 5, 11, 12, 13, 24])
def test_rolled_levinson_model_matches_oracle(M):
    """The order-independent form of the recursion (reversed predictor) equals the reference's Levinson-Durbin."""
    rng = np.random.default_rng(20 + M)
    fr = O.window(O.frame(rng.standard_normal(1200), 400, 80), None)
    want = O.lpc(fr, M, eps=1e-5)
    for f, w in zip(fr, want):
        r = np.array([np.dot(f[:400 - k], f[k:]) for k in range(M + 1)])
        np.testing.assert_allclose(KM.levinson_rolled_model(r, M, 1e-5), w, rtol=1e-8, atol=1e-10)


def test_istft512_lane_model_matches_oracle():
    rng = np.random.default_rng(3)
    Y = rng.standard_normal((4, 257)) + 1j * rng.standard_normal((4, 257))
    want = O.ifftr(Y)
    got = np.stack([KM.istft512_frame_model(y) for y in Y])
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-12)


def test_stft512_bwd_lane_model_matches_the_adjoint():
    rng = np.random.default_rng(4)
    xw = rng.standard_normal(512)
    xw[400:] = 0
    g = rng.standard_normal(257)
    G = 2 * g * np.fft.rfft(xw)
    kk = np.arange(257)
    want = np.array([np.real(np.sum(np.conj(G) * np.exp(-2j * np.pi * j * kk / 512))) for j in range(512)])
    np.testing.assert_allclose(KM.stft512_bwd_frame_model(xw, g), want, rtol=1e-9, atol=1e-9)
    # and it is the gradient: finite differences of sum(g |X|^2)
    f = lambda v: float(np.sum(g * np.abs(np.fft.rfft(v)) ** 2))  # noqa: E731
    for j in (0, 7, 399):
        e = np.zeros(512)
        e[j] = 1e-6
        assert abs((f(xw + e) - f(xw - e)) / 2e-6 - want[j]) < 1e-4 * max(1.0, abs(want[j]))


def test_overlap_add_incremental_indexing():
    for L, P, N, s, T_out, tile in ((400, 80, 30, 200, 2300, 2160), (100, 50, 7, 0, 399, 1500), (6, 4, 9, 3, 35, 8),
                                    (512, 128, 5, 256, 600, 3584), (398, 7, 11, 0, 400, 64)):
        for q, na, nb, j in KM.overlap_add_ranges_model(L, P, N, s, T_out, tile):
            a_ref = 0 if q - L + 1 <= 0 else (q - L + P) // P
            assert (na, nb) == (a_ref, min(q // P, N - 1)) and j == q - na * P


def test_stft512_pair2_reads_the_right_samples():
    # frame f + 2 starts 2 P = 160 floats = 5 columns of 32 floats after frame f: same lane, column j + 5
    for read_a, want_a, read_b, want_b in KM.stft512_pair2_columns():
        assert read_a == want_a and read_b == want_b


def test_stft512_staged_stores_are_bank_conflict_free():
    for d in (1, 2):
        for banks in KM.stft512_staged_store_banks(d):
            assert len(set(banks)) == 32, (d, sorted(banks))


def test_lsp_model_matches_reference():
    """The numerical recipe of lsp.cu (unit-circle zeros by Chebyshev series + bisection) against the reference's
    companion-matrix eigenvalues (golden vectors, float64)."""
    import helpers as H
    for name in ("lpc2lsp_m1_o0", "lpc2lsp_m7_o0", "lpc2lsp_m8_o0", "lpc2lsp_speech_m24", "lpc2lsp_noise_m24"):
        op, params, ins, outs = H.load_case(name, "f64")
        a = ins[0].reshape(-1, ins[0].shape[-1])
        want = outs[0].reshape(-1, outs[0].shape[-1])
        for r in range(0, a.shape[0], max(1, a.shape[0] // 12)):
            got = KM.lsp_model(a[r])
            assert np.allclose(got, want[r, 1:], rtol=1e-9, atol=1e-11), (name, r, np.abs(got - want[r, 1:]).max())


def test_gc2gc_model_matches_oracle():
    """Direct-transform recipe of gc2gc.cu against the oracle's FFT route, even and odd n_fft."""
    rng = np.random.default_rng(3)
    for n, M1, M2, g1, g2 in ((64, 6, 9, 0.0, -0.5), (64, 6, 4, -0.5, 0.0), (63, 5, 7, -1.0, -0.3), (32, 8, 8, 0.2, 0.1),
                              (128, 12, 64, 0.0, 0.0)):
        c1 = 0.2 * rng.standard_normal(M1 + 1)
        c1[0] = 0.7
        want = O._gc2gc(c1[None], M2, g1, g2, n)[0]
        got = KM.gc2gc_model(c1, M2, g1, g2, n)
        assert np.allclose(got, want, rtol=1e-10, atol=1e-12), (n, M1, M2, g1, g2, np.abs(got - want).max())
