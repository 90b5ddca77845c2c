"""The numpy lane-level models of the specialised kernels agree with the oracle (CPU test)."""

import numpy as np

import kernel_models as KM
from oracle import np_oracle as O


def test_fft16_model():
    rng = np.random.default_rng(0)
    a = rng.standard_normal(16) + 1j * rng.standard_normal(16)
    np.testing.assert_allclose(KM.fft16(list(a)), np.fft.fft(a), rtol=1e-12, atol=1e-12)


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
