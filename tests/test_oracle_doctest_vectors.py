"""The 13 known answers stored in the reference's own docstrings, restated as
explicit vectors and checked against the oracle.  Inputs follow the doctests: ``ramp(a, b)`` is
``arange(a, b+1)``, ``ramp(n)`` is ``arange(0, n+1)``, ``step(n)`` is ``ones(n+1)``."""

import numpy as np

from oracle import np_oracle as O


def ramp(a, b=None):
    return np.arange(0, a + 1, dtype=np.float64) if b is None else np.arange(a, b + 1, dtype=np.float64)


def close(a, b, tol=5e-4):
    assert np.allclose(np.asarray(a), np.asarray(b), atol=tol, rtol=0), (a, b)


def test_frame_doctest():  # diffsptk/modules/frame.py:75-84
    y = O.frame(ramp(1, 9), 5, 2)
    assert np.array_equal(y, [[0, 0, 1, 2, 3], [1, 2, 3, 4, 5], [3, 4, 5, 6, 7], [5, 6, 7, 8, 9], [7, 8, 9, 0, 0]])


def test_window_doctest():  # window.py:97-102
    close(O.window(np.ones(5), 7, window="hamming", norm="none"), [0.08, 0.54, 1.0, 0.54, 0.08, 0, 0])


def test_stft_doctest():  # stft.py:124-133
    y = O.stft(ramp(1, 3), frame_length=3, frame_period=1, fft_length=8)
    close(y, np.outer([1.0, 4.0, 9.0], np.ones(5)))


def test_fftr_doctest():  # fftr.py:78-83
    close(O.fftr(ramp(1, 3), 8, "real"), [6.0, 2.4142, -2.0, -0.4142, 2.0])


def test_spec_doctest():  # spec.py:82-89
    close(O.spec(ramp(1, 3), fft_length=8), [36.0, 25.3137, 8.0, 2.6863, 4.0])


def test_acorr_doctest():  # acorr.py:67-72
    close(O.acorr(ramp(4), 3), [30.0, 20.0, 11.0, 4.0])


def test_levdur_lpc_doctest():  # levdur.py:74-80, lpc.py:80-85 (float32 default eps = 1e-5)
    x = (ramp(1, 5) * 0.1).astype(np.float32)
    close(O.levdur(O.acorr(x, 2)), [0.5054, -0.8140, 0.1193])
    close(O.lpc(x, 2), [0.5054, -0.8140, 0.1193])


def test_freqt_doctest():  # freqt.py:82-93
    c2 = O.freqt(ramp(1, 4), 4, 0.02)
    close(c2, [1.0412, 2.1240, 3.1949, 3.8666, -0.2358])
    close(O.freqt(c2, 3, -0.02), [1.0, 2.0, 3.0, 4.0])


def test_mcep_doctest():  # mcep.py:93-102
    X = O.stft(ramp(19), frame_length=10, frame_period=10, fft_length=16)
    close(O.mcep(X, 3, 0.1, 1), [[-0.885, 0.792, -0.174, 0.017], [-0.352, 4.422, -1.088, -0.051]], 2e-3)


def test_fbank_doctest():  # fbank.py:145-154
    X = O.stft(ramp(19), frame_length=10, frame_period=10, fft_length=32)
    close(O.fbank(X, 4, 8000), [[0.1214, 0.4825, 0.6072, 0.3589], [3.3640, 3.4518, 2.7717, 0.5088]])


def test_mfcc_doctest():  # mfcc.py:134-143
    X = O.stft(ramp(19), frame_length=10, frame_period=10, fft_length=32)
    close(O.mfcc(X, 4, 8, 8000), [[-0.340, -0.658, -0.021, -0.194], [4.553, -3.210, 1.319, -0.930]], 2e-3)


def test_dct_doctest():  # dct.py:73-78
    close(O.dct(ramp(3)), [3.0, -2.2304, 0.0, -0.1585])
