"""numpy restatement of the reference's frame-rate analysis path (CPU oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Parity status: pinned
(doctest known answers + reference-generated golden vectors).

Every function follows the *algorithm the reference uses* (FFT-based
autocorrelation, dense Toeplitz solve, matrix frequency transform, ...), not
the algorithm the CUDA kernels use, so a kernel bug and an oracle bug cannot
cancel.  All ``file:line`` citations are relative to ``/root/reference``.

Arithmetic runs in the dtype of the input array (float32 or float64), like the
reference.  Host tables (window, warping matrices, filter bank, DCT basis,
lifter) are built in float64 and cast, as the reference does for every table
except the window, which the reference builds directly in the module dtype
(``diffsptk/modules/window.py:134-183``); the resulting <= 2e-8 table
difference is far inside the stated tolerances.
"""

from __future__ import annotations

import math

import numpy as np

__all__ = [
    "frame", "window_table", "window", "fftr", "spec", "stft", "acorr", "levdur",
    "lpc", "freqt_matrix", "freqt", "coef_freqt_matrix", "mcep", "fbank_matrix",
    "fbank", "dct_matrix", "dct", "mfcc", "lifter_vector",
]

_PAD_MODE = {"constant": "constant", "reflect": "reflect", "replicate": "edge", "circular": "wrap"}


def _as_float(x):
    x = np.asarray(x)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float32)
    return x


# ----------------------------------------------------------------------------- frame
def frame(x, frame_length=400, frame_period=80, center=True, zmean=False, mode="constant"):
    """Overlapping frames of a waveform.  diffsptk/modules/frame.py:120-141.

    ``y[..., i, j] = xpad[..., i*P + j]`` with ``xpad = pad(x, (L//2, (L-1)//2))``
    when centred, ``(0, L-1)`` otherwise (frame.py:130-137); frame count is the
    ``unfold`` count ``(T-1)//P + 1`` (frame.py:138).
    """
    if frame_length <= 0:
        raise ValueError("frame_length must be positive.")
    if frame_period <= 0:
        raise ValueError("frame_period must be positive.")
    x = np.asarray(x)
    L, P = frame_length, frame_period
    left, right = (L // 2, (L - 1) // 2) if center else (0, L - 1)
    widths = [(0, 0)] * (x.ndim - 1) + [(left, right)]
    xp = np.pad(x, widths, mode=_PAD_MODE[mode])
    n = (xp.shape[-1] - L) // P + 1
    idx = (np.arange(n) * P)[:, None] + np.arange(L)[None, :]
    y = xp[..., idx]
    if zmean:
        y = y - y.mean(-1, keepdims=True)
    return y


# ----------------------------------------------------------------------------- window
def _cos_sum(L, periodic, coefs):
    n_den = L if periodic else L - 1
    if L == 1:
        return np.ones(1)
    n = np.arange(L, dtype=np.float64)
    w = np.zeros(L)
    for k, c in enumerate(coefs):
        w += c * np.cos(2.0 * math.pi * k * n / n_den)
    return w


def _sine_window(L, symmetric):
    # torch.signal.windows.cosine: sin(pi * (n + 0.5) / M), M = L (sym) or L + 1.
    if L == 1:
        return np.ones(1)
    M = L if symmetric else L + 1
    n = np.arange(L, dtype=np.float64)
    return np.sin(math.pi * (n + 0.5) / M)


def window_table(in_length, window="blackman", norm="power", symmetric=True, dtype=np.float64):
    """The window table.  diffsptk/modules/window.py:122-183."""
    L = in_length
    if L <= 0:
        raise ValueError("in_length must be positive.")
    periodic = not symmetric
    if window in (0, "blackman"):
        w = _cos_sum(L, periodic, (0.42, -0.5, 0.08))
    elif window in (1, "hamming"):
        w = _cos_sum(L, periodic, (0.54, -0.46))
    elif window in (2, "hanning"):
        w = _cos_sum(L, periodic, (0.5, -0.5))
    elif window in (3, "bartlett", 4, "trapezoidal"):
        if L == 1:
            w = np.ones(1)
        else:
            den = L if periodic else L - 1
            w = 1.0 - np.abs(2.0 * np.arange(L, dtype=np.float64) / den - 1.0)
        if window in (4, "trapezoidal"):
            w = np.minimum(2.0 * w, 1.0)
    elif window in (5, "rectangular"):
        w = np.ones(L)
    elif window in (6, "nuttall"):
        size = L if periodic else L - 1
        c1 = np.array([0.355768, -0.487396, 0.144232, -0.012604])
        c2 = np.arange(0, 8, 2, dtype=np.float64) * (math.pi / size)
        w = (c1 * np.cos(np.outer(np.arange(L, dtype=np.float64), c2))).sum(1)
    elif window == "povey":
        w = _cos_sum(L, periodic, (0.5, -0.5)) ** 0.85
    elif window == "sine":
        w = _sine_window(L, symmetric)
    elif window == "vorbis":
        w = np.sin(0.5 * math.pi * _sine_window(L, symmetric) ** 2)
    elif window == "kbd":
        if periodic:
            raise ValueError("periodic is not supported for kbd window.")
        seed = np.kaiser(L // 2 + 1, 12.0)
        cs = np.cumsum(seed)
        half = np.sqrt(cs[:-1] / cs[-1])
        w = np.concatenate([half, half[::-1]])
    else:
        raise ValueError(f"window {window} is not supported.")
    if norm in (0, "none"):
        pass
    elif norm in (1, "power"):
        w = w / math.sqrt(float((w * w).sum()))
    elif norm in (2, "magnitude"):
        w = w / float(w.sum())
    else:
        raise ValueError(f"norm {norm} is not supported.")
    return w.astype(dtype)


def window(x, out_length=None, *, window="blackman", norm="power", symmetric=True, table=None):
    """Multiply by the window, then right zero-pad.  window.py:185-193."""
    x = _as_float(x)
    L = x.shape[-1]
    w = window_table(L, window, norm, symmetric, x.dtype) if table is None else np.asarray(table)
    y = x * w
    if out_length is not None:
        if out_length >= L:
            widths = [(0, 0)] * (y.ndim - 1) + [(0, out_length - L)]
            y = np.pad(y, widths)
        else:  # F.pad with a negative width truncates
            y = y[..., :out_length]
    return y


# ----------------------------------------------------------------------------- fftr / spec / stft
def fftr(x, fft_length=None, out_format="complex"):
    """Real FFT + output formatter.  diffsptk/modules/fftr.py:110-121,136-151."""
    if fft_length is not None and (fft_length <= 0 or fft_length % 2 == 1):
        raise ValueError("fft_length must be positive even.")
    x = _as_float(x)
    y = np.fft.rfft(x, n=fft_length, axis=-1)
    if out_format in (0, "complex"):
        return y
    if out_format in (1, "real"):
        return y.real
    if out_format in (2, "imaginary"):
        return y.imag
    if out_format in (3, "amplitude"):
        return np.abs(y)
    if out_format in (4, "power"):
        return np.square(np.abs(y))
    raise ValueError(f"out_format {out_format} is not supported.")


def _remove_gain(a):
    # diffsptk/utils/private.py:200-209
    K = a[..., :1]
    a = np.concatenate([np.ones_like(K), a[..., 1:]], axis=-1)
    return K, a


def spec(b=None, a=None, *, fft_length=512, eps=0.0, relative_floor=None, out_format="power"):
    """Power spectrum of b / a.  diffsptk/modules/spec.py:152-178."""
    if fft_length <= 1:
        raise ValueError("fft_length must be greater than 1.")
    if eps < 0:
        raise ValueError("eps must be non-negative.")
    if relative_floor is not None and 0 <= relative_floor:
        raise ValueError("relative_floor must be negative.")
    if b is not None and a is not None:
        K, a = _remove_gain(_as_float(a))
        X = K * (fftr(b, fft_length, "amplitude") / fftr(a, fft_length, "amplitude"))
    elif b is not None:
        X = fftr(b, fft_length, "amplitude")
    elif a is not None:
        K, a = _remove_gain(_as_float(a))
        X = K / fftr(a, fft_length, "amplitude")
    else:
        raise ValueError("Either b or a must be specified.")
    dt = X.dtype
    s = np.square(X) + dt.type(eps)
    if relative_floor is not None:
        rf = 10 ** (relative_floor / 10)  # spec.py:121-122
        s = np.maximum(s, s.max(-1, keepdims=True) * dt.type(rf))
    if out_format in (0, "db"):
        return (10 * np.log10(s)).astype(dt)
    if out_format in (1, "log-magnitude"):
        return (0.5 * np.log(s)).astype(dt)
    if out_format in (2, "magnitude"):
        return np.sqrt(s)
    if out_format in (3, "power"):
        return s
    raise ValueError(f"out_format {out_format} is not supported.")


def stft(x, *, frame_length=400, frame_period=80, fft_length=512, center=True, zmean=False,
         mode="constant", window="blackman", norm="power", symmetric=True, eps=1e-9,
         relative_floor=None, out_format="power", window_table_override=None):
    """spec(window(frame(x))).  diffsptk/modules/stft.py:158-241 (defaults :92-100)."""
    x = _as_float(x)
    f = frame(x, frame_length, frame_period, center, zmean, mode)
    g = globals()["window"](f, fft_length, window=window, norm=norm, symmetric=symmetric,
                            table=window_table_override)
    if out_format == "complex":
        return fftr(g, fft_length, "complex")
    return spec(g, fft_length=fft_length, eps=eps, relative_floor=relative_floor, out_format=out_format)


# ----------------------------------------------------------------------------- delta
# regression features over the frame axis.
def delta_window(seed, static_out=True, dtype=np.float64):
    """Regression windows (H, W).  diffsptk/modules/delta.py:98-170."""
    if not isinstance(seed, (tuple, list)):
        raise ValueError("seed must be tuple or list.")
    if isinstance(seed[0], (tuple, list)):
        rows = ([[1.0]] if static_out else []) + [list(map(float, c)) for c in seed]
        max_len = max(len(c) for c in rows)
        max_len += 1 if max_len % 2 == 0 else 0
        out = []
        for c in rows:
            diff = max_len - len(c)
            left = diff // 2 if diff % 2 == 0 else (diff - 1) // 2
            out.append(np.array([0.0] * left + c + [0.0] * (diff - left)))
    else:
        if min(seed) <= 0:
            raise ValueError("The width of regression coefficients must be positive.")
        if len(seed) >= 3:
            raise ValueError("3rd order regression is not supported.")
        max_len = max(seed) * 2 + 1
        out = []
        if static_out:
            w = np.zeros(max_len)
            w[(max_len - 1) // 2] = 1
            out.append(w)
        n = seed[0]
        j = np.arange(-n, n + 1, dtype=np.float64)
        p = (max_len - (2 * n + 1)) // 2
        out.append(np.pad(j * (1 / (n * (n + 1) * (2 * n + 1) / 3)), (p, p)))
        if len(seed) >= 2:
            n = seed[1]
            a0 = 2 * n + 1
            a1 = a0 * n * (n + 1) / 3
            a2 = a1 * (3 * n * n + 3 * n - 1) / 5
            j = np.arange(-n, n + 1, dtype=np.float64)
            p = (max_len - (2 * n + 1)) // 2
            out.append(np.pad((a0 * j * j - a1) * (1 / (2 * (a2 * a0 - a1 * a1))), (p, p)))
    return np.stack(out).astype(dtype)


def delta(x, seed=[[-0.5, 0, 0.5]], static_out=True):  # noqa: B006
    """y[b, t, h D + d] = sum_w window[h, w] x[b, clamp(t + w - (W-1)/2), d].  delta.py:172-194."""
    x = _as_float(x)
    if x.ndim not in (2, 3):
        raise ValueError("Input must be 2D or 3D tensor.")
    win = delta_window(seed, static_out, x.dtype)
    Hn, W = win.shape
    T, D = x.shape[-2], x.shape[-1]
    pad = (W - 1) // 2
    idx = np.clip(np.arange(T)[:, None] + np.arange(W)[None, :] - pad, 0, T - 1)      # [T, W]
    taps = x[..., idx, :]                                                            # [..., T, W, D]
    y = np.einsum("hw,...twd->...thd", win, taps).astype(x.dtype)
    return y.reshape(*x.shape[:-1], Hn * D)


# ----------------------------------------------------------------------------- fftcep
# a direct consumer of the STFT power spectrum.
def fftcep(x, cep_order, accel=0.0, n_iter=0):
    """Cepstral analysis by the improved cepstral method.  diffsptk/modules/fftcep.py:94-136."""
    x = _as_float(x)
    H = x.shape[-1]
    n = 2 * (H - 1)
    if n <= 1:
        raise ValueError("fft_length must be greater than 1.")
    if cep_order < 0:
        raise ValueError("cep_order must be non-negative.")
    if n < 2 * cep_order:
        raise ValueError("cep_order must be less than or equal to fft_length // 2.")
    if accel < 0:
        raise ValueError("accel must be non-negative.")
    if n_iter < 0:
        raise ValueError("n_iter must be non-negative.")
    N = cep_order + 1
    dt = x.dtype
    e = np.fft.irfft(np.log(x), axis=-1).astype(dt)
    v = e[..., :N].copy()
    pad = [(0, 0)] * (x.ndim - 1)
    e = np.pad(e[..., N:H], pad + [(N, 0)])
    for _ in range(n_iter):
        e = np.fft.hfft(e, axis=-1).astype(dt)
        e[e < 0] = 0
        e = np.fft.ihfft(e, axis=-1).real.astype(dt)
        t = e[..., :N] * (1 + accel)
        v += t
        e -= np.pad(t, pad + [(0, H - N)])
    idx = [0, N - 1] if H == N else [0]
    v[..., idx] *= 0.5
    return v


# ----------------------------------------------------------------------------- ifftr / unframe / istft
# the inverse of the hot path.
def ifftr(y, out_length=None):
    """Inverse real FFT, truncated to ``out_length``.  diffsptk/modules/ifftr.py:103-108,130-143.

    x[n] = (1/N) (Re Y[0] + (-1)^n Re Y[N/2] + 2 sum_{0<k<N/2} Re(Y[k] e^{+2 pi i k n / N})), N = 2 (K - 1):
    the imaginary parts of the DC and Nyquist bins do not contribute (torch.fft.irfft / numpy agree).
    """
    y = np.asarray(y)
    n = 2 * (y.shape[-1] - 1)
    if n <= 0 or n % 2 == 1:
        raise ValueError("fft_length must be positive even.")
    if out_length is not None and (out_length <= 0 or n < out_length):
        raise ValueError("out_length must be in [1, fft_length].")
    dt = np.float32 if y.dtype == np.complex64 else np.float64
    x = np.fft.irfft(y.astype(np.complex128), n=n, axis=-1)[..., :out_length]
    return x.astype(dt)


def unframe(y, out_length=None, *, frame_period=80, center=True, window="rectangular", norm="none",
            symmetric=True, table=None):
    """Windowed overlap-add with sum-of-squares normalisation.  diffsptk/modules/unframe.py:128-211.

    out[t] = sum_n y[n, j] w[j] / (sum_n w[j]^2 + 1e-16), j = t + s - n P, s = L // 2 if center else 0;
    default length N P when centred, the whole folded span otherwise.
    """
    y = _as_float(y)
    if y.ndim <= 1:
        raise ValueError("Input must be at least 2D tensor.")
    N, L = y.shape[-2], y.shape[-1]
    if L <= 0:
        raise ValueError("frame_length must be positive.")
    if L < frame_period:
        raise ValueError("frame_period must be less than or equal to frame_length.")
    w = window_table(L, window, norm, symmetric, y.dtype) if table is None else np.asarray(table, dtype=y.dtype)
    span = (N - 1) * frame_period + L
    num = np.zeros(y.shape[:-2] + (span,), dtype=y.dtype)
    den = np.zeros(span, dtype=y.dtype)
    for n in range(N):   # F.fold sums the overlapping columns in frame order
        num[..., n * frame_period:n * frame_period + L] += y[..., n, :] * w
        den[n * frame_period:n * frame_period + L] += w * w
    x = num / (den + np.asarray(1e-16, dtype=y.dtype))
    if out_length is None and center:
        out_length = N * frame_period
    s = L // 2 if center else 0
    e = None if out_length is None else s + out_length
    return x[..., s:e]


def istft(y, *, out_length=None, frame_length=400, frame_period=80, fft_length=512, center=True,
          window="blackman", norm="power", symmetric=True, table=None):
    """unframe(ifftr(y)[..., :frame_length]).  diffsptk/modules/istft.py:146-193."""
    fr = ifftr(y, frame_length)
    if fr.shape[-1] != frame_length or 2 * (np.asarray(y).shape[-1] - 1) != fft_length:
        raise ValueError("dimension of spectrum does not match fft_length")
    return unframe(fr, out_length, frame_period=frame_period, center=center, window=window, norm=norm,
                   symmetric=symmetric, table=table)


# ----------------------------------------------------------------------------- acorr / levdur / lpc
def acorr(x, acr_order, out_format="naive"):
    """FFT-based autocorrelation.  diffsptk/modules/acorr.py:95-120."""
    x = _as_float(x)
    L = x.shape[-1]
    if L <= 0:
        raise ValueError("frame_length must be positive.")
    if L <= acr_order:
        raise ValueError("acr_order must be less than frame_length.")
    n = L + acr_order
    n += n % 2
    X = np.square(np.abs(np.fft.rfft(x, n=n, axis=-1)))
    r = np.fft.irfft(X, axis=-1)[..., : acr_order + 1].astype(x.dtype)
    if out_format in (0, "naive"):
        return r
    if out_format in (1, "normalized"):
        return r / r[..., :1]
    if out_format in (2, "biased"):
        return r / x.dtype.type(L)
    if out_format in (3, "unbiased"):
        return (r / np.arange(L, L - acr_order - 1, -1)).astype(x.dtype)
    raise ValueError(f"out_format {out_format} is not supported.")


def _toeplitz(r):
    # diffsptk/utils/private.py:291-295: R[i, j] = r[|i - j|]
    d = r.shape[-1]
    idx = np.abs(np.arange(d)[:, None] - np.arange(d)[None, :])
    return r[..., idx]


def _hankel(x):
    # diffsptk/utils/private.py:298-302: Q[i, j] = x[i + j], n = (d + 1) // 2
    n = (x.shape[-1] + 1) // 2
    idx = np.arange(n)[:, None] + np.arange(n)[None, :]
    return x[..., idx]


def levdur(r, eps=None):
    """Yule-Walker solve by a dense Toeplitz system.  diffsptk/modules/levdur.py:98-127."""
    r = _as_float(r)
    M = r.shape[-1] - 1
    if eps is None:
        eps = 1e-5 if r.dtype == np.float32 else 0.0  # levdur.py:108-110
    if eps < 0:
        raise ValueError("eps must be non-negative.")
    r0, r1 = r[..., :1], r[..., 1:]
    if M == 0:
        return np.sqrt(r0)
    R = _toeplitz(r[..., :-1]) + (np.eye(M, dtype=r.dtype) * r.dtype.type(eps))
    a = np.linalg.solve(R, -r1[..., None])[..., 0].astype(r.dtype)
    K = np.sqrt((r1 * a).sum(-1, keepdims=True) + r0)
    return np.concatenate([K, a], axis=-1)


def lpc(x, lpc_order, eps=None):
    """levdur(acorr(x)).  diffsptk/modules/lpc.py:106-139."""
    return levdur(acorr(x, lpc_order), eps)


# ----------------------------------------------------------------------------- freqt
def freqt_matrix(in_order, out_order, alpha):
    """All-pass warping matrix in float64, shape (M1+1, M2+1).  freqt.py:115-139."""
    if in_order < 0:
        raise ValueError("in_order must be non-negative.")
    if out_order < 0:
        raise ValueError("out_order must be non-negative.")
    if 1 <= abs(alpha):
        raise ValueError("alpha must be in (-1, 1).")
    L1, L2 = in_order + 1, out_order + 1
    beta = 1 - alpha * alpha
    ramp = np.arange(L1, dtype=np.float64)
    A = np.zeros((L2, L1))
    A[0, :] = alpha ** ramp
    if 1 < L2 and 1 < L1:
        A[1, 1:] = A[0, :-1] * beta * ramp[1:]
    for i in range(2, L2):
        for j in range(1, L1):
            A[i, j] = A[i - 1, j - 1] + alpha * (A[i, j - 1] - A[i - 1, j])
    return np.ascontiguousarray(A.T)


def freqt(c, out_order, alpha=0.0):
    """c @ A.  diffsptk/modules/freqt.py:141-143."""
    c = _as_float(c)
    A = freqt_matrix(c.shape[-1] - 1, out_order, alpha).astype(c.dtype)
    return c @ A


def coef_freqt_matrix(in_order, out_order, alpha):
    """mcep-internal warping of autocorrelation-like sequences.  mcep.py:264-288."""
    L1, L2 = in_order + 1, out_order + 1
    A = np.zeros((L2, L1))
    A[:, 0] = (-alpha) ** np.arange(L2, dtype=np.float64)
    for i in range(1, L2):
        for j in range(1, L1):
            A[i, j] = A[i - 1, j - 1] + alpha * (A[i, j - 1] - A[i - 1, j])
    return np.ascontiguousarray(A.T)


# ----------------------------------------------------------------------------- mcep
def mcep(x, cep_order, alpha=0.0, n_iter=0):
    """Mel-cepstral analysis of a power spectrum.  diffsptk/modules/mcep.py:189-224."""
    x = _as_float(x)
    dt = x.dtype
    H = x.shape[-1] - 1
    fft_length = 2 * H
    M = cep_order
    if fft_length <= 1:
        raise ValueError("fft_length must be greater than 1.")
    if M < 0:
        raise ValueError("cep_order must be non-negative.")
    if fft_length < 2 * M:
        raise ValueError("cep_order must be less than or equal to fft_length // 2.")
    if 1 <= abs(alpha):
        raise ValueError("alpha must be in (-1, 1).")
    if n_iter < 0:
        raise ValueError("n_iter must be non-negative.")
    A_f = freqt_matrix(H, M, alpha).astype(dt)
    A_i = freqt_matrix(M, H, -alpha).astype(dt)
    A_r = coef_freqt_matrix(H, 2 * M, alpha).astype(dt)
    alpha_vector = ((-alpha) ** np.arange(M + 1, dtype=np.float64)).astype(dt)  # mcep.py:179-181

    log_x = np.log(x)
    c = np.fft.irfft(log_x, axis=-1).astype(dt)
    c[..., 0] *= 0.5
    c[..., H] *= 0.5
    mc = c[..., : H + 1] @ A_f
    for _ in range(n_iter):
        c = mc @ A_i
        d = np.fft.rfft(c, n=fft_length, axis=-1).real.astype(dt)
        d = np.exp(log_x - d - d)
        rd = np.fft.irfft(d, axis=-1).astype(dt)
        rt = rd[..., : H + 1] @ A_r
        r = rt[..., : M + 1]
        ra = r - alpha_vector
        RQ = _toeplitz(r) + _hankel(rt)
        grad = np.linalg.solve(RQ, ra[..., None])[..., 0].astype(dt)
        mc = mc + grad
    return mc


# ----------------------------------------------------------------------------- fbank / dct / mfcc
def _hz_to_auditory(f, scale):
    # diffsptk/utils/private.py:241-274
    if scale == "htk":
        return 1127 * np.log1p(f / 700)
    if scale in ("oshaughnessy", "mel"):
        return 2595 * np.log10(1 + f / 700)
    if scale in ("chakroborty", "inverted-mel"):
        return 2195.286 - 2595 * np.log10(1 + (4031.25 - f) / 700)
    if scale in ("traunmuller", "bark"):
        return (26.81 * f) / (1960 + f) - 0.53
    if scale == "linear":
        return f
    raise ValueError(f"scale {scale} is not supported.")


def _auditory_to_hz(z, scale):
    # diffsptk/utils/private.py:277-288
    if scale == "htk":
        return 700 * np.expm1(z / 1127)
    if scale in ("oshaughnessy", "mel"):
        return 700 * (np.power(10, z / 2595) - 1)
    if scale in ("chakroborty", "inverted-mel"):
        return 4031.25 - 700 * (np.power(10, (2195.286 - z) / 2595) - 1)
    if scale in ("traunmuller", "bark"):
        return 1960 * (z + 0.53) / (26.28 - z)
    if scale == "linear":
        return z
    raise ValueError(f"scale {scale} is not supported.")


def fbank_matrix(fft_length, n_channel, sample_rate, f_min=0.0, f_max=None, scale="htk", erb_factor=None):
    """Triangular filter-bank weights H[L/2+1, C] in float64.  fbank.py:233-293."""
    if fft_length <= 1:
        raise ValueError("fft_length must be greater than 1.")
    if n_channel <= 0:
        raise ValueError("n_channel must be positive.")
    if sample_rate <= 0:
        raise ValueError("sample_rate must be positive.")
    if f_min < 0 or sample_rate / 2 <= f_min:
        raise ValueError("Invalid f_min.")
    if f_max is not None and not (f_min < f_max <= sample_rate / 2):
        raise ValueError("Invalid f_min and f_max.")
    if erb_factor is not None and erb_factor <= 0:
        raise ValueError("erb_factor must be positive.")
    if f_max is None:
        f_max = sample_rate / 2
    K = fft_length // 2 + 1
    C = n_channel
    H = np.zeros((K, C))
    if erb_factor is None:
        z_lo = _hz_to_auditory(np.asarray(f_min, dtype=np.float64), scale)
        z_hi = _hz_to_auditory(np.asarray(f_max, dtype=np.float64), scale)
        k_lo = max(1, int(f_min / sample_rate * fft_length + 1.5))
        k_hi = min(fft_length // 2, int(f_max / sample_rate * fft_length + 0.5))
        centers = (z_hi - z_lo) / (C + 1) * np.arange(1, C + 2) + z_lo
        widths = centers - np.concatenate([[z_lo], centers[:-1]])
        for k in range(k_lo, k_hi):
            z = _hz_to_auditory(np.float64(sample_rate * k / fft_length), scale)
            m = int(np.argmax(z <= centers))  # first centre at or above this bin
            w = (centers[m] - z) / widths[m]
            if 0 < m:
                H[k, m - 1] = w
            if m < C:
                H[k, m] = 1 - w
    else:
        a = erb_factor * 6.23e-6
        b = erb_factor * 93.39e-3
        c = erb_factor * 28.52

        def centre(f, first):
            s = 1 if first else -1
            a_h = s * 0.5 / (700 + f)
            b_h = s * 700 / (700 + f)
            c_h = -s * 0.5 * f * (1 + 700 / (700 + f))
            b_b = (b - b_h) / (a - a_h)
            c_b = (c - c_h) / (a - a_h)
            return 0.5 * (-b_b + np.sqrt(b_b ** 2 - 4 * c_b))

        z1 = _hz_to_auditory(centre(f_min, True), scale)
        zc = _hz_to_auditory(centre(f_max, False), scale)
        fc = _auditory_to_hz(np.linspace(z1, zc, C), scale)
        erb = a * fc ** 2 + b * fc + c
        fl = -(700 + erb) + np.sqrt(erb ** 2 + (700 + fc) ** 2)
        fh = fl + 2 * erb
        f = np.linspace(0, sample_rate / 2, K)
        for m in range(C):
            up = (fl[m] <= f) & (f < fc[m])
            H[up, m] = (f[up] - fl[m]) / (fc[m] - fl[m])
            dn = (fc[m] <= f) & (f <= fh[m])
            H[dn, m] = (fh[m] - f[dn]) / (fh[m] - fc[m])
    return H


def fbank(x, n_channel, sample_rate, f_min=0.0, f_max=None, floor=1e-5, gamma=0.0, scale="htk",
          erb_factor=None, use_power=False, out_format="y"):
    """Mel filter-bank analysis of a power spectrum.  fbank.py:305-321."""
    x = _as_float(x)
    dt = x.dtype
    if floor <= 0:
        raise ValueError("floor must be positive.")
    if 1 < abs(gamma):
        raise ValueError("gamma must be in [-1, 1].")
    fft_length = 2 * x.shape[-1] - 2
    H = fbank_matrix(fft_length, n_channel, sample_rate, f_min, f_max, scale, erb_factor).astype(dt)
    y = x if use_power else np.sqrt(x)
    y = np.maximum(y @ H, dt.type(floor))
    y = np.log(y) if gamma == 0 else (np.power(y, dt.type(gamma)) - 1) / dt.type(gamma)
    E = (2 * x[..., 1:-1]).sum(-1) + x[..., 0] + x[..., -1]
    E = np.log(E / dt.type(2 * (x.shape[-1] - 1)))[..., None]
    if out_format in (0, "y"):
        return y
    if out_format in (1, "yE"):
        return np.concatenate([y, E], axis=-1)
    if out_format in (2, "y,E"):
        return y, E
    raise ValueError(f"out_format {out_format} is not supported.")


def dct_matrix(dct_length, dct_type=2):
    """DCT basis W[n, k] in float64 (y = x @ W).  diffsptk/modules/dct.py:98-133."""
    L = dct_length
    if L <= 0:
        raise ValueError("dct_length must be positive.")
    if not 1 <= dct_type <= 4:
        raise ValueError("dct_type must be in [1, 4].")
    n = np.arange(L, dtype=np.float64)
    k = np.arange(L, dtype=np.float64)
    if dct_type in (2, 4):
        n = n + 0.5
    if dct_type in (3, 4):
        k = k + 0.5
    n = n * (math.pi / ((L - 1) if dct_type == 1 else L))
    if dct_type == 1:
        c = 0.5 ** 0.5
        z0 = np.full(L, 1.0); z0[0] = c; z0[-1] = c
        z1 = np.full(L, 2.0); z1[0] = 1; z1[-1] = 1
        z = z0[None, :] * np.sqrt(z1 / (L - 1))[:, None]
    elif dct_type == 2:
        z = np.full(L, 2.0); z[0] = 1
        z = np.sqrt(z / L)[None, :]
    elif dct_type == 3:
        z = np.full(L, 2.0); z[0] = 1
        z = np.sqrt(z / L)[:, None]
    else:
        z = (2 / L) ** 0.5
    return z * np.cos(k[None, :] * n[:, None])


def dct(x, dct_type=2):
    """x @ W.  diffsptk/modules/dct.py:135-137."""
    x = _as_float(x)
    return x @ dct_matrix(x.shape[-1], dct_type).astype(x.dtype)


def lifter_vector(mfcc_order, lifter):
    """1 + (lifter/2) sin(pi k / lifter), [0] = sqrt(2).  mfcc.py:233-235."""
    ramp = np.arange(mfcc_order + 1, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = 1 + (lifter / 2) * np.sin((math.pi / lifter) * ramp)
    v[0] = 2 ** 0.5
    return v


def mfcc(x, mfcc_order, n_channel, sample_rate, lifter=1, f_min=0.0, f_max=None, floor=1e-5,
         gamma=0.0, scale="htk", erb_factor=None, out_format="y"):
    """fbank -> DCT-II -> lifter -> split.  diffsptk/modules/mfcc.py:243-256."""
    x = _as_float(x)
    if mfcc_order < 0:
        raise ValueError("mfcc_order must be non-negative.")
    if n_channel <= mfcc_order:
        raise ValueError("mfcc_order must be less than n_channel.")
    if lifter < 0:
        raise ValueError("lifter must be non-negative.")
    y, E = fbank(x, n_channel, sample_rate, f_min, f_max, floor, gamma, scale, erb_factor,
                 use_power=False, out_format="y,E")
    y = dct(y, 2)
    y = y[..., : mfcc_order + 1] * lifter_vector(mfcc_order, lifter).astype(x.dtype)
    c, y = y[..., :1], y[..., 1:]
    if out_format in (0, "y"):
        return y
    if out_format in (1, "yE"):
        return np.concatenate([y, E], axis=-1)
    if out_format in (2, "yc"):
        return np.concatenate([y, c], axis=-1)
    if out_format in (3, "ycE"):
        return np.concatenate([y, c, E], axis=-1)
    raise ValueError(f"out_format {out_format} is not supported.")


# ----------------------------------------------------------------------------- per-row converters (8f rank 4)
def _get_gamma(gamma, c):
    """diffsptk/utils/private.py:233-238."""
    if c is None or c == 0:
        return gamma
    if not 1 <= c:
        raise ValueError("c must be an integer greater than or equal to 1.")
    return -1 / c


def _check_gamma(order, gamma, c, what):
    if order < 0:
        raise ValueError(f"{what} must be non-negative.")
    if 1 < abs(gamma):
        raise ValueError("gamma must be in [-1, 1].")
    if c is not None and c < 1:
        raise ValueError("c must be greater than or equal to 1.")


def lpc2par(a, gamma=1, c=None):
    """Step-down recursion [K, a_1..a_M] -> [K, k_1..k_M].  diffsptk/modules/lpc2par.py:104-120."""
    a = _as_float(a)
    M = a.shape[-1] - 1
    _check_gamma(M, gamma, c, "lpc_order")
    g = a.dtype.type(_get_gamma(gamma, c))
    K, a = a[..., :1], a[..., 1:]
    ks = []
    a = a * g
    for m in reversed(range(M)):
        km = a[..., m:m + 1]
        ks.append(km)
        if m == 0:
            break
        z = 1 - km * km
        k = a[..., :-1]
        a = (k - km * k[..., ::-1]) / z
    ks.append(K)
    return np.concatenate(ks[::-1], axis=-1)


def par2lpc(k, gamma=1, c=None):
    """Step-up recursion [K, k_1..k_M] -> [K, a_1..a_M] / gamma.  diffsptk/modules/par2lpc.py:100-107."""
    k = _as_float(k)
    _check_gamma(k.shape[-1] - 1, gamma, c, "lpc_order")
    g = k.dtype.type(_get_gamma(gamma, c))
    a = k / g
    for m in range(2, k.shape[-1]):
        km = k[..., m:m + 1]
        am = a[..., 1:m].copy()
        a[..., 1:m] = am + km * am[..., ::-1]
    return a


def gnorm(x, gamma=0, c=None):
    """Gain normalisation.  diffsptk/modules/gnorm.py:101-112."""
    x = _as_float(x)
    _check_gamma(x.shape[-1] - 1, gamma, c, "cep_order")
    g = _get_gamma(gamma, c)
    x0, x1 = x[..., :1], x[..., 1:]
    if g == 0:
        K, y = np.exp(x0), x1
    else:
        z = 1 + x.dtype.type(g) * x0
        K, y = np.power(z, x.dtype.type(1 / g)), x1 / z
    return np.concatenate([K, y], axis=-1)


def ignorm(y, gamma=0, c=None):
    """Inverse gain normalisation.  diffsptk/modules/ignorm.py:98-109."""
    y = _as_float(y)
    _check_gamma(y.shape[-1] - 1, gamma, c, "cep_order")
    g = _get_gamma(gamma, c)
    K, y1 = y[..., :1], y[..., 1:]
    if g == 0:
        x0, x1 = np.log(K), y1
    else:
        z = np.power(K, y.dtype.type(g))
        x0, x1 = (z - 1) / y.dtype.type(g), y1 * z
    return np.concatenate([x0, x1], axis=-1)


def norm0(a):
    """[K, a_1..a_M] -> [1/K, a_1/K..a_M/K].  diffsptk/modules/norm0.py:88-94."""
    a = _as_float(a)
    b0 = 1 / a[..., :1]
    return np.concatenate([b0, a[..., 1:] * b0], axis=-1)


def mc2b(mc, alpha=0):
    """b_M = mc_M, b_m = mc_m - alpha b_{m+1}.  diffsptk/modules/mc2b.py:93-101."""
    mc = _as_float(mc)
    if 1 <= abs(alpha):
        raise ValueError("alpha must be in (-1, 1).")
    M = mc.shape[-1] - 1
    b = np.zeros_like(mc)
    b[..., M] = mc[..., M]
    for m in reversed(range(M)):
        b[..., m] = mc[..., m] - mc.dtype.type(alpha) * b[..., m + 1]
    return b


def b2mc(b, alpha=0):
    """mc_m = b_m + alpha b_{m+1}.  diffsptk/modules/b2mc.py:92-95."""
    b = _as_float(b)
    if 1 <= abs(alpha):
        raise ValueError("alpha must be in (-1, 1).")
    mc = b.copy()
    mc[..., :-1] += b.dtype.type(alpha) * b[..., 1:]
    return mc


# ----------------------------------------------------------------------------- mgc2mgc / mgc2sp / plp (8f rank 3)
def _gc2gc(c1, out_order, in_gamma, out_gamma, n_fft):
    """Gamma conversion in the spectral domain.  diffsptk/modules/mgc2mgc.py:327-364."""
    c01 = np.concatenate([np.zeros_like(c1[..., :1]), c1[..., 1:]], axis=-1)
    C1 = np.fft.fft(c01.astype(np.float64), n=n_fft, axis=-1)
    if in_gamma == 0:
        sC1 = np.exp(C1.real) * np.exp(1j * C1.imag)
    else:
        C1 = C1 * in_gamma
        C1 = C1 + 1
        sC1 = (np.abs(C1) ** (1 / in_gamma)) * np.exp(1j * (np.angle(C1) / in_gamma))
    if out_gamma == 0:
        C2 = np.log(np.abs(sC1))
    else:
        C2 = ((np.abs(sC1) ** out_gamma) * np.cos(np.angle(sC1) * out_gamma) - 1) / out_gamma
    c02 = np.fft.ifft(C2, axis=-1).real[..., : out_order + 1]
    return np.concatenate([c1[..., :1], (2 * c02[..., 1:]).astype(c1.dtype)], axis=-1)


def mgc2mgc(mc, out_order, in_alpha=0, out_alpha=0, in_gamma=0, out_gamma=0, in_norm=False, out_norm=False,
            in_mul=False, out_mul=False, n_fft=512):
    """The reference's step sequence.  diffsptk/modules/mgc2mgc.py:209-297 (the FFTs of the gamma conversion run
    in float64 here; the reference's run in the input dtype)."""
    mc = _as_float(mc)
    in_order = mc.shape[-1] - 1
    if in_order < 0 or out_order < 0:
        raise ValueError("order must be non-negative.")
    if 1 <= abs(in_alpha) or 1 <= abs(out_alpha):
        raise ValueError("alpha must be in (-1, 1).")
    if 1 < abs(in_gamma) or 1 < abs(out_gamma):
        raise ValueError("gamma must be in [-1, 1].")
    if n_fft <= max(in_order, out_order) + 1:
        raise ValueError("n_fft must be much larger than order of cepstrum.")
    if 0 == in_gamma and in_mul:
        raise ValueError("Invalid combination of in_gamma and in_mul.")
    t = mc.dtype.type
    tail = lambda c, f: np.concatenate([c[..., :1], f(c[..., 1:])], axis=-1)      # noqa: E731
    head = lambda c, f: np.concatenate([f(c[..., :1]), c[..., 1:]], axis=-1)      # noqa: E731
    c = mc
    if not in_norm and in_mul:
        c = head(c, lambda v: (v - 1) / t(in_gamma))
    alpha = (out_alpha - in_alpha) / (1 - in_alpha * out_alpha)
    if 0 == alpha:
        if in_order == out_order and in_gamma == out_gamma:
            if not in_mul and out_mul:
                c = tail(c, lambda v: v * t(in_gamma))
            if not in_norm and out_norm:
                c = gnorm(c, in_gamma)
            if in_norm and not out_norm:
                c = ignorm(c, out_gamma)
            if in_mul and not out_mul:
                c = tail(c, lambda v: v / t(out_gamma))
        else:
            if in_mul:
                c = tail(c, lambda v: v / t(in_gamma))
            if not in_norm:
                c = gnorm(c, in_gamma)
            c = _gc2gc(c, out_order, in_gamma, out_gamma, n_fft)
            if not out_norm:
                c = ignorm(c, out_gamma)
            if out_mul:
                c = tail(c, lambda v: v * t(out_gamma))
    else:
        if in_mul:
            c = tail(c, lambda v: v / t(in_gamma))
        if in_norm:
            c = ignorm(c, in_gamma)
        c = freqt(c, out_order, alpha)
        if out_norm or in_gamma != out_gamma:
            c = gnorm(c, in_gamma)
        if in_gamma != out_gamma:
            c = _gc2gc(c, out_order, in_gamma, out_gamma, n_fft)
        if not out_norm and in_gamma != out_gamma:
            c = ignorm(c, out_gamma)
        if out_mul:
            c = tail(c, lambda v: v * t(out_gamma))
    if not out_norm and out_mul:
        c = head(c, lambda v: v * t(out_gamma) + 1)
    return c


def mgc2sp(mc, fft_length, alpha=0, gamma=0, norm=False, mul=False, n_fft=512, out_format="power"):
    """mgc2mgc to a plain cepstrum of order L/2, rfft, formatter.  diffsptk/modules/mgc2sp.py:131-202."""
    mc = _as_float(mc)
    c = mgc2mgc(mc, fft_length // 2, in_alpha=alpha, out_alpha=0, in_gamma=gamma, out_gamma=0, in_norm=norm,
                out_norm=False, in_mul=mul, out_mul=False, n_fft=n_fft)
    sp = np.fft.rfft(c.astype(np.float64), n=(c.shape[-1] - 1) * 2, axis=-1)
    if out_format in (0, "db"):
        out = sp.real * (20 / math.log(10))
    elif out_format in (1, "log-magnitude"):
        out = sp.real
    elif out_format in (2, "magnitude"):
        out = np.exp(sp.real)
    elif out_format in (3, "power"):
        out = np.exp(2 * sp.real)
    elif out_format in (4, "cycle"):
        out = sp.imag / math.pi
    elif out_format in (5, "radian"):
        out = sp.imag
    elif out_format in (6, "degree"):
        out = sp.imag * (180 / math.pi)
    elif out_format == "complex":
        return (np.exp(sp.real) * np.exp(1j * sp.imag)).astype(np.complex64 if mc.dtype == np.float32 else np.complex128)
    else:
        raise ValueError(f"out_format {out_format} is not supported.")
    return out.astype(mc.dtype)


def plp(x, plp_order, n_channel, sample_rate, compression_factor=0.33, lifter=1, f_min=0.0, f_max=None, floor=1e-5,
        gamma=0.0, scale="htk", erb_factor=None, n_fft=512, out_format="y"):
    """fbank (power) -> equal loudness -> compression -> inverse DFT -> levdur -> LPC cepstrum -> lifter.
    diffsptk/modules/plp.py:192-320."""
    x = _as_float(x)
    if plp_order < 0:
        raise ValueError("plp_order must be non-negative.")
    if n_channel <= plp_order:
        raise ValueError("plp_order must be less than n_channel.")
    if compression_factor <= 0:
        raise ValueError("compression_factor must be positive.")
    if lifter < 0:
        raise ValueError("lifter must be non-negative.")
    y, E = fbank(x, n_channel, sample_rate, f_min, f_max, floor, gamma, scale, erb_factor, use_power=True,
                 out_format="y,E")
    fmax = sample_rate / 2 if f_max is None else f_max
    mel_min = _hz_to_auditory(np.asarray(f_min, dtype=np.float64), scale)
    mel_max = _hz_to_auditory(np.asarray(fmax, dtype=np.float64), scale)
    centre = (mel_max - mel_min) / (n_channel + 1) * np.arange(1, n_channel + 2) + mel_min
    f = _auditory_to_hz(centre, scale)[:-1] ** 2
    elc = ((f / (f + 1.6e5)) ** 2 * (f + 1.44e6) / (f + 9.61e6)).astype(x.dtype)
    y = (np.exp(y) * elc) ** x.dtype.type(compression_factor)
    y = np.concatenate([y[..., :1], y, y[..., -1:]], axis=-1)
    r = np.fft.hfft(y.astype(np.float64), norm="forward", axis=-1).real[..., : plp_order + 1].astype(x.dtype)
    a = levdur(r, eps=0)
    c = mgc2mgc(a, plp_order, in_alpha=0, out_alpha=0, in_gamma=-1, out_gamma=0, in_norm=True, out_norm=False,
                in_mul=True, out_mul=False, n_fft=n_fft)
    ramp = np.arange(plp_order + 1, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        lv = 1 + (lifter / 2) * np.sin((math.pi / lifter) * ramp)
    lv[0] = 2
    c = c * lv.astype(x.dtype)
    c0, y = c[..., :1], c[..., 1:]
    if out_format in (0, "y"):
        return y
    if out_format in (1, "yE"):
        return np.concatenate([y, E], axis=-1)
    if out_format in (2, "yc"):
        return np.concatenate([y, c0], axis=-1)
    if out_format in (3, "ycE"):
        return np.concatenate([y, c0, E], axis=-1)
    raise ValueError(f"out_format {out_format} is not supported.")


# ----------------------------------------------------------------------------- mgcep (8f rank 3)
def _mgcep_freqt_matrix(in_order, out_order, alpha):
    """CoefficientsFrequencyTransform of diffsptk/modules/mgcep.py:251-283, shape (L1, L2)."""
    beta = 1 - alpha * alpha
    L1, L2 = in_order + 1, out_order + 1
    A = np.zeros((L2, L1))
    A[0, 0] = 1
    if 1 < L2 and 1 < L1:
        A[1, 1:] = alpha ** np.arange(L1 - 1, dtype=np.float64) * beta
    for i in range(2, L2):
        for j in range(1, L1):
            A[i, j] = A[i - 1, j - 1] + alpha * (A[i, j - 1] - A[i - 1, j])
    return np.ascontiguousarray(A.T)


def mgcep(x, *, fft_length, cep_order, alpha=0, gamma=0, c=None, n_iter=0):
    """Newton iteration on the mel-generalized cepstrum.  diffsptk/modules/mgcep.py:173-246 (module-only API)."""
    x = _as_float(x)
    gamma = _get_gamma(gamma, c)
    if fft_length < 2 * cep_order:
        raise ValueError("cep_order must be less than or equal to fft_length // 2.")
    if gamma < -1 or 0 < gamma:
        raise ValueError("gamma must be in [-1, 0].")
    if gamma == 0:
        return mcep(x, cep_order, alpha, n_iter)
    M, L, dt = cep_order, fft_length, x.dtype
    cf = _mgcep_freqt_matrix(M, L - 1, -alpha).astype(dt)
    pf = _mgcep_freqt_matrix(L - 1, 2 * M, alpha).astype(dt)
    rf = _mgcep_freqt_matrix(L - 1, M, alpha).astype(dt)
    P = np.eye(2 * M + 1)
    P[np.arange(2 * M), np.arange(1, 2 * M + 1)] = alpha
    P[0, 0] -= alpha * alpha
    P[0, 1] += alpha
    P[-1, -1] += alpha
    Q = np.eye(2 * M + 1)
    Q[np.arange(1, 2 * M + 1), np.arange(2 * M)] = alpha
    Q[1, 0] = 0
    Q[1, 1] += alpha
    P, Q = P.T.astype(dt), Q.T.astype(dt)

    def irfft(z):
        return np.fft.irfft(z, axis=-1).astype(dt)

    def newton(g, b1):
        b = np.concatenate([np.zeros_like(b1[..., :1]), b1], axis=-1)
        C = np.fft.rfft((b @ cf).astype(np.float64), n=L, axis=-1)
        Cr, Ci = C.real.astype(dt), C.imag.astype(dt)
        if g == -1:
            p_re = x
        else:
            X, Y = 1 + dt.type(g) * Cr, dt.type(g) * Ci
            XX, YY = X * X, Y * Y
            D = XX + YY
            p_re = x * np.power(D, dt.type(-1 / g)) / D
            q = p_re / D
            q_c = q * (XX - YY) + 1j * (q * (2 * X * Y))
            r_c = p_re * X + 1j * (p_re * Y)
        p = irfft(p_re) @ pf
        if g == -1:
            q, r = p, p[..., : M + 1]
        else:
            q, r = irfft(q_c) @ pf, irfft(r_c) @ rf
        p, q = p @ P, q @ Q
        if g != -1:
            eps = r[..., 0] + dt.type(g) * (r[..., 1:] * b1).sum(-1)
        R = _toeplitz(p[..., :M]) + _hankel(q[..., 2:] * dt.type(1 + g))
        b1 = b1 + np.linalg.solve(R, r[..., 1:, None])[..., 0].astype(dt)
        if g == -1:
            eps = r[..., 0] + dt.type(g) * (r[..., 1:] * b1).sum(-1)
        return np.sqrt(eps)[..., None], b1

    b0, b1 = newton(-1, np.zeros((*x.shape[:-1], M), dtype=dt))
    if gamma != -1:
        b = np.concatenate([b0, b1], axis=-1)
        b = ignorm(b, -1)
        b = b2mc(b, alpha)
        b = mgc2mgc(b, M, in_gamma=-1, out_gamma=gamma)
        b = mc2b(b, alpha)
        b = gnorm(b, gamma)
        b1 = b[..., 1:]
        for _ in range(n_iter):
            b0, b1 = newton(gamma, b1)
    return b2mc(ignorm(np.concatenate([b0, b1], axis=-1), gamma), alpha)


# ----------------------------------------------------------------------------- lpc2lsp (8f rank 4)
def lpc2lsp(a, log_gain=False, sample_rate=None, out_format="radian"):
    """LPC -> line spectral pairs: the roots of the deflated symmetric / antisymmetric polynomials, found as the
    eigenvalues of the companion matrix like the reference (diffsptk/modules/lpc2lsp.py:159-197,
    root_pol.py:130-146); one member of every conjugate pair is kept (the reference takes every other eigenvalue,
    which LAPACK returns pair by pair)."""
    a = _as_float(a)
    M = a.shape[-1] - 1
    if out_format in (2, 3, "hz", "khz") and (sample_rate is None or sample_rate <= 0):
        raise ValueError("sample_rate must be positive.")
    tau = 2 * math.pi
    if out_format in (0, "radian"):
        scale = 1.0
    elif out_format in (1, "cycle"):
        scale = 1 / tau
    elif out_format in (2, "khz"):
        scale = 1 / (tau / sample_rate * 1000)
    elif out_format in (3, "hz"):
        scale = 1 / (tau / sample_rate)
    else:
        raise ValueError(f"out_format {out_format} is not supported.")
    K = a[..., :1]
    if log_gain:
        K = np.log(K)
    if M == 0:
        return K
    flat = a.reshape(-1, M + 1).astype(np.float64)
    out = np.empty((flat.shape[0], M))
    for n, row in enumerate(flat):
        a1 = np.concatenate([[1.0], row[1:], [0.0]])
        a2 = a1[::-1]
        p, q = a1 - a2, a1 + a2
        if M == 1:
            r = np.roots(q)
            out[n] = np.abs(np.angle(r[:1]))
            continue
        if M % 2 == 0:
            p = np.polydiv(p, [1.0, -1.0])[0]
            q = np.polydiv(q, [1.0, 1.0])[0]
        else:
            p = np.polydiv(p, [1.0, 0.0, -1.0])[0]
        # conjugate pairs: one angle of each pair (sorted |angle| come in equal twos)
        ang = [np.sort(np.abs(np.angle(np.roots(c))))[0::2] for c in (p, q)]
        out[n] = np.sort(np.concatenate(ang))
    w = (out * scale).astype(a.dtype).reshape(*a.shape[:-1], M)
    return np.concatenate([K, w], axis=-1)
