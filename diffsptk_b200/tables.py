"""Host-side tables of the analysis path, memoised by parameter tuple.

The reference rebuilds these on every functional call (``_precompute`` runs per
call, e.g. diffsptk/modules/window.py:109-113 and the Python double loops of
diffsptk/modules/freqt.py:133-137).  Here each table is built once per
parameter tuple, in the precision the reference uses (float64 recursions cast
to the module dtype; the window directly in the module dtype with the same
torch generators so the table is bit-identical), and cached per device.
"""

from __future__ import annotations

import functools
import math

import numpy as np
import torch
from torch.signal.windows import cosine as _cosine

_WINDOW_IDS = {0: "blackman", 1: "hamming", 2: "hanning", 3: "bartlett", 4: "trapezoidal",
               5: "rectangular", 6: "nuttall"}
_NORM_IDS = {0: "none", 1: "power", 2: "magnitude"}


def _default_dtype(dtype):
    return torch.get_default_dtype() if dtype is None else dtype


def _cast(x, device, dtype):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=device, dtype=_default_dtype(dtype))


def _dev_key(device):
    return None if device is None else str(torch.device(device))


# ------------------------------------------------------------------------------------ window
def make_window(length, kind="blackman", norm="power", symmetric=True, device=None, dtype=None):
    """Window table of diffsptk/modules/window.py:122-183 (same generators, same dtype)."""
    kind = _WINDOW_IDS.get(kind, kind) if not isinstance(kind, str) else kind
    norm = _NORM_IDS.get(norm, norm) if not isinstance(norm, str) else norm
    return _window_cached(int(length), kind, norm, bool(symmetric), _dev_key(device), dtype).clone()


@functools.lru_cache(maxsize=256)
def _window_cached(L, kind, norm, symmetric, device, dtype):
    periodic = not symmetric
    kw = {"dtype": dtype, "device": device}
    if kind == "blackman":
        w = torch.blackman_window(L, periodic=periodic, **kw)
    elif kind == "hamming":
        w = torch.hamming_window(L, periodic=periodic, **kw)
    elif kind == "hanning":
        w = torch.hann_window(L, periodic=periodic, **kw)
    elif kind in ("bartlett", "trapezoidal"):
        w = torch.bartlett_window(L, periodic=periodic, **kw)
        if kind == "trapezoidal":
            w = (2 * w).clip(max=1)
    elif kind == "rectangular":
        w = torch.ones(L, **kw)
    elif kind == "nuttall":
        size = L if periodic else L - 1
        coef = torch.tensor([0.355768, -0.487396, 0.144232, -0.012604], **kw)
        ang = torch.arange(0, 8, 2, **kw) * (torch.pi / size)
        w = (coef * torch.cos(torch.outer(torch.arange(L, **kw), ang))).sum(dim=1)
    elif kind == "povey":
        w = torch.hann_window(L, periodic=periodic, **kw).pow(0.85)
    elif kind == "sine":
        w = _cosine(L, sym=symmetric, **kw)
    elif kind == "vorbis":
        w = torch.sin(torch.pi * 0.5 * _cosine(L, sym=symmetric, **kw) ** 2)
    elif kind == "kbd":
        if periodic:
            raise ValueError("periodic is not supported for kbd window.")
        seed = torch.kaiser_window(L // 2 + 1, periodic=False, **kw)
        cs = torch.cumsum(seed, dim=0)
        half = torch.sqrt(cs[:-1] / cs[-1])
        w = torch.cat([half, half.flip(0)])
    else:
        raise ValueError(f"window {kind} is not supported.")
    if norm == "none":
        pass
    elif norm == "power":
        w = w / torch.sqrt(torch.sum(w ** 2))
    elif norm == "magnitude":
        w = w / torch.sum(w)
    else:
        raise ValueError(f"norm {norm} is not supported.")
    return w.to(dtype=_default_dtype(dtype))


# ------------------------------------------------------------------------------ warping (freqt)
@functools.lru_cache(maxsize=64)
def _freqt_np(in_order, out_order, alpha):
    """Oppenheim recursion, float64, stored transposed (M1+1, M2+1).  freqt.py:124-139."""
    L1, L2 = in_order + 1, out_order + 1
    A = np.zeros((L2, L1))
    A[0, :] = (alpha ** torch.arange(L1, dtype=torch.double)).numpy()  # torch pow: bit-identical to the reference
    if L2 > 1 and L1 > 1:
        A[1, 1:] = A[0, :-1] * (1 - alpha * alpha) * np.arange(1, L1, dtype=np.float64)
    for i in range(2, L2):
        prev, cur = A[i - 1], A[i]
        for j in range(1, L1):
            cur[j] = prev[j - 1] + alpha * (cur[j - 1] - prev[j])
    return np.ascontiguousarray(A.T)


@functools.lru_cache(maxsize=64)
def _coef_freqt_np(in_order, out_order, alpha):
    """mcep-internal warping of correlation-like sequences, float64.  mcep.py:264-288."""
    L1, L2 = in_order + 1, out_order + 1
    A = np.zeros((L2, L1))
    A[:, 0] = ((-alpha) ** torch.arange(L2, dtype=torch.double)).numpy()
    for i in range(1, L2):
        prev, cur = A[i - 1], A[i]
        for j in range(1, L1):
            cur[j] = prev[j - 1] + alpha * (cur[j - 1] - prev[j])
    return np.ascontiguousarray(A.T)


def make_freqt_matrix(in_order, out_order, alpha, device=None, dtype=None):
    return _cast(_freqt_np(int(in_order), int(out_order), float(alpha)).copy(), device, dtype)


def make_coef_freqt_matrix(in_order, out_order, alpha, device=None, dtype=None):
    return _cast(_coef_freqt_np(int(in_order), int(out_order), float(alpha)).copy(), device, dtype)


@functools.lru_cache(maxsize=16)
def _mcep_fused_np(fft_length, cep_order, alpha):
    """The three dense tables of the fused mcep kernel, float64 (see include/diffsptk_b200.h).

    With H = L/2, k, n in 0..H:
      irfft of a real even spectrum :  c[n] = sum_k s[k] * w_k / L * cos(2 pi k n / L), w_0 = w_H = 1, else 2
      real part of rfft(c, n=L)     :  d[k] = sum_n c[n] * cos(2 pi n k / L)
    """
    L, M = fft_length, cep_order
    H = L // 2
    k = np.arange(H + 1, dtype=np.float64)
    cosm = np.cos(2.0 * math.pi * np.outer(k, k) / L)              # [k, n], symmetric
    wk = np.full(H + 1, 2.0); wk[0] = 1.0; wk[H] = 1.0
    C1 = (wk / L)[:, None] * cosm                                  # spectrum -> cepstrum/correlation
    halve = np.ones(H + 1); halve[0] = 0.5; halve[H] = 0.5         # mcep.py:205-206
    P0 = (C1 * halve[None, :]) @ _freqt_np(H, M, alpha)            # [H+1, M+1]
    G = _freqt_np(M, H, -alpha) @ cosm                             # [M+1, H+1]
    Hm = C1 @ _coef_freqt_np(H, 2 * M, alpha)                      # [H+1, 2M+1]
    return np.ascontiguousarray(P0), np.ascontiguousarray(G), np.ascontiguousarray(Hm)


def make_mcep_tables(fft_length, cep_order, alpha, device=None, dtype=None):
    P0, G, Hm = _mcep_fused_np(int(fft_length), int(cep_order), float(alpha))
    return tuple(_cast(t.copy(), device, dtype) for t in (P0, G, Hm))


# -------------------------------------------------------------------------------- filter bank
def _to_auditory(f, scale):
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


def _from_auditory(z, scale):
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


@functools.lru_cache(maxsize=64)
def _fbank_np(fft_length, n_channel, sample_rate, f_min, f_max, scale, erb_factor):
    """H[L/2+1, C] of diffsptk/modules/fbank.py:233-293, float64."""
    K, C = fft_length // 2 + 1, n_channel
    H = np.zeros((K, C))
    if erb_factor is None:
        z0 = _to_auditory(np.asarray(f_min, dtype=np.float64), scale)
        z1 = _to_auditory(np.asarray(f_max, dtype=np.float64), scale)
        first = max(1, int(f_min / sample_rate * fft_length + 1.5))
        last = min(fft_length // 2, int(f_max / sample_rate * fft_length + 0.5))
        centre = (z1 - z0) / (C + 1) * np.arange(1, C + 2) + z0
        span = np.diff(np.concatenate([[z0], centre]))
        bins = np.arange(first, last)
        zb = _to_auditory(sample_rate * bins / fft_length, scale)
        for k, z in zip(bins, zb):
            m = int(np.argmax(z <= centre))
            w = (centre[m] - z) / span[m]
            if m > 0:
                H[k, m - 1] = w
            if m < C:
                H[k, m] = 1 - w
    else:
        a, b, c = erb_factor * 6.23e-6, erb_factor * 93.39e-3, erb_factor * 28.52

        def edge_centre(f, sign):
            ah = sign * 0.5 / (700 + f)
            bh = sign * 700 / (700 + f)
            ch = -sign * 0.5 * f * (1 + 700 / (700 + f))
            bb, cc = (b - bh) / (a - ah), (c - ch) / (a - ah)
            return 0.5 * (-bb + np.sqrt(bb * bb - 4 * cc))

        zc = np.linspace(_to_auditory(edge_centre(f_min, 1), scale),
                         _to_auditory(edge_centre(f_max, -1), scale), C)
        fc = _from_auditory(zc, scale)
        erb = a * fc ** 2 + b * fc + c
        lo = -(700 + erb) + np.sqrt(erb ** 2 + (700 + fc) ** 2)
        hi = lo + 2 * erb
        f = np.linspace(0, sample_rate / 2, K)
        for m in range(C):
            rise = (lo[m] <= f) & (f < fc[m])
            H[rise, m] = (f[rise] - lo[m]) / (fc[m] - lo[m])
            fall = (fc[m] <= f) & (f <= hi[m])
            H[fall, m] = (hi[m] - f[fall]) / (hi[m] - fc[m])
    return H


def make_fbank_matrix(fft_length, n_channel, sample_rate, f_min=0.0, f_max=None, scale="htk",
                      erb_factor=None, device=None, dtype=None):
    if f_max is None:
        f_max = sample_rate / 2
    H = _fbank_np(int(fft_length), int(n_channel), sample_rate, float(f_min), float(f_max), scale,
                  None if erb_factor is None else float(erb_factor))
    return _cast(H.copy(), device, dtype)


def column_support(H: torch.Tensor):
    """[begin, end) of the non-zero rows of every column of H, as int32 tensors on H's device."""
    nz = (H != 0)
    K = H.shape[0]
    idx = torch.arange(K, device=H.device).unsqueeze(1)
    big = torch.where(nz, idx, torch.full_like(idx, K))
    small = torch.where(nz, idx + 1, torch.zeros_like(idx))
    begin = big.min(dim=0).values
    end = small.max(dim=0).values
    begin = torch.minimum(begin, end)
    return begin.to(torch.int32).contiguous(), end.to(torch.int32).contiguous()


# ----------------------------------------------------------------------------------- DCT / lifter
@functools.lru_cache(maxsize=64)
def _dct_np(L, dct_type):
    """W[n, k] of diffsptk/modules/dct.py:98-133, float64 (y = x @ W).

    Built with torch element-wise kernels (cos, sqrt) because their last-bit results differ from
    numpy's and the table is meant to be bit-identical to the reference's.
    """
    f64 = torch.double
    n = torch.arange(L, dtype=f64) + (0.5 if dct_type in (2, 4) else 0.0)
    k = torch.arange(L, dtype=f64) + (0.5 if dct_type in (3, 4) else 0.0)
    n = n * (math.pi / ((L - 1) if dct_type == 1 else L))

    def ends(mid, first, last=None):
        v = torch.full((L,), float(mid), dtype=f64)
        v[0] = first
        if last is not None:
            v[-1] = last
        return v

    if dct_type == 1:
        z = ends(1.0, 0.5 ** 0.5, 0.5 ** 0.5)[None, :] * torch.sqrt(ends(2.0, 1.0, 1.0) / (L - 1))[:, None]
    elif dct_type == 2:
        z = torch.sqrt(ends(2.0, 1.0) / L)[None, :]
    elif dct_type == 3:
        z = torch.sqrt(ends(2.0, 1.0) / L)[:, None]
    else:
        z = (2.0 / L) ** 0.5
    return (z * torch.cos(k[None, :] * n[:, None])).numpy()


def make_dct_matrix(dct_length, dct_type=2, device=None, dtype=None):
    return _cast(_dct_np(int(dct_length), int(dct_type)).copy(), device, dtype)


def make_lifter(mfcc_order, lifter, device=None, dtype=None):
    """1 + (lifter/2) sin(pi k / lifter), element 0 = sqrt(2).  mfcc.py:233-235."""
    ramp = torch.arange(mfcc_order + 1, device=device, dtype=torch.double)
    v = 1 + (lifter / 2) * torch.sin((torch.pi / lifter) * ramp)
    v[0] = 2 ** 0.5
    return v.to(dtype=_default_dtype(dtype))


@functools.lru_cache(maxsize=64)
def _delta_window_cached(seed_key, static_out: bool, device, dtype):
    """Regression windows (H, W), built in double exactly like diffsptk/modules/delta.py:98-170."""
    seed = [list(s) if isinstance(s, tuple) else s for s in seed_key]
    if isinstance(seed[0], list):
        rows = ([[1.0]] if static_out else []) + [list(map(float, c)) for c in seed]
        max_len = max(len(c) for c in rows)
        if max_len % 2 == 0:
            max_len += 1
        window = []
        for c in rows:
            diff = max_len - len(c)
            left, right = (diff // 2, diff // 2) if diff % 2 == 0 else ((diff - 1) // 2, (diff + 1) // 2)
            window.append(torch.tensor([0.0] * left + c + [0.0] * right, dtype=torch.double))
    else:
        if min(seed) <= 0:
            raise ValueError("The width of regression coefficients must be positive.")
        if len(seed) >= 3:
            raise ValueError("3rd order regression is not supported.")
        max_len = max(seed) * 2 + 1
        window = []
        if static_out:
            w = torch.zeros(max_len, dtype=torch.double)
            w[(max_len - 1) // 2] = 1
            window.append(w)
        n = seed[0]
        z = 1 / (n * (n + 1) * (2 * n + 1) / 3)
        j = torch.arange(-n, n + 1, dtype=torch.double)
        p = (max_len - (n * 2 + 1)) // 2
        window.append(torch.nn.functional.pad(j * z, (p, p)))
        if len(seed) >= 2:
            n = seed[1]
            a0 = 2 * n + 1
            a1 = a0 * n * (n + 1) / 3
            a2 = a1 * (3 * n * n + 3 * n - 1) / 5
            z = 1 / (2 * (a2 * a0 - a1 * a1))
            j = torch.arange(-n, n + 1, dtype=torch.double)
            p = (max_len - (n * 2 + 1)) // 2
            window.append(torch.nn.functional.pad((a0 * j * j - a1) * z, (p, p)))
    return _cast(torch.stack(window), device, dtype)


def make_delta_window(seed, static_out: bool = True, device=None, dtype=None) -> Tensor:
    if not isinstance(seed, (tuple, list)):
        raise ValueError("seed must be tuple or list.")
    key = tuple(tuple(s) if isinstance(s, (tuple, list)) else s for s in seed)
    return _delta_window_cached(key, bool(static_out), _dev_key(device), dtype).clone()


# ------------------------------------------------------------------------- mc2b / b2mc (section 8f rank 4)
def make_mc2b_matrix(cep_order, alpha, device=None, dtype=None):
    """``A`` of mc2b.py:107-119 (stored transposed): ``b = mc @ A``, ``A[m + d, m] = (-alpha)^d``."""
    a = 1
    A = torch.eye(int(cep_order) + 1, dtype=torch.double)
    for m in range(1, len(A)):
        a *= -alpha
        A[:, m:].fill_diagonal_(a)
    return _cast(A.T.contiguous(), device, dtype)


def make_b2mc_matrix(cep_order, alpha, device=None, dtype=None):
    """``A`` of b2mc.py:104-115 (stored transposed): ``mc = b @ A``, ones on the diagonal, alpha below it."""
    A = torch.eye(int(cep_order) + 1, dtype=torch.double)
    A[:, 1:].fill_diagonal_(alpha)
    return _cast(A.T.contiguous(), device, dtype)


# ------------------------------------------------------------------------- mgcep tables (section 8f rank 3)
@functools.lru_cache(maxsize=32)
def _mgcep_freqt_np(in_order, out_order, alpha):
    """``CoefficientsFrequencyTransform`` of mgcep.py:251-283 (NOT the one of mcep.py), stored transposed."""
    beta = 1 - alpha * alpha
    L1, L2 = in_order + 1, out_order + 1
    A = np.zeros((L2, L1))
    A[0, 0] = 1
    if 1 < L2 and 1 < L1:
        A[1, 1:] = (alpha ** torch.arange(L1 - 1, dtype=torch.double)).numpy() * beta
    for i in range(2, L2):
        prev, cur = A[i - 1], A[i]
        for j in range(1, L1):
            cur[j] = prev[j - 1] + alpha * (cur[j - 1] - prev[j])
    return np.ascontiguousarray(A.T)


def make_mgcep_freqt_matrix(in_order, out_order, alpha, device=None, dtype=None):
    return _cast(_mgcep_freqt_np(int(in_order), int(out_order), float(alpha)).copy(), device, dtype)


def make_mgcep_ptrans(order, alpha, device=None, dtype=None):
    """``PTransform`` of mgcep.py:286-308, stored transposed."""
    A = torch.eye(order + 1, dtype=torch.double)
    A[:, 1:].fill_diagonal_(alpha)
    A[0, 0] -= alpha * alpha
    A[0, 1] += alpha
    A[-1, -1] += alpha
    return _cast(A.T.contiguous(), device, dtype)


def make_mgcep_qtrans(order, alpha, device=None, dtype=None):
    """``QTransform`` of mgcep.py:311-332, stored transposed."""
    A = torch.eye(order + 1, dtype=torch.double)
    A[1:].fill_diagonal_(alpha)
    A[1, 0] = 0
    A[1, 1] += alpha
    return _cast(A.T.contiguous(), device, dtype)
