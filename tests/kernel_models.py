"""numpy models of the data flow of the specialised CUDA kernels (same index arithmetic, lane by
lane).  They let the CPU test-suite check the *algorithms* -- decomposition, twiddles, exchange
patterns, lane-0 special cases -- against the oracle before the CUDA code ever runs.

stft512 model (diffsptk_b200/csrc/stft512.cu): a 512-point real FFT as a 256-point complex FFT
(z[m] = x[2m] + i x[2m+1]) factored 16 x 16 across the 16 lanes of a half-warp:

  pass 1   lane m1 holds z[m1 + 16 j], j = 0..15, and does a radix-16 FFT over j -> index k2
  twiddle  C[m1][k2] *= W256^(m1 k2)
  exch. 1  transpose through shared memory: lane k2 receives C[0..15][k2]
  pass 2   radix-16 FFT over m1 -> k1;  lane k2 holds Z[16 k1 + k2]
  exch. 2  lane l swaps registers 8..15 with lane (16 - l) % 16 (lane 0 and 8 with themselves)
  split    X[k], X[256-k] from Z[k], Z[256-k];  lane 0 also owns the bins 0, 128, 256
"""

from __future__ import annotations

import numpy as np


def w(n, k):
    return np.exp(-2j * np.pi * k / n)


def radix4(x0, x1, x2, x3):
    t0, t1, t2, t3 = x0 + x2, x0 - x2, x1 + x3, x1 - x3
    return t0 + t2, t1 - 1j * t3, t0 - t2, t1 + 1j * t3


def fft16(a):
    """Natural-order 16-point DFT as 4 x 4 (j = 4 s + c, k = r + 4 t), as in the kernel."""
    b = [[None] * 4 for _ in range(4)]
    for c in range(4):
        y = radix4(a[c], a[c + 4], a[c + 8], a[c + 12])
        for r in range(4):
            b[c][r] = y[r] * w(16, c * r)
    out = [None] * 16
    for r in range(4):
        y = radix4(b[0][r], b[1][r], b[2][r], b[3][r])
        for t in range(4):
            out[r + 4 * t] = y[t]
    return out


def fft16_fma(a):
    """fft16.cuh, round 2: the W16^(c r) twiddles of the second radix-4 layer folded into the butterflies'
    multiply-adds -- a twiddle cos (1 - i tan) is a rotation-by-tan (two multiply-adds) whose cosine rides on the
    following +/-; R (1 -+ i) is two additions whose R rides likewise.  Real-level formulas as in the kernel
    (layer2<R>), result in natural order."""
    R, C8, T8 = np.sqrt(0.5), np.cos(np.pi / 8), np.tan(np.pi / 8)
    a = [complex(v) for v in a]
    for c in range(4):
        a[c], a[c + 4], a[c + 8], a[c + 12] = radix4(a[c], a[c + 4], a[c + 8], a[c + 12])

    def pm_s(x, y, s_):
        return x + s_ * y, x - s_ * y

    def pm_is(x, y, s_):
        return x - 1j * s_ * y, x + 1j * s_ * y

    def layer2(r, x0, x1, x2, x3):
        if r == 0:
            return radix4(x0, x1, x2, x3)
        if r == 2:
            t0, t1 = x0 - 1j * x2, x0 + 1j * x2
            s_, d = x1 + x3, x1 - x3
            o0, o2 = pm_s(t0, d - 1j * s_, R)
            o1, o3 = pm_is(t1, s_ - 1j * d, R)
            return o0, o1, o2, o3
        if r == 1:
            q = complex(x2.real + x2.imag, x2.imag - x2.real)                  # (1 - i) b2
            t0, t1 = pm_s(x0, q, R)
            u = complex(x1.imag * T8 + x1.real, -x1.real * T8 + x1.imag)       # (1 - i t) b1
            v = complex(x3.real * T8 + x3.imag, x3.imag * T8 - x3.real)        # (t - i) b3
            p, m = u + v, u - v
        else:
            q = complex(x2.real - x2.imag, x2.imag + x2.real)                  # (1 + i) b2
            t1, t0 = pm_s(x0, q, R)
            u = complex(x1.real * T8 + x1.imag, x1.imag * T8 - x1.real)        # (t - i) b1
            v = complex(x3.imag * T8 + x3.real, -x3.real * T8 + x3.imag)       # (1 - i t) b3
            p, m = u - v, u + v
        o0, o2 = pm_s(t0, p, C8)
        o1, o3 = pm_is(t1, m, C8)
        return o0, o1, o2, o3

    out = [None] * 16
    for r in range(4):
        y = layer2(r, a[4 * r], a[4 * r + 1], a[4 * r + 2], a[4 * r + 3])
        for t in range(4):
            out[r + 4 * t] = y[t]
    return out


def stft512_frame_model(frame512):
    """512 windowed (zero-padded) real samples -> 257 complex bins, following the kernel's lanes."""
    x = np.asarray(frame512, dtype=np.float64)
    z = x[0::2] + 1j * x[1::2]
    # pass 1 + twiddle, per lane m1
    C = np.zeros((16, 16), dtype=np.complex128)
    for m1 in range(16):
        A = fft16([z[m1 + 16 * j] for j in range(16)])
        for k2 in range(16):
            C[m1][k2] = A[k2] * w(256, m1 * k2)
    # exchange 1 + pass 2, per lane k2: reg[k1] = Z[16 k1 + k2]
    reg = np.zeros((16, 16), dtype=np.complex128)
    for k2 in range(16):
        reg[k2] = fft16([C[m1][k2] for m1 in range(16)])
    X = np.zeros(257, dtype=np.complex128)

    def butterfly(a, b, k):
        """a = Z[k], b = Z[256-k] -> X[k], X[256-k] (k in 0..128)."""
        E = 0.5 * (a + np.conj(b))
        O = -0.5j * (a - np.conj(b))
        T = w(512, k) * O
        return E + T, np.conj(E - T)

    for l in range(16):
        partner = (16 - l) % 16
        send = [reg[partner][8 + j] for j in range(8)]  # what lane l receives: partner's regs 8..15
        if l == 0:
            # lane 0 pre-permutes what it sends to itself: (reg9..reg15, reg0)
            send = [reg[0][9 + j] for j in range(7)] + [reg[0][0]]
        for k1 in range(8):
            k = 16 * k1 + l
            xa, xb = butterfly(reg[l][k1], send[7 - k1], k)
            X[k] = xa
            X[256 - k] = xb
        if l == 0:  # the one extra butterfly: bin 128 pairs with itself
            xa, _ = butterfly(reg[0][8], reg[0][8], 128)
            X[128] = xa
    return X


def stftn_frame_model(xw):
    """stftn.cu: fft_length 1024 / 2048 as an in-place decimation-in-frequency FFT of the Nc = n/2 packed points in
    shared memory -- radix 16, radix 16, radix R3 -- with one pad element per M = Nc/16 elements, digit-reversed
    result, and the real-input split reading Z[k] and Z[Nc - k] through the position map.  Same index arithmetic as
    the kernel (pass loops, twiddle indices into the W_n table, padded positions)."""
    x = np.asarray(xw, dtype=np.float64)
    n = len(x)
    Nc, M = n // 2, n // 32
    R3, pitch = M // 16, M + 1
    tw = np.exp(-2j * np.pi * np.arange(n) / n)                  # the library's table W_n^k
    z = x[0::2] + 1j * x[1::2]
    work = np.zeros(16 * pitch, dtype=np.complex128)
    for j in range(M):                                            # pass 1: lane = j
        u = np.array(fft16_fma([z[j + M * s] for s in range(16)]))
        for t in range(16):
            work[j + pitch * t] = u[t] * (tw[(2 * j * t) & (n - 1)] if t else 1.0)
    for t in range(16):                                           # pass 2: lane = (t, j2)
        for j2 in range(R3):
            base = j2 + pitch * t
            v = np.array(fft16_fma([work[base + R3 * s2] for s2 in range(16)]))
            for t2 in range(16):
                work[base + R3 * t2] = v[t2] * (tw[((n // M) * j2 * t2) & (n - 1)] if t2 else 1.0)
    for t in range(16):                                           # pass 3: the contiguous groups (t, t2)
        for t2 in range(16):
            base = R3 * t2 + pitch * t
            work[base:base + R3] = np.fft.fft(work[base:base + R3])

    def zpos(k):
        return (k >> 8) + R3 * ((k >> 4) & 15) + pitch * (k & 15)

    X = np.zeros(Nc + 1, dtype=np.complex128)
    for k in range(Nc + 1):
        a, b = work[zpos(k & (Nc - 1))], work[zpos((Nc - k) & (Nc - 1))]
        sr, dr, si, di = a.real + b.real, a.real - b.real, a.imag + b.imag, a.imag - b.imag
        hx, hy = 0.5 * tw[k].real, 0.5 * tw[k].imag
        X[k] = complex(0.5 * sr + (dr * hy + si * hx), 0.5 * di + (si * hy - dr * hx))
    return X


def lpc_wave_model(frame, M, eps):
    """Windowed frame -> [K, a_1..a_M] with time-domain lag sums and the Levinson recursion."""
    x = np.asarray(frame, dtype=np.float64)
    L = len(x)
    r = np.array([np.dot(x[:L - k], x[k:]) for k in range(M + 1)])
    a = np.zeros(M + 1)
    E = r[0] + eps
    for i in range(1, M + 1):
        k = -(r[i] + np.dot(a[1:i], r[i - 1:0:-1])) / E
        a[1:i] = a[1:i] + k * a[i - 1:0:-1]
        a[i] = k
        E *= 1 - k * k
    out = a.copy()
    out[0] = np.sqrt(r[0] + np.dot(r[1:], a[1:]))
    return out


def stftn_prefetch_scatter_model(L, P, n, k_regs=80):
    """stftn_kernel's register prefetch: lane l, register i hold span element j = l + 32 i of the contiguous span
    x[s0 .. s0 + L + P); at the top of the next iteration element j goes to frame A at j (j < L) and to frame B at
    j - P (0 <= j - P < L).  Returns (A, B) index arrays into the span (-1 = zero padding) for fft_length n."""
    assert L + P <= 32 * k_regs
    A = -np.ones(n, dtype=np.int64)
    B = -np.ones(n, dtype=np.int64)
    for i in range(k_regs):
        for lane in range(32):
            j = lane + 32 * i
            if j >= L + P:
                continue
            if j < L:
                A[j] = j
            if 0 <= j - P < L:
                B[j - P] = j
    return A, B


def lpc_lagpair_autocorr_model(xw):
    """Lag sums of one windowed 400-sample frame as lpc_wave2_kernel forms them: lane l owns samples [26 l, 26 l + 26)
    and 25 ALIGNED pairs E[j] = (x[2j], x[2j+1]) of its 50-sample reach; an even sample multiplies E[t + m] into
    (r[2m], r[2m+1]), an odd sample multiplies E[t + 1 + m] into a second set (r[2m+1], r[2m+2]), its lag 0 is a
    scalar product; samples past the frame end are zeros.  Returns r[0..24]."""
    x = np.zeros(16 * 26 + 50)
    x[:400] = np.asarray(xw, dtype=np.float64)
    r = np.zeros(26)
    for l in range(16):
        E = [x[26 * l + 2 * j: 26 * l + 2 * j + 2] for j in range(25)]
        ae, ao, r0o = np.zeros((13, 2)), np.zeros((12, 2)), 0.0
        for t in range(13):
            xe, xo = E[t]
            for m in range(13):
                ae[m] += xe * E[t + m]
            for m in range(12):
                ao[m] += xo * E[t + 1 + m]
            r0o += xo * xo
        for m in range(13):
            r[2 * m] += ae[m, 0] + (ao[m - 1, 1] if m > 0 else r0o)
            r[2 * m + 1] += ae[m, 1] + (ao[m, 0] if m < 12 else 0.0)
    return r[:25]


def levinson_rolled_model(r, M, eps):
    """The rolled recursion of lpc_wave_kernel (variant bit 8): no array is indexed by the order.  Besides a[j] the
    lane keeps the reversed predictor ar[j] = a[i - j] (ar[i] = a[0] = 1), so that every order runs the same statements:
    acc = sum_m ar[m] r[m];  a'[j] = a[j] + k ar[j];  ar'[j] = ar[j-1] + k a[j-1], ar'[1] = k.  Orders 1..11 use the
    half-width body (positions <= 12).  Returns [K, a_1..a_M] like lpc_wave_model."""
    r = np.asarray(r, dtype=np.float64)
    a, ar, rd = np.zeros(25), np.zeros(25), np.zeros(25)
    rd[1:len(r)] = r[1:]
    ar[1] = 1.0
    E = r[0] + eps

    def orders(W, i0, i1):
        nonlocal E
        for _ in range(i0, i1 + 1):
            s = [0.0, 0.0, 0.0, 0.0]
            for m in range(1, W + 1):
                s[(m - 1) & 3] += ar[m] * rd[m]
            k = -((s[0] + s[1]) + (s[2] + s[3])) / E
            for j in range(W, 1, -1):          # descending: positions j - 1 are still the old ones
                ta = a[j] + k * ar[j]
                ar[j] = ar[j - 1] + k * a[j - 1]
                a[j] = ta
            a[1] = a[1] + k * ar[1]
            ar[1] = k
            E *= 1 - k * k

    orders(12, 1, min(M, 11))
    orders(24, 12, M)
    assert not a[M + 1:].any(), "positions beyond the order must stay zero"
    out = a[:M + 1].copy()
    out[0] = np.sqrt(r[0] + np.dot(a[1:], rd[1:]))
    return out


# ---------------------------------------------------------------------------------------------------------
# istft512.cu / stft512_bwd.cu: lane-level data flow of the synthesis half.  The 256-point FFT itself is the
# forward kernel's (modelled above); what is new is how its input conj(E + i O) is built from the spectrum rows,
# how the outputs map back to samples, how the backward kernel forms G = g X in the split domain and returns
# the mirrored bins to their owner lanes, and the incremental overlap-add indexing.
def istft512_frame_model(X):
    """One frame: X[0..256] complex -> 512 real samples, mirroring istft512_kernel's steps A and C."""
    tw = np.exp(-2j * np.pi * np.arange(512) / 512)
    c = np.zeros(256, complex)
    for l in range(16):           # lane within the half-warp
        for j in range(16):       # register
            k = 16 * j + l
            a, b = X[k], X[256 - k]
            if k == 0:            # irfft ignores the imaginary parts of the DC and Nyquist bins
                a, b = complex(a.real, 0), complex(b.real, 0)
            b = np.conj(b)
            E, D = a + b, a - b
            wr, wi = tw[k].real, tw[k].imag
            Or, Oi = D.real * wr + D.imag * wi, D.imag * wr - D.real * wi     # O = D conj(W512^k)
            c[16 * j + l] = complex(E.real - Oi, -(E.imag + Or))              # conj(E + i O)
    R = np.fft.fft(c)
    x = np.zeros(512)
    for l in range(16):
        for k1 in range(16):
            m = 16 * k1 + l
            x[2 * m], x[2 * m + 1] = R[m].real / 512, -R[m].imag / 512
    return x


def stft512_bwd_frame_model(xw, g):
    """Gradient of sum(g * |rfft(xw)|^2) wrt the 512 windowed samples, mirroring stft512_bwd_kernel's data flow
    (split, G = g X, inverse split, mirrored halves returned by shuffle with lane 0's special cases)."""
    tw = np.exp(-2j * np.pi * np.arange(512) / 512)
    Z = np.fft.fft(xw[0::2] + 1j * xw[1::2])          # Z[16 k1 + l] lives in lane l, register k1
    c = np.zeros(256, complex)
    received = {}
    for l in range(16):
        cm = {}
        for k1 in range(8):
            k, kp = 16 * k1 + l, 256 - (16 * k1 + l)
            z = Z[k]
            m = Z[0] if (l == 0 and k1 == 0) else Z[kp % 256]
            sr, dr, si, di = z.real + m.real, z.real - m.real, z.imag + m.imag, z.imag - m.imag
            wr, wi = tw[k].real, tw[k].imag
            tr, ti = 0.5 * (dr * wi + si * wr), 0.5 * (-dr * wr + si * wi)
            Xk, Xm = complex(0.5 * sr + tr, 0.5 * di + ti), complex(0.5 * sr - tr, -0.5 * di + ti)
            Gk, Gm = g[k] * Xk, g[kp] * Xm            # the factor 2 of d|X|^2 is folded into the final scale
            if k == 0:
                Gk, Gm = complex(2 * Gk.real, 0), complex(2 * Gm.real, 0)
            Er, Ei, Dr, Di = Gk.real + Gm.real, Gk.imag - Gm.imag, Gk.real - Gm.real, Gk.imag + Gm.imag
            Or, Oi = Dr * wr + Di * wi, Di * wr - Dr * wi
            c[k] = complex(Er - Oi, -(Ei + Or))
            cm[k1] = complex(Er + Oi, Ei - Or)        # c[256 - k], owned by lane 16 - l
        c128 = None
        if l == 0:
            G128 = g[128] * np.conj(Z[128])
            c128 = complex(2 * G128.real, 2 * G128.imag)
        for j in range(8):                            # shuffle step j: the receiver stores into register 8 + j
            s = cm[7 - j]
            if l == 0:
                s = c128 if j == 0 else cm[8 - j]
            received[((16 - l) % 16, 8 + j)] = s
    for (lane, reg), v in received.items():
        c[16 * reg + lane] = v
    R = np.fft.fft(c)
    out = np.zeros(512)
    for l in range(16):
        for k1 in range(16):
            m = 16 * k1 + l
            out[2 * m], out[2 * m + 1] = R[m].real, -R[m].imag
    return out


def overlap_add_ranges_model(L, P, N, s, T_out, tile, threads=256):
    """Yield (q, na, nb, j) per output sample from the incremental indexing of the fused kernels' gather loop."""
    c, m = (L - 1) // P, (L - 1) % P
    for t0 in range(0, T_out, tile):
        t1, q0 = min(t0 + tile, T_out), t0 + s
        for tid in range(threads):
            q_first = q0 + tid
            ne, r = q_first // P, q_first % P
            dq, dr = threads // P, threads % P
            i = tid
            while i < t1 - t0:
                na = max(ne - c + (1 if r > m else 0), 0)
                nb = min(ne, N - 1)
                yield q0 + i, na, nb, r + (ne - na) * P
                ne, r = ne + dq, r + dr
                if r >= P:
                    r, ne = r - P, ne + 1
                i += threads


# stft512.cu, variant kVPair2: index model of the (f, f + 2) frame pairing and of the staged-row store order.
def stft512_pair2_columns(P=80, L=400, NJ=13, shift=5):
    """For every half-warp h, lane l and column j: the span offsets the kernel reads for frames A = h and B = h + 2
    (as raw[j] and raw[j + shift] of the SAME lane) next to the offsets the frames really start at."""
    out = []
    for h in range(2):
        for l in range(16):
            base = h * P + 2 * l                      # pa = span + hf * P + 2 * l
            for j in range(NJ):
                read_a = base + 32 * j                # raw[j]
                read_b = base + 32 * (j + shift)      # raw[j + shift]
                want_a = h * P + 2 * l + 32 * j       # sample 2 l + 32 j of frame h
                want_b = (h + 2) * P + 2 * l + 32 * j  # the same sample of frame h + 2
                out.append((read_a, want_a, read_b, want_b))
    return out


def stft512_staged_store_banks(d):
    """Banks touched by each of the four store instructions of one `u` step of the staged split (stft512.cu):
    d = bank distance between the two half-warps' rows (2: frames (2h, 2h+1); 1: frames (h, h+2))."""
    rows = []
    for u in range(4):
        for which in range(4):
            banks = []
            for h in range(2):
                for l in range(16):
                    rowA = h * (257 * (2 if d == 2 else 1))
                    swF = bool(h) and l < 16 - d
                    swM = bool(h) and l >= d
                    if which == 0:
                        addr = rowA + l + (16 if swF else 0) + 32 * u
                    elif which == 1:
                        addr = rowA + l - (16 if swF else 0) + 32 * u + 16
                    elif which == 2:
                        addr = rowA + 256 - l - (16 if swM else 0) - 32 * u
                    else:
                        addr = rowA + 256 - l + (16 if swM else 0) - 32 * u - 16
                    banks.append(addr % 32)
            rows.append(banks)
    return rows


# lsp.cu: the numerical steps of lpc2lsp_kernel for one row (deflation, Chebyshev series, grid + refinement in x).
def lsp_model(row, G=None):
    row = np.asarray(row, dtype=np.float64)
    M = row.size - 1
    if M == 0:
        return np.zeros(0)
    if G is None:
        G = min(2048, max(256, 32 * M))
    a1 = np.concatenate([[1.0], row[1:], [0.0]])
    p, q = a1 - a1[::-1], a1 + a1[::-1]
    if M % 2 == 0:
        nP = nQ = M // 2
        for i in range(1, M + 1):
            p[i] += p[i - 1]
            q[i] -= q[i - 1]
    else:
        nP, nQ = (M - 1) // 2, (M + 1) // 2
        for i in range(2, M):
            p[i] += p[i - 2]

    def cheb(g, n, x):
        b1 = b2 = 0.0
        for k in range(n, 0, -1):
            b1, b2 = 2.0 * x * b1 + (g[k] - b2), b1
        return x * b1 + (g[0] - b2)

    out = []
    for c, n in ((p, nP), (q, nQ)):
        if n == 0:
            continue
        g = np.array([(1.0 if k == 0 else 2.0) * c[n - k] for k in range(n + 1)])
        brackets = []
        for fine in (1, 16):
            Gn = G * fine
            xs = np.cos(np.pi * np.arange(Gn + 1) / Gn)
            fs = np.array([cheb(g, n, x) for x in xs])
            idx = np.nonzero((fs[:-1] < 0) != (fs[1:] < 0))[0]
            brackets = [(xs[i], xs[i + 1]) for i in idx]
            if len(brackets) == n:
                break
        for xa, xb in brackets[:n]:
            fa, fb = cheb(g, n, xa), cheb(g, n, xb)
            for _ in range(4):                       # bisection
                xm = 0.5 * (xa + xb)
                fm = cheb(g, n, xm)
                if (fm < 0) == (fa < 0):
                    xa, fa = xm, fm
                else:
                    xb, fb = xm, fm
            for _ in range(12):                      # regula falsi with the Illinois damping
                d = fb - fa
                if d == 0:
                    break
                xc = (xa * fb - xb * fa) / d
                fc = cheb(g, n, xc)
                if (fc < 0) != (fb < 0):
                    xa, fa = xb, fb
                else:
                    fa *= 0.5
                xb, fb = xc, fc
            out.append(np.arccos(xb))
        out += [np.nan] * (n - min(n, len(brackets)))
    return np.sort(np.array(out))


# gc2gc.cu: both transforms as trigonometric series evaluated by Clenshaw recurrences in float64 (one recurrence
# yields the cosine AND the sine sum of the forward transform), half-spectrum weights in the inverse.
def gc2gc_model(c1, out_order, g1, g2, n):
    c1 = np.asarray(c1, dtype=np.float64)
    D1, D2, K = c1.size, out_order + 1, n // 2 + 1
    C2 = np.zeros(K)
    for k in range(K):
        x, sn = np.cos(2 * np.pi * k / n), np.sin(2 * np.pi * k / n)
        b1 = b2 = 0.0
        for m in range(D1 - 1, 0, -1):              # b_m = c_m + 2 x b_{m+1} - b_{m+2}
            b1, b2 = c1[m] + 2 * x * b1 - b2, b1
        re, im = x * b1 - b2, -sn * b1              # sum c_m cos(m t), -sum c_m sin(m t)
        if g1 == 0:
            mag, ang = np.exp(re), np.arctan2(np.sin(im), np.cos(im))
        else:
            zr, zi = 1 + g1 * re, g1 * im
            mag = np.hypot(zr, zi) ** (1 / g1)
            th = np.arctan2(zi, zr) / g1
            ang = np.arctan2(np.sin(th), np.cos(th))
        C2[k] = np.log(mag) if g2 == 0 else (mag ** g2 * np.cos(ang * g2) - 1) / g2
    a = C2.copy()
    a[1:] *= 2
    if n % 2 == 0:
        a[K - 1] *= 0.5
    out = np.zeros(D2)
    out[0] = c1[0]
    for m in range(1, D2):
        x = np.cos(2 * np.pi * m / n)
        b1 = b2 = 0.0
        for k in range(K - 1, 0, -1):
            b1, b2 = a[k] + 2 * x * b1 - b2, b1
        out[m] = 2 * (a[0] + x * b1 - b2) / n
    return out
