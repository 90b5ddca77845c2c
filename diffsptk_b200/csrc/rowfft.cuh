// Warp-cooperative FFT building blocks shared by the generic forward kernels (spectral.cu) and the
// backward kernels (spectral_bwd.cu): Stockham radix-2 in shared memory, table-driven direct DFT, the
// real-input split and the spectrum formatters.
#pragma once

#include "common.cuh"

namespace dsb200 {

// Stockham autosort FFT of Nc = 2^m complex points held in shared memory by one warp: radix-4 passes (half the
// shared-memory round trips of radix-2), then one radix-2 pass when m is odd.
// Returns the buffer that holds the natural-order result.
// `tw_n` is the number of points on the unit circle the table samples: tw[k] = exp(-2 pi i k / tw_n);
// the real-FFT packing uses tw_n = 2 Nc, a plain complex FFT of Nc points uses tw_n = Nc.
template <typename T>
__device__ cx_t<T>* warp_fft_pow2(cx_t<T>* in, cx_t<T>* out, int Nc, const cx_t<T>* tw, int lane, int tw_n) {
  using C = cx_t<T>;
  int Ns = 1;
  const int quarter = Nc >> 2;
  for (; Ns * 4 <= Nc; Ns <<= 2) {
    // y[(j - k) 4 + k + t Ns] = sum_s (-i)^(s t) w^(s k) x[j + s Nc/4],  w = exp(-2 pi i / (4 Ns)), k = j mod Ns
    const int tstride = tw_n / (4 * Ns);
    for (int j = lane; j < quarter; j += 32) {
      const int k = j & (Ns - 1);
      const C a = in[j];
      C b = in[j + quarter], c = in[j + 2 * quarter], d = in[j + 3 * quarter];
      if (Ns > 1) {  // k == 0 in the first pass: all twiddles are 1
        b = cmul(b, tw[k * tstride]);
        c = cmul(c, tw[2 * k * tstride]);
        d = cmul(d, tw[3 * k * tstride]);
      }
      const C ac = cadd(a, c), amc = csub(a, c), bd = cadd(b, d), bmd = csub(b, d);
      const int j0 = ((j - k) << 2) + k;
      out[j0] = cadd(ac, bd);
      out[j0 + Ns] = mk<T>(amc.x + bmd.y, amc.y - bmd.x);       // a - i b - c + i d
      out[j0 + 2 * Ns] = csub(ac, bd);
      out[j0 + 3 * Ns] = mk<T>(amc.x - bmd.y, amc.y + bmd.x);   // a + i b - c - i d
    }
    __syncwarp();
    C* t = in; in = out; out = t;
  }
  if (Ns < Nc) {  // one radix-2 pass left (Ns == Nc / 2)
    const int half = Nc >> 1;
    const int tstride = tw_n / (2 * Ns);
    for (int j = lane; j < half; j += 32) {
      const int k = j & (Ns - 1);
      const C u = in[j];
      const C v = cmul(in[j + half], tw[k * tstride]);
      const int j0 = ((j - k) << 1) + k;
      out[j0] = cadd(u, v);
      out[j0 + Ns] = csub(u, v);
    }
    __syncwarp();
    C* t = in; in = out; out = t;
  }
  return in;
}

// Direct DFT of `len` real samples (zero beyond) to bins 0..Nc, table-driven, one warp.
template <typename T>
__device__ void warp_dft_direct(const T* xin, int len, cx_t<T>* X, int n, int Nc, const cx_t<T>* tw, int lane) {
  for (int k = lane; k <= Nc; k += 32) {
    T re = 0, im = 0;
    int idx = 0;
    for (int j = 0; j < len; ++j) {
      const cx_t<T> w = tw[idx];
      re = dfma(xin[j], w.x, re);
      im = dfma(xin[j], w.y, im);
      idx += k;
      if (idx >= n) idx -= n;
    }
    X[k] = mk<T>(re, im);
  }
  __syncwarp();
}

// Bin k of the length-n real FFT from the length-Nc complex FFT Z of the even/odd packing.
template <typename T>
__device__ __forceinline__ cx_t<T> real_split(const cx_t<T>* Z, int k, int Nc, const cx_t<T>* tw) {
  const int k1 = (k == Nc) ? 0 : k;
  const int k2 = (k == 0 || k == Nc) ? 0 : Nc - k;
  const cx_t<T> zk = Z[k1];
  cx_t<T> zc = Z[k2];
  zc.y = -zc.y;
  const T half = static_cast<T>(0.5);
  const cx_t<T> E = mk<T>(half * (zk.x + zc.x), half * (zk.y + zc.y));
  const cx_t<T> O = mk<T>(half * (zk.y - zc.y), -half * (zk.x - zc.x));
  const cx_t<T> w = (k == Nc) ? mk<T>(static_cast<T>(-1), static_cast<T>(0)) : tw[k];
  return cadd(E, cmul(w, O));
}

template <typename T>
__device__ __forceinline__ T spec_format(T s, int fmt) {
  switch (fmt) {
    case DSB200_SPEC_DB: return static_cast<T>(10) * dlog10(s);
    case DSB200_SPEC_LOGMAG: return static_cast<T>(0.5) * dlog(s);
    case DSB200_SPEC_MAGNITUDE: return dsqrt(s);
    default: return s;
  }
}

}  // namespace dsb200
