// Warp-cooperative FFT building blocks shared by the generic forward kernels (spectral.cu) and the
// backward kernels (spectral_bwd.cu): Stockham radix-2 in shared memory, table-driven direct DFT, the
// real-input split and the spectrum formatters.
#pragma once

#include "common.cuh"

namespace dsb200 {

// Stockham autosort radix-2 FFT of Nc complex points held in shared memory by one warp.
// Returns the buffer that holds the natural-order result.
// `tw_n` is the number of points on the unit circle the table samples: tw[k] = exp(-2 pi i k / tw_n);
// the real-FFT packing uses tw_n = 2 Nc, a plain complex FFT of Nc points uses tw_n = Nc.
template <typename T>
__device__ cx_t<T>* warp_fft_pow2(cx_t<T>* in, cx_t<T>* out, int Nc, const cx_t<T>* tw, int lane, int tw_n) {
  const int half = Nc >> 1;
  for (int Ns = 1; Ns < Nc; Ns <<= 1) {
    const int tstride = tw_n / (2 * Ns);
    for (int j = lane; j < half; j += 32) {
      const int k = j & (Ns - 1);
      const cx_t<T> w = tw[k * tstride];
      const cx_t<T> u = in[j];
      const cx_t<T> v = cmul(in[j + half], w);
      const int j0 = ((j - k) << 1) + k;
      out[j0] = cadd(u, v);
      out[j0 + Ns] = csub(u, v);
    }
    __syncwarp();
    cx_t<T>* t = in; in = out; out = t;
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
