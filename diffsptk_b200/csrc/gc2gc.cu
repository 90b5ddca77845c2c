// Generalized cepstrum -> generalized cepstrum (gamma conversion of gain-normalised cepstra) in ONE kernel.
// Reference: GeneralizedCepstrumToGeneralizedCepstrum._forward, diffsptk/modules/mgc2mgc.py:327-364:
//     C1 = fft([0, c1_1..c1_M1], n)            (the spectrum of the cepstrum without its gain term)
//     sC1 = exp(C1)            (in_gamma = 0)   or  (1 + g1 C1)^(1/g1)   (polar form)
//     C2 = log|sC1|            (out_gamma = 0)  or  (|sC1|^g2 cos(g2 arg sC1) - 1) / g2
//     c2 = [c1_0, 2 ifft(C2).real[1..M2]]
// The input has M1 <= ~50 non-zero terms and only M2 + 1 outputs are kept, so both transforms are evaluated
// DIRECTLY as trigonometric series, by Clenshaw recurrences in float64: per bin one recurrence over the M1 input
// terms gives the cosine and the sine sum at once (M1 DFMA), per output one recurrence over the n/2+1 bins (the
// operands are broadcast reads of the warp's shared memory; the first version looked twiddles up by index and was
// bound by the bank conflicts of those look-ups: 9.8 ms per 1 024 000 rows).  Fewer operations than two length-n
// FFTs for the usual orders, no power-of-two or even-length restriction, float64 accuracy for float32 rows, no
// intermediate in HBM.  One warp per row: lanes own the bins k = lane + 32 t of the forward transform and the
// pointwise map, park the weighted, real, even C2 in shared memory, then own the outputs m = lane + 32 t.
#include <algorithm>

#include "common.cuh"

namespace dsb200 {
namespace {

__device__ __forceinline__ float datan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double datan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ void dsincos(float a, float* s, float* c) { sincosf(a, s, c); }
__device__ __forceinline__ void dsincos(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ float dcos(float a) { return cosf(a); }
__device__ __forceinline__ double dcos(double a) { return cos(a); }

template <typename T>
__global__ void __launch_bounds__(256) gc2gc_kernel(const T* __restrict__ c1, T* __restrict__ c2, int64_t rows, int D1,
                                                    int D2, T g1, T g2, int n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int K = n / 2 + 1;                        // bins 0..floor(n/2)
  double* xc = reinterpret_cast<double*>(smem_raw);   // [K] cos(2 pi k / n)
  double* xs = xc + K;                                // [K] sin(2 pi k / n)
  for (int k = threadIdx.x; k < K; k += blockDim.x) sincospi(2.0 * k / n, &xs[k], &xc[k]);
  double* cin = xs + K + static_cast<size_t>(warp) * (D1 + K);   // [D1] input row (float64)
  double* A2 = cin + D1;                                         // [K]  weighted mapped spectrum w_k C2[k]
  __syncthreads();
  const T one = static_cast<T>(1), zero = static_cast<T>(0);
  const double inv_n = 1.0 / n;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + warp; row < rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    for (int i = lane; i < D1; i += 32) cin[i] = static_cast<double>(c1[row * D1 + i]);
    __syncwarp();
    // ---- forward transform of [0, c_1..c_M1] at bin k: ONE Clenshaw recurrence b_m = c_m + 2 x b_{m+1} - b_{m+2}
    //      yields sum c_m cos(m t) = x b_1 - b_2 and sum c_m sin(m t) = sin(t) b_1; then the pointwise map -----------
    for (int k = lane; k < K; k += 32) {
      const double x = xc[k], x2 = 2.0 * x;
      double b1 = 0.0, b2 = 0.0;
      for (int m = D1 - 1; m >= 1; --m) {           // cin[m]: the same address in every lane (broadcast)
        const double b0 = fma(x2, b1, cin[m] - b2);
        b2 = b1;
        b1 = b0;
      }
      const T re = static_cast<T>(fma(x, b1, -b2)), im = static_cast<T>(-xs[k] * b1);
      T mag, ang;                                   // |sC1| and arg sC1 in (-pi, pi]
      if (g1 == zero) {
        mag = dexp(re);
        T s, c;
        dsincos(im, &s, &c);
        ang = datan2(s, c);
      } else {
        const T zr = dfma(g1, re, one), zi = g1 * im;
        mag = dpow(dsqrt(zr * zr + zi * zi), one / g1);
        T s, c;
        dsincos(datan2(zi, zr) / g1, &s, &c);
        ang = datan2(s, c);
      }
      const T v = (g2 == zero) ? dlog(mag) : (dpow(mag, g2) * dcos(ang * g2) - one) / g2;
      A2[k] = ((k == 0 || 2 * k == n) ? 1.0 : 2.0) * static_cast<double>(v);
    }
    __syncwarp();
    // ---- inverse transform of the real, even C2: c[m] = (1/n) sum_k w_k C2[k] cos(k t_m), a cosine series in
    //      x = cos(t_m) by the same recurrence (A2[k] is a broadcast read) ---------------------------------------
    T* out = c2 + row * D2;
    for (int m = lane; m < D2; m += 32) {
      if (m == 0) {
        out[0] = static_cast<T>(cin[0]);
        continue;
      }
      const int mm = (m < K) ? m : n - m;           // cos(2 pi m / n) = cos(2 pi (n - m) / n)
      const double x = xc[mm], x2 = 2.0 * x;
      double b1 = 0.0, b2 = 0.0;
      for (int k = K - 1; k >= 1; --k) {
        const double b0 = fma(x2, b1, A2[k] - b2);
        b2 = b1;
        b1 = b0;
      }
      out[m] = static_cast<T>(2.0 * inv_n * (fma(x, b1, -b2) + A2[0]));
    }
    __syncwarp();
  }
}

template <typename T>
int gc2gc_impl(const void* c1, void* c2, int64_t rows, int32_t in_order, int32_t out_order, double in_gamma,
               double out_gamma, int32_t n_fft, int device, void* stream) {
  DSB_REQUIRE(in_order >= 0, "in_order must be non-negative.");
  DSB_REQUIRE(out_order >= 0, "out_order must be non-negative.");
  DSB_REQUIRE(in_gamma >= -1.0 && in_gamma <= 1.0, "in_gamma must be in [-1, 1].");
  DSB_REQUIRE(out_gamma >= -1.0 && out_gamma <= 1.0, "out_gamma must be in [-1, 1].");
  DSB_REQUIRE(n_fft > std::max(in_order, out_order) + 1, "n_fft must be much larger than order of cepstrum.");
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(c1 != nullptr && c2 != nullptr, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  const int D1 = in_order + 1, D2 = out_order + 1, K = n_fft / 2 + 1;
  const size_t fixed = static_cast<size_t>(2 * K) * sizeof(double), per_warp = static_cast<size_t>(D1 + K) * sizeof(double);
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  if (fixed + per_warp > cap) return fail(DSB200_E_UNSUPPORTED, "n_fft=%d does not fit in shared memory", n_fft);
  const int wpb = static_cast<int>(std::min<size_t>(8, (cap - fixed) / per_warp));
  DSB_CUDA(cudaFuncSetAttribute(gc2gc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int64_t need = (rows + wpb - 1) / wpb;
  const int blocks = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(sm_count(device)) * 8));
  gc2gc_kernel<T><<<blocks, wpb * 32, fixed + wpb * per_warp, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const T*>(c1), static_cast<T*>(c2), rows, D1, D2, static_cast<T>(in_gamma), static_cast<T>(out_gamma),
      n_fft);
  return after_launch("gc2gc_kernel");
}

}  // namespace
}  // namespace dsb200

extern "C" {

int dsb200_gc2gc_f32(const void* c1, void* c2, int64_t rows, int32_t in_order, int32_t out_order, double in_gamma,
                     double out_gamma, int32_t n_fft, int device, void* stream) {
  return dsb200::gc2gc_impl<float>(c1, c2, rows, in_order, out_order, in_gamma, out_gamma, n_fft, device, stream);
}
int dsb200_gc2gc_f64(const void* c1, void* c2, int64_t rows, int32_t in_order, int32_t out_order, double in_gamma,
                     double out_gamma, int32_t n_fft, int device, void* stream) {
  return dsb200::gc2gc_impl<double>(c1, c2, rows, in_order, out_order, in_gamma, out_gamma, n_fft, device, stream);
}

}  // extern "C"
