// Autocorrelation, Levinson-Durbin and LPC kernels (generic path, fp32 + fp64, any order).
//
// Reference semantics: diffsptk/modules/acorr.py:110-120 (linear autocorrelation, lags 0..M, the
// reference computes it through a zero-padded FFT), levdur.py:113-127 (dense solve of the
// regularised Yule-Walker system; the Levinson recursion on r with r0+eps is the same system),
// lpc.py:137-139.  The fused waveform->LPC kernel for the BASELINE configuration lives in
// lpc_wave.cu; this file keeps every (L, M) of the reference API on the GPU.
#include <algorithm>

#include "common.cuh"

namespace dsb200 {
namespace {

// ---- autocorrelation: one warp per row, row staged in shared memory ---------------------
template <typename T>
struct AcorrArgs {
  const T* x;       // framed rows [rows,L]  (wave == 0)  or waveform [batch,T] (wave == 1)
  const T* window;  // wave only
  T* r;             // [rows, ld_out] (first M+1 entries of each row are written)
  int64_t rows, T_len, n_frames;
  int L, M, ld_out, out_format;
  int wave, P, left, zmean, pad_mode;
};

template <typename T>
__global__ void __launch_bounds__(256) acorr_kernel(AcorrArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int Lp = A.L + A.M;  // zero tail so that x[n+k] never leaves the buffer
  T* xs = reinterpret_cast<T*>(smem_raw) + static_cast<size_t>(warp) * Lp;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + warp; row < A.rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    if (A.wave) {
      const int64_t b = row / A.n_frames, i = row - b * A.n_frames;
      const T* xb = A.x + b * A.T_len;
      const int64_t start = i * A.P - A.left;
      T mean = 0;
      if (A.zmean) {
        T acc = 0;
        for (int j = lane; j < A.L; j += 32) {
          const int64_t q = pad_index(start + j, A.T_len, A.pad_mode);
          acc += q < 0 ? static_cast<T>(0) : xb[q];
        }
        mean = warp_sum(acc) / static_cast<T>(A.L);
      }
      for (int j = lane; j < Lp; j += 32) {
        T v = 0;
        if (j < A.L) {
          const int64_t q = pad_index(start + j, A.T_len, A.pad_mode);
          v = ((q < 0 ? static_cast<T>(0) : xb[q]) - mean) * A.window[j];
        }
        xs[j] = v;
      }
    } else {
      const T* xr = A.x + row * A.L;
      for (int j = lane; j < Lp; j += 32) xs[j] = j < A.L ? xr[j] : static_cast<T>(0);
    }
    __syncwarp();
    T r0 = 0;
    for (int k0 = 0; k0 <= A.M; k0 += 32) {
      const int k = k0 + lane;
      // float64 accumulators for both element types: the LPC systems built from these sums are
      // ill-conditioned on speech, and a 400-term float32 chain costs ~10x the reference's FFT route
      // in accuracy; this generic kernel is not the throughput path (see fused_wave.cu).
      double acc0 = 0, acc1 = 0;
      if (k <= A.M) {
        int n = 0;
        for (; n + 1 < A.L; n += 2) {
          acc0 = fma(static_cast<double>(xs[n]), static_cast<double>(xs[n + k]), acc0);
          acc1 = fma(static_cast<double>(xs[n + 1]), static_cast<double>(xs[n + 1 + k]), acc1);
        }
        if (n < A.L) acc0 = fma(static_cast<double>(xs[n]), static_cast<double>(xs[n + k]), acc0);
      }
      T v = static_cast<T>(acc0 + acc1);
      if (k0 == 0) r0 = __shfl_sync(0xffffffffu, v, 0);
      if (k <= A.M) {
        switch (A.out_format) {
          case DSB200_ACORR_NORMALIZED: v = v / r0; break;
          case DSB200_ACORR_BIASED: v = v / static_cast<T>(A.L); break;
          case DSB200_ACORR_UNBIASED: v = v / static_cast<T>(A.L - k); break;
          default: break;
        }
        A.r[row * A.ld_out + k] = v;
      }
    }
    __syncwarp();
  }
}

template <typename T>
int launch_acorr(AcorrArgs<T>& A, int device, cudaStream_t stream) {
  if (A.rows == 0) return DSB200_OK;
  const size_t per_warp = static_cast<size_t>(A.L + A.M) * sizeof(T);
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  if (per_warp > cap) return fail(DSB200_E_UNSUPPORTED, "frame_length=%d is too long for the shared-memory acorr kernel", A.L);
  int wpb = static_cast<int>(std::min<size_t>(8, cap / per_warp));
  while (wpb > 1 && wpb * per_warp > 48 * 1024) --wpb;
  DSB_CUDA(cudaFuncSetAttribute(acorr_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int64_t need = (A.rows + wpb - 1) / wpb;
  const int blocks = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(sm_count(device)) * 16));
  acorr_kernel<T><<<blocks, wpb * 32, wpb * per_warp, stream>>>(A);
  return after_launch("acorr_kernel");
}

// ---- Levinson-Durbin: one thread per row, state in shared memory [j][thread] ----------------
//   rho = r, rho0 = r0 + eps;  E = rho0
//   for i = 1..M:  k = -(rho_i + sum_{j<i} a_j rho_{i-j}) / E;  a_j += k a_{i-j};  a_i = k;  E *= 1 - k^2
//   K = sqrt(r0 + sum_j r_j a_j)          (un-regularised r0, levdur.py:124)
// The recursion state is float64 for both element types: on speech-like (ill-conditioned) frames a
// float32 recursion loses ~4x more than the reference's pivoted float32 LU does, float64 state keeps
// the result at the accuracy of the float32 autocorrelation itself (DFMA runs at half FP32 rate on
// B200 and this kernel is ~5 % of the LPC path).  In-place safe: a row is read before it is written.
template <typename T>
__global__ void levdur_kernel(const T* r, T* out, int64_t rows, int M, double eps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tpb = blockDim.x, ld = tpb + 1, D = M + 1;
  double* rs = reinterpret_cast<double*>(smem_raw);  // [D][ld]
  double* as = rs + static_cast<size_t>(D) * ld;     // [D][ld]
  double* ts = as + static_cast<size_t>(D) * ld;     // [D][ld]
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * tpb; base < rows; base += static_cast<int64_t>(gridDim.x) * tpb) {
    const int nrow = static_cast<int>(rows - base < tpb ? rows - base : tpb);
    for (int idx = threadIdx.x; idx < nrow * D; idx += tpb) {
      const int rl = idx / D, j = idx - rl * D;
      rs[j * ld + rl] = static_cast<double>(r[base * D + idx]);
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t < nrow) {
      const double r0 = rs[t];
      double E = r0 + eps;
      for (int i = 1; i <= M; ++i) {
        double acc = rs[i * ld + t];
        for (int j = 1; j < i; ++j) acc = fma(as[j * ld + t], rs[(i - j) * ld + t], acc);
        const double k = -acc / E;
        for (int j = 1; j < i; ++j) ts[j * ld + t] = fma(k, as[(i - j) * ld + t], as[j * ld + t]);
        for (int j = 1; j < i; ++j) as[j * ld + t] = ts[j * ld + t];
        as[i * ld + t] = k;
        E *= (1.0 - k * k);
      }
      double g = r0;
      for (int j = 1; j <= M; ++j) g = fma(rs[j * ld + t], as[j * ld + t], g);
      as[t] = sqrt(g);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nrow * D; idx += tpb) {
      const int rl = idx / D, j = idx - rl * D;
      out[base * D + idx] = static_cast<T>(as[j * ld + rl]);
    }
    __syncthreads();
  }
}

template <typename T>
int launch_levdur(const T* r, T* out, int64_t rows, int M, double eps, int device, cudaStream_t stream) {
  if (rows == 0) return DSB200_OK;
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  const size_t D = static_cast<size_t>(M) + 1;
  int tpb = 128;
  auto bytes = [&](int t) { return 3 * D * (t + 1) * sizeof(double); };
  while (tpb > 32 && bytes(tpb) > std::min<size_t>(cap, 96 * 1024)) tpb -= 32;
  if (bytes(tpb) > cap) return fail(DSB200_E_UNSUPPORTED, "lpc_order=%d is too large for the shared-memory Levinson kernel", M);
  DSB_CUDA(cudaFuncSetAttribute(levdur_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int64_t need = (rows + tpb - 1) / tpb;
  const int blocks = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(sm_count(device)) * 8));
  levdur_kernel<T><<<blocks, tpb, bytes(tpb), stream>>>(r, out, rows, M, eps);
  return after_launch("levdur_kernel");
}

template <typename T>
int acorr_impl(const void* x, void* r, int64_t rows, int32_t L, int32_t M, int32_t fmt, int device, void* stream) {
  DSB_REQUIRE(L > 0, "frame_length must be positive.");
  DSB_REQUIRE(M >= 0 && M < L, "acr_order must be less than frame_length.");
  DSB_REQUIRE(fmt >= DSB200_ACORR_NAIVE && fmt <= DSB200_ACORR_UNBIASED, "out_format %d is not supported.", fmt);
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(x != nullptr && r != nullptr, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  AcorrArgs<T> A{};
  A.x = static_cast<const T*>(x);
  A.r = static_cast<T*>(r);
  A.rows = rows;
  A.L = L;
  A.M = M;
  A.ld_out = M + 1;
  A.out_format = fmt;
  return launch_acorr<T>(A, device, static_cast<cudaStream_t>(stream));
}

template <typename T>
int levdur_impl(const void* r, void* a, int64_t rows, int32_t M, double eps, int device, void* stream) {
  DSB_REQUIRE(M >= 0, "lpc_order must be non-negative.");
  DSB_REQUIRE(eps >= 0, "eps must be non-negative.");
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(r != nullptr && a != nullptr, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  return launch_levdur<T>(static_cast<const T*>(r), static_cast<T*>(a), rows, M, eps, device, static_cast<cudaStream_t>(stream));
}

template <typename T>
int lpc_impl(const void* x, void* a, int64_t rows, int32_t L, int32_t M, double eps, int device, void* stream) {
  // acorr straight into the output buffer, then the recursion in place (no scratch allocation).
  if (int rc = acorr_impl<T>(x, a, rows, L, M, DSB200_ACORR_NAIVE, device, stream)) return rc;
  return levdur_impl<T>(a, a, rows, M, eps, device, stream);
}

}  // namespace

// Fast fused kernel for the BASELINE configuration (lpc_wave.cu); DSB200_E_UNSUPPORTED outside it.
int lpc_wave_fast_try(const float* x, const float* window, float* a, int64_t batch, int64_t T_len,
                      const dsb200_frame_params* fp, int32_t M, double eps, int device, cudaStream_t stream);

template <typename T>
int lpc_wave_impl(const void* x, const void* window, void* a, int64_t batch, int64_t T_len,
                  const dsb200_frame_params* fp, int32_t M, double eps, int device, void* stream) {
  DSB_REQUIRE(fp != nullptr, "frame params are NULL");
  DSB_REQUIRE(fp->frame_length > 0, "frame_length must be positive.");
  DSB_REQUIRE(fp->frame_period > 0, "frame_period must be positive.");
  DSB_REQUIRE(M >= 0 && M < fp->frame_length, "acr_order must be less than frame_length.");
  DSB_REQUIRE(eps >= 0, "eps must be non-negative.");
  DSB_REQUIRE(T_len >= 1, "waveform length must be at least 1");
  DSB_REQUIRE(batch >= 0, "batch must be non-negative");
  {
    const int L = fp->frame_length;
    const int left = fp->center ? L / 2 : 0, right = fp->center ? (L - 1) / 2 : L - 1;
    const int big = left > right ? left : right;
    if (fp->pad_mode == DSB200_PAD_REFLECT) DSB_REQUIRE(big < T_len, "reflect padding must be smaller than the waveform length");
    if (fp->pad_mode == DSB200_PAD_CIRCULAR) DSB_REQUIRE(big <= T_len, "circular padding must not exceed the waveform length");
  }
  if (batch == 0) return DSB200_OK;
  DSB_REQUIRE(x != nullptr && a != nullptr && window != nullptr, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  if (sizeof(T) == 4) {
    const int rc = lpc_wave_fast_try(static_cast<const float*>(x), static_cast<const float*>(window),
                                     static_cast<float*>(a), batch, T_len, fp, M, eps, device,
                                     static_cast<cudaStream_t>(stream));
    if (rc != DSB200_E_UNSUPPORTED) return rc;
  }
  AcorrArgs<T> A{};
  A.x = static_cast<const T*>(x);
  A.window = static_cast<const T*>(window);
  A.r = static_cast<T*>(a);
  A.T_len = T_len;
  A.n_frames = dsb200_num_frames(T_len, fp->frame_period);
  A.rows = batch * A.n_frames;
  A.L = fp->frame_length;
  A.M = M;
  A.ld_out = M + 1;
  A.out_format = DSB200_ACORR_NAIVE;
  A.wave = 1;
  A.P = fp->frame_period;
  A.left = fp->center ? fp->frame_length / 2 : 0;
  A.zmean = fp->zmean;
  A.pad_mode = fp->pad_mode;
  if (int rc = launch_acorr<T>(A, device, static_cast<cudaStream_t>(stream))) return rc;
  return launch_levdur<T>(static_cast<const T*>(a), static_cast<T*>(a), A.rows, M, eps, device, static_cast<cudaStream_t>(stream));
}

}  // namespace dsb200

using namespace dsb200;

extern "C" {

int dsb200_acorr_f32(const void* x, void* r, int64_t rows, int32_t L, int32_t M, int32_t fmt, int device, void* stream) {
  return acorr_impl<float>(x, r, rows, L, M, fmt, device, stream);
}
int dsb200_acorr_f64(const void* x, void* r, int64_t rows, int32_t L, int32_t M, int32_t fmt, int device, void* stream) {
  return acorr_impl<double>(x, r, rows, L, M, fmt, device, stream);
}
int dsb200_levdur_f32(const void* r, void* a, int64_t rows, int32_t M, double eps, int device, void* stream) {
  return levdur_impl<float>(r, a, rows, M, eps, device, stream);
}
int dsb200_levdur_f64(const void* r, void* a, int64_t rows, int32_t M, double eps, int device, void* stream) {
  return levdur_impl<double>(r, a, rows, M, eps, device, stream);
}
int dsb200_lpc_f32(const void* x, void* a, int64_t rows, int32_t L, int32_t M, double eps, int device, void* stream) {
  return lpc_impl<float>(x, a, rows, L, M, eps, device, stream);
}
int dsb200_lpc_f64(const void* x, void* a, int64_t rows, int32_t L, int32_t M, double eps, int device, void* stream) {
  return lpc_impl<double>(x, a, rows, L, M, eps, device, stream);
}
int dsb200_lpc_wave_f32(const void* x, const void* w, void* a, int64_t batch, int64_t T, const dsb200_frame_params* fp,
                        int32_t M, double eps, int device, void* stream) {
  return lpc_wave_impl<float>(x, w, a, batch, T, fp, M, eps, device, stream);
}
int dsb200_lpc_wave_f64(const void* x, const void* w, void* a, int64_t batch, int64_t T, const dsb200_frame_params* fp,
                        int32_t M, double eps, int device, void* stream) {
  return lpc_wave_impl<double>(x, w, a, batch, T, fp, M, eps, device, stream);
}

}  // extern "C"
