// Vector-Jacobian products of the feature ops behind the spectrum: mel filter bank (fbank.py:305-330),
// autocorrelation (acorr.py:112-121) and the Levinson-Durbin solve (levdur.py:113-127).
// the reference is differentiable end to end, so the drop-in needs the adjoints of
// its fused kernels.  These are coverage kernels (one warp or one thread per row, float64 recursion state),
// not tuned ones; nothing is saved by the forward pass, intermediate values are recomputed per row.
#include <algorithm>

#include "common.cuh"

namespace dsb200 {
namespace {

// ------------------------------------------------------------------------------------ fbank
// Forward per row: amp = x or sqrt(x); z = amp @ H; y = log(max(z, floor)) or (max(z, floor)^g - 1) / g;
// E = log((x_0 + x_{K-1} + 2 sum_{0<k<K-1} x_k) / fft_length).
template <typename T>
struct FbankBwdArgs {
  const T* x;    // [rows, K]
  const T* H;    // [K, C]
  const int32_t* cb;
  const int32_t* ce;
  const T* gy;   // [rows, C]
  const T* gE;   // [rows] or null
  T* gx;         // [rows, K]
  int64_t rows;
  int K, C, use_power;
  T floor, gamma;
};

template <typename T>
__global__ void __launch_bounds__(256) fbank_bwd_kernel(FbankBwdArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  T* amp = reinterpret_cast<T*>(smem_raw) + static_cast<size_t>(warp) * (2 * A.K + A.C);
  T* ga = amp + A.K;   // [K] gradient wrt the amplitudes
  T* gz = ga + A.K;    // [C] gradient wrt the filter-bank sums
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + warp; row < A.rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    const T* xr = A.x + row * A.K;
    T esum = 0;
    for (int k = lane; k < A.K; k += 32) {
      const T v = xr[k];
      esum += (k == 0 || k == A.K - 1) ? v : static_cast<T>(2) * v;
      amp[k] = A.use_power ? v : dsqrt(v);
      ga[k] = 0;
    }
    __syncwarp();
    const T ge = A.gE ? A.gE[row] / warp_sum(esum) : static_cast<T>(0);   // dE/dx_k = w_k / sum
    for (int c = lane; c < A.C; c += 32) {
      const int lo = A.cb ? A.cb[c] : 0, hi = A.ce ? A.ce[c] : A.K;
      T z = 0;
      for (int k = lo; k < hi; ++k) z = dfma(amp[k], A.H[static_cast<size_t>(k) * A.C + c], z);
      T d = 0;  // d y / d z; the floor clamps the gradient to zero (torch.clip backward)
      if (z >= A.floor) d = (A.gamma == static_cast<T>(0)) ? static_cast<T>(1) / z : dpow(z, A.gamma - static_cast<T>(1));
      gz[c] = A.gy[row * A.C + c] * d;
    }
    __syncwarp();
    // ga[k] = sum_c H[k][c] gz[c]: walk every column's support, lanes over its rows (distinct k: no conflicts)
    for (int c = 0; c < A.C; ++c) {
      const int lo = A.cb ? A.cb[c] : 0, hi = A.ce ? A.ce[c] : A.K;
      const T g = gz[c];
      for (int k = lo + lane; k < hi; k += 32) ga[k] = dfma(A.H[static_cast<size_t>(k) * A.C + c], g, ga[k]);
      __syncwarp();
    }
    T* gxr = A.gx + row * A.K;
    for (int k = lane; k < A.K; k += 32) {
      T g = ga[k];
      if (!A.use_power) g = amp[k] > static_cast<T>(0) ? g * static_cast<T>(0.5) / amp[k] : static_cast<T>(0);
      g += ge * ((k == 0 || k == A.K - 1) ? static_cast<T>(1) : static_cast<T>(2));
      gxr[k] = g;
    }
    __syncwarp();
  }
}

template <typename T>
int fbank_bwd_impl(const void* x, const void* H, const int32_t* cb, const int32_t* ce, const void* gy, const void* gE,
                   void* gx, int64_t rows, const dsb200_fbank_params* p, int device, void* stream) {
  DSB_REQUIRE(p != nullptr, "fbank params are NULL");
  DSB_REQUIRE(p->fft_length > 1 && p->fft_length % 2 == 0, "fft_length must be positive even.");
  DSB_REQUIRE(p->n_channel > 0, "n_channel must be positive.");
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(x && H && gy && gx, "NULL data pointer");
  DSB_REQUIRE((cb == nullptr) == (ce == nullptr), "col_begin and col_end must be given together");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  FbankBwdArgs<T> A{};
  A.x = static_cast<const T*>(x);
  A.H = static_cast<const T*>(H);
  A.cb = cb;
  A.ce = ce;
  A.gy = static_cast<const T*>(gy);
  A.gE = static_cast<const T*>(gE);
  A.gx = static_cast<T*>(gx);
  A.rows = rows;
  A.K = p->fft_length / 2 + 1;
  A.C = p->n_channel;
  A.use_power = p->use_power;
  A.floor = static_cast<T>(p->floor);
  A.gamma = static_cast<T>(p->gamma);
  const size_t per_warp = static_cast<size_t>(2 * A.K + A.C) * sizeof(T);
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  if (per_warp > cap) return fail(DSB200_E_UNSUPPORTED, "spectrum row does not fit in shared memory");
  int wpb = static_cast<int>(std::min<size_t>(8, cap / per_warp));
  while (wpb > 1 && wpb * per_warp > 48 * 1024) --wpb;
  DSB_CUDA(cudaFuncSetAttribute(fbank_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int64_t need = (rows + wpb - 1) / wpb;
  const int blocks = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(sm_count(device)) * 16));
  fbank_bwd_kernel<T><<<blocks, wpb * 32, wpb * per_warp, static_cast<cudaStream_t>(stream)>>>(A);
  return after_launch("fbank_bwd_kernel");
}

// ------------------------------------------------------------------------------------ acorr
// Forward per row: r_k = sum_n x_n x_{n+k}, k = 0..M, then naive | normalized (r / r_0) | biased (r / L) |
// unbiased (r_k / (L - k)).  Backward: gx_n = sum_k gr_k (x_{n+k} + x_{n-k}).
template <typename T>
__global__ void __launch_bounds__(256) acorr_bwd_kernel(const T* __restrict__ x, const T* __restrict__ gy,
                                                        T* __restrict__ gx, int64_t rows, int L, int M, int fmt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int D = M + 1;
  double* xs = reinterpret_cast<double*>(smem_raw) + static_cast<size_t>(warp) * (L + 2 * D);
  double* gr = xs + L;   // [D] gradient wrt the raw lags
  double* rr = gr + D;   // [D] raw lags (normalized format only)
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + warp; row < rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    for (int n = lane; n < L; n += 32) xs[n] = static_cast<double>(x[row * L + n]);
    for (int k = lane; k < D; k += 32) {
      double g = static_cast<double>(gy[row * D + k]);
      if (fmt == DSB200_ACORR_BIASED) g /= static_cast<double>(L);
      if (fmt == DSB200_ACORR_UNBIASED) g /= static_cast<double>(L - k);
      gr[k] = g;
    }
    __syncwarp();
    if (fmt == DSB200_ACORR_NORMALIZED) {
      for (int k = lane; k < D; k += 32) {
        double acc = 0;
        for (int n = 0; n + k < L; ++n) acc = fma(xs[n], xs[n + k], acc);
        rr[k] = acc;
      }
      __syncwarp();
      const double r0 = rr[0];
      double part = 0;
      for (int k = 1 + lane; k < D; k += 32) part += gr[k] * rr[k];
      part = warp_sum(part);
      __syncwarp();
      for (int k = lane; k < D; k += 32) gr[k] = (k == 0) ? -part / (r0 * r0) : gr[k] / r0;
      __syncwarp();
    }
    for (int n = lane; n < L; n += 32) {
      double acc = 0;
      for (int k = 0; k < D; ++k) {
        const double hi = (n + k < L) ? xs[n + k] : 0.0, lo = (n - k >= 0) ? xs[n - k] : 0.0;
        acc = fma(gr[k], hi + lo, acc);
      }
      gx[row * L + n] = static_cast<T>(acc);
    }
    __syncwarp();
  }
}

template <typename T>
int acorr_bwd_impl(const void* x, const void* gy, void* gx, int64_t rows, int32_t L, int32_t M, int32_t fmt,
                   int device, void* stream) {
  DSB_REQUIRE(L > 0, "frame_length must be positive.");
  DSB_REQUIRE(M >= 0 && M < L, "acr_order must be less than frame_length.");
  DSB_REQUIRE(fmt >= DSB200_ACORR_NAIVE && fmt <= DSB200_ACORR_UNBIASED, "out_format %d is not supported.", fmt);
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(x && gy && gx, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  const size_t per_warp = static_cast<size_t>(L + 2 * (M + 1)) * sizeof(double);
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  if (per_warp > cap) return fail(DSB200_E_UNSUPPORTED, "frame_length=%d does not fit in shared memory", L);
  int wpb = static_cast<int>(std::min<size_t>(8, cap / per_warp));
  while (wpb > 1 && wpb * per_warp > 64 * 1024) --wpb;
  DSB_CUDA(cudaFuncSetAttribute(acorr_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int64_t need = (rows + wpb - 1) / wpb;
  const int blocks = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(sm_count(device)) * 16));
  acorr_bwd_kernel<T><<<blocks, wpb * 32, wpb * per_warp, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const T*>(x), static_cast<const T*>(gy), static_cast<T*>(gx), rows, L, M, fmt);
  return after_launch("acorr_bwd_kernel");
}

// ------------------------------------------------------------------------------------ levdur
// Forward (levdur.py:113-127): R = Toeplitz(r_0..r_{M-1}) + eps I, a = -R^{-1} r_{1..M}, K = sqrt(r_0 + r_1.a).
// With gs = gK / (2 K) and the total gradient on a, ga' = ga + gs r_1:
//     lambda = R^{-1} ga' = R^{-1} ga - gs a          (R is symmetric; R^{-1} r_1 = -a)
//     g r_0   = gs - sum_i lambda_i a_i
//     g r_d   = gs a_d - lambda_{d-1} - sum_{|i-j|=d} lambda_i a_j        (Toeplitz part only for d < M)
// mu = R^{-1} ga comes from the general-right-hand-side Levinson recursion that runs alongside the predictor
// recursion (the order-n backward vector is the reversed predictor), one row per thread, float64 state.
template <typename T>
__global__ void levdur_bwd_kernel(const T* r, const T* ga, T* gr, int64_t rows, int M, double eps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tpb = blockDim.x, ld = tpb + 1, D = M + 1;
  double* rs = reinterpret_cast<double*>(smem_raw);  // [D][ld] lags
  double* as = rs + static_cast<size_t>(D) * ld;     // [D][ld] predictor (as[0] unused -> K at the end)
  double* ts = as + static_cast<size_t>(D) * ld;     // [D][ld] scratch, then lambda
  double* ys = ts + static_cast<size_t>(D) * ld;     // [D][ld] upstream gradient (gK, ga_1..ga_M)
  double* xs = ys + static_cast<size_t>(D) * ld;     // [D][ld] mu, then the result
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * tpb; base < rows; base += static_cast<int64_t>(gridDim.x) * tpb) {
    const int nrow = static_cast<int>(rows - base < tpb ? rows - base : tpb);
    for (int idx = threadIdx.x; idx < nrow * D; idx += tpb) {
      const int rl = idx / D, j = idx - rl * D;
      rs[j * ld + rl] = static_cast<double>(r[base * D + idx]);
      ys[j * ld + rl] = static_cast<double>(ga[base * D + idx]);
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t < nrow) {
      const double r0 = rs[t];
      double E = r0 + eps;
      if (M >= 1) xs[t] = ys[1 * ld + t] / E;    // mu of the 1 x 1 system (mu_i is stored at xs[i])
      for (int i = 1; i <= M; ++i) {
        double acc = rs[i * ld + t];
        for (int j = 1; j < i; ++j) acc = fma(as[j * ld + t], rs[(i - j) * ld + t], acc);
        const double k = -acc / E;
        for (int j = 1; j < i; ++j) ts[j * ld + t] = fma(k, as[(i - j) * ld + t], as[j * ld + t]);
        for (int j = 1; j < i; ++j) as[j * ld + t] = ts[j * ld + t];
        as[i * ld + t] = k;
        E *= (1.0 - k * k);
        if (i < M) {  // grow mu from i to i + 1 unknowns with the order-i backward vector (a_i, ..., a_1, 1)
          double res = ys[(i + 1) * ld + t];
          for (int j = 0; j < i; ++j) res = fma(-rs[(i - j) * ld + t], xs[j * ld + t], res);
          const double q = res / E;
          for (int j = 0; j < i; ++j) xs[j * ld + t] = fma(q, as[(i - j) * ld + t], xs[j * ld + t]);
          xs[i * ld + t] = q;
        }
      }
      double s = r0;
      for (int j = 1; j <= M; ++j) s = fma(rs[j * ld + t], as[j * ld + t], s);
      const double K = sqrt(s);
      const double gs = K > 0.0 ? ys[t] / (2.0 * K) : 0.0;
      // lambda_i = mu_i - gs a_{i+1}, kept in ts[1..M] as lambda_{i} -> ts[i + 1]
      for (int i = 0; i < M; ++i) ts[(i + 1) * ld + t] = xs[i * ld + t] - gs * as[(i + 1) * ld + t];
      // result into xs (mu is dead once lambda exists): walk d downwards is not needed, xs is only written
      double g0 = gs;
      for (int i = 1; i <= M; ++i) g0 = fma(-ts[i * ld + t], as[i * ld + t], g0);
      ys[t] = g0;  // ys is dead (all of it was consumed above): reuse it for the output
      for (int d = 1; d <= M; ++d) {
        double g = fma(gs, as[d * ld + t], -ts[d * ld + t]);
        if (d < M)
          for (int i = 1; i + d <= M; ++i)
            g -= ts[i * ld + t] * as[(i + d) * ld + t] + ts[(i + d) * ld + t] * as[i * ld + t];
        ys[d * ld + t] = g;
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nrow * D; idx += tpb) {
      const int rl = idx / D, j = idx - rl * D;
      gr[base * D + idx] = static_cast<T>(ys[j * ld + rl]);
    }
    __syncthreads();
  }
}

template <typename T>
int levdur_bwd_impl(const void* r, const void* ga, void* gr, int64_t rows, int32_t M, double eps, int device,
                    void* stream) {
  DSB_REQUIRE(M >= 0, "lpc_order must be non-negative.");
  DSB_REQUIRE(eps >= 0, "eps must be non-negative.");
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(r && ga && gr, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  const size_t D = static_cast<size_t>(M) + 1;
  int tpb = 128;
  auto bytes = [&](int t) { return 5 * D * (t + 1) * sizeof(double); };
  while (tpb > 32 && bytes(tpb) > std::min<size_t>(cap, 96 * 1024)) tpb -= 32;
  if (bytes(tpb) > cap) return fail(DSB200_E_UNSUPPORTED, "lpc_order=%d is too large for the shared-memory Levinson kernel", M);
  DSB_CUDA(cudaFuncSetAttribute(levdur_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int64_t need = (rows + tpb - 1) / tpb;
  const int blocks = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(sm_count(device)) * 8));
  levdur_bwd_kernel<T><<<blocks, tpb, bytes(tpb), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const T*>(r), static_cast<const T*>(ga), static_cast<T*>(gr), rows, M, eps);
  return after_launch("levdur_bwd_kernel");
}

}  // namespace
}  // namespace dsb200

using namespace dsb200;

extern "C" {

int dsb200_fbank_backward_f32(const void* x, const void* H, const int32_t* cb, const int32_t* ce, const void* gy,
                              const void* gE, void* gx, int64_t rows, const dsb200_fbank_params* p, int device,
                              void* stream) {
  return fbank_bwd_impl<float>(x, H, cb, ce, gy, gE, gx, rows, p, device, stream);
}
int dsb200_fbank_backward_f64(const void* x, const void* H, const int32_t* cb, const int32_t* ce, const void* gy,
                              const void* gE, void* gx, int64_t rows, const dsb200_fbank_params* p, int device,
                              void* stream) {
  return fbank_bwd_impl<double>(x, H, cb, ce, gy, gE, gx, rows, p, device, stream);
}
int dsb200_acorr_backward_f32(const void* x, const void* gy, void* gx, int64_t rows, int32_t frame_length,
                              int32_t acr_order, int32_t out_format, int device, void* stream) {
  return acorr_bwd_impl<float>(x, gy, gx, rows, frame_length, acr_order, out_format, device, stream);
}
int dsb200_acorr_backward_f64(const void* x, const void* gy, void* gx, int64_t rows, int32_t frame_length,
                              int32_t acr_order, int32_t out_format, int device, void* stream) {
  return acorr_bwd_impl<double>(x, gy, gx, rows, frame_length, acr_order, out_format, device, stream);
}
int dsb200_levdur_backward_f32(const void* r, const void* ga, void* gr, int64_t rows, int32_t lpc_order, double eps,
                               int device, void* stream) {
  return levdur_bwd_impl<float>(r, ga, gr, rows, lpc_order, eps, device, stream);
}
int dsb200_levdur_backward_f64(const void* r, const void* ga, void* gr, int64_t rows, int32_t lpc_order, double eps,
                               int device, void* stream) {
  return levdur_bwd_impl<double>(r, ga, gr, rows, lpc_order, eps, device, stream);
}

}  // extern "C"
