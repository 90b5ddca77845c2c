// Backward (vector-Jacobian) kernels of the spectral part of the path: STFT, fftr, Spectrum (numerator
// only) and Frame.  fp32 + fp64, every parameter combination of the forward kernels.
//
// the reference is differentiable end to end (every reference test calls
// check_differentiability, e.g. tests/test_stft.py:62).  The forward ops are fused, so autograd cannot
// see inside them; these kernels are their adjoints, registered through torch.library.register_autograd.
//
// One warp owns one frame-rate row.  It re-computes the forward spectrum X of the row on chip (nothing is
// saved by the forward pass), turns the incoming gradient into dL/dX, applies the adjoint of the real FFT
//     dL/dx_j = Re sum_{k=0}^{n/2} conj(g_k) W_n^{jk},      g_k = dL/dRe X_k + i dL/dIm X_k
// as one length-n complex FFT, multiplies by the window, removes the mean (zmean) and scatters the frame
// back into the waveform gradient with atomics -- the adjoint of padding + unfold, for all four pad modes.
#include <algorithm>
#include <cstdlib>

#include "rowfft.cuh"

namespace dsb200 {
namespace {

enum BwdMode { BWD_STFT = 0, BWD_RFFT = 1, BWD_SPEC = 2 };

template <typename T>
struct BwdArgs {
  int mode;
  const T* x;        // STFT: waveform [batch,T]; RFFT/SPEC: rows [rows,in_len]
  const T* window;   // STFT: [L]
  const T* tw;       // W_n^k, n entries
  const T* gy;       // gradient of the output (layout of the forward output)
  T* gx;             // gradient of x (same layout as x); must be zero-initialised for STFT
  T* gw;             // STFT: gradient of the window [L] or null; zero-initialised
  int64_t rows, T_len, n_frames;
  int L, P, left, zmean, pad_mode;
  int in_len;
  int n, Nc, pow2;
  int out_format, has_floor;
  T eps, rel_floor;
};

template <typename T>
__device__ __forceinline__ T fmt_derivative(T s, int fmt) {  // d format(s) / d s
  switch (fmt) {
    case DSB200_SPEC_DB: return static_cast<T>(4.3429448190325175) / s;  // 10 / ln 10
    case DSB200_SPEC_LOGMAG: return static_cast<T>(0.5) / s;
    case DSB200_SPEC_MAGNITUDE: return static_cast<T>(0.5) / dsqrt(s);
    default: return static_cast<T>(1);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) rowfft_bwd_kernel(BwdArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using C = cx_t<T>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int Nc = A.Nc, n = A.n, K = Nc + 1;

  C* tw = reinterpret_cast<C*>(smem_raw);  // full circle: n entries
  for (int i = threadIdx.x; i < n; i += blockDim.x) tw[i] = reinterpret_cast<const C*>(A.tw)[i];
  const int gw_len = (A.mode == BWD_STFT && A.gw != nullptr) ? A.L : 0;
  // K + gw_len rounded up to even: every warp's complex buffers stay aligned to sizeof(C)
  const size_t per_warp = static_cast<size_t>(2 * n) * sizeof(C) + static_cast<size_t>((K + gw_len + 1) & ~1) * sizeof(T);
  unsigned char* wb = smem_raw + static_cast<size_t>(n) * sizeof(C) + warp * per_warp;
  C* buf0 = reinterpret_cast<C*>(wb);
  C* buf1 = buf0 + n;
  T* aux = reinterpret_cast<T*>(buf1 + n);   // [K]
  T* gws = aux + K;                          // [L] window-gradient partial sums of this warp
  for (int j = lane; j < gw_len; j += 32) gws[j] = 0;
  __syncthreads();

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + warp; row < A.rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    // ---- forward: stage the row and transform it -------------------------------------------------
    T* xin = reinterpret_cast<T*>(buf0);
    int len;
    int64_t b = 0, start = 0;
    T mean = 0;
    const T* xb = A.x;
    if (A.mode == BWD_STFT) {
      b = row / A.n_frames;
      const int64_t i = row - b * A.n_frames;
      xb = A.x + b * A.T_len;
      start = i * A.P - A.left;
      if (A.zmean) {
        T acc = 0;
        for (int j = lane; j < A.L; j += 32) {
          const int64_t q = pad_index(start + j, A.T_len, A.pad_mode);
          acc += q < 0 ? static_cast<T>(0) : xb[q];
        }
        mean = warp_sum(acc) / static_cast<T>(A.L);
      }
      len = A.L < n ? A.L : n;
      for (int j = lane; j < n; j += 32) {
        T v = 0;
        if (j < len) {
          const int64_t q = pad_index(start + j, A.T_len, A.pad_mode);
          v = ((q < 0 ? static_cast<T>(0) : xb[q]) - mean) * A.window[j];
        }
        xin[j] = v;
      }
    } else {
      len = A.in_len < n ? A.in_len : n;
      const T* xr = A.x + row * A.in_len;
      for (int j = lane; j < n; j += 32) xin[j] = j < len ? xr[j] : static_cast<T>(0);
    }
    __syncwarp();
    const C* S;
    if (A.pow2) {
      S = warp_fft_pow2<T>(buf0, buf1, Nc, tw, lane, n);
    } else {
      warp_dft_direct<T>(xin, len, buf1, n, Nc, tw, lane);
      S = buf1;
    }
    C* Y = (S == buf0) ? buf1 : buf0;  // the buffer the spectrum does not occupy (n entries)

    // ---- dL/dX from the output gradient ----------------------------------------------------------
    const bool cplx_out = (A.mode == BWD_RFFT) ? (A.out_format == DSB200_FFTR_COMPLEX)
                                               : (A.out_format == DSB200_SPEC_COMPLEX);
    const T* gr = A.gy + row * static_cast<int64_t>(K) * (cplx_out ? 2 : 1);
    T fl = 0, floored = 0;
    int arg = 0;
    if (A.mode != BWD_RFFT && !cplx_out && A.has_floor) {
      // s' = max(s, max(s) * rf): gradients of floored bins flow to the arg-max bin (torch.amax backward)
      T mx = 0;
      for (int k = lane; k < K; k += 32) {
        const C X = A.pow2 ? real_split<T>(S, k, Nc, tw) : S[k];
        const T s = X.x * X.x + X.y * X.y + A.eps;
        aux[k] = s;
        mx = dmax(mx, s);
      }
      mx = warp_max(mx);
      fl = mx * A.rel_floor;
      T part = 0;
      int cand = K;
      for (int k = lane; k < K; k += 32) {
        if (aux[k] == mx && k < cand) cand = k;
        if (aux[k] < fl) part += gr[k] * fmt_derivative<T>(fl, A.out_format);
      }
      floored = warp_sum(part) * A.rel_floor;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const int other = __shfl_xor_sync(0xffffffffu, cand, o);
        cand = other < cand ? other : cand;
      }
      arg = cand;
      __syncwarp();
    }
    for (int k = lane; k < n; k += 32) {
      C y = mk<T>(0, 0);
      if (k < K) {
        const C X = A.pow2 ? real_split<T>(S, k, Nc, tw) : S[k];
        C g;  // dL/dRe X + i dL/dIm X
        if (cplx_out) {
          g = reinterpret_cast<const C*>(gr)[k];
        } else if (A.mode == BWD_RFFT) {
          const T G = gr[k];
          switch (A.out_format) {
            case DSB200_FFTR_REAL: g = mk<T>(G, 0); break;
            case DSB200_FFTR_IMAG: g = mk<T>(0, G); break;
            case DSB200_FFTR_AMPLITUDE: {
              const T amp = dsqrt(X.x * X.x + X.y * X.y);
              const T c = amp > 0 ? G / amp : static_cast<T>(0);
              g = mk<T>(c * X.x, c * X.y);
              break;
            }
            default: g = mk<T>(2 * G * X.x, 2 * G * X.y); break;  // power
          }
        } else {
          const T s = X.x * X.x + X.y * X.y + A.eps;
          T ds;
          if (A.has_floor) {
            ds = (s < fl) ? static_cast<T>(0) : gr[k] * fmt_derivative<T>(s, A.out_format);
            if (k == arg) ds += floored;
          } else {
            ds = gr[k] * fmt_derivative<T>(s, A.out_format);
          }
          g = mk<T>(2 * ds * X.x, 2 * ds * X.y);
        }
        y = mk<T>(g.x, -g.y);  // conj(g)
      }
      Y[k] = y;
    }
    __syncwarp();

    // ---- adjoint of the real FFT: z_j = sum_k conj(g_k) W_n^(jk), dL/dx_j = Re z_j ------------------
    C* other = (Y == buf0) ? buf1 : buf0;  // the spectrum is dead now
    const C* Z;
    if (A.pow2) {
      Z = warp_fft_pow2<T>(Y, other, n, tw, lane, n);
    } else {
      for (int j = lane; j < len; j += 32) {
        T re = 0;
        int idx = 0;
        for (int k = 0; k < K; ++k) {
          const C w = tw[idx];
          re += Y[k].x * w.x - Y[k].y * w.y;
          idx += j;
          if (idx >= n) idx -= n;
        }
        other[j] = mk<T>(re, 0);
      }
      __syncwarp();
      Z = other;
    }

    // ---- window, mean removal, scatter ---------------------------------------------------------------
    if (A.mode == BWD_STFT) {
      T gmean = 0;
      if (A.zmean) {
        T acc = 0;
        for (int j = lane; j < len; j += 32) acc += Z[j].x * A.window[j];
        gmean = warp_sum(acc) / static_cast<T>(A.L);
      }
      for (int j = lane; j < A.L; j += 32) {
        T gxw = 0;  // gradient wrt the windowed sample (zero beyond the FFT length)
        if (j < len) gxw = Z[j].x;
        const int64_t q = pad_index(start + j, A.T_len, A.pad_mode);
        if (gw_len) {
          const T xv = (q < 0 ? static_cast<T>(0) : xb[q]) - mean;
          gws[j] += gxw * xv;
        }
        const T gf = (j < len ? gxw * A.window[j] : static_cast<T>(0)) - gmean;
        if (q >= 0 && gf != static_cast<T>(0)) atomicAdd(A.gx + b * A.T_len + q, gf);
      }
    } else {
      T* gxr = A.gx + row * A.in_len;
      for (int j = lane; j < A.in_len; j += 32) gxr[j] = j < len ? Z[j].x : static_cast<T>(0);
    }
    __syncwarp();
  }
  if (gw_len) {
    __syncwarp();
    for (int j = lane; j < gw_len; j += 32) atomicAdd(A.gw + j, gws[j]);
  }
}

template <typename T>
int launch_bwd(BwdArgs<T>& A, int device, cudaStream_t stream) {
  if (A.rows == 0) return DSB200_OK;
  A.Nc = A.n / 2;
  A.pow2 = is_pow2(A.n) ? 1 : 0;
  const void* tw = twiddle_table(device, A.n, sizeof(T) == 8, stream);
  if (tw == nullptr) return fail(DSB200_E_CUDA, "could not build the twiddle table for fft_length=%d", A.n);
  A.tw = static_cast<const T*>(tw);
  const int K = A.Nc + 1;
  const int gw_len = (A.mode == BWD_STFT && A.gw != nullptr) ? A.L : 0;
  const size_t tw_bytes = static_cast<size_t>(A.n) * 2 * sizeof(T);
  const size_t per_warp = static_cast<size_t>(2 * A.n) * 2 * sizeof(T) + static_cast<size_t>((K + gw_len + 1) & ~1) * sizeof(T);
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  if (tw_bytes + per_warp > cap)
    return fail(DSB200_E_UNSUPPORTED, "fft_length=%d is too long for the backward kernel's shared memory", A.n);
  int wpb = static_cast<int>(std::min<size_t>(8, (cap - tw_bytes) / per_warp));
  while (wpb > 1 && tw_bytes + wpb * per_warp > 96 * 1024 && A.n <= 2048) --wpb;
  const size_t smem = tw_bytes + wpb * per_warp;
  DSB_CUDA(cudaFuncSetAttribute(rowfft_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int64_t need = (A.rows + wpb - 1) / wpb;
  const int blocks = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(sm_count(device)) * 8));
  rowfft_bwd_kernel<T><<<blocks, wpb * 32, smem, stream>>>(A);
  return after_launch("rowfft_bwd_kernel");
}

// ---- Frame backward: scatter-add of the frame gradients (adjoint of pad + unfold [+ mean removal]) ----
template <typename T>
__global__ void frame_bwd_kernel(const T* __restrict__ gy, T* __restrict__ gx, int64_t rows, int64_t T_len,
                                 int64_t n_frames, int L, int P, int left, int zmean, int pad_mode) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5); row < rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    const int64_t b = row / n_frames, i = row - b * n_frames;
    const T* g = gy + row * L;
    const int64_t start = i * P - left;
    T gmean = 0;
    if (zmean) {
      T acc = 0;
      for (int j = lane; j < L; j += 32) acc += g[j];
      gmean = warp_sum(acc) / static_cast<T>(L);
    }
    for (int j = lane; j < L; j += 32) {
      const int64_t q = pad_index(start + j, T_len, pad_mode);
      if (q >= 0) atomicAdd(gx + b * T_len + q, g[j] - gmean);
    }
  }
}

int check_common(const dsb200_frame_params* f, int64_t T_len) {
  DSB_REQUIRE(f != nullptr, "frame params are NULL");
  DSB_REQUIRE(f->frame_length > 0, "frame_length must be positive.");
  DSB_REQUIRE(f->frame_period > 0, "frame_period must be positive.");
  DSB_REQUIRE(T_len >= 1, "waveform length must be at least 1");
  return DSB200_OK;
}

}  // namespace

// Fused backward for the BASELINE shape (stft512_bwd.cu); DSB200_E_UNSUPPORTED outside its envelope.
int stft512_bwd_try(const float* x, const float* window, const float* gy, float* gx, int64_t batch, int64_t T_len,
                    const dsb200_stft_params* p, int device, cudaStream_t stream);

namespace {

template <typename T>
int stft_bwd_impl(const void* x, const void* window, const void* gy, void* gx, void* gw, int64_t batch, int64_t T_len,
                  const dsb200_stft_params* p, int device, void* stream) {
  DSB_REQUIRE(p != nullptr, "stft params are NULL");
  if (int rc = check_common(&p->frame, T_len)) return rc;
  DSB_REQUIRE(p->spec.fft_length > 1 && p->spec.fft_length % 2 == 0, "fft_length must be positive even.");
  if (batch == 0) return DSB200_OK;
  DSB_REQUIRE(x && window && gy && gx, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (sizeof(T) == 4 && gw == nullptr && getenv("DSB200_STFT_BWD_GENERIC") == nullptr) {
    const int rc = stft512_bwd_try(static_cast<const float*>(x), static_cast<const float*>(window),
                                   static_cast<const float*>(gy), static_cast<float*>(gx), batch, T_len, p, device, s);
    if (rc != DSB200_E_UNSUPPORTED) return rc;
  }
  DSB_CUDA(cudaMemsetAsync(gx, 0, static_cast<size_t>(batch) * T_len * sizeof(T), s));
  if (gw) DSB_CUDA(cudaMemsetAsync(gw, 0, static_cast<size_t>(p->frame.frame_length) * sizeof(T), s));
  BwdArgs<T> A{};
  A.mode = BWD_STFT;
  A.x = static_cast<const T*>(x);
  A.window = static_cast<const T*>(window);
  A.gy = static_cast<const T*>(gy);
  A.gx = static_cast<T*>(gx);
  A.gw = static_cast<T*>(gw);
  A.T_len = T_len;
  A.n_frames = dsb200_num_frames(T_len, p->frame.frame_period);
  A.rows = batch * A.n_frames;
  A.L = p->frame.frame_length;
  A.P = p->frame.frame_period;
  A.left = p->frame.center ? p->frame.frame_length / 2 : 0;
  A.zmean = p->frame.zmean;
  A.pad_mode = p->frame.pad_mode;
  A.n = p->spec.fft_length;
  A.out_format = p->spec.out_format;
  A.has_floor = p->spec.has_relative_floor;
  A.eps = static_cast<T>(p->spec.eps);
  A.rel_floor = static_cast<T>(p->spec.relative_floor);
  return launch_bwd<T>(A, device, s);
}

template <typename T>
int rows_bwd_impl(int mode, const void* x, const void* gy, void* gx, int64_t rows, int32_t in_length,
                  int32_t fft_length, int32_t out_format, double eps, int has_floor, double rel_floor, int device,
                  void* stream) {
  DSB_REQUIRE(fft_length > 0 && fft_length % 2 == 0, "fft_length must be positive even.");
  DSB_REQUIRE(in_length > 0, "input length must be positive");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(x && gy && gx, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  BwdArgs<T> A{};
  A.mode = mode;
  A.x = static_cast<const T*>(x);
  A.gy = static_cast<const T*>(gy);
  A.gx = static_cast<T*>(gx);
  A.rows = rows;
  A.in_len = in_length;
  A.n = fft_length;
  A.out_format = out_format;
  A.has_floor = has_floor;
  A.eps = static_cast<T>(eps);
  A.rel_floor = static_cast<T>(rel_floor);
  return launch_bwd<T>(A, device, static_cast<cudaStream_t>(stream));
}

template <typename T>
int frame_bwd_impl(const void* gy, void* gx, int64_t batch, int64_t T_len, const dsb200_frame_params* p, int device,
                   void* stream) {
  if (int rc = check_common(p, T_len)) return rc;
  if (batch == 0) return DSB200_OK;
  DSB_REQUIRE(gy && gx, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DSB_CUDA(cudaMemsetAsync(gx, 0, static_cast<size_t>(batch) * T_len * sizeof(T), s));
  const int64_t N = dsb200_num_frames(T_len, p->frame_period);
  const int64_t rows = batch * N;
  const int threads = 256, wpb = threads / 32;
  const int blocks = static_cast<int>(std::min<int64_t>((rows + wpb - 1) / wpb, static_cast<int64_t>(sm_count(device)) * 32));
  frame_bwd_kernel<T><<<blocks, threads, 0, s>>>(static_cast<const T*>(gy), static_cast<T*>(gx), rows, T_len, N,
                                                 p->frame_length, p->frame_period,
                                                 p->center ? p->frame_length / 2 : 0, p->zmean, p->pad_mode);
  return after_launch("frame_bwd_kernel");
}

}  // namespace
}  // namespace dsb200

using namespace dsb200;

extern "C" {

int dsb200_stft_backward_f32(const void* x, const void* w, const void* gy, void* gx, void* gw, int64_t batch, int64_t T,
                             const dsb200_stft_params* p, int device, void* stream) {
  return stft_bwd_impl<float>(x, w, gy, gx, gw, batch, T, p, device, stream);
}
int dsb200_stft_backward_f64(const void* x, const void* w, const void* gy, void* gx, void* gw, int64_t batch, int64_t T,
                             const dsb200_stft_params* p, int device, void* stream) {
  return stft_bwd_impl<double>(x, w, gy, gx, gw, batch, T, p, device, stream);
}
int dsb200_rfft_backward_f32(const void* x, const void* gy, void* gx, int64_t rows, int32_t in_length, int32_t fft_length,
                             int32_t out_format, int device, void* stream) {
  return rows_bwd_impl<float>(BWD_RFFT, x, gy, gx, rows, in_length, fft_length, out_format, 0, 0, 0, device, stream);
}
int dsb200_rfft_backward_f64(const void* x, const void* gy, void* gx, int64_t rows, int32_t in_length, int32_t fft_length,
                             int32_t out_format, int device, void* stream) {
  return rows_bwd_impl<double>(BWD_RFFT, x, gy, gx, rows, in_length, fft_length, out_format, 0, 0, 0, device, stream);
}
int dsb200_spec_backward_f32(const void* b, int32_t b_length, const void* gy, void* gb, int64_t rows,
                             const dsb200_spec_params* p, int device, void* stream) {
  if (p == nullptr) return fail(DSB200_E_BAD_PARAM, "spec params are NULL");
  return rows_bwd_impl<float>(BWD_SPEC, b, gy, gb, rows, b_length, p->fft_length, p->out_format, p->eps,
                              p->has_relative_floor, p->relative_floor, device, stream);
}
int dsb200_spec_backward_f64(const void* b, int32_t b_length, const void* gy, void* gb, int64_t rows,
                             const dsb200_spec_params* p, int device, void* stream) {
  if (p == nullptr) return fail(DSB200_E_BAD_PARAM, "spec params are NULL");
  return rows_bwd_impl<double>(BWD_SPEC, b, gy, gb, rows, b_length, p->fft_length, p->out_format, p->eps,
                               p->has_relative_floor, p->relative_floor, device, stream);
}
int dsb200_frame_backward_f32(const void* gy, void* gx, int64_t batch, int64_t T, const dsb200_frame_params* p,
                              int device, void* stream) {
  return frame_bwd_impl<float>(gy, gx, batch, T, p, device, stream);
}
int dsb200_frame_backward_f64(const void* gy, void* gx, int64_t batch, int64_t T, const dsb200_frame_params* p,
                              int device, void* stream) {
  return frame_bwd_impl<double>(gy, gx, batch, T, p, device, stream);
}

}  // extern "C"
