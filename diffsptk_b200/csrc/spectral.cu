// Generic frame / window / real-FFT / spectrum / STFT kernels (any even FFT length, fp32 + fp64).
//
// This file is the *general* path: one warp owns one frame-rate row, stages it in shared
// memory, runs a Stockham radix-2 FFT on the half-length complex packing (power-of-two
// lengths) or a table-driven direct DFT (other even lengths), and applies the reference's
// output formatter while the spectrum is still on chip.  The fl=400/fp=80/n_fft=512 headline
// configuration is served by the specialised kernel in stft512.cu; everything here is what
// keeps the remaining parameter space of the reference API on the GPU.
//
// Reference semantics: diffsptk/modules/frame.py:120-141, window.py:185-193,
// fftr.py:136-151, spec.py:152-178, stft.py:237-241.
#include <algorithm>
#include <cstdlib>

#include "rowfft.cuh"

namespace dsb200 {
namespace {

enum RowMode { MODE_STFT = 0, MODE_RFFT = 1, MODE_SPEC = 2 };

template <typename T>
struct RowArgs {
  int mode;
  // inputs
  const T* x;       // STFT: waveform [batch,T]; RFFT: rows [rows,in_len]; SPEC: b rows or null
  const T* a;       // SPEC only: denominator rows or null
  const T* window;  // STFT only: [L]
  const T* tw;      // twiddles W_n^k interleaved, n entries
  T* y;
  int64_t rows;
  // framing (STFT)
  int64_t T_len, n_frames;
  int L, P, left, zmean, pad_mode;
  // row lengths (RFFT/SPEC)
  int in_len, a_len;
  // transform
  int n, Nc, pow2;
  // formatter
  int out_format;  // DSB200_SPEC_* for STFT/SPEC, DSB200_FFTR_* for RFFT
  int has_floor;
  T eps, rel_floor;
};

template <typename T>
__device__ __forceinline__ cx_t<T> ld_tw(const cx_t<T>* tw, int k) { return tw[k]; }

// Transform the `len` real samples staged in buf0[0..n) (zero padded) and return a pointer to
// the spectrum accessor state: for pow2, Z (length Nc, needs real_split); for direct, X itself.
template <typename T>
__device__ __forceinline__ const cx_t<T>* transform_row(const RowArgs<T>& A, cx_t<T>* buf0, cx_t<T>* buf1,
                                                        int len, const cx_t<T>* tw, int lane) {
  if (A.pow2) return warp_fft_pow2<T>(buf0, buf1, A.Nc, tw, lane, A.n);
  warp_dft_direct<T>(reinterpret_cast<const T*>(buf0), len, buf1, A.n, A.Nc, tw, lane);
  return buf1;
}

template <typename T>
__device__ __forceinline__ cx_t<T> bin(const RowArgs<T>& A, const cx_t<T>* S, int k, const cx_t<T>* tw) {
  return A.pow2 ? real_split<T>(S, k, A.Nc, tw) : S[k];
}

// Stage a contiguous row (zero padded / truncated to n) into shared memory; optionally force
// element 0 to one (remove_gain, utils/private.py:200-209).  Returns the staged length.
template <typename T>
__device__ __forceinline__ int stage_row(const T* src, int len, T* dst, int n, bool unit_first, int lane) {
  const int m = len < n ? len : n;
  for (int j = lane; j < n; j += 32) {
    T v = j < m ? src[j] : static_cast<T>(0);
    if (unit_first && j == 0) v = static_cast<T>(1);
    dst[j] = v;
  }
  __syncwarp();
  return m;
}

template <typename T>
__global__ void __launch_bounds__(256) rowfft_kernel(RowArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using C = cx_t<T>;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int Nc = A.Nc, n = A.n, K = Nc + 1;

  // block-shared twiddles, then per-warp buffers
  C* tw = reinterpret_cast<C*>(smem_raw);
  const int n_tw = n;   // the radix-4 passes index up to 3/4 of the circle, the direct DFT all of it
  for (int i = threadIdx.x; i < n_tw; i += blockDim.x) tw[i] = reinterpret_cast<const C*>(A.tw)[i];
  __syncthreads();
  C* buf0 = tw + n_tw + static_cast<size_t>(warp) * (2 * K) ;
  C* buf1 = buf0 + K;
  T* aux = reinterpret_cast<T*>(tw + n_tw + static_cast<size_t>(wpb) * (2 * K)) + static_cast<size_t>(warp) * K;

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + warp; row < A.rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    T* xin = reinterpret_cast<T*>(buf0);
    int len;
    bool ratio = false;  // SPEC with a denominator
    T gainK = static_cast<T>(1);

    if (A.mode == MODE_STFT) {
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
      len = A.L < n ? A.L : n;
      for (int j = lane; j < n; j += 32) {
        T v = 0;
        if (j < len) {
          const int64_t q = pad_index(start + j, A.T_len, A.pad_mode);
          v = q < 0 ? static_cast<T>(0) : xb[q];
          v = (v - mean) * A.window[j];
        }
        xin[j] = v;
      }
      __syncwarp();
    } else if (A.mode == MODE_RFFT) {
      len = stage_row<T>(A.x + row * A.in_len, A.in_len, xin, n, false, lane);
    } else {  // MODE_SPEC
      if (A.x != nullptr) {
        len = stage_row<T>(A.x + row * A.in_len, A.in_len, xin, n, false, lane);
        if (A.a != nullptr) {
          // numerator amplitude -> aux, then transform the denominator
          const C* S = transform_row<T>(A, buf0, buf1, len, tw, lane);
          for (int k = lane; k < K; k += 32) {
            const C X = bin<T>(A, S, k, tw);
            aux[k] = dsqrt(X.x * X.x + X.y * X.y);
          }
          __syncwarp();
          ratio = true;
          gainK = A.a[row * A.a_len];
          len = stage_row<T>(A.a + row * A.a_len, A.a_len, xin, n, true, lane);
        }
      } else {
        gainK = A.a[row * A.a_len];
        len = stage_row<T>(A.a + row * A.a_len, A.a_len, xin, n, true, lane);
      }
    }

    const C* S = transform_row<T>(A, buf0, buf1, len, tw, lane);

    if (A.mode == MODE_RFFT) {
      T* yr = A.y + row * static_cast<int64_t>(K) * (A.out_format == DSB200_FFTR_COMPLEX ? 2 : 1);
      for (int k = lane; k < K; k += 32) {
        const C X = bin<T>(A, S, k, tw);
        switch (A.out_format) {
          case DSB200_FFTR_COMPLEX: reinterpret_cast<C*>(yr)[k] = X; break;
          case DSB200_FFTR_REAL: yr[k] = X.x; break;
          case DSB200_FFTR_IMAG: yr[k] = X.y; break;
          case DSB200_FFTR_AMPLITUDE: yr[k] = dsqrt(X.x * X.x + X.y * X.y); break;
          default: yr[k] = X.x * X.x + X.y * X.y; break;
        }
      }
    } else if (A.out_format == DSB200_SPEC_COMPLEX) {
      C* yr = reinterpret_cast<C*>(A.y) + row * static_cast<int64_t>(K);
      for (int k = lane; k < K; k += 32) yr[k] = bin<T>(A, S, k, tw);
    } else {
      T* yr = A.y + row * static_cast<int64_t>(K);
      const bool a_only = (A.mode == MODE_SPEC && A.x == nullptr);
      T mx = 0;
      // pass 1: power (+eps); keep in aux when a relative floor needs the row maximum
      for (int k = lane; k < K; k += 32) {
        const C X = bin<T>(A, S, k, tw);
        T p = X.x * X.x + X.y * X.y;
        if (ratio) {
          const T amp = gainK * (aux[k] / dsqrt(p));
          p = amp * amp;
        } else if (a_only) {
          const T amp = gainK / dsqrt(p);
          p = amp * amp;
        }
        const T s = p + A.eps;
        if (A.has_floor) {
          aux[k] = s;
          mx = dmax(mx, s);
        } else {
          yr[k] = spec_format<T>(s, A.out_format);
        }
      }
      if (A.has_floor) {
        mx = warp_max(mx);
        const T fl = mx * A.rel_floor;
        for (int k = lane; k < K; k += 32) yr[k] = spec_format<T>(dmax(aux[k], fl), A.out_format);
      }
    }
    __syncwarp();
  }
}

template <typename T>
int launch_rowfft(RowArgs<T>& A, int device, cudaStream_t stream) {
  if (A.rows == 0) return DSB200_OK;
  A.Nc = A.n / 2;
  A.pow2 = is_pow2(A.n) ? 1 : 0;
  const void* tw = twiddle_table(device, A.n, sizeof(T) == 8, stream);
  if (tw == nullptr) return fail(DSB200_E_CUDA, "could not build the twiddle table for fft_length=%d", A.n);
  A.tw = static_cast<const T*>(tw);
  const int K = A.Nc + 1;
  const size_t tw_bytes = static_cast<size_t>(A.n) * 2 * sizeof(T);
  const size_t per_warp = static_cast<size_t>(2 * K) * 2 * sizeof(T) + static_cast<size_t>(K) * sizeof(T);
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  if (tw_bytes + per_warp > cap)
    return fail(DSB200_E_UNSUPPORTED, "fft_length=%d needs %zu B of shared memory per warp (limit %zu)", A.n,
                tw_bytes + per_warp, cap);
  int wpb = static_cast<int>(std::min<size_t>(8, (cap - tw_bytes) / per_warp));
  // keep a few blocks resident per SM when the row is small
  const size_t target = 64 * 1024;
  while (wpb > 1 && tw_bytes + wpb * per_warp > target && A.n <= 2048) --wpb;
  const size_t smem = tw_bytes + wpb * per_warp;
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "float or double");
  DSB_CUDA(cudaFuncSetAttribute(rowfft_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int64_t need = (A.rows + wpb - 1) / wpb;
  const int64_t max_blocks = static_cast<int64_t>(sm_count(device)) * 16;
  const int blocks = static_cast<int>(std::min<int64_t>(need, max_blocks));
  rowfft_kernel<T><<<blocks, wpb * 32, smem, stream>>>(A);
  return after_launch("rowfft_kernel");
}

// ---------------------------------------------------------------------------------- frame
template <typename T>
__global__ void frame_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int64_t T_len,
                             int64_t n_frames, int L, int P, int left, int zmean, int pad_mode) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5); row < rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    const int64_t b = row / n_frames, i = row - b * n_frames;
    const T* xb = x + b * T_len;
    T* yr = y + row * L;
    const int64_t start = i * P - left;
    if (!zmean) {  // pure gather: bit-exact with pad + unfold
      for (int j = lane; j < L; j += 32) {
        const int64_t q = pad_index(start + j, T_len, pad_mode);
        yr[j] = q < 0 ? static_cast<T>(0) : xb[q];
      }
    } else {
      T acc = 0;
      for (int j = lane; j < L; j += 32) {
        const int64_t q = pad_index(start + j, T_len, pad_mode);
        acc += q < 0 ? static_cast<T>(0) : xb[q];
      }
      const T mean = warp_sum(acc) / static_cast<T>(L);
      for (int j = lane; j < L; j += 32) {
        const int64_t q = pad_index(start + j, T_len, pad_mode);
        yr[j] = (q < 0 ? static_cast<T>(0) : xb[q]) - mean;
      }
    }
  }
}

template <typename T>
__global__ void window_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ y,
                              int64_t total, int L1, int L2) {
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = idx / L2;
    const int j = static_cast<int>(idx - r * L2);
    y[idx] = j < L1 ? x[r * L1 + j] * w[j] : static_cast<T>(0);
  }
}

int check_frame_params(const dsb200_frame_params* p, int64_t T_len) {
  DSB_REQUIRE(p != nullptr, "frame params are NULL");
  DSB_REQUIRE(p->frame_length > 0, "frame_length must be positive.");
  DSB_REQUIRE(p->frame_period > 0, "frame_period must be positive.");
  DSB_REQUIRE(p->pad_mode >= DSB200_PAD_CONSTANT && p->pad_mode <= DSB200_PAD_CIRCULAR, "unknown pad mode %d", p->pad_mode);
  DSB_REQUIRE(T_len >= 1, "waveform length must be at least 1");
  const int L = p->frame_length;
  const int left = p->center ? L / 2 : 0;
  const int right = p->center ? (L - 1) / 2 : L - 1;
  const int big = left > right ? left : right;
  if (p->pad_mode == DSB200_PAD_REFLECT)
    DSB_REQUIRE(big < T_len, "reflect padding (%d) must be smaller than the waveform length (%lld)", big, (long long)T_len);
  if (p->pad_mode == DSB200_PAD_CIRCULAR)
    DSB_REQUIRE(big <= T_len, "circular padding (%d) must not exceed the waveform length (%lld)", big, (long long)T_len);
  return DSB200_OK;
}

int check_spec_params(const dsb200_spec_params* p, bool allow_complex) {
  DSB_REQUIRE(p != nullptr, "spec params are NULL");
  DSB_REQUIRE(p->fft_length > 1, "fft_length must be greater than 1.");
  DSB_REQUIRE(p->fft_length % 2 == 0, "fft_length must be positive even.");
  DSB_REQUIRE(p->eps >= 0, "eps must be non-negative.");
  DSB_REQUIRE(p->out_format >= DSB200_SPEC_DB && p->out_format <= (allow_complex ? DSB200_SPEC_COMPLEX : DSB200_SPEC_POWER),
              "out_format %d is not supported.", p->out_format);
  if (p->has_relative_floor) DSB_REQUIRE(p->relative_floor > 0 && p->relative_floor < 1, "relative_floor must be negative (dB).");
  return DSB200_OK;
}

template <typename T>
int frame_impl(const void* x, void* y, int64_t batch, int64_t T_len, const dsb200_frame_params* p, int device, void* stream) {
  if (int rc = check_frame_params(p, T_len)) return rc;
  DSB_REQUIRE(batch >= 0, "batch must be non-negative");
  if (batch == 0) return DSB200_OK;
  DSB_REQUIRE(x != nullptr && y != nullptr, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  const int64_t N = dsb200_num_frames(T_len, p->frame_period);
  const int64_t rows = batch * N;
  const int threads = 256, wpb = threads / 32;
  const int blocks = static_cast<int>(std::min<int64_t>((rows + wpb - 1) / wpb, static_cast<int64_t>(sm_count(device)) * 32));
  frame_kernel<T><<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const T*>(x), static_cast<T*>(y), rows, T_len, N, p->frame_length, p->frame_period,
      p->center ? p->frame_length / 2 : 0, p->zmean, p->pad_mode);
  return after_launch("frame_kernel");
}

template <typename T>
int window_impl(const void* x, const void* w, void* y, int64_t rows, int32_t L1, int32_t L2, int device, void* stream) {
  DSB_REQUIRE(L1 > 0, "in_length must be positive.");
  DSB_REQUIRE(L2 > 0, "out_length must be positive.");
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(x != nullptr && y != nullptr && w != nullptr, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  const int64_t total = rows * L2;
  const int threads = 256;
  const int blocks = static_cast<int>(std::min<int64_t>((total + threads - 1) / threads, static_cast<int64_t>(sm_count(device)) * 32));
  window_kernel<T><<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const T*>(x), static_cast<const T*>(w), static_cast<T*>(y), total, L1, L2);
  return after_launch("window_kernel");
}

template <typename T>
int rfft_impl(const void* x, void* y, int64_t rows, int32_t in_length, int32_t fft_length, int32_t out_format,
              int device, void* stream) {
  DSB_REQUIRE(fft_length > 0 && fft_length % 2 == 0, "fft_length must be positive even.");
  DSB_REQUIRE(in_length > 0, "input length must be positive");
  DSB_REQUIRE(out_format >= DSB200_FFTR_COMPLEX && out_format <= DSB200_FFTR_POWER, "out_format %d is not supported.", out_format);
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(x != nullptr && y != nullptr, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  RowArgs<T> A{};
  A.mode = MODE_RFFT;
  A.x = static_cast<const T*>(x);
  A.y = static_cast<T*>(y);
  A.rows = rows;
  A.in_len = in_length;
  A.n = fft_length;
  A.out_format = out_format;
  return launch_rowfft<T>(A, device, static_cast<cudaStream_t>(stream));
}

template <typename T>
int spec_impl(const void* b, int32_t b_length, const void* a, int32_t a_length, void* y, int64_t rows,
              const dsb200_spec_params* p, int device, void* stream) {
  if (int rc = check_spec_params(p, false)) return rc;
  DSB_REQUIRE(b != nullptr || a != nullptr, "Either b or a must be specified.");
  DSB_REQUIRE(b == nullptr || b_length > 0, "b length must be positive");
  DSB_REQUIRE(a == nullptr || a_length > 0, "a length must be positive");
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(y != nullptr, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  RowArgs<T> A{};
  A.mode = MODE_SPEC;
  A.x = static_cast<const T*>(b);
  A.a = static_cast<const T*>(a);
  A.y = static_cast<T*>(y);
  A.rows = rows;
  A.in_len = b_length;
  A.a_len = a_length;
  A.n = p->fft_length;
  A.out_format = p->out_format;
  A.has_floor = p->has_relative_floor;
  A.eps = static_cast<T>(p->eps);
  A.rel_floor = static_cast<T>(p->relative_floor);
  return launch_rowfft<T>(A, device, static_cast<cudaStream_t>(stream));
}

}  // namespace

// The specialised fl<=512 / n_fft=512 fp32 kernel (stft512.cu).  Returns DSB200_E_UNSUPPORTED
// when the configuration is outside its envelope, in which case the generic kernel runs.
int stft512_try(const float* x, const float* window, float* y, int64_t batch, int64_t T_len,
                const dsb200_stft_params* p, int device, cudaStream_t stream);

// The shared-memory radix-16 kernel for fft_length 1024 / 2048 (stftn.cu), same contract.
int stftn_try(const float* x, const float* window, float* y, int64_t batch, int64_t T_len,
              const dsb200_stft_params* p, int device, cudaStream_t stream);

template <typename T>
int stft_generic(const void* x, const void* window, void* y, int64_t batch, int64_t T_len,
                 const dsb200_stft_params* p, int device, void* stream) {
  RowArgs<T> A{};
  A.mode = MODE_STFT;
  A.x = static_cast<const T*>(x);
  A.window = static_cast<const T*>(window);
  A.y = static_cast<T*>(y);
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
  return launch_rowfft<T>(A, device, static_cast<cudaStream_t>(stream));
}

template <typename T>
int stft_impl(const void* x, const void* window, void* y, int64_t batch, int64_t T_len,
              const dsb200_stft_params* p, int device, void* stream) {
  DSB_REQUIRE(p != nullptr, "stft params are NULL");
  if (int rc = check_frame_params(&p->frame, T_len)) return rc;
  if (int rc = check_spec_params(&p->spec, true)) return rc;
  DSB_REQUIRE(batch >= 0, "batch must be non-negative");
  if (batch == 0) return DSB200_OK;
  DSB_REQUIRE(x != nullptr && y != nullptr && window != nullptr, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  if (sizeof(T) == 4) {
    const int rc = stft512_try(static_cast<const float*>(x), static_cast<const float*>(window),
                               static_cast<float*>(y), batch, T_len, p, device, static_cast<cudaStream_t>(stream));
    if (rc != DSB200_E_UNSUPPORTED) return rc;
    static const bool generic_only = getenv("DSB200_STFT_GENERIC") != nullptr;   // A/B knob (read once)
    if (!generic_only) {
      const int rn = stftn_try(static_cast<const float*>(x), static_cast<const float*>(window), static_cast<float*>(y),
                               batch, T_len, p, device, static_cast<cudaStream_t>(stream));
      if (rn != DSB200_E_UNSUPPORTED) return rn;
    }
  }
  return stft_generic<T>(x, window, y, batch, T_len, p, device, stream);
}

}  // namespace dsb200

using namespace dsb200;

extern "C" {

int dsb200_frame_f32(const void* x, void* y, int64_t batch, int64_t T, const dsb200_frame_params* p, int device, void* stream) {
  return frame_impl<float>(x, y, batch, T, p, device, stream);
}
int dsb200_frame_f64(const void* x, void* y, int64_t batch, int64_t T, const dsb200_frame_params* p, int device, void* stream) {
  return frame_impl<double>(x, y, batch, T, p, device, stream);
}
int dsb200_window_f32(const void* x, const void* w, void* y, int64_t rows, int32_t L1, int32_t L2, int device, void* stream) {
  return window_impl<float>(x, w, y, rows, L1, L2, device, stream);
}
int dsb200_window_f64(const void* x, const void* w, void* y, int64_t rows, int32_t L1, int32_t L2, int device, void* stream) {
  return window_impl<double>(x, w, y, rows, L1, L2, device, stream);
}
int dsb200_rfft_f32(const void* x, void* y, int64_t rows, int32_t in_length, int32_t fft_length, int32_t out_format, int device, void* stream) {
  return rfft_impl<float>(x, y, rows, in_length, fft_length, out_format, device, stream);
}
int dsb200_rfft_f64(const void* x, void* y, int64_t rows, int32_t in_length, int32_t fft_length, int32_t out_format, int device, void* stream) {
  return rfft_impl<double>(x, y, rows, in_length, fft_length, out_format, device, stream);
}
int dsb200_spec_f32(const void* b, int32_t bl, const void* a, int32_t al, void* y, int64_t rows, const dsb200_spec_params* p, int device, void* stream) {
  return spec_impl<float>(b, bl, a, al, y, rows, p, device, stream);
}
int dsb200_spec_f64(const void* b, int32_t bl, const void* a, int32_t al, void* y, int64_t rows, const dsb200_spec_params* p, int device, void* stream) {
  return spec_impl<double>(b, bl, a, al, y, rows, p, device, stream);
}
int dsb200_stft_f32(const void* x, const void* w, void* y, int64_t batch, int64_t T, const dsb200_stft_params* p, int device, void* stream) {
  return stft_impl<float>(x, w, y, batch, T, p, device, stream);
}
int dsb200_stft_f64(const void* x, const void* w, void* y, int64_t batch, int64_t T, const dsb200_stft_params* p, int device, void* stream) {
  return stft_impl<double>(x, w, y, batch, T, p, device, stream);
}

}  // extern "C"
