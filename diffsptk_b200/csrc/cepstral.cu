// Frame-rate cepstral kernels: row-matrix product (freqt, DCT), mel filter bank, MFCC and
// mel-cepstral analysis (fp32 + fp64).
//
// Reference semantics: diffsptk/modules/freqt.py:141-143, dct.py:135-137, fbank.py:305-321,
// mfcc.py:243-256, mcep.py:189-224.  All dense tables are built on the host from the
// reference's own float64 recursions and passed in as device pointers (see the header).
#include <algorithm>

#include "common.cuh"

namespace dsb200 {

// Fast fp32 path for fft_length = 512, cep_order <= 24 (mcep_fast.cu); DSB200_E_UNSUPPORTED outside it.
int mcep_fast_try(const float* x, float* y, int64_t rows, const dsb200_mcep_params* p, const float* P0,
                  const float* G, const float* Hm, const float* av, int device, cudaStream_t stream);

namespace {


// ------------------------------------------------------------------------------- rowmat
// y[r, :] = x[r, :] @ W.  A block stages RB rows of x (and W when it fits) in shared memory;
// thread -> (row, column) so that W reads are conflict-free and x reads are broadcasts.
template <typename T>
__global__ void __launch_bounds__(256) rowmat_kernel(const T* __restrict__ x, const T* __restrict__ W,
                                                     T* __restrict__ y, int64_t rows, int Din, int Dout,
                                                     int RB, int w_in_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* xs = reinterpret_cast<T*>(smem_raw);                 // [RB][Din]
  T* ws = xs + static_cast<size_t>(RB) * Din;             // [Din][Dout] (optional)
  const T* Wp = W;
  if (w_in_smem) {
    for (int i = threadIdx.x; i < Din * Dout; i += blockDim.x) ws[i] = W[i];
    Wp = ws;
  }
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * RB; base < rows; base += static_cast<int64_t>(gridDim.x) * RB) {
    const int nr = static_cast<int>(rows - base < RB ? rows - base : RB);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * Din; i += blockDim.x) xs[i] = x[base * Din + i];
    __syncthreads();
    for (int idx = threadIdx.x; idx < nr * Dout; idx += blockDim.x) {
      const int rl = idx / Dout, c = idx - rl * Dout;
      const T* xr = xs + static_cast<size_t>(rl) * Din;
      T a0 = 0, a1 = 0;
      int d = 0;
      for (; d + 1 < Din; d += 2) {
        a0 = dfma(xr[d], Wp[static_cast<size_t>(d) * Dout + c], a0);
        a1 = dfma(xr[d + 1], Wp[static_cast<size_t>(d + 1) * Dout + c], a1);
      }
      if (d < Din) a0 = dfma(xr[d], Wp[static_cast<size_t>(d) * Dout + c], a0);
      y[base * Dout + idx] = a0 + a1;
    }
  }
}

template <typename T>
int rowmat_impl(const void* x, const void* W, void* y, int64_t rows, int32_t Din, int32_t Dout, int device, void* stream) {
  DSB_REQUIRE(Din > 0 && Dout > 0, "matrix dimensions must be positive");
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(x != nullptr && W != nullptr && y != nullptr, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  int RB = std::max(1, std::min(64, 1024 / Dout));
  const size_t wbytes = static_cast<size_t>(Din) * Dout * sizeof(T);
  int w_in_smem = wbytes <= 96 * 1024 ? 1 : 0;
  auto bytes = [&](int rb) { return static_cast<size_t>(rb) * Din * sizeof(T) + (w_in_smem ? wbytes : 0); };
  while (RB > 1 && bytes(RB) > cap) RB /= 2;
  if (bytes(RB) > cap) return fail(DSB200_E_UNSUPPORTED, "row length %d does not fit in shared memory", Din);
  DSB_CUDA(cudaFuncSetAttribute(rowmat_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int64_t need = (rows + RB - 1) / RB;
  const int blocks = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(sm_count(device)) * 8));
  rowmat_kernel<T><<<blocks, 256, bytes(RB), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const T*>(x), static_cast<const T*>(W), static_cast<T*>(y), rows, Din, Dout, RB, w_in_smem);
  return after_launch("rowmat_kernel");
}

// ------------------------------------------------------------------------- fbank / mfcc
template <typename T>
struct FbankArgs {
  const T* x;   // [rows, K] power spectrum
  const T* H;   // [K, C]
  const int32_t* cb;  // [C] first non-zero row of each column (or null)
  const int32_t* ce;  // [C] one past the last non-zero row (or null)
  const T* W;   // mfcc: DCT basis [C, C]
  const T* lifter;  // mfcc: [M+1]
  T* y;
  T* E;
  int64_t rows;
  int K, C, use_power, want_energy, mfcc, M, out_format, D;
  T floor, gamma;
};

// Filter-bank stage for one row held by one warp: amplitude row `amp` (shared), result in `ych`.
template <typename T>
__device__ __forceinline__ void fbank_row(const FbankArgs<T>& A, const T* amp, T* ych, int lane) {
  for (int c = lane; c < A.C; c += 32) {
    const int lo = A.cb ? A.cb[c] : 0, hi = A.ce ? A.ce[c] : A.K;
    T a0 = 0, a1 = 0;
    int k = lo;
    for (; k + 1 < hi; k += 2) {
      a0 = dfma(amp[k], A.H[static_cast<size_t>(k) * A.C + c], a0);
      a1 = dfma(amp[k + 1], A.H[static_cast<size_t>(k + 1) * A.C + c], a1);
    }
    if (k < hi) a0 = dfma(amp[k], A.H[static_cast<size_t>(k) * A.C + c], a0);
    T v = dmax(a0 + a1, A.floor);
    v = (A.gamma == static_cast<T>(0)) ? dlog(v) : (dpow(v, A.gamma) - static_cast<T>(1)) / A.gamma;
    ych[c] = v;
  }
  __syncwarp();
}

template <typename T>
__global__ void __launch_bounds__(256) fbank_kernel(FbankArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  T* amp = reinterpret_cast<T*>(smem_raw) + static_cast<size_t>(warp) * (A.K + A.C);
  T* ych = amp + A.K;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + warp; row < A.rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    const T* xr = A.x + row * A.K;
    T esum = 0;
    for (int k = lane; k < A.K; k += 32) {
      const T v = xr[k];
      esum += (k == 0 || k == A.K - 1) ? v : static_cast<T>(2) * v;
      amp[k] = A.use_power ? v : dsqrt(v);
    }
    __syncwarp();
    T En = 0;
    if (A.want_energy) En = dlog(warp_sum(esum) / static_cast<T>(2 * (A.K - 1)));
    fbank_row<T>(A, amp, ych, lane);
    if (!A.mfcc) {
      T* yr = A.y + row * A.D;
      for (int c = lane; c < A.C; c += 32) yr[c] = ych[c];
      if (A.want_energy && lane == 0) {
        if (A.E) A.E[row] = En; else yr[A.C] = En;
      }
    } else {
      // DCT-II (first M+1 columns) * lifter, then pack:  y | yE | yc | ycE
      T* yr = A.y + row * A.D;
      for (int m = lane; m <= A.M; m += 32) {
        T acc = 0;
        for (int c = 0; c < A.C; ++c) acc = dfma(ych[c], A.W[static_cast<size_t>(c) * A.C + m], acc);
        acc *= A.lifter[m];
        if (m > 0) yr[m - 1] = acc;
        else if (A.out_format == DSB200_MFCC_YC || A.out_format == DSB200_MFCC_YCE) yr[A.M] = acc;
      }
      if (lane == 0) {
        if (A.out_format == DSB200_MFCC_YE) yr[A.M] = En;
        if (A.out_format == DSB200_MFCC_YCE) yr[A.M + 1] = En;
      }
    }
    __syncwarp();
  }
}

template <typename T>
int launch_fbank(FbankArgs<T>& A, int device, cudaStream_t stream) {
  const size_t per_warp = static_cast<size_t>(A.K + A.C) * sizeof(T);
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  if (per_warp > cap) return fail(DSB200_E_UNSUPPORTED, "spectrum row does not fit in shared memory");
  int wpb = static_cast<int>(std::min<size_t>(8, cap / per_warp));
  while (wpb > 1 && wpb * per_warp > 48 * 1024) --wpb;
  DSB_CUDA(cudaFuncSetAttribute(fbank_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int64_t need = (A.rows + wpb - 1) / wpb;
  const int blocks = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(sm_count(device)) * 16));
  fbank_kernel<T><<<blocks, wpb * 32, wpb * per_warp, stream>>>(A);
  return after_launch("fbank_kernel");
}

int check_fbank(const dsb200_fbank_params* p) {
  DSB_REQUIRE(p != nullptr, "fbank params are NULL");
  DSB_REQUIRE(p->fft_length > 1, "fft_length must be greater than 1.");
  DSB_REQUIRE(p->fft_length % 2 == 0, "fft_length must be even");
  DSB_REQUIRE(p->n_channel > 0, "n_channel must be positive.");
  DSB_REQUIRE(p->floor > 0, "floor must be positive.");
  DSB_REQUIRE(p->gamma >= -1 && p->gamma <= 1, "gamma must be in [-1, 1].");
  return DSB200_OK;
}

template <typename T>
int fbank_impl(const void* x, const void* H, const int32_t* cb, const int32_t* ce, void* y, void* E, int64_t rows,
               const dsb200_fbank_params* p, int device, void* stream) {
  if (int rc = check_fbank(p)) return rc;
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(x != nullptr && H != nullptr && y != nullptr, "NULL data pointer");
  DSB_REQUIRE((cb == nullptr) == (ce == nullptr), "col_begin and col_end must be given together");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  FbankArgs<T> A{};
  A.x = static_cast<const T*>(x);
  A.H = static_cast<const T*>(H);
  A.cb = cb;
  A.ce = ce;
  A.y = static_cast<T*>(y);
  A.E = static_cast<T*>(E);
  A.rows = rows;
  A.K = p->fft_length / 2 + 1;
  A.C = p->n_channel;
  A.use_power = p->use_power;
  A.want_energy = p->want_energy;
  A.D = A.C + ((p->want_energy && E == nullptr) ? 1 : 0);
  A.floor = static_cast<T>(p->floor);
  A.gamma = static_cast<T>(p->gamma);
  return launch_fbank<T>(A, device, static_cast<cudaStream_t>(stream));
}

template <typename T>
int mfcc_impl(const void* x, const void* H, const int32_t* cb, const int32_t* ce, const void* W, const void* lifter,
              void* y, int64_t rows, const dsb200_mfcc_params* p, int device, void* stream) {
  DSB_REQUIRE(p != nullptr, "mfcc params are NULL");
  if (int rc = check_fbank(&p->fbank)) return rc;
  DSB_REQUIRE(p->mfcc_order >= 0, "mfcc_order must be non-negative.");
  DSB_REQUIRE(p->mfcc_order < p->fbank.n_channel, "mfcc_order must be less than n_channel.");
  DSB_REQUIRE(p->out_format >= DSB200_MFCC_Y && p->out_format <= DSB200_MFCC_YCE, "out_format %d is not supported.", p->out_format);
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(x != nullptr && H != nullptr && W != nullptr && lifter != nullptr && y != nullptr, "NULL data pointer");
  DSB_REQUIRE((cb == nullptr) == (ce == nullptr), "col_begin and col_end must be given together");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  FbankArgs<T> A{};
  A.x = static_cast<const T*>(x);
  A.H = static_cast<const T*>(H);
  A.cb = cb;
  A.ce = ce;
  A.W = static_cast<const T*>(W);
  A.lifter = static_cast<const T*>(lifter);
  A.y = static_cast<T*>(y);
  A.rows = rows;
  A.K = p->fbank.fft_length / 2 + 1;
  A.C = p->fbank.n_channel;
  A.use_power = 0;
  A.want_energy = (p->out_format == DSB200_MFCC_YE || p->out_format == DSB200_MFCC_YCE);
  A.mfcc = 1;
  A.M = p->mfcc_order;
  A.out_format = p->out_format;
  A.D = A.M + (p->out_format == DSB200_MFCC_Y ? 0 : (p->out_format == DSB200_MFCC_YCE ? 2 : 1));
  A.floor = static_cast<T>(p->fbank.floor);
  A.gamma = static_cast<T>(p->fbank.gamma);
  return launch_fbank<T>(A, device, static_cast<cudaStream_t>(stream));
}

// --------------------------------------------------------------------------------- mcep
// One warp per spectrum row.  Per Newton step (mcep.py:209-222), with the FFTs folded into the
// host-built matrices G and Hm:
//   d  = mc @ G                 (K values, lanes over bins)
//   e  = exp(log x - 2 d)
//   rt = e @ Hm                 (2M+1 values, lanes over bins + warp reduction)
//   solve (Toeplitz(rt[:M+1]) + Hankel(rt)) g = rt[:M+1] - alpha_vector ;  mc += g
template <typename T>
struct McepArgs {
  const T* x;
  T* mc;
  const T* P0;  // [K, D]
  const T* G;   // [D, K]
  const T* Hm;  // [K, J]
  const T* av;  // [D]
  int64_t rows;
  int K, D, J, n_iter;
  int g_in_smem, h_in_smem;
};

template <typename T, int JMAX>
__global__ void __launch_bounds__(256) mcep_kernel(McepArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int K = A.K, D = A.D, J = A.J;
  const int ldm = (D + 1) | 1;  // odd row stride of the augmented system -> conflict-free columns
  T* p = reinterpret_cast<T*>(smem_raw);
  const T* Gp = A.G;
  const T* Hp = A.Hm;
  if (A.g_in_smem) {
    for (int i = threadIdx.x; i < D * K; i += blockDim.x) p[i] = A.G[i];
    Gp = p;
    p += static_cast<size_t>(D) * K;
  }
  if (A.h_in_smem) {
    for (int i = threadIdx.x; i < K * J; i += blockDim.x) p[i] = A.Hm[i];
    Hp = p;
    p += static_cast<size_t>(K) * J;
  }
  __syncthreads();
  const size_t per_warp = static_cast<size_t>(2 * K + D + J + D * ldm);
  T* logx = p + warp * per_warp;
  T* e = logx + K;
  T* mc = e + K;
  T* rt = mc + D;
  T* Aug = rt + J;  // [D][ldm]

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + warp; row < A.rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    const T* xr = A.x + row * K;
    for (int k = lane; k < K; k += 32) logx[k] = dlog(xr[k]);
    __syncwarp();
    // initial estimate: mc = log_x @ P0
    for (int m = lane; m < D; m += 32) {
      T a0 = 0, a1 = 0;
      int k = 0;
      for (; k + 1 < K; k += 2) {
        a0 = dfma(logx[k], A.P0[static_cast<size_t>(k) * D + m], a0);
        a1 = dfma(logx[k + 1], A.P0[static_cast<size_t>(k + 1) * D + m], a1);
      }
      if (k < K) a0 = dfma(logx[k], A.P0[static_cast<size_t>(k) * D + m], a0);
      mc[m] = a0 + a1;
    }
    __syncwarp();

    for (int it = 0; it < A.n_iter; ++it) {
      // e = exp(log x - 2 * (mc @ G))
      for (int k = lane; k < K; k += 32) {
        T d = 0;
        for (int m = 0; m < D; ++m) d = dfma(mc[m], Gp[static_cast<size_t>(m) * K + k], d);
        e[k] = dexp(logx[k] - d - d);
      }
      __syncwarp();
      // rt = e @ Hm : per-lane partial sums over its bins, then a warp reduction
      T acc[JMAX];
#pragma unroll
      for (int j = 0; j < JMAX; ++j) acc[j] = 0;
      for (int k = lane; k < K; k += 32) {
        const T ek = e[k];
        const T* hr = Hp + static_cast<size_t>(k) * J;
#pragma unroll
        for (int j = 0; j < JMAX; ++j)
          if (j < J) acc[j] = dfma(ek, hr[j], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < JMAX; ++j) {
        if (j < J) {
          const T s = warp_sum(acc[j]);
          if (lane == 0) rt[j] = s;
        }
      }
      __syncwarp();
      // augmented system [R + Q | ra]
      for (int idx = lane; idx < D * (D + 1); idx += 32) {
        const int i = idx / (D + 1), j = idx - i * (D + 1);
        T v;
        if (j < D) v = rt[i > j ? i - j : j - i] + rt[i + j];
        else v = rt[i] - A.av[i];
        Aug[i * ldm + j] = v;
      }
      __syncwarp();
      // Gaussian elimination (the Newton matrix is symmetric positive definite: no pivoting)
      for (int pc = 0; pc < D - 1; ++pc) {
        const T piv = Aug[pc * ldm + pc];
        for (int i = pc + 1 + lane; i < D; i += 32) {
          const T f = Aug[i * ldm + pc] / piv;
          for (int c = pc + 1; c <= D; ++c) Aug[i * ldm + c] = dfma(-f, Aug[pc * ldm + c], Aug[i * ldm + c]);
        }
        __syncwarp();
      }
      // back substitution; solution overwrites column D
      for (int i = D - 1; i >= 0; --i) {
        T s = 0;
        for (int c = i + 1 + lane; c < D; c += 32) s = dfma(Aug[i * ldm + c], Aug[c * ldm + D], s);
        s = warp_sum(s);
        if (lane == 0) Aug[i * ldm + D] = (Aug[i * ldm + D] - s) / Aug[i * ldm + i];
        __syncwarp();
      }
      for (int m = lane; m < D; m += 32) mc[m] += Aug[m * ldm + D];
      __syncwarp();
    }
    T* out = A.mc + row * D;
    for (int m = lane; m < D; m += 32) out[m] = mc[m];
    __syncwarp();
  }
}

template <typename T, int JMAX>
int launch_mcep(McepArgs<T>& A, int device, cudaStream_t stream) {
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  const int ldm = (A.D + 1) | 1;
  const size_t per_warp = static_cast<size_t>(2 * A.K + A.D + A.J + A.D * ldm) * sizeof(T);
  const size_t gb = static_cast<size_t>(A.D) * A.K * sizeof(T), hb = static_cast<size_t>(A.K) * A.J * sizeof(T);
  int wpb = 8;
  while (wpb > 1 && wpb * per_warp > cap / 2) --wpb;
  if (wpb * per_warp > cap) return fail(DSB200_E_UNSUPPORTED, "mcep working set does not fit in shared memory");
  size_t used = wpb * per_warp;
  A.h_in_smem = (used + hb <= cap) ? 1 : 0;
  if (A.h_in_smem) used += hb;
  A.g_in_smem = (used + gb <= cap) ? 1 : 0;
  if (A.g_in_smem) used += gb;
  DSB_CUDA(cudaFuncSetAttribute(mcep_kernel<T, JMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int64_t need = (A.rows + wpb - 1) / wpb;
  const int blocks = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(sm_count(device)) * 2));
  mcep_kernel<T, JMAX><<<blocks, wpb * 32, used, stream>>>(A);
  return after_launch("mcep_kernel");
}

template <typename T>
int mcep_impl(const void* x, void* mc, int64_t rows, const dsb200_mcep_params* p, const void* P0, const void* G,
              const void* Hm, const void* av, int device, void* stream) {
  DSB_REQUIRE(p != nullptr, "mcep params are NULL");
  DSB_REQUIRE(p->fft_length > 1, "fft_length must be greater than 1.");
  DSB_REQUIRE(p->fft_length % 2 == 0, "fft_length must be even");
  DSB_REQUIRE(p->cep_order >= 0, "cep_order must be non-negative.");
  DSB_REQUIRE(p->fft_length >= 2 * p->cep_order, "cep_order must be less than or equal to fft_length // 2.");
  DSB_REQUIRE(p->n_iter >= 0, "n_iter must be non-negative.");
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(x != nullptr && mc != nullptr && P0 != nullptr && G != nullptr && Hm != nullptr && av != nullptr, "NULL data pointer");
  if (p->cep_order > 63) return fail(DSB200_E_UNSUPPORTED, "cep_order > 63 is not implemented in the fused mcep kernel");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  if (sizeof(T) == 4 && getenv("DSB200_MCEP_GENERIC") == nullptr) {
    const int rc = mcep_fast_try(static_cast<const float*>(x), static_cast<float*>(mc), rows, p,
                                 static_cast<const float*>(P0), static_cast<const float*>(G),
                                 static_cast<const float*>(Hm), static_cast<const float*>(av), device,
                                 static_cast<cudaStream_t>(stream));
    if (rc != DSB200_E_UNSUPPORTED) return rc;
  }
  McepArgs<T> A{};
  A.x = static_cast<const T*>(x);
  A.mc = static_cast<T*>(mc);
  A.P0 = static_cast<const T*>(P0);
  A.G = static_cast<const T*>(G);
  A.Hm = static_cast<const T*>(Hm);
  A.av = static_cast<const T*>(av);
  A.rows = rows;
  A.K = p->fft_length / 2 + 1;
  A.D = p->cep_order + 1;
  A.J = 2 * p->cep_order + 1;
  A.n_iter = p->n_iter;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (A.J <= 16) return launch_mcep<T, 16>(A, device, s);
  if (A.J <= 32) return launch_mcep<T, 32>(A, device, s);
  if (A.J <= 64) return launch_mcep<T, 64>(A, device, s);
  return launch_mcep<T, 128>(A, device, s);
}

}  // namespace
}  // namespace dsb200

using namespace dsb200;

extern "C" {

int dsb200_rowmat_f32(const void* x, const void* W, void* y, int64_t rows, int32_t Din, int32_t Dout, int device, void* stream) {
  return rowmat_impl<float>(x, W, y, rows, Din, Dout, device, stream);
}
int dsb200_rowmat_f64(const void* x, const void* W, void* y, int64_t rows, int32_t Din, int32_t Dout, int device, void* stream) {
  return rowmat_impl<double>(x, W, y, rows, Din, Dout, device, stream);
}
int dsb200_fbank_f32(const void* x, const void* H, const int32_t* cb, const int32_t* ce, void* y, void* E, int64_t rows,
                     const dsb200_fbank_params* p, int device, void* stream) {
  return fbank_impl<float>(x, H, cb, ce, y, E, rows, p, device, stream);
}
int dsb200_fbank_f64(const void* x, const void* H, const int32_t* cb, const int32_t* ce, void* y, void* E, int64_t rows,
                     const dsb200_fbank_params* p, int device, void* stream) {
  return fbank_impl<double>(x, H, cb, ce, y, E, rows, p, device, stream);
}
int dsb200_mfcc_f32(const void* x, const void* H, const int32_t* cb, const int32_t* ce, const void* W, const void* lifter,
                    void* y, int64_t rows, const dsb200_mfcc_params* p, int device, void* stream) {
  return mfcc_impl<float>(x, H, cb, ce, W, lifter, y, rows, p, device, stream);
}
int dsb200_mfcc_f64(const void* x, const void* H, const int32_t* cb, const int32_t* ce, const void* W, const void* lifter,
                    void* y, int64_t rows, const dsb200_mfcc_params* p, int device, void* stream) {
  return mfcc_impl<double>(x, H, cb, ce, W, lifter, y, rows, p, device, stream);
}
int dsb200_mcep_f32(const void* x, void* mc, int64_t rows, const dsb200_mcep_params* p, const void* P0, const void* G,
                    const void* Hm, const void* av, int device, void* stream) {
  return mcep_impl<float>(x, mc, rows, p, P0, G, Hm, av, device, stream);
}
int dsb200_mcep_f64(const void* x, void* mc, int64_t rows, const dsb200_mcep_params* p, const void* P0, const void* G,
                    const void* Hm, const void* av, int device, void* stream) {
  return mcep_impl<double>(x, mc, rows, p, P0, G, Hm, av, device, stream);
}

}  // extern "C"
