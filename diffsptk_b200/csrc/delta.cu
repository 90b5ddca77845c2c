// Delta (regression) features over the frame axis: diffsptk/modules/delta.py:172-194.
//   y[b, t, h D + d] = sum_w window[h, w] x[b, clamp(t + w - (W-1)/2, 0, T-1), d]      (replicate padding)
// A pure HBM-bound stencil: D * 4 B read (the W-row halo stays in L1/L2) and H * D * 4 B written per frame.
// Forward: CTA tiles of frames staged in shared memory, fully coalesced loads and stores.  The backward kernel is
// the adjoint gather, one thread per (frame, feature) pair (border frames collect the clamped taps).
#include <algorithm>

#include "common.cuh"

namespace dsb200 {
namespace {

constexpr int kMaxTaps = 1024;   // H * W window coefficients kept in shared memory
constexpr int kPF = 8;           // staged input elements per thread (256 threads): a tile holds <= 2048 of them

// A CTA owns a tile of `tt` consecutive frames of one utterance: the tt + W - 1 input rows it needs (clamped at the
// utterance ends = replicate padding) are staged in shared memory with coalesced loads, and the tile's
// tt x (H D) outputs, one contiguous span of y, are written with coalesced stores.
// WC > 0: the window length is a compile-time constant (taps in registers, unrolled); WC == 0: any odd length.
template <typename T, int WC>
__global__ void __launch_bounds__(256) delta_kernel(const T* __restrict__ x, const T* __restrict__ win,
                                                    T* __restrict__ y, int64_t batch, int64_t Tn, int D, int Hn,
                                                    int Wrt, int tt, int64_t tiles_per_utt) {
  const int W = WC > 0 ? WC : Wrt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* ws = reinterpret_cast<T*>(smem_raw);      // [H W]
  T* xs = ws + Hn * W;                         // [tt + W - 1][D] staged input rows
  T* ys = xs + (tt + W - 1) * D;               // [tt][H D] staged outputs
  for (int i = threadIdx.x; i < Hn * W; i += blockDim.x) ws[i] = win[i];
  const int pad = (W - 1) / 2, HD = Hn * D;
  // (row, feature) of flat index tid, advanced by blockDim per iteration without divisions
  const int nthr = blockDim.x, step_r = nthr / D, step_d = nthr - step_r * D;
  const int r0 = threadIdx.x / D, d0 = threadIdx.x - r0 * D;
  const int64_t n_tiles = batch * tiles_per_utt;
  // Software pipeline: the rows of the NEXT tile are fetched into registers (<= kPF elements per thread, the host
  // sizes the tile accordingly) before the current tile is computed, so their HBM latency hides behind it.
  T pf[kPF];
  auto fetch = [&](int64_t tile) {
    const int64_t b = tile / tiles_per_utt;
    const int64_t t0 = (tile - b * tiles_per_utt) * tt;
    const int nt = static_cast<int>(Tn - t0 < tt ? Tn - t0 : tt);
    const T* xb = x + b * Tn * D;
    int r = r0, d = d0;
#pragma unroll
    for (int k = 0; k < kPF; ++k) {
      if (threadIdx.x + k * nthr < (nt + W - 1) * D) {
        int64_t t = t0 + r - pad;
        t = t < 0 ? 0 : (t > Tn - 1 ? Tn - 1 : t);
        pf[k] = xb[t * D + d];
      }
      r += step_r;
      d += step_d;
      if (d >= D) { d -= D; ++r; }
    }
  };
  if (blockIdx.x < n_tiles) fetch(blockIdx.x);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t b = tile / tiles_per_utt;
    const int64_t t0 = (tile - b * tiles_per_utt) * tt;
    const int nt = static_cast<int>(Tn - t0 < tt ? Tn - t0 : tt);
    __syncthreads();                           // previous tile's copy-out (and the window table) is done
#pragma unroll
    for (int k = 0; k < kPF; ++k)
      if (threadIdx.x + k * nthr < (nt + W - 1) * D) xs[threadIdx.x + k * nthr] = pf[k];
    if (tile + gridDim.x < n_tiles) fetch(tile + gridDim.x);
    __syncthreads();
    for (int i = threadIdx.x, r = r0, d = d0; i < nt * D; i += nthr) {
      const T* xp = xs + i;                    // tap w of frame r, feature d is xs[(r + w) D + d]
      T* yp = ys + r * HD + d;
      if constexpr (WC > 0) {
        T tap[WC];
#pragma unroll
        for (int w = 0; w < WC; ++w) tap[w] = xp[w * D];
        for (int h = 0; h < Hn; ++h) {
          T acc = 0;
#pragma unroll
          for (int w = 0; w < WC; ++w) acc = dfma(ws[h * WC + w], tap[w], acc);
          yp[h * D] = acc;
        }
      } else {
        for (int h = 0; h < Hn; ++h) {
          T acc = 0;
          for (int w = 0; w < W; ++w) acc = dfma(ws[h * W + w], xp[w * D], acc);
          yp[h * D] = acc;
        }
      }
      r += step_r;
      d += step_d;
      if (d >= D) { d -= D; ++r; }
    }
    __syncthreads();
    T* yo = y + (b * Tn + t0) * HD;            // the tile's outputs are one contiguous span
    for (int o = threadIdx.x; o < nt * HD; o += nthr) yo[o] = ys[o];
  }
}

// Fallback for feature dimensions too large for the tile kernel: one thread per (frame, feature) pair.
template <typename T>
__global__ void __launch_bounds__(256) delta_wide_kernel(const T* __restrict__ x, const T* __restrict__ win,
                                                         T* __restrict__ y, int64_t batch, int64_t Tn, int D, int Hn,
                                                         int W) {
  __shared__ T ws[kMaxTaps];
  for (int i = threadIdx.x; i < Hn * W; i += blockDim.x) ws[i] = win[i];
  __syncthreads();
  const int pad = (W - 1) / 2;
  const int64_t total = batch * Tn * D;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t bt = i / D;
    const int d = static_cast<int>(i - bt * D);
    const int64_t b = bt / Tn, t = bt - b * Tn;
    const T* xb = x + b * Tn * D + d;
    T* yo = y + bt * (static_cast<int64_t>(Hn) * D) + d;
    for (int h = 0; h < Hn; ++h) {
      T acc = 0;
      for (int w = 0; w < W; ++w) {
        int64_t tt = t + w - pad;
        tt = tt < 0 ? 0 : (tt > Tn - 1 ? Tn - 1 : tt);
        acc = dfma(ws[h * W + w], xb[tt * D], acc);
      }
      yo[static_cast<int64_t>(h) * D] = acc;
    }
  }
}

// gx[b, t', d] = sum_{h, w} window[h, w] sum_{t : clamp(t + w - pad) = t'} gy[b, t, h D + d]
template <typename T>
__global__ void __launch_bounds__(256) delta_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ win,
                                                        T* __restrict__ gx, int64_t batch, int64_t Tn, int D,
                                                        int Hn, int W) {
  __shared__ T ws[kMaxTaps];
  for (int i = threadIdx.x; i < Hn * W; i += blockDim.x) ws[i] = win[i];
  __syncthreads();
  const int pad = (W - 1) / 2;
  const int64_t total = batch * Tn * D;
  const int64_t HD = static_cast<int64_t>(Hn) * D;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t bt = i / D;
    const int d = static_cast<int>(i - bt * D);
    const int64_t b = bt / Tn, tp = bt - b * Tn;
    const T* gb = gy + b * Tn * HD + d;
    T acc = 0;
    auto collect = [&](int64_t t, int w) {     // all H windows of output frame t, tap w
      for (int h = 0; h < Hn; ++h) acc = dfma(ws[h * W + w], gb[t * HD + static_cast<int64_t>(h) * D], acc);
    };
    for (int w = 0; w < W; ++w) {
      const int64_t off = w - pad;             // output frame t reads input frame clamp(t + off, 0, T - 1)
      const int64_t t = tp - off;              // the unclamped reader
      if (t >= 0 && t <= Tn - 1) collect(t, w);
      if (tp == 0) {                           // readers clamped up onto the first frame: t + off < 0
        const int64_t hi = (-off - 1 < Tn - 1) ? -off - 1 : Tn - 1;
        for (int64_t u = 0; u <= hi; ++u) collect(u, w);
      }
      if (tp == Tn - 1) {                      // readers clamped down onto the last frame: t + off > T - 1
        const int64_t lo = (Tn - off > 0) ? Tn - off : 0;
        for (int64_t u = lo; u <= Tn - 1; ++u) collect(u, w);
      }
    }
    gx[i] = acc;
  }
}

int check_delta(int64_t batch, int64_t Tn, int32_t D, int32_t Hn, int32_t W) {
  DSB_REQUIRE(batch >= 0 && Tn >= 1, "need at least one frame");
  DSB_REQUIRE(D >= 1 && Hn >= 1, "feature and window counts must be positive");
  DSB_REQUIRE(W >= 1 && (W & 1), "the regression window length must be odd");
  if (Hn * W > kMaxTaps) return fail(DSB200_E_UNSUPPORTED, "more than %d window coefficients", kMaxTaps);
  return DSB200_OK;
}

template <typename T, bool BWD>
int delta_impl(const void* in, const void* win, void* out, int64_t batch, int64_t Tn, int32_t D, int32_t Hn,
               int32_t W, int device, void* stream) {
  if (int rc = check_delta(batch, Tn, D, Hn, W)) return rc;
  if (batch == 0) return DSB200_OK;
  DSB_REQUIRE(in && win && out, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (BWD) {
    const int64_t total = batch * Tn * D;
    const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(sm_count(device)) * 16));
    delta_bwd_kernel<T><<<blocks, 256, 0, s>>>(static_cast<const T*>(in), static_cast<const T*>(win),
                                               static_cast<T*>(out), batch, Tn, D, Hn, W);
  } else {
    // frames per tile: (tt + W - 1) D staged elements must fit the kPF x 256 register prefetch; at most 128 frames
    if (static_cast<int64_t>(W) * D > kPF * 256) {   // a single frame's halo does not fit the tile kernel
      const int64_t total = batch * Tn * D;
      const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(sm_count(device)) * 16));
      delta_wide_kernel<T><<<blocks, 256, 0, s>>>(static_cast<const T*>(in), static_cast<const T*>(win),
                                                  static_cast<T*>(out), batch, Tn, D, Hn, W);
      return after_launch("delta_wide_kernel");
    }
    int tt = static_cast<int>(std::min<int64_t>(128, (kPF * 256) / D - (W - 1)));
    if (tt > Tn) tt = static_cast<int>(Tn);
    const size_t smem = (static_cast<size_t>(Hn) * W + static_cast<size_t>(tt + W - 1) * D +
                         static_cast<size_t>(tt) * Hn * D) * sizeof(T);
    if (smem > static_cast<size_t>(max_dynamic_smem(device)))
      return fail(DSB200_E_UNSUPPORTED, "feature dimension %d is too large for the delta kernel's shared memory", D);
    const int64_t tiles_per_utt = (Tn + tt - 1) / tt;
    const int blocks = static_cast<int>(std::min<int64_t>(batch * tiles_per_utt, static_cast<int64_t>(sm_count(device)) * 8));
    auto launch = [&](auto kern) -> int {
      DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      kern<<<blocks, 256, smem, s>>>(static_cast<const T*>(in), static_cast<const T*>(win), static_cast<T*>(out), batch,
                                     Tn, D, Hn, W, tt, tiles_per_utt);
      return DSB200_OK;
    };
    int rc;
    switch (W) {
      case 1: rc = launch(delta_kernel<T, 1>); break;
      case 3: rc = launch(delta_kernel<T, 3>); break;
      case 5: rc = launch(delta_kernel<T, 5>); break;
      case 7: rc = launch(delta_kernel<T, 7>); break;
      case 9: rc = launch(delta_kernel<T, 9>); break;
      default: rc = launch(delta_kernel<T, 0>); break;
    }
    if (rc != DSB200_OK) return rc;
  }
  return after_launch(BWD ? "delta_bwd_kernel" : "delta_kernel");
}

}  // namespace
}  // namespace dsb200

using namespace dsb200;

extern "C" {

int dsb200_delta_f32(const void* x, const void* window, void* y, int64_t batch, int64_t n_frames, int32_t dim,
                     int32_t n_windows, int32_t width, int device, void* stream) {
  return delta_impl<float, false>(x, window, y, batch, n_frames, dim, n_windows, width, device, stream);
}
int dsb200_delta_f64(const void* x, const void* window, void* y, int64_t batch, int64_t n_frames, int32_t dim,
                     int32_t n_windows, int32_t width, int device, void* stream) {
  return delta_impl<double, false>(x, window, y, batch, n_frames, dim, n_windows, width, device, stream);
}
int dsb200_delta_backward_f32(const void* gy, const void* window, void* gx, int64_t batch, int64_t n_frames,
                              int32_t dim, int32_t n_windows, int32_t width, int device, void* stream) {
  return delta_impl<float, true>(gy, window, gx, batch, n_frames, dim, n_windows, width, device, stream);
}
int dsb200_delta_backward_f64(const void* gy, const void* window, void* gx, int64_t batch, int64_t n_frames,
                              int32_t dim, int32_t n_windows, int32_t width, int device, void* stream) {
  return delta_impl<double, true>(gy, window, gx, batch, n_frames, dim, n_windows, width, device, stream);
}

}  // extern "C"
