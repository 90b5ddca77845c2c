// Delta (regression) features over the frame axis: diffsptk/modules/delta.py:172-194.
//   y[b, t, h D + d] = sum_w window[h, w] x[b, clamp(t + w - (W-1)/2, 0, T-1), d]      (replicate padding)
// A pure HBM-bound stencil: D * 4 B read (the W-row halo stays in L1/L2) and H * D * 4 B written per frame.
// One thread per (frame, feature) pair, consecutive threads on consecutive features: every load and store is
// coalesced.  The backward kernel is the adjoint gather (border frames collect the clamped taps).
#include <algorithm>

#include "common.cuh"

namespace dsb200 {
namespace {

constexpr int kMaxTaps = 1024;   // H * W window coefficients kept in shared memory

template <typename T>
__global__ void __launch_bounds__(256) delta_kernel(const T* __restrict__ x, const T* __restrict__ win,
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
  const int64_t total = batch * Tn * D;
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(sm_count(device)) * 16));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (BWD)
    delta_bwd_kernel<T><<<blocks, 256, 0, s>>>(static_cast<const T*>(in), static_cast<const T*>(win),
                                               static_cast<T*>(out), batch, Tn, D, Hn, W);
  else
    delta_kernel<T><<<blocks, 256, 0, s>>>(static_cast<const T*>(in), static_cast<const T*>(win),
                                           static_cast<T*>(out), batch, Tn, D, Hn, W);
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
