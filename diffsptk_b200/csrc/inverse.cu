// The inverse of the hot path: inverse real FFT (ifftr.py:130-143), windowed
// overlap-add with sum-of-squares normalisation (unframe.py:164-211) and their fusion, the inverse STFT
// (istft.py:186-193).
//
// istft_kernel is output-stationary: a CTA owns a tile of consecutive output samples of one utterance, inverse-
// transforms the <= tile_frames + ceil(L / P) - 1 frames that overlap it (one warp per frame, half-length complex
// FFT in shared memory), keeps the windowed frames on chip and sums them per output sample in frame order
// (F.fold's order), so the [B, N, L] frame tensor never exists in HBM: 2 (K) * 8 B read + P * 4 B written per
// frame instead of an extra 2 * L * 4 B round trip.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "rowfft.cuh"

namespace dsb200 {

// Fast fused kernel for fft_length = 512 (istft512.cu); DSB200_E_UNSUPPORTED outside its envelope.
int istft512_try(const float* Y, const float* w, float* out, int64_t batch, int64_t N, int64_t T_out, int L, int P,
                 int n, int center, int device, cudaStream_t stream);

namespace {

// One spectrum row Y[0..Nc] (global memory) -> min(out_len, n) real samples dst[j] (* win[j] if win != nullptr).
// buf0 / buf1: Nc + 1 complex entries each, private to the warp.  The imaginary parts of the DC and Nyquist
// bins are ignored, as torch.fft.irfft does.
template <typename T>
__device__ void warp_irfft(const cx_t<T>* __restrict__ Y, cx_t<T>* buf0, cx_t<T>* buf1, int n, int Nc, int pow2,
                           const cx_t<T>* tw, int lane, int out_len, T* dst, const T* win) {
  using C = cx_t<T>;
  if (pow2 && Nc >= 2) {
    // x[2m] + i x[2m+1] = IFFT_Nc(E + i O), E = (Y[k] + conj Y[Nc-k]) / 2, O = (Y[k] - conj Y[Nc-k]) / 2 * W_n^-k.
    // The inverse transform runs as conj(FFT(conj .)) on the forward butterflies.
    for (int k = lane; k < Nc; k += 32) {
      C a = Y[k], b = Y[Nc - k];
      if (k == 0) { a.y = 0; b.y = 0; }
      b.y = -b.y;
      const T h = static_cast<T>(0.5);
      const C E = mk<T>(h * (a.x + b.x), h * (a.y + b.y));
      const C D = mk<T>(h * (a.x - b.x), h * (a.y - b.y));
      C w = tw[k];
      w.y = -w.y;                       // W_n^-k
      const C O = cmul(D, w);
      buf0[k] = mk<T>(E.x - O.y, -(E.y + O.x));   // conj(E + i O)
    }
    __syncwarp();
    const C* R = warp_fft_pow2<T>(buf0, buf1, Nc, tw, lane, n);
    const T sc = static_cast<T>(1) / static_cast<T>(Nc);
    for (int m = lane; m < Nc; m += 32) {
      const C r = R[m];
      const int j = 2 * m;
      if (j < out_len) dst[j] = r.x * sc * (win ? win[j] : static_cast<T>(1));
      if (j + 1 < out_len) dst[j + 1] = -r.y * sc * (win ? win[j + 1] : static_cast<T>(1));
    }
  } else {
    for (int k = lane; k <= Nc; k += 32) buf0[k] = Y[k];
    __syncwarp();
    const T sc = static_cast<T>(1) / static_cast<T>(n);
    for (int j = lane; j < out_len; j += 32) {
      T acc = buf0[0].x + ((j & 1) ? -buf0[Nc].x : buf0[Nc].x);
      int idx = j % n;
      for (int k = 1; k < Nc; ++k) {
        const C w = tw[idx];            // (cos, -sin)(2 pi k j / n)
        acc += static_cast<T>(2) * (buf0[k].x * w.x + buf0[k].y * w.y);
        idx += j;
        if (idx >= n) idx -= n;
      }
      dst[j] = acc * sc * (win ? win[j] : static_cast<T>(1));
    }
  }
  __syncwarp();
}

template <typename T>
__global__ void __launch_bounds__(256) ifftr_kernel(const cx_t<T>* __restrict__ Y, T* __restrict__ x, int64_t rows,
                                                    int n, int out_len, int pow2, const cx_t<T>* __restrict__ twg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using C = cx_t<T>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int Nc = n / 2, K = Nc + 1;
  C* tw = reinterpret_cast<C*>(smem_raw);
  for (int i = threadIdx.x; i < n; i += blockDim.x) tw[i] = twg[i];
  C* buf0 = tw + n + static_cast<size_t>(warp) * 2 * K;
  C* buf1 = buf0 + K;
  __syncthreads();
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + warp; row < rows;
       row += static_cast<int64_t>(gridDim.x) * wpb)
    warp_irfft<T>(Y + row * K, buf0, buf1, n, Nc, pow2, tw, lane, out_len, x + row * out_len, nullptr);
}

// out[b, t] = sum_n fr[b, n, q - n P] w[q - n P] / (sum_n w[q - n P]^2 + 1e-16), q = t + s.
template <typename T>
__global__ void __launch_bounds__(256) unframe_kernel(const T* __restrict__ fr, const T* __restrict__ w,
                                                      T* __restrict__ out, int64_t batch, int64_t N, int64_t T_out,
                                                      int L, int P, int s) {
  const int64_t total = batch * T_out;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = i / T_out, t = i - b * T_out, q = t + s;
    int64_t n_lo = (q - L + P) / P;            // ceil((q - L + 1) / P) for q - L + 1 > 0
    if (q - L + 1 <= 0) n_lo = 0;
    int64_t n_hi = q / P;
    if (n_hi > N - 1) n_hi = N - 1;
    T num = 0, den = 0;
    for (int64_t n = n_lo; n <= n_hi; ++n) {
      const int j = static_cast<int>(q - n * P);
      const T wj = w[j];
      num = dfma(fr[(b * N + n) * L + j], wj, num);
      den = dfma(wj, wj, den);
    }
    out[i] = num / (den + static_cast<T>(1e-16));
  }
}

template <typename T>
struct IstftArgs {
  const cx_t<T>* Y;   // [batch, N, K]
  const T* w;         // [L]
  T* out;             // [batch, T_out]
  const cx_t<T>* tw;  // [n]
  int64_t batch, N, T_out;
  int n, L, P, s, pow2;
  int tile;           // output samples per CTA tile
  int max_frames;     // frames that can overlap one tile
  int64_t tiles_per_utt;
};

template <typename T>
__global__ void __launch_bounds__(256) istft_kernel(const IstftArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using C = cx_t<T>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int Nc = A.n / 2, K = Nc + 1;
  C* tw = reinterpret_cast<C*>(smem_raw);
  C* bufs = tw + A.n;                                            // [wpb][2][K]
  T* ws = reinterpret_cast<T*>(bufs + static_cast<size_t>(wpb) * 2 * K);   // [L] window
  T* w2 = ws + A.L;                                              // [L] window squared
  T* fbuf = w2 + A.L;                                            // [max_frames][L] windowed frames of the tile
  for (int i = threadIdx.x; i < A.n; i += blockDim.x) tw[i] = A.tw[i];
  for (int i = threadIdx.x; i < A.L; i += blockDim.x) {
    const T v = A.w[i];
    ws[i] = v;
    w2[i] = v * v;
  }
  __syncthreads();
  C* buf0 = bufs + static_cast<size_t>(warp) * 2 * K;
  C* buf1 = buf0 + K;
  const int64_t n_tiles = A.batch * A.tiles_per_utt;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t b = tile / A.tiles_per_utt;
    const int64_t t0 = (tile - b * A.tiles_per_utt) * A.tile;
    const int64_t t1 = (t0 + A.tile < A.T_out) ? t0 + A.tile : A.T_out;
    const int64_t q0 = t0 + A.s, q1 = t1 - 1 + A.s;             // folded positions covered by the tile
    int64_t n_lo = (q0 - A.L + 1 <= 0) ? 0 : (q0 - A.L + A.P) / A.P;
    int64_t n_hi = q1 / A.P;
    if (n_hi > A.N - 1) n_hi = A.N - 1;
    const int nf = static_cast<int>(n_hi - n_lo + 1);            // <= max_frames by construction
    for (int f = warp; f < nf; f += wpb)
      warp_irfft<T>(A.Y + ((b * A.N + n_lo + f) * K), buf0, buf1, A.n, Nc, A.pow2, tw, lane, A.L,
                    fbuf + static_cast<size_t>(f) * A.L, ws);
    __syncthreads();
    for (int64_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
      const int64_t q = t + A.s;
      int64_t a = (q - A.L + 1 <= 0) ? 0 : (q - A.L + A.P) / A.P;
      int64_t e = q / A.P;
      if (e > A.N - 1) e = A.N - 1;
      T num = 0, den = 0;
      for (int64_t n = a; n <= e; ++n) {
        const int j = static_cast<int>(q - n * A.P);
        num += fbuf[static_cast<size_t>(n - n_lo) * A.L + j];
        den += w2[j];
      }
      A.out[b * A.T_out + t] = num / (den + static_cast<T>(1e-16));
    }
    __syncthreads();
  }
}

int check_lengths(int32_t fft_length, int32_t out_length) {
  DSB_REQUIRE(fft_length > 0 && fft_length % 2 == 0, "fft_length must be positive even.");
  DSB_REQUIRE(out_length > 0 && out_length <= fft_length, "out_length must be in [1, fft_length].");
  return DSB200_OK;
}

template <typename T>
int ifftr_impl(const void* y, void* x, int64_t rows, int32_t n, int32_t out_length, int device, void* stream) {
  if (int rc = check_lengths(n, out_length)) return rc;
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(y && x, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const void* tw = twiddle_table(device, n, sizeof(T) == 8, s);
  if (tw == nullptr) return fail(DSB200_E_CUDA, "could not build the twiddle table for fft_length=%d", n);
  const size_t K = static_cast<size_t>(n) / 2 + 1;
  const size_t tw_bytes = static_cast<size_t>(n) * 2 * sizeof(T), per_warp = 2 * K * 2 * sizeof(T);
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  if (tw_bytes + per_warp > cap) return fail(DSB200_E_UNSUPPORTED, "fft_length=%d does not fit in shared memory", n);
  int wpb = static_cast<int>(std::min<size_t>(8, (cap - tw_bytes) / per_warp));
  while (wpb > 1 && tw_bytes + wpb * per_warp > 64 * 1024) --wpb;
  DSB_CUDA(cudaFuncSetAttribute(ifftr_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int64_t need = (rows + wpb - 1) / wpb;
  const int blocks = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(sm_count(device)) * 8));
  ifftr_kernel<T><<<blocks, wpb * 32, tw_bytes + wpb * per_warp, s>>>(
      static_cast<const cx_t<T>*>(y), static_cast<T*>(x), rows, n, out_length, is_pow2(n) ? 1 : 0,
      static_cast<const cx_t<T>*>(tw));
  return after_launch("ifftr_kernel");
}

int check_unframe(int64_t batch, int64_t N, int64_t T_out, int32_t L, int32_t P, int32_t center) {
  DSB_REQUIRE(L > 0, "frame_length must be positive.");
  DSB_REQUIRE(P > 0 && P <= L, "frame_period must be less than or equal to frame_length.");
  DSB_REQUIRE(batch >= 0 && N >= 1, "need at least one frame");
  const int64_t avail = (N - 1) * P + L - (center ? L / 2 : 0);
  DSB_REQUIRE(T_out >= 1 && T_out <= avail, "out_length exceeds the overlap-added span");
  return DSB200_OK;
}

template <typename T>
int unframe_impl(const void* fr, const void* w, void* out, int64_t batch, int64_t N, int64_t T_out, int32_t L,
                 int32_t P, int32_t center, int device, void* stream) {
  if (int rc = check_unframe(batch, N, T_out, L, P, center)) return rc;
  if (batch == 0) return DSB200_OK;
  DSB_REQUIRE(fr && w && out, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  const int64_t total = batch * T_out;
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(sm_count(device)) * 16));
  unframe_kernel<T><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const T*>(fr), static_cast<const T*>(w), static_cast<T*>(out), batch, N, T_out, L, P,
      center ? L / 2 : 0);
  return after_launch("unframe_kernel");
}

template <typename T>
int istft_impl(const void* Y, const void* w, void* out, int64_t batch, int64_t N, int64_t T_out, int32_t L,
               int32_t P, int32_t n, int32_t center, int device, void* stream) {
  if (int rc = check_lengths(n, L)) return rc;   // frame_length <= fft_length
  if (int rc = check_unframe(batch, N, T_out, L, P, center)) return rc;
  if (batch == 0) return DSB200_OK;
  DSB_REQUIRE(Y && w && out, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (sizeof(T) == 4 && getenv("DSB200_ISTFT_GENERIC") == nullptr) {
    const int rc = istft512_try(static_cast<const float*>(Y), static_cast<const float*>(w), static_cast<float*>(out),
                                batch, N, T_out, L, P, n, center, device, s);
    if (rc != DSB200_E_UNSUPPORTED) return rc;
  }
  const void* tw = twiddle_table(device, n, sizeof(T) == 8, s);
  if (tw == nullptr) return fail(DSB200_E_CUDA, "could not build the twiddle table for fft_length=%d", n);
  IstftArgs<T> A{};
  A.Y = static_cast<const cx_t<T>*>(Y);
  A.w = static_cast<const T*>(w);
  A.out = static_cast<T*>(out);
  A.tw = static_cast<const cx_t<T>*>(tw);
  A.batch = batch;
  A.N = N;
  A.T_out = T_out;
  A.n = n;
  A.L = L;
  A.P = P;
  A.s = center ? L / 2 : 0;
  A.pow2 = is_pow2(n) ? 1 : 0;
  const int wpb = 8;
  const size_t K = static_cast<size_t>(n) / 2 + 1;
  const size_t fixed = static_cast<size_t>(n) * 2 * sizeof(T) + wpb * 2 * K * 2 * sizeof(T) + 2 * static_cast<size_t>(L) * sizeof(T);
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  // tile = tf periods of output; the frames overlapping it number at most tf + ceil(L / P) (+1 for the offset s)
  const int halo = (L + P - 1) / P + 1;
  int tf = 32;
  auto bytes = [&](int f) { return fixed + static_cast<size_t>(f + halo) * L * sizeof(T); };
  while (tf > 1 && bytes(tf) > std::min<size_t>(cap, 160 * 1024)) tf /= 2;
  if (bytes(tf) > cap) return fail(DSB200_E_UNSUPPORTED, "frame_length=%d / fft_length=%d do not fit in shared memory", L, n);
  A.tile = tf * P;
  A.max_frames = tf + halo;
  A.tiles_per_utt = (T_out + A.tile - 1) / A.tile;
  const int64_t n_tiles = batch * A.tiles_per_utt;
  DSB_CUDA(cudaFuncSetAttribute(istft_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int blocks = static_cast<int>(std::min<int64_t>(n_tiles, static_cast<int64_t>(sm_count(device)) * 4));
  istft_kernel<T><<<blocks, wpb * 32, bytes(tf), s>>>(A);
  return after_launch("istft_kernel");
}

}  // namespace
}  // namespace dsb200

using namespace dsb200;

extern "C" {

int dsb200_ifftr_f32(const void* y, void* x, int64_t rows, int32_t fft_length, int32_t out_length, int device,
                     void* stream) {
  return ifftr_impl<float>(y, x, rows, fft_length, out_length, device, stream);
}
int dsb200_ifftr_f64(const void* y, void* x, int64_t rows, int32_t fft_length, int32_t out_length, int device,
                     void* stream) {
  return ifftr_impl<double>(y, x, rows, fft_length, out_length, device, stream);
}
int dsb200_unframe_f32(const void* frames, const void* window, void* out, int64_t batch, int64_t n_frames,
                       int64_t out_length, int32_t frame_length, int32_t frame_period, int32_t center, int device,
                       void* stream) {
  return unframe_impl<float>(frames, window, out, batch, n_frames, out_length, frame_length, frame_period, center,
                             device, stream);
}
int dsb200_unframe_f64(const void* frames, const void* window, void* out, int64_t batch, int64_t n_frames,
                       int64_t out_length, int32_t frame_length, int32_t frame_period, int32_t center, int device,
                       void* stream) {
  return unframe_impl<double>(frames, window, out, batch, n_frames, out_length, frame_length, frame_period, center,
                              device, stream);
}
int dsb200_istft_f32(const void* Y, const void* window, void* out, int64_t batch, int64_t n_frames,
                     int64_t out_length, int32_t frame_length, int32_t frame_period, int32_t fft_length,
                     int32_t center, int device, void* stream) {
  return istft_impl<float>(Y, window, out, batch, n_frames, out_length, frame_length, frame_period, fft_length,
                           center, device, stream);
}
int dsb200_istft_f64(const void* Y, const void* window, void* out, int64_t batch, int64_t n_frames,
                     int64_t out_length, int32_t frame_length, int32_t frame_period, int32_t fft_length,
                     int32_t center, int device, void* stream) {
  return istft_impl<double>(Y, window, out, batch, n_frames, out_length, frame_length, frame_period, fft_length,
                            center, device, stream);
}

}  // extern "C"
