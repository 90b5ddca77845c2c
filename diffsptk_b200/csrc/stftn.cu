// Fused Frame + Window + real FFT + spectrum formatter for fft_length 1024 and 2048 (fp32, sm_100a).
//
// Round 2: the reference's other users of the path sit at these sizes -- pitch.py:245-256 (CREPE front end: frame
// length = fft length = 1024, zmean, hanning, dB), yingram.py:97 (2048 / hop 441), pitch_spec.py:244-248 and
// ap.py:542-543 (WORLD spectra) -- and in round 1 they all fell to the general one-row-per-warp radix-4 kernel
// (spectral.cu, 0.065 of the HBM roofline at fft_length 512).  stft512.cu keeps a whole 256-point transform in the
// registers of a half-warp; at 512 / 1024 complex points that no longer fits, so this kernel keeps the transform of a
// frame PAIR in shared memory and runs it as an in-place decimation-in-frequency FFT whose passes are the same
// packed radix-16 register butterflies (fft16.cuh, float2 = (frame A, frame B)):
//
//   Nc = fft_length / 2 complex points  z[m] = x[2m] + i x[2m+1],   Nc = 16 * 16 * R3   (R3 = 2 or 4), M = Nc / 16
//   pass 1  for j < M:            16-point DFT over s of z[j + M s]                -> t,   times W_Nc^(j t)
//   pass 2  for t, j2 < R3:       16-point DFT over s2 of y1[j2 + R3 s2 + M t]     -> t2,  times W_M^(j2 t2)
//   pass 3  for t, t2:            R3-point DFT over j2 of y2[j2 + R3 t2 + M t]     -> t3
//   Z[t + 16 t2 + 256 t3] ends up at position t3 + R3 t2 + M t (digit-reversed); the real-input split reads Z[k] and
//   Z[Nc - k] through that map, applies W_n^k and formats the bins, which leave for HBM as coalesced row stores.
//
// One 16-byte pad per M elements makes every pass and the split free of shared-memory bank conflicts (pass 1 walks
// consecutive j, passes 2 / 3 and the split walk consecutive t at pitch M + 1).  tests/kernel_models.py holds the
// numpy model of this index arithmetic (stftn_frame_model).  The frames of a pair are staged with guarded loads (any
// hop, all four pad modes, utterance edges; interior pairs are prefetched into registers one iteration ahead), so the
// envelope is: float32, fft_length 1024 or 2048, frame_length <=
// fft_length; zmean, relative floor and every output format included.
#include <algorithm>

#include "fft16.cuh"

namespace dsb200 {
namespace {

using namespace fft16_detail;

struct NArgs {
  const float* x;
  const float* window;   // [L]
  const float2* tw;      // W_n^k, k < n
  float* y;
  int64_t T;
  int n_frames, pairs_per_utt;
  int64_t n_pairs;
  int L, P, left, pad_mode, zmean, fmt, has_floor;
  float eps, rel_floor;
};

template <int FMT>
__device__ __forceinline__ float fmt_real(float s) {
  if (FMT == DSB200_SPEC_DB) return 10.0f * log10f(s);
  if (FMT == DSB200_SPEC_LOGMAG) return 0.5f * logf(s);
  if (FMT == DSB200_SPEC_MAGNITUDE) return sqrtf(s);
  return s;
}

__device__ __forceinline__ C2 ld_c2(const float4* p) {
  const float4 v = *p;
  return {make_float2(v.x, v.y), make_float2(v.z, v.w)};
}
__device__ __forceinline__ void st_c2(float4* p, C2 v) { *p = make_float4(v.re.x, v.re.y, v.im.x, v.im.y); }

template <int LOGN, int W, int FMT>
__global__ void __launch_bounds__(W * 32, 1) stftn_kernel(const NArgs A) {
  constexpr int n = 1 << LOGN, Nc = n / 2, M = Nc / 16, R3 = M / 16, K = Nc + 1;
  constexpr int kPitch = M + 1;                      // elements per t-block incl. one pad
  constexpr int kWork = 16 * kPitch;                 // float4 units
  static_assert(R3 == 2 || R3 == 4, "fft_length 1024 or 2048");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  float2* tw1 = reinterpret_cast<float2*>(smem_raw);             // [15][M]  W_Nc^(j t), t = 1..15
  float2* tw2 = tw1 + 15 * M;                                    // [R3][16] W_M^(j2 t2)
  float* win = reinterpret_cast<float*>(tw2 + R3 * 16);          // [n], zero beyond L
  unsigned char* wbase = reinterpret_cast<unsigned char*>(win + n) + static_cast<size_t>(warp) * (2 * n * 4 + kWork * 16);
  float* sA = reinterpret_cast<float*>(wbase);                   // [n] samples of frame A (later: its power row)
  float* sB = sA + n;
  float4* work = reinterpret_cast<float4*>(sB + n);              // [kWork]

  for (int i = tid; i < 15 * M; i += W * 32) {
    const int t = i / M + 1, j = i - (t - 1) * M;
    tw1[i] = A.tw[(2 * j * t) & (n - 1)];
  }
  for (int i = tid; i < R3 * 16; i += W * 32) tw2[i] = A.tw[((n / M) * (i >> 4) * (i & 15)) & (n - 1)];
  for (int i = tid; i < n; i += W * 32) win[i] = i < A.L ? A.window[i] : 0.0f;
  __syncthreads();

  const int64_t stride = static_cast<int64_t>(gridDim.x) * W;
  // Software pipeline of the staging: the samples of the NEXT pair (one contiguous span of L + P floats when both of
  // its frames lie inside the utterance) are fetched into registers before the transform of the current pair and
  // written to shared memory at the top of the next iteration, so their HBM / L2 latency hides behind three FFT
  // passes (round 2: long_scoreboard was 29 % of the stall samples with loads issued right before their use).
  // Measured (config of bench.py's stft1024 / stft2048): 2048 at 6 warps and 255 registers 0.742 -> 0.687 ms; 1024 at 12
  // warps has 168 registers, the 38 span registers squeeze the butterflies and it LOSES (0.601 -> 0.644 ms): off there.
  constexpr bool kPrefetch = (LOGN == 11);
  constexpr int kSR = kPrefetch ? 80 : 1;            // span registers per lane: L + P <= 32 kSR is prefetched
  float sp[kSR];
  bool have = false;                                 // sp holds the span of this iteration's pair (warp-uniform)
  const bool can_prefetch = kPrefetch && A.L + A.P <= 32 * kSR;
  auto pair_geometry = [&](int64_t pr, int64_t& b, int& f, bool& vB, int64_t& s0, int64_t& s1) {
    b = pr / A.pairs_per_utt;
    f = 2 * static_cast<int>(pr - b * A.pairs_per_utt);
    vB = f + 1 < A.n_frames;
    s0 = static_cast<int64_t>(f) * A.P - A.left;
    s1 = vB ? s0 + A.P : s0;
  };
  for (int64_t pr = static_cast<int64_t>(blockIdx.x) * W + warp; pr < A.n_pairs; pr += stride) {
    int64_t b, s0, s1;
    int f;
    bool vB;
    pair_geometry(pr, b, f, vB, s0, s1);
    const float* xb = A.x + b * A.T;

    // ---- stage the two frames (frame.py:130-141: padding modes, utterance edges); sums for zmean ----------------
    float sumA = 0.0f, sumB = 0.0f;
    if (have) {
      // prefetched span: register i holds x[s0 + lane + 32 i]; frame A takes [0, L), frame B [P, P + L)
#pragma unroll
      for (int i = 0; i < kSR; ++i) {
        const int j = lane + 32 * i, jb = j - A.P;
        const float v = sp[i];
        if (j < A.L) { sA[j] = v; sumA += v; }
        if (jb >= 0 && jb < A.L) { sB[jb] = v; sumB += v; }
      }
      for (int j = A.L + lane; j < n; j += 32) {
        sA[j] = 0.0f;
        sB[j] = 0.0f;
      }
    } else {
      if (s0 >= 0 && s1 + A.L <= A.T) {
        // both frames inside the utterance: plain coalesced loads, no index map
        const float* pa = xb + s0;
        const float* pb = xb + s1;
#pragma unroll 8
        for (int j = lane; j < A.L; j += 32) {
          const float va = __ldg(pa + j), vb = __ldg(pb + j);
          sA[j] = va;
          sB[j] = vb;
          sumA += va;
          sumB += vb;
        }
        for (int j = A.L + lane; j < n; j += 32) {     // zero padding up to fft_length (and last pair's power rows)
          sA[j] = 0.0f;
          sB[j] = 0.0f;
        }
      } else {
        for (int j = lane; j < n; j += 32) {
          float va = 0.0f, vb = 0.0f;
          if (j < A.L) {
            const int64_t qa = pad_index(s0 + j, A.T, A.pad_mode), qb = pad_index(s1 + j, A.T, A.pad_mode);
            va = (qa < 0 || qa >= A.T) ? 0.0f : xb[qa];
            vb = (qb < 0 || qb >= A.T) ? 0.0f : xb[qb];
          }
          sA[j] = va;
          sB[j] = vb;
          sumA += va;
          sumB += vb;
        }
      }
    }
    // fetch the next pair's span now; it is consumed at the top of the next iteration
    have = false;
    if (can_prefetch && pr + stride < A.n_pairs) {
      int64_t bn, t0, t1;
      int fn;
      bool vn;
      pair_geometry(pr + stride, bn, fn, vn, t0, t1);
      if (vn && t0 >= 0 && t1 + A.L <= A.T) {
        const float* pn = A.x + bn * A.T + t0;
        const int len = A.L + A.P;
#pragma unroll
        for (int i = 0; i < kSR; ++i) {
          const int j = lane + 32 * i;
          sp[i] = j < len ? __ldg(pn + j) : 0.0f;
        }
        have = true;
      }
    }
    float2 mean2 = make_float2(0.0f, 0.0f);
    if (A.zmean) {
      const float inv_len = 1.0f / static_cast<float>(A.L);
      mean2 = make_float2(warp_sum(sumA) * inv_len, warp_sum(sumB) * inv_len);
    }
    __syncwarp();

    // ---- pass 1: lane = j (consecutive samples: conflict-free 64-bit loads), radix 16 over s ----------------------
#pragma unroll 1
    for (int j = lane; j < M; j += 32) {
      C2 a[16];
#pragma unroll
      for (int s = 0; s < 16; ++s) {
        const int m = j + M * s;
        float2 xa = *reinterpret_cast<const float2*>(sA + 2 * m);
        float2 xb2 = *reinterpret_cast<const float2*>(sB + 2 * m);
        const float2 wv = *reinterpret_cast<const float2*>(win + 2 * m);      // zero past the frame end
        xa.x -= mean2.x; xa.y -= mean2.x;
        xb2.x -= mean2.y; xb2.y -= mean2.y;
        a[s].re = make_float2(xa.x * wv.x, xb2.x * wv.x);
        a[s].im = make_float2(xa.y * wv.y, xb2.y * wv.y);
      }
      fft16<16>(a);
#pragma unroll
      for (int t = 1; t < 16; ++t) {
        const float2 w = tw1[(t - 1) * M + j];
        a[dig(t)] = cmul_s(a[dig(t)], w.x, w.y);
      }
#pragma unroll
      for (int t = 0; t < 16; ++t) st_c2(work + j + kPitch * t, a[dig(t)]);
    }
    __syncwarp();

    // ---- pass 2: lane = (t, j2), radix 16 over s2 inside block t, in place ----------------------------------------
#pragma unroll 1
    for (int j2 = lane >> 4; j2 < R3; j2 += 2) {
      float4* base = work + j2 + kPitch * (lane & 15);
      C2 a[16];
#pragma unroll
      for (int s2 = 0; s2 < 16; ++s2) a[s2] = ld_c2(base + R3 * s2);
      fft16<16>(a);
#pragma unroll
      for (int t2 = 1; t2 < 16; ++t2) {
        const float2 w = tw2[j2 * 16 + t2];
        a[dig(t2)] = cmul_s(a[dig(t2)], w.x, w.y);
      }
#pragma unroll
      for (int t2 = 0; t2 < 16; ++t2) st_c2(base + R3 * t2, a[dig(t2)]);
    }
    __syncwarp();

    // ---- pass 3: R3-point transforms of the contiguous groups (t, t2), in place -----------------------------------
#pragma unroll 2
    for (int t2 = lane >> 4; t2 < 16; t2 += 2) {
      float4* base = work + R3 * t2 + kPitch * (lane & 15);
      if (R3 == 2) {
        const C2 u = ld_c2(base), v = ld_c2(base + 1);
        st_c2(base, cadd(u, v));
        st_c2(base + 1, csub(u, v));
      } else {
        C2 x0 = ld_c2(base), x1 = ld_c2(base + 1), x2 = ld_c2(base + 2), x3 = ld_c2(base + 3);
        radix4<false>(x0, x1, x2, x3);
        st_c2(base, x0);
        st_c2(base + 1, x1);
        st_c2(base + 2, x2);
        st_c2(base + 3, x3);
      }
    }
    __syncwarp();

    // ---- real-input split, formatter, row stores ------------------------------------------------------------------
    const int64_t row = b * A.n_frames + f;
    auto zpos = [&](int k) { return (k >> 8) + R3 * ((k >> 4) & 15) + kPitch * (k & 15); };
    float2 mx = make_float2(0.0f, 0.0f);
    for (int k = lane; k <= Nc; k += 32) {
      const C2 z = ld_c2(work + zpos(k & (Nc - 1))), m = ld_c2(work + zpos((Nc - k) & (Nc - 1)));
      const float2 w = A.tw[k];                                               // W_n^k (k = Nc: -1)
      const float2 sr = add2(z.re, m.re), dr = sub2(z.re, m.re);
      const float2 si = add2(z.im, m.im), di = sub2(z.im, m.im);
      const float hx = 0.5f * w.x, hy = 0.5f * w.y;
      const float2 xr = fma2s(sr, 0.5f, fma2s(dr, hy, mul2s(si, hx)));        // X = E + W O, see stft512.cu
      const float2 xi = fma2s(di, 0.5f, fma2s(dr, -hx, mul2s(si, hy)));
      if (FMT == DSB200_SPEC_COMPLEX) {
        float2* ya = reinterpret_cast<float2*>(A.y) + row * K + k;
        ya[0] = make_float2(xr.x, xi.x);
        if (vB) ya[K] = make_float2(xr.y, xi.y);
      } else {
        const float2 s = fma2(xr, xr, fma2(xi, xi, make_float2(A.eps, A.eps)));
        if (A.has_floor) {                                                    // rows parked on chip until the maxima are known
          sA[k] = s.x;
          sB[k] = s.y;
          mx = make_float2(fmaxf(mx.x, s.x), fmaxf(mx.y, s.y));
        } else {
          float* ya = A.y + row * K + k;
          ya[0] = fmt_real<FMT>(s.x);
          if (vB) ya[K] = fmt_real<FMT>(s.y);
        }
      }
    }
    if (FMT != DSB200_SPEC_COMPLEX && A.has_floor) {                          // spec.py:174-176
      const float fa = warp_max(mx.x) * A.rel_floor, fb = warp_max(mx.y) * A.rel_floor;
      __syncwarp();
      for (int k = lane; k <= Nc; k += 32) {
        float* ya = A.y + row * K + k;
        ya[0] = fmt_real<FMT>(fmaxf(sA[k], fa));
        if (vB) ya[K] = fmt_real<FMT>(fmaxf(sB[k], fb));
      }
    }
    __syncwarp();
  }
}

template <int LOGN, int W>
int launch_n(const NArgs& A, int device, cudaStream_t stream) {
  constexpr int n = 1 << LOGN, M = n / 32, R3 = M / 16;
  const size_t smem = (15 * M + R3 * 16) * sizeof(float2) + n * sizeof(float) +
                      static_cast<size_t>(W) * (2 * n * 4 + 16 * (M + 1) * 16);
  if (smem > static_cast<size_t>(max_dynamic_smem(device))) return DSB200_E_UNSUPPORTED;
  const int blocks = static_cast<int>(std::min<int64_t>((A.n_pairs + W - 1) / W, sm_count(device)));
#define DSB_LAUNCH_N(F)                                                                                     \
  case F: {                                                                                                 \
    DSB_CUDA(cudaFuncSetAttribute(stftn_kernel<LOGN, W, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                  static_cast<int>(smem)));                                                 \
    stftn_kernel<LOGN, W, F><<<blocks, W * 32, smem, stream>>>(A);                                          \
    break;                                                                                                  \
  }
  switch (A.fmt) {
    DSB_LAUNCH_N(DSB200_SPEC_DB)
    DSB_LAUNCH_N(DSB200_SPEC_LOGMAG)
    DSB_LAUNCH_N(DSB200_SPEC_MAGNITUDE)
    DSB_LAUNCH_N(DSB200_SPEC_POWER)
    DSB_LAUNCH_N(DSB200_SPEC_COMPLEX)
    default:
      return fail(DSB200_E_BAD_PARAM, "out_format %d is not supported.", A.fmt);
  }
#undef DSB_LAUNCH_N
  return after_launch("stftn_kernel");
}

}  // namespace

// Returns DSB200_E_UNSUPPORTED outside the envelope (the caller then runs the general kernel).
int stftn_try(const float* x, const float* window, float* y, int64_t batch, int64_t T_len,
              const dsb200_stft_params* p, int device, cudaStream_t stream) {
  const dsb200_frame_params& f = p->frame;
  const dsb200_spec_params& s = p->spec;
  if ((s.fft_length != 1024 && s.fft_length != 2048) || f.frame_length > s.fft_length || T_len > (1LL << 40))
    return DSB200_E_UNSUPPORTED;
  const int64_t N = dsb200_num_frames(T_len, f.frame_period);
  if (N > (1 << 30)) return DSB200_E_UNSUPPORTED;
  const void* tw = twiddle_table(device, s.fft_length, false, stream);
  if (tw == nullptr) return fail(DSB200_E_CUDA, "could not build the twiddle table for fft_length=%d", s.fft_length);
  NArgs A{};
  A.x = x;
  A.window = window;
  A.tw = static_cast<const float2*>(tw);
  A.y = y;
  A.T = T_len;
  A.n_frames = static_cast<int>(N);
  A.pairs_per_utt = static_cast<int>((N + 1) / 2);
  A.n_pairs = batch * A.pairs_per_utt;
  A.L = f.frame_length;
  A.P = f.frame_period;
  A.left = f.center ? f.frame_length / 2 : 0;
  A.pad_mode = f.pad_mode;
  A.zmean = f.zmean;
  A.fmt = s.out_format;
  A.has_floor = s.has_relative_floor && s.out_format != DSB200_SPEC_COMPLEX;
  A.eps = static_cast<float>(s.eps);
  A.rel_floor = static_cast<float>(s.relative_floor);
  if (s.fft_length == 1024) return launch_n<10, 12>(A, device, stream);
  return launch_n<11, 6>(A, device, stream);
}

}  // namespace dsb200
