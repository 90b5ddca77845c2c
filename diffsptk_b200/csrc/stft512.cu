// Specialised fused Frame + Window + 512-point real FFT + spectrum formatter (fp32, sm_100a).
//
// This is the headline kernel of BASELINE.json (fl=400, fp=80, n_fft=512): one pass over HBM
// (each waveform sample read once, each output written once), everything else on chip.
//
// Mapping (see tests/kernel_models.py for a lane-by-lane numpy model of the same data flow):
//   * A persistent CTA walks tiles of F = 2 * (half-warps per CTA) consecutive frames of one
//     utterance.  The contiguous sample span of a tile is staged in shared memory by ONE bulk
//     async copy (cp.async.bulk -> UBLKCP, completion on an mbarrier, double buffered); tiles that
//     touch the padded ends of an utterance (or are not 16-byte aligned) are staged by guarded
//     loads that implement the reference's pad modes (frame.py:130-137).
//   * Each half-warp (16 lanes) transforms a PAIR of frames at once: every arithmetic value is a
//     float2 holding (frame A, frame B), so butterflies issue as packed FADD2 / FMUL2 / FFMA2 --
//     the FP32 pipe is the binding resource of this kernel (ubench: 128 lanes/clk/SM for scalar
//     and packed alike) and packed issue frees the slots the LDS/STS/SHFL/STG traffic needs.
//   * 512-point real FFT = 256-point complex FFT of z[m] = x[2m] + i x[2m+1], factored 16 x 16:
//     radix-16 in registers (pruned: samples >= frame_length are zero), twiddle, transpose through
//     padded shared memory, radix-16 in registers, then the real-input split.  The split needs
//     Z[k] and Z[256-k], which live in lanes l and 16-l: they swap 8 registers by shuffle.
//   * |X|^2 + eps (or another spectrum format) is formed in registers and stored straight to HBM;
//     a half-warp writes 64 contiguous bytes per instruction.
//
// Envelope: float32, fft_length == 512, frame_length <= 512, even frame_period, no zmean, no
// relative floor.  Anything else returns DSB200_E_UNSUPPORTED and the generic kernel runs.
#include <algorithm>

#include "common.cuh"

namespace dsb200 {
namespace {

constexpr int kThreads = 384;           // 12 warps = 24 half-warps -> 48 frames per tile
constexpr int kHalfWarps = kThreads / 16;
constexpr int kTileFrames = 2 * kHalfWarps;
constexpr int kXchStride = 17;           // float4 units per row of the transpose buffer (16 + 1 pad)

struct C2 {  // one complex value for each of the two frames of a pair
  float2 re, im;
};

__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 mul2s(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ float2 fma2s(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

__device__ __forceinline__ C2 cadd(C2 a, C2 b) { return {add2(a.re, b.re), add2(a.im, b.im)}; }
__device__ __forceinline__ C2 csub(C2 a, C2 b) { return {sub2(a.re, b.re), sub2(a.im, b.im)}; }
// a - i b  and  a + i b
__device__ __forceinline__ C2 csub_i(C2 a, C2 b) { return {add2(a.re, b.im), sub2(a.im, b.re)}; }
__device__ __forceinline__ C2 cadd_i(C2 a, C2 b) { return {sub2(a.re, b.im), add2(a.im, b.re)}; }
__device__ __forceinline__ C2 cmul_s(C2 a, float wr, float wi) {
  C2 r;
  r.re = fma2s(a.im, -wi, mul2s(a.re, wr));
  r.im = fma2s(a.im, wr, mul2s(a.re, wi));
  return r;
}

constexpr float kR = 0.70710678118654752440f;   // cos(pi/4)
constexpr float kC8 = 0.92387953251128675613f;  // cos(pi/8)
constexpr float kS8 = 0.38268343236508977173f;  // sin(pi/8)

// y_r = sum_s x_s W4^(s r); X3ZERO prunes the additions with a structurally-zero fourth input.
template <bool X3ZERO>
__device__ __forceinline__ void radix4(C2& x0, C2& x1, C2& x2, C2& x3) {
  const C2 t0 = cadd(x0, x2), t1 = csub(x0, x2);
  if (X3ZERO) {
    const C2 u = x1;
    x0 = cadd(t0, u);
    x2 = csub(t0, u);
    x1 = csub_i(t1, u);
    x3 = cadd_i(t1, u);
  } else {
    const C2 t2 = cadd(x1, x3), t3 = csub(x1, x3);
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1 = csub_i(t1, t3);
    x3 = cadd_i(t1, t3);
  }
}

// In-place natural-order 16-point DFT of a[0..15]; inputs a[j], j >= NJ, are structurally zero.
// j = 4 s + c, k = r + 4 t:  W16^(jk) = W4^(s r) W16^(c r) W4^(c t).
template <int NJ>
__device__ __forceinline__ void fft16(C2 (&a)[16]) {
  radix4<(12 >= NJ)>(a[0], a[4], a[8], a[12]);
  radix4<(13 >= NJ)>(a[1], a[5], a[9], a[13]);
  radix4<(14 >= NJ)>(a[2], a[6], a[10], a[14]);
  radix4<(15 >= NJ)>(a[3], a[7], a[11], a[15]);
  // a[c + 4 r] now holds b_c[r]; multiply by W16^(c r)
  a[5] = cmul_s(a[5], kC8, -kS8);                       // c=1 r=1 : W16^1
  {                                                    // c=1 r=2 : W16^2 = R (1 - i)
    const C2 v = a[9];
    a[9].re = mul2s(add2(v.re, v.im), kR);
    a[9].im = mul2s(sub2(v.im, v.re), kR);
  }
  a[13] = cmul_s(a[13], kS8, -kC8);                     // c=1 r=3 : W16^3
  {                                                    // c=2 r=1 : W16^2
    const C2 v = a[6];
    a[6].re = mul2s(add2(v.re, v.im), kR);
    a[6].im = mul2s(sub2(v.im, v.re), kR);
  }
  {                                                    // c=2 r=2 : W16^4 = -i
    const C2 v = a[10];
    a[10].re = v.im;
    a[10].im = make_float2(-v.re.x, -v.re.y);
  }
  {                                                    // c=2 r=3 : W16^6 = R (-1 - i)
    const C2 v = a[14];
    a[14].re = mul2s(sub2(v.im, v.re), kR);
    a[14].im = mul2s(add2(v.re, v.im), -kR);
  }
  a[7] = cmul_s(a[7], kS8, -kC8);                       // c=3 r=1 : W16^3
  {                                                    // c=3 r=2 : W16^6
    const C2 v = a[11];
    a[11].re = mul2s(sub2(v.im, v.re), kR);
    a[11].im = mul2s(add2(v.re, v.im), -kR);
  }
  a[15] = cmul_s(a[15], -kC8, kS8);                     // c=3 r=3 : W16^9
  // second stage over c for every r: inputs a[0+4r], a[1+4r], a[2+4r], a[3+4r] -> outputs k = r + 4 t
  radix4<false>(a[0], a[1], a[2], a[3]);      // r = 0 -> k = 0, 4, 8, 12
  radix4<false>(a[4], a[5], a[6], a[7]);      // r = 1 -> k = 1, 5, 9, 13
  radix4<false>(a[8], a[9], a[10], a[11]);    // r = 2 -> k = 2, 6, 10, 14
  radix4<false>(a[12], a[13], a[14], a[15]);  // r = 3 -> k = 3, 7, 11, 15
  // a[4 r + t] holds A[r + 4 t]: transpose the 4 x 4 register block to natural order
  C2 t;
  t = a[1]; a[1] = a[4]; a[4] = t;
  t = a[2]; a[2] = a[8]; a[8] = t;
  t = a[3]; a[3] = a[12]; a[12] = t;
  t = a[6]; a[6] = a[9]; a[9] = t;
  t = a[7]; a[7] = a[13]; a[13] = t;
  t = a[11]; a[11] = a[14]; a[14] = t;
}

struct Args {
  const float* x;
  const float* window;  // [L]
  const float* tw512;   // W512^k interleaved (re, im), 512 entries
  float* y;
  int64_t T;
  int64_t n_frames;     // frames per utterance
  int64_t n_tiles;      // batch * tiles_per_utt
  int tiles_per_utt;
  int L, P, left, pad_mode;
  int tile_floats;      // floats staged per tile (multiple of 4)
  float eps;
};

// ---- mbarrier / bulk-copy helpers (PTX) -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int FMT>
__device__ __forceinline__ float fmt1(float s) {
  if (FMT == DSB200_SPEC_DB) return 10.0f * log10f(s);
  if (FMT == DSB200_SPEC_LOGMAG) return 0.5f * logf(s);
  if (FMT == DSB200_SPEC_MAGNITUDE) return sqrtf(s);
  return s;
}

// Store one bin of both frames of the pair.
template <int FMT>
__device__ __forceinline__ void store_bin(float* yA, float* yB, bool vA, bool vB, int k, float2 re, float2 im, float eps) {
  if (FMT == DSB200_SPEC_COMPLEX) {
    if (vA) reinterpret_cast<float2*>(yA)[k] = make_float2(re.x, im.x);
    if (vB) reinterpret_cast<float2*>(yB)[k] = make_float2(re.y, im.y);
  } else {
    const float2 s = fma2(re, re, fma2(im, im, make_float2(eps, eps)));
    if (vA) yA[k] = fmt1<FMT>(s.x);
    if (vB) yB[k] = fmt1<FMT>(s.y);
  }
}

template <int NJ, bool MASK_ALL, int FMT>
__global__ void __launch_bounds__(kThreads, 1) stft512_kernel(const Args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);            // [2]
  float* win = reinterpret_cast<float*>(smem_raw + 16);              // [512], zero padded
  float* tile0 = win + 512;                                          // [2][tile_floats]
  float4* xch = reinterpret_cast<float4*>(tile0 + 2 * A.tile_floats) + (threadIdx.x >> 4) * (16 * kXchStride);

  const int tid = threadIdx.x;
  const int l = tid & 15;          // lane within the half-warp
  const int hw = tid >> 4;         // half-warp = frame pair within the tile
  const int partner = (tid & 16) | ((16 - l) & 15);  // lane (within the warp) that holds Z[256 - k]
  const unsigned hmask = 0xFFFFu << (tid & 16);      // this half-warp (the two halves may diverge on a partial tile)

  for (int i = tid; i < 512; i += kThreads) win[i] = i < A.L ? A.window[i] : 0.0f;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // per-lane twiddles: W256^(l k2) for the inter-pass rotation, W512^(16 k1 + l) / 2 for the split
  const float2* tw = reinterpret_cast<const float2*>(A.tw512);
  float twr[16], twi[16], hwr[8], hwi[8];
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) {
    const float2 v = tw[2 * l * k2];
    twr[k2] = v.x;
    twi[k2] = v.y;
  }
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    const float2 v = tw[16 * k1 + l];
    hwr[k1] = 0.5f * v.x;
    hwi[k1] = 0.5f * v.y;
  }
  __syncthreads();

  // ---- tile staging ---------------------------------------------------------------------
  const bool x_aligned = (reinterpret_cast<uintptr_t>(A.x) & 15) == 0;
  auto tile_geom = [&](int64_t t, int64_t& b, int& n0, int64_t& s0, bool& bulk) {
    b = t / A.tiles_per_utt;
    n0 = static_cast<int>(t - b * A.tiles_per_utt) * kTileFrames;
    s0 = static_cast<int64_t>(n0) * A.P - A.left;
    bulk = x_aligned && s0 >= 0 && s0 + A.tile_floats <= A.T && (((b * A.T + s0) & 3) == 0);
  };
  auto stage = [&](int64_t t, int buf) {  // called by all threads
    int64_t b, s0;
    int n0;
    bool bulk;
    tile_geom(t, b, n0, s0, bulk);
    float* dst = tile0 + buf * A.tile_floats;
    const float* xb = A.x + b * A.T;
    if (bulk) {
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&mbar[buf], static_cast<uint32_t>(A.tile_floats) * 4u);
        bulk_g2s(dst, xb + s0, static_cast<uint32_t>(A.tile_floats) * 4u, &mbar[buf]);
      }
    } else {
      for (int i = tid; i < A.tile_floats; i += kThreads) {
        const int64_t q = pad_index(s0 + i, A.T, A.pad_mode);
        dst[i] = q < 0 ? 0.0f : xb[q];
      }
    }
    return bulk;
  };

  uint32_t phase[2] = {0u, 0u};
  int64_t t = blockIdx.x;
  bool cur_bulk = false;
  if (t < A.n_tiles) cur_bulk = stage(t, 0);
  __syncthreads();

  for (int it = 0; t < A.n_tiles; ++it, t += gridDim.x) {
    const int buf = it & 1;
    const int64_t tn = t + gridDim.x;
    bool next_bulk = false;
    if (tn < A.n_tiles) next_bulk = stage(tn, buf ^ 1);
    if (cur_bulk) {
      mbar_wait(&mbar[buf], phase[buf]);
      phase[buf] ^= 1u;
    }
    int64_t b, s0;
    int n0;
    bool dummy;
    tile_geom(t, b, n0, s0, dummy);
    const float* tile = tile0 + buf * A.tile_floats;

    const int fA = n0 + 2 * hw;
    if (fA < A.n_frames) {  // half-warp uniform
      const bool vA = true, vB = (fA + 1) < A.n_frames;
      const float* pa = tile + (2 * hw) * A.P + 2 * l;
      const float* pb = pa + A.P;

      C2 a[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < NJ) {
          float2 xa = *reinterpret_cast<const float2*>(pa + 32 * j);
          float2 xb2 = *reinterpret_cast<const float2*>(pb + 32 * j);
          const float2 wv = *reinterpret_cast<const float2*>(win + 2 * l + 32 * j);
          if (MASK_ALL || j == NJ - 1) {  // never let samples past the frame end in (0 * inf = nan)
            const int p0 = 2 * l + 32 * j;
            if (p0 >= A.L) { xa.x = 0.0f; xb2.x = 0.0f; }
            if (p0 + 1 >= A.L) { xa.y = 0.0f; xb2.y = 0.0f; }
          }
          a[j].re = make_float2(xa.x * wv.x, xb2.x * wv.x);
          a[j].im = make_float2(xa.y * wv.y, xb2.y * wv.y);
        } else {
          a[j].re = make_float2(0.0f, 0.0f);
          a[j].im = make_float2(0.0f, 0.0f);
        }
      }

      fft16<NJ>(a);  // over j -> k2
#pragma unroll
      for (int k2 = 1; k2 < 16; ++k2) a[k2] = cmul_s(a[k2], twr[k2], twi[k2]);

      // transpose through shared memory: lane m1 writes column m1, lane k2 reads row k2
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2)
        xch[k2 * kXchStride + l] = make_float4(a[k2].re.x, a[k2].re.y, a[k2].im.x, a[k2].im.y);
      __syncwarp(hmask);
#pragma unroll
      for (int m1 = 0; m1 < 16; ++m1) {
        const float4 v = xch[l * kXchStride + m1];
        a[m1].re = make_float2(v.x, v.y);
        a[m1].im = make_float2(v.z, v.w);
      }
      __syncwarp(hmask);

      fft16<16>(a);  // over m1 -> k1 : a[k1] = Z[16 k1 + l]

      // swap the upper registers with the lane that holds the mirrored bins
      C2 r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        C2 s = a[8 + j];
        if (l == 0) s = (j < 7) ? a[9 + j] : a[0];  // lane 0 pairs k1 with 16 - k1 (and bin 0 with itself)
        r[j].re.x = __shfl_sync(hmask, s.re.x, partner);
        r[j].re.y = __shfl_sync(hmask, s.re.y, partner);
        r[j].im.x = __shfl_sync(hmask, s.im.x, partner);
        r[j].im.y = __shfl_sync(hmask, s.im.y, partner);
      }

      const int64_t rowA = b * A.n_frames + fA;
      const int64_t stride = (FMT == DSB200_SPEC_COMPLEX) ? 514 : 257;
      float* yA = A.y + rowA * stride;
      float* yB = yA + stride;
#pragma unroll
      for (int k1 = 0; k1 < 8; ++k1) {
        // a = Z[k], m = Z[256 - k], k = 16 k1 + l
        const C2 z = a[k1], m = r[7 - k1];
        const float2 sr = add2(z.re, m.re), dr = sub2(z.re, m.re);
        const float2 si = add2(z.im, m.im), di = sub2(z.im, m.im);
        // T = (W/2) * (si, -dr)
        const float2 tr = fma2s(dr, hwi[k1], mul2s(si, hwr[k1]));
        const float2 ti = fma2s(dr, -hwr[k1], mul2s(si, hwi[k1]));
        const float2 xr = fma2s(sr, 0.5f, tr), xi = fma2s(di, 0.5f, ti);  // X[k]      = E + T
        const float2 mr = fma2s(sr, 0.5f, make_float2(-tr.x, -tr.y));      // X[256-k]  = conj(E - T)
        const float2 mi = fma2s(di, -0.5f, ti);
        const int k = 16 * k1 + l;
        store_bin<FMT>(yA, yB, vA, vB, k, xr, xi, A.eps);
        store_bin<FMT>(yA, yB, vA, vB, 256 - k, mr, mi, A.eps);
      }
      if (l == 0) {  // bin 128 pairs with itself: X[128] = conj(Z[128])
        store_bin<FMT>(yA, yB, vA, vB, 128, a[8].re, make_float2(-a[8].im.x, -a[8].im.y), A.eps);
      }
    }
    __syncthreads();  // tile buffer `buf` may be overwritten by the staging of iteration it + 1
    cur_bulk = next_bulk;
  }
}

template <int NJ, bool MASK_ALL>
int launch_fmt(const Args& A, int fmt, int blocks, size_t smem, cudaStream_t stream) {
#define DSB_LAUNCH(F)                                                                                      \
  case F: {                                                                                                \
    DSB_CUDA(cudaFuncSetAttribute(stft512_kernel<NJ, MASK_ALL, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  static_cast<int>(smem)));                                                \
    stft512_kernel<NJ, MASK_ALL, F><<<blocks, kThreads, smem, stream>>>(A);                                \
    break;                                                                                                 \
  }
  switch (fmt) {
    DSB_LAUNCH(DSB200_SPEC_DB)
    DSB_LAUNCH(DSB200_SPEC_LOGMAG)
    DSB_LAUNCH(DSB200_SPEC_MAGNITUDE)
    DSB_LAUNCH(DSB200_SPEC_POWER)
    DSB_LAUNCH(DSB200_SPEC_COMPLEX)
    default:
      return fail(DSB200_E_BAD_PARAM, "out_format %d is not supported.", fmt);
  }
#undef DSB_LAUNCH
  return after_launch("stft512_kernel");
}

}  // namespace

int stft512_try(const float* x, const float* window, float* y, int64_t batch, int64_t T_len,
                const dsb200_stft_params* p, int device, cudaStream_t stream) {
  const dsb200_frame_params& f = p->frame;
  const dsb200_spec_params& s = p->spec;
  if (s.fft_length != 512 || f.frame_length > 512 || (f.frame_period & 1) || f.zmean || s.has_relative_floor)
    return DSB200_E_UNSUPPORTED;
  const int64_t N = dsb200_num_frames(T_len, f.frame_period);
  const int nj_exact = (f.frame_length + 31) / 32;
  const bool fast13 = nj_exact == 13;
  const int NJ = fast13 ? 13 : 16;
  // floats staged per tile: last frame starts at (F-1) P and the loader touches 32 NJ samples of it
  int64_t tile_floats = static_cast<int64_t>(kTileFrames - 1) * f.frame_period + 32 * NJ;
  tile_floats = (tile_floats + 3) & ~static_cast<int64_t>(3);
  const size_t smem = 16 + 512 * sizeof(float) + 2 * static_cast<size_t>(tile_floats) * sizeof(float) +
                      static_cast<size_t>(kHalfWarps) * 16 * kXchStride * sizeof(float4);
  if (smem > static_cast<size_t>(max_dynamic_smem(device))) return DSB200_E_UNSUPPORTED;
  const void* tw = twiddle_table(device, 512, false, stream);
  if (tw == nullptr) return fail(DSB200_E_CUDA, "could not build the twiddle table for fft_length=512");

  Args A{};
  A.x = x;
  A.window = window;
  A.tw512 = static_cast<const float*>(tw);
  A.y = y;
  A.T = T_len;
  A.n_frames = N;
  A.tiles_per_utt = static_cast<int>((N + kTileFrames - 1) / kTileFrames);
  A.n_tiles = batch * A.tiles_per_utt;
  A.L = f.frame_length;
  A.P = f.frame_period;
  A.left = f.center ? f.frame_length / 2 : 0;
  A.pad_mode = f.pad_mode;
  A.tile_floats = static_cast<int>(tile_floats);
  A.eps = static_cast<float>(s.eps);
  const int blocks = static_cast<int>(std::min<int64_t>(A.n_tiles, sm_count(device)));
  if (fast13) return launch_fmt<13, false>(A, s.out_format, blocks, smem, stream);
  return launch_fmt<16, true>(A, s.out_format, blocks, smem, stream);
}

}  // namespace dsb200
