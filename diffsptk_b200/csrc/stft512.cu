// Specialised fused Frame + Window + 512-point real FFT + spectrum formatter (fp32, sm_100a).
//
// This is the headline kernel of BASELINE.json (fl=400, fp=80, n_fft=512): one pass over HBM
// (each waveform sample read once, each output written once), everything else on chip.
//
// Mapping (tests/kernel_models.py holds a lane-by-lane numpy model of the same data flow):
//   * Persistent CTAs, one per SM.  Every WARP runs its own software pipeline over "quads" of four
//     consecutive frames of one utterance -- there is no CTA-wide barrier in the main loop:
//       - the contiguous sample span of a quad is staged in shared memory by ONE bulk async copy
//         (cp.async.bulk -> UBLKCP) completing on a warp-private mbarrier, double buffered; the
//         part of the span that falls outside the utterance is zero-filled (constant padding,
//         frame.py:134); other pad modes / unaligned waveforms use guarded loads for those quads;
//       - the four finished output rows (4 x 257 floats, contiguous in HBM) are staged in shared
//         memory and leave with ONE bulk async store (UBLKCP) per quad.
//   * Each half-warp (16 lanes) transforms a PAIR of frames at once: every arithmetic value is a
//     float2 = (frame A, frame B), so butterflies issue as packed FADD2 / FMUL2 / FFMA2 with
//     scalar-broadcast twiddles.  The FP32 pipe is the binding resource (tools/ubench: 128 results
//     per clock per SM, scalar or packed); packed issue frees the slots LDS/STS/SHFL need.
//   * 512-point real FFT = 256-point complex FFT of z[m] = x[2m] + i x[2m+1], factored 16 x 16:
//     radix-16 in registers (pruned: samples >= frame_length are structural zeros), W256 twiddles,
//     16 x 16 transpose through padded shared memory (conflict-free 64-bit accesses), radix-16,
//     then the real-input split.  The split needs Z[k] and Z[256-k], which live in lanes l and
//     16-l: they swap 8 registers by shuffle (lane 0 pairs bins with itself).
//
// Envelope: float32, fft_length == 512, frame_length <= 512, even frame_period; zmean and the relative floor are
// template variants since round 2 (bulk copies additionally need 16-byte aligned spans, else guarded loads are
// used).  Anything else returns DSB200_E_UNSUPPORTED and the generic kernel (spectral.cu) runs.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "bulk.cuh"
#include "fft16.cuh"
#include "mfcc_plan.h"

namespace dsb200 {
namespace {

using namespace fft16_detail;

// Warps per CTA (template parameter W of the kernel).  The register file is split per scheduler, so
// 12 warps allow 168 registers/thread and 16 warps 128.  Measured on B200 (BASELINE config 2): the
// spectrum kernel is latency-bound and gains 6 % from 16 warps; the MFCC epilogue needs the shared memory
// of four warps for its tables and runs with 12.
constexpr int kWarpsSpectrum = 16;
constexpr int kWarpsMfcc = 12;
constexpr int kOutFloats = 4 * 257;                    // one quad of real-valued output rows
constexpr int kDefaultBulkStore = 1;                   // see stft512_try (DSB200_STFT_STORE)
constexpr int kDefaultWarpsV7 = 20;                    // see stft512_try (DSB200_STFT_W)
constexpr int kDefaultVariant = 1;                     // see stft512_try (DSB200_STFT_V)

constexpr int kMaxDst = 8;     // destinations of one feature row (ranks of the fused all-gather)

struct Args {
  const float* x;
  const float* window;  // [L]
  const float* tw512;   // W512^k interleaved (re, im), 512 entries
  float* y;
  int T;                // samples per utterance
  int n_frames;         // frames per utterance
  int quads_per_utt;    // ceil(n_frames / 4)
  int n_quads;          // batch * quads_per_utt
  int L, P, left, pad_mode;
  int span;             // floats staged per quad: 3 P + L, rounded up to 4
  int in_floats;        // floats per input buffer (>= span, multiple of 4)
  int bulk_in;          // waveform layout allows bulk copies (alignment)
  int bulk_out;         // output layout allows bulk stores
  float eps, rel_floor;  // rel_floor: linear relative floor (RF variants)
  // MFCC epilogue (FMT == kFmtMfcc): fbank.py:315-320, dct.py:135-137, mfcc.py:252-256
  const float* mf_H;       // [257, C] filter bank
  const int32_t* mf_cb;    // [C] first non-zero row of each filter
  const int32_t* mf_ce;    // [C] one past the last non-zero row
  const float* mf_W;       // [C, C] DCT-II basis
  const float* mf_lifter;  // [M + 1]
  int mf_C, mf_M, mf_format, mf_D;
  float mf_floor, mf_gamma;
  const int32_t* mf_plan;  // device copy of the host-built segment plan (mfcc_plan.cu), or nullptr
  // Destinations of the feature rows: y_dst[0 .. n_dst) all receive every row of this launch at row offset
  // mf_row_off.  One entry (the caller's tensor) normally; with the all-gather fused into the epilogue, the same
  // tensor in every rank's memory (NVLink peer mappings) or its NVSwitch multicast address.
  float* y_dst[kMaxDst];
  int n_dst, mf_vec;       // mf_vec: every destination is 16-byte aligned (128-bit stores of whole quads)
  int mf_run;              // consecutive quads a warp takes before it stores their rows in one burst (1 or 4)
  int mf_tail;             // floats per warp BEHIND the exchange planes: the run's parked feature rows (the planes
                           // are overwritten by the next quad's transposes, the rows must survive them)
  int64_t mf_row_off;
  int stagger;             // start-up offset between the warps of one scheduler, cycles (stagger_start)
};

constexpr int kStftStagger = 0;  // defaults of the knobs STFT_STAGGER / MFCC_STAGGER
constexpr int kMfccStagger = 0;
constexpr int kInPad = 16;       // default of the knob STFT_INPAD (floats, multiple of 4)

constexpr int kFmtMfcc = 5;  // internal: stage amplitudes, then filter bank + DCT + lifter on chip
constexpr int kSegLen = kPlanSegLen;     // bins per filter-bank segment
constexpr int kMaxSeg = kPlanMaxSeg;     // segment slots of the fast filter-bank path (40 mel filters at 512 bins: ~85)
constexpr int kAmpPitch = kPlanAmpPitch; // float2 units per staged amplitude row: 257 bins + zero padding for segment tails
constexpr int kSlotsPerCh = kPlanSlotsPerCh;

// shared-memory tables of the MFCC epilogue: W^T [C][M+1] | lifter [M+1] | seg_start [kMaxSeg] | channel_slots
// [C][kSlotsPerCh] (bytes) | segment weights [kSegLen][kMaxSeg] | info [4]
__host__ __device__ constexpr int mf_slot_words(int C) { return (C * kSlotsPerCh + 3) / 4; }
__host__ __device__ constexpr int mf_table_floats(int C, int M) {
  return (C * (M + 1) + (M + 1) + kMaxSeg + mf_slot_words(C) + kSegLen * kMaxSeg + 4 + 3) & ~3;
}
// per-warp scratch of the epilogue inside the exchange planes, in floats: two amplitude rows, segment sums (+ one
// zero entry), mel rows
__host__ __device__ constexpr int mf_warp_floats(int C) {
  return 4 * kAmpPitch + 4 * (kMaxSeg + 1) + 4 * C;
}

template <int FMT>
__device__ __forceinline__ float fmt1(float s) {
  if (FMT == DSB200_SPEC_DB) return 10.0f * log10f(s);
  if (FMT == DSB200_SPEC_LOGMAG) return 0.5f * logf(s);
  if (FMT == DSB200_SPEC_MAGNITUDE) return sqrtf(s);
  if (FMT == kFmtMfcc) {  // amplitude fed to the filter bank: one MUFU op, 2^-23 relative error
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    return r;
  }
  return s;
}

// ln(x) for normal positive x: one MUFU.LG2 (no denormal pre-scaling, unlike __logf without -ftz).
__device__ __forceinline__ float fast_ln(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 0.6931471805599453f;
}

// One bin of both frames of the pair: to the shared staging rows (real formats, STAGED) or to HBM.
template <int FMT, bool STAGED>
__device__ __forceinline__ void put_bin(float* rowA, float* rowB, bool vB, int k, float2 re, float2 im, float eps) {
  if (FMT == DSB200_SPEC_COMPLEX) {
    reinterpret_cast<float2*>(rowA)[k] = make_float2(re.x, im.x);
    if (vB) reinterpret_cast<float2*>(rowB)[k] = make_float2(re.y, im.y);
  } else if (FMT == kFmtMfcc) {  // rowA = this pair's float2 row: (amplitude of frame A, of frame B)
    const float2 s = fma2(re, re, fma2(im, im, make_float2(eps, eps)));
    reinterpret_cast<float2*>(rowA)[k] = make_float2(fmt1<FMT>(s.x), fmt1<FMT>(s.y));
  } else {
    const float2 s = fma2(re, re, fma2(im, im, make_float2(eps, eps)));
    rowA[k] = fmt1<FMT>(s.x);
    if (STAGED || vB) rowB[k] = fmt1<FMT>(s.y);
  }
}

// V (variant bits): 1 = a half-warp pairs frames (f, f + 2) instead of (f, f + 1): with 2 P = 32 kShift floats the
// second frame's column j is the first frame's column j + kShift of the SAME lane, so a lane reads NJ + kShift
// instead of 2 NJ sample pairs from shared memory (frame_period 80 only): -8 LDS.64 and -16 shared-memory
// wavefronts per quad, 0.1946 -> 0.1926 ms at BASELINE config 2 (profiles/r1_stft512_v4_sweep.json).
// Tried on the same sweep and dropped: staging spans / output rows at the 128-byte phase of their global address
// (the bank conflicts ncu attributes to the bulk copies do not come from misalignment: no change), dropping the
// proxy fence in front of the span copies (no change), and moving (re A, re B, im A, im B) through the transposes
// as one 128-bit access (32 fewer LDS/STS but 60 more MOVs to build aligned register quads).
constexpr int kVPair2 = 1;
// 2 = ONE staging buffer per warp: the next span's bulk copy is issued as soon as the current samples sit in
// registers and lands while the butterflies run (saves the second 2.5 KB buffer per warp); 4 = the inter-pass
// twiddles W256^(l k2) come from a shared-memory table instead of 30 registers per thread.  Both together make
// room -- 96 registers, 11 KB of shared memory per warp -- for 20 warps per SM instead of 16.  Measured
// (profiles/r1_stft512_v5_sweep.json): 0.1884 ms with 20 warps against 0.1894 ms for variant 1 with 16 -- the
// kernel does not respond to occupancy (nor to a start stagger of the warps that share a scheduler, tried and
// removed), so variant 1 stays the default and this one is kept as an A/B knob (DSB200_STFT_V=7).
constexpr int kVSingleBuf = 2, kVTwSmem = 4;
// 8 = the window slice and the split twiddles of a lane live in registers (26 + 16) instead of being re-read from
// shared memory for every quad: -21 LDS.64 = -42 shared-memory wavefronts per quad (of ~300); needs the 168-register
// budget of a 12-warp CTA (round 2 experiment, DSB200_STFT_V=9).
constexpr int kVRegTables = 8;
// 16 = zmean (frame.py:139-140: every frame minus its own mean over the frame_length samples, before the window):
// the lanes of a half-warp sum their samples of both frames, four shuffles spread the sums, one subtraction per
// sample.  32 = relative floor (spec.py:174-176: s = max(s, amax(s) * floor) per frame, before the formatter): the
// power row is staged unformatted, every half-warp takes the maxima of its two rows (17 values per lane, four
// shuffles) and formats the row in place.  Both were outside the fast envelope in round 1 (pitch.py:245-256 and
// the WORLD spectra use them).
constexpr int kVZmean = 16, kVRelFloor = 32;
constexpr int kShift = 5;   // 2 * 80 / 32

template <int NJ, bool MASK_ALL, int FMT, int W, int V = 0>
__global__ void __launch_bounds__(W * 32, 1) stft512_kernel(const Args A) {
  constexpr int kWarps = W, kThreads = W * 32;
  constexpr bool PAIR2 = (V & kVPair2) != 0, SB = (V & kVSingleBuf) != 0, TWS = (V & kVTwSmem) != 0;
  constexpr bool RT = (V & kVRegTables) != 0;
  constexpr bool ZM = (V & kVZmean) != 0, RF = (V & kVRelFloor) != 0;
  constexpr int kBufs = SB ? 1 : 2;
  constexpr int kFB = PAIR2 ? 2 : 1;        // frame B = frame A + kFB
  constexpr int kRowB = kFB * 257;          // its staged row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int l = lane & 15;   // lane within the half-warp
  const int h = lane >> 4;   // half-warp within the warp = frame pair within the quad
  const int partner = (lane & 16) | ((16 - l) & 15);  // lane that holds Z[256 - k]

  // shared-memory carve-up: [mbar 2/warp][window 512][per warp: in0 | in1 | exchange/out-staging]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw) + 2 * warp;
  float* win = reinterpret_cast<float*>(smem_raw + 16 * kWarps);
  float2* htw = reinterpret_cast<float2*>(win + 512);  // [128]
  // MFCC tables (FMT == kFmtMfcc only), see mf_table_floats()
  float2* tw16 = htw + 128;                            // [16 k2][16 l] (TWS only)
  float* mfW = reinterpret_cast<float*>(tw16 + (TWS ? 256 : 0));
  const int mf_floats = (FMT == kFmtMfcc) ? mf_table_floats(A.mf_C, A.mf_M) : 0;
  unsigned char* wbase = smem_raw + 16 * kWarps + 512 * sizeof(float) + (128 + (TWS ? 256 : 0)) * sizeof(float2) +
                         static_cast<size_t>(mf_floats) * 4 +
                         static_cast<size_t>(warp) * (kBufs * static_cast<size_t>(A.in_floats) * 4 + kXchBytesPerWarp +
                                                      static_cast<size_t>(FMT == kFmtMfcc ? A.mf_tail : 0) * 4);
  float* in0 = reinterpret_cast<float*>(wbase);
  float2* xch = reinterpret_cast<float2*>(wbase + kBufs * static_cast<size_t>(A.in_floats) * 4);
  float* ostage = reinterpret_cast<float*>(xch);            // aliases the exchange planes (see loop)
  float2* xr = xch + h * (2 * kPlane);
  float2* xi = xr + kPlane;

  for (int i = tid; i < 512; i += kThreads) win[i] = i < A.L ? A.window[i] : 0.0f;
  if (lane == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  // per-lane twiddles: W256^(l k2) for the inter-pass rotation, W512^(16 k1 + l) / 2 for the split
  const float2* tw = reinterpret_cast<const float2*>(A.tw512);
  float twr[TWS ? 1 : 16], twi[TWS ? 1 : 16];
  if (TWS) {
    for (int i = tid; i < 256; i += kThreads) tw16[i] = tw[2 * (i & 15) * (i >> 4)];
  } else {
#pragma unroll
    for (int k2 = 1; k2 < 16; ++k2) {
      const float2 v = tw[2 * l * k2];
      twr[TWS ? 0 : k2] = v.x;
      twi[TWS ? 0 : k2] = v.y;
    }
  }
  for (int i = tid; i < 128; i += kThreads) {  // half split-twiddles W512^k / 2, k = 16 k1 + l < 128
    const float2 v = tw[i];
    htw[i] = make_float2(0.5f * v.x, 0.5f * v.y);
  }
  float* mfL = mfW + A.mf_C * (A.mf_M + 1);
  int* seg_k0 = reinterpret_cast<int*>(mfL + (A.mf_M + 1));   // [kMaxSeg] first bin a segment reads
  uint8_t* ch_slots = reinterpret_cast<uint8_t*>(seg_k0 + kMaxSeg);   // [C][kSlotsPerCh] segments of each filter (kMaxSeg = none)
  float* wT = reinterpret_cast<float*>(seg_k0 + kMaxSeg + mf_slot_words(A.mf_C));   // [kSegLen][kMaxSeg] segment weights
  int* mf_info = reinterpret_cast<int*>(wT + kSegLen * kMaxSeg);   // [0] number of slots, 0 = dense fallback
  if (FMT == kFmtMfcc) {
    const int C = A.mf_C, M1 = A.mf_M + 1;
    // scratch for the plan while the tables are built: the staging buffers of warp 0 (idle until the main loop)
    int* seg_a = reinterpret_cast<int*>(smem_raw + 16 * kWarps + 512 * sizeof(float) +
                                        (128 + (TWS ? 256 : 0)) * sizeof(float2) + static_cast<size_t>(mf_floats) * 4);
    int* seg_b = seg_a + kMaxSeg;
    int* seg_c = seg_b + kMaxSeg;
    for (int i = tid; i < C * M1; i += kThreads) mfW[i] = A.mf_W[(i / M1) * C + (i % M1)];
    for (int i = tid; i < M1; i += kThreads) mfL[i] = A.mf_lifter[i];
    if (A.mf_plan != nullptr) {
      // host-built plan (mfcc_plan.cu): starts and slots chosen so that a half-warp's 64-bit loads of the
      // amplitude rows fall into distinct banks
      const int32_t* P = A.mf_plan;
      for (int i = tid; i < kMaxSeg; i += kThreads) {
        seg_k0[i] = P[4 + i];
        seg_a[i] = P[4 + kMaxSeg + i];
        seg_b[i] = P[4 + 2 * kMaxSeg + i];
        seg_c[i] = P[4 + 3 * kMaxSeg + i];
      }
      for (int i = tid; i < C * kSlotsPerCh; i += kThreads) ch_slots[i] = static_cast<uint8_t>(P[4 + 4 * kMaxSeg + i]);
      if (tid == 0) mf_info[0] = (P[1] == C) ? P[0] : 0;
    } else if (tid == 0) {
      // no plan: cut every filter's support [cb, ce) in order into segments of <= kSegLen bins
      int n = 0;
      bool fits = true;
      for (int c = 0; c < C; ++c) {
        int t = 0;
        const int cb = A.mf_cb[c], ce = A.mf_ce[c];
        for (int k = cb; k < ce; k += kSegLen, ++t, ++n) {
          if (n < kMaxSeg && t < kSlotsPerCh) {
            seg_k0[n] = k;
            seg_a[n] = k;
            seg_b[n] = (k + kSegLen < ce) ? k + kSegLen : ce;
            seg_c[n] = c;
            ch_slots[c * kSlotsPerCh + t] = static_cast<uint8_t>(n);
          } else {
            fits = false;
          }
        }
        for (; t < kSlotsPerCh; ++t) ch_slots[c * kSlotsPerCh + t] = kMaxSeg;
      }
      const int ns = (n + 31) & ~31;
      for (int i = n; i < ns && i < kMaxSeg; ++i) { seg_k0[i] = 0; seg_c[i] = -1; }
      mf_info[0] = (fits && n > 0) ? ns : 0;
    }
    __syncthreads();
    // segment weights: H[k, c] for the bins the segment owns, zero for the bins it merely reads past
    for (int i = tid; i < kSegLen * kMaxSeg; i += kThreads) {
      const int sg = i % kMaxSeg, k = seg_k0[sg] + i / kMaxSeg, c = seg_c[sg];
      wT[i] = (sg < mf_info[0] && c >= 0 && k >= seg_a[sg] && k < seg_b[sg]) ? A.mf_H[k * C + c] : 0.0f;
    }
  }
  __syncthreads();  // last CTA-wide barrier (tables + mbarrier init); the main loop has none

  float2 wreg[RT ? NJ : 1], hreg[RT ? 8 : 1];
  if (RT) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) wreg[RT ? j : 0] = *reinterpret_cast<const float2*>(win + 2 * l + 32 * j);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) hreg[RT ? k1 : 0] = htw[16 * k1 + l];
  }
  stagger_start(A.stagger);
  const int n_warps = gridDim.x * kWarps;
  // consecutive warps take consecutive quads (L2 locality); the MFCC build takes them in RUNS of `run` quads per warp
  // so that the feature rows of a run (contiguous in the output) leave in one burst -- with the all-gather fused in,
  // NVLink sees 832-byte writes instead of 208-byte ones (round 2: at N = 8 the 208-byte multicast writes arrived at
  // ~400 GB/s per GPU and the step was bound by them)
  const int run = (FMT == kFmtMfcc) ? A.mf_run : 1;
  int q = run * (blockIdx.x * kWarps + warp);
  int b = q / A.quads_per_utt;
  int g = q - b * A.quads_per_utt;
  const int db = n_warps / A.quads_per_utt, dg = n_warps - db * A.quads_per_utt;
  int64_t run_off = 0;   // MFCC: element offset of the rows parked in the staging tile, and how many floats
  int run_n = 0;

  auto stage = [&](int bq, int gq, float* dst, uint64_t* bar) {
    stage_span_fast(A.x + static_cast<int64_t>(bq) * A.T, A.T, 4 * gq * A.P - A.left, A.span, A.pad_mode,
                    A.bulk_in != 0, dst, bar, lane);
  };

  // Every staging completes one phase of its buffer's barrier (bulk copy or plain arrive), so the parity of
  // buffer `it & 1` at iteration `it` is (it >> 1) & 1: no phase bits are carried through the loop.
  if (q < A.n_quads) stage(b, g, in0, &mbar[0]);
  bool store_pending = false;

  for (int it = 0; q < A.n_quads; ++it) {
    const int buf = it & 1;
    // next quad of this warp
    int qn, bn, gn;
    if (FMT == kFmtMfcc && run > 1) {
      const bool last_of_run = (it % run) == run - 1;
      qn = last_of_run ? q + run * n_warps - (run - 1) : q + 1;
      bn = qn / A.quads_per_utt;
      gn = qn - bn * A.quads_per_utt;
    } else {
      qn = q + n_warps;
      bn = b + db;
      gn = g + dg;
      if (gn >= A.quads_per_utt) { gn -= A.quads_per_utt; ++bn; }
    }
    if (!SB && qn < A.n_quads) stage(bn, gn, in0 + (buf ^ 1) * A.in_floats, &mbar[buf ^ 1]);
    mbar_wait(&mbar[SB ? 0 : buf], static_cast<uint32_t>(SB ? it : (it >> 1)) & 1u);
    __syncwarp();  // zero-fill / guarded stores of the other lanes
    const float* span = in0 + (SB ? 0 : buf) * A.in_floats;

    const int hf = PAIR2 ? h : 2 * h;         // first frame of this half-warp's pair within the quad
    const int fA = 4 * g + hf;
    const int rows = (A.n_frames - 4 * g) < 4 ? (A.n_frames - 4 * g) : 4;   // valid frames in the quad
    const bool vA = fA < A.n_frames, vB = (fA + kFB) < A.n_frames;
    const float* pa = span + hf * A.P + 2 * l;
    const float* pb = pa + A.P;

    C2 a[16];
    float2 raw[PAIR2 ? NJ + kShift : 1];
    if (PAIR2) {
#pragma unroll
      for (int j = 0; j < NJ + kShift; ++j) raw[PAIR2 ? j : 0] = *reinterpret_cast<const float2*>(pa + 32 * j);
    }
    float2 mean2 = make_float2(0.0f, 0.0f);   // (mean of frame A, mean of frame B), ZM only
    if (ZM) {
      float2 sum2 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float2 xa, xb2;
        if (PAIR2) {
          xa = raw[PAIR2 ? j : 0];
          xb2 = raw[PAIR2 ? j + kShift : 0];
        } else {
          xa = *reinterpret_cast<const float2*>(pa + 32 * j);
          xb2 = *reinterpret_cast<const float2*>(pb + 32 * j);
        }
        if (MASK_ALL || j == NJ - 1) {
          const int p0 = 2 * l + 32 * j;
          if (p0 >= A.L) { xa.x = 0.0f; xb2.x = 0.0f; }
          if (p0 + 1 >= A.L) { xa.y = 0.0f; xb2.y = 0.0f; }
        }
        sum2 = add2(sum2, make_float2(xa.x + xa.y, xb2.x + xb2.y));
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        sum2.x += __shfl_xor_sync(0xffffffffu, sum2.x, o);
        sum2.y += __shfl_xor_sync(0xffffffffu, sum2.y, o);
      }
      const float inv_len = 1.0f / static_cast<float>(A.L);
      mean2 = make_float2(sum2.x * inv_len, sum2.y * inv_len);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < NJ) {
        float2 xa, xb2;
        if (PAIR2) {
          xa = raw[PAIR2 ? j : 0];
          xb2 = raw[PAIR2 ? j + kShift : 0];
        } else {
          xa = *reinterpret_cast<const float2*>(pa + 32 * j);
          xb2 = *reinterpret_cast<const float2*>(pb + 32 * j);
        }
        if (ZM) {
          xa.x -= mean2.x; xa.y -= mean2.x;
          xb2.x -= mean2.y; xb2.y -= mean2.y;
        }
        const float2 wv = RT ? wreg[RT ? j : 0] : *reinterpret_cast<const float2*>(win + 2 * l + 32 * j);
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

    if (SB) {
      // every sample of this quad is in registers (the multiplies above consumed the loads): refill the buffer
      __syncwarp();
      if (qn < A.n_quads) stage(bn, gn, in0, &mbar[0]);
    }

    fft16<NJ>(a);  // over j -> k2, result in digit-swapped order
#pragma unroll
    for (int k2 = 1; k2 < 16; ++k2) {
      if (TWS) {
        const float2 t = tw16[16 * k2 + l];
        a[dig(k2)] = cmul_s(a[dig(k2)], t.x, t.y);
      } else {
        a[dig(k2)] = cmul_s(a[dig(k2)], twr[TWS ? 0 : k2], twi[TWS ? 0 : k2]);
      }
    }

    // the previous quad's bulk store reads the staging rows that alias the exchange planes
    if (store_pending) {
      if (lane == 0) bulk_wait_read();
      store_pending = false;
    }
    __syncwarp();
    // transpose through shared memory: lane m1 writes column m1, lane k2 reads row k2
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      xr[k2 * kXRow + l] = a[dig(k2)].re;
      xi[k2 * kXRow + l] = a[dig(k2)].im;
    }
    __syncwarp();
#pragma unroll
    for (int m1 = 0; m1 < 16; ++m1) {
      a[m1].re = xr[l * kXRow + m1];
      a[m1].im = xi[l * kXRow + m1];
    }
    __syncwarp();

    fft16<16>(a);  // over m1 -> k1 : a[dig(k1)] = Z[16 k1 + l]

    // swap the upper registers with the lane that holds the mirrored bins
    C2 r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      C2 s = a[dig(8 + j)];
      if (l == 0) s = (j < 7) ? a[dig(9 + j)] : a[dig(0)];  // lane 0 pairs k1 with 16 - k1 (and bin 0 with itself)
      r[j].re.x = __shfl_sync(0xffffffffu, s.re.x, partner);
      r[j].re.y = __shfl_sync(0xffffffffu, s.re.y, partner);
      r[j].im.x = __shfl_sync(0xffffffffu, s.im.x, partner);
      r[j].im.y = __shfl_sync(0xffffffffu, s.im.y, partner);
    }

    const bool bulk_ok = (FMT != DSB200_SPEC_COMPLEX) && A.bulk_out && rows == 4 &&
                         (((static_cast<int64_t>(b) * A.n_frames + 4 * g) & 3) == 0);
    // the relative floor needs whole rows on chip: those builds always stage (and copy out by hand when the bulk
    // store's alignment conditions do not hold)
    const bool staged = (FMT == kFmtMfcc) || bulk_ok || (RF && FMT != DSB200_SPEC_COMPLEX);
    const int64_t row0 = static_cast<int64_t>(b) * A.n_frames + 4 * g;
    constexpr int kStride = (FMT == DSB200_SPEC_COMPLEX) ? 514 : 257;
    float* rowA;
    float* rowB;
    if (FMT == kFmtMfcc) {
      rowA = ostage + h * (2 * kAmpPitch);   // float2 amp2[kAmpPitch] of this half-warp's frame pair
      rowB = rowA;
    } else if (staged) {
      rowA = ostage + hf * 257;
      rowB = rowA + kRowB;
    } else {
      rowA = A.y + (row0 + hf) * kStride;
      rowB = rowA + kFB * kStride;
    }

    // X[k] and X[256 - k] (k = 16 k1 + l) of both frames from Z[k] = a[dig(k1)] and Z[256 - k] = r[7 - k1]
    auto bin_pair = [&](int k1, float2& xr_, float2& xi_, float2& mr, float2& mi) {
      const C2 z = a[dig(k1)], m = r[7 - k1];
      const float2 sr = add2(z.re, m.re), dr = sub2(z.re, m.re);
      const float2 si = add2(z.im, m.im), di = sub2(z.im, m.im);
      const float2 hw = RT ? hreg[RT ? k1 : 0] : htw[16 * k1 + l];   // W512^k / 2
      const float2 tr = fma2s(dr, hw.y, mul2s(si, hw.x));          // T = (W/2) (si, -dr)
      const float2 ti = fma2s(dr, -hw.x, mul2s(si, hw.y));
      xr_ = fma2s(sr, 0.5f, tr);                                   // X[k]     = E + T
      xi_ = fma2s(di, 0.5f, ti);
      mr = fma2s(sr, 0.5f, make_float2(-tr.x, -tr.y));            // X[256-k] = conj(E - T)
      mi = fma2s(di, -0.5f, ti);
    };
    auto split = [&](auto staged_tag) {
      constexpr bool ST = decltype(staged_tag)::value;
      if constexpr (ST && FMT != DSB200_SPEC_COMPLEX && FMT != kFmtMfcc) {
        // Staged real-valued rows.  The two half-warps write rows that start 2 banks apart (514 floats), so a
        // store of "bin 16 k1 + l" from both would always collide.  Bins are therefore written two k1 at a
        // time and the upper half-warp swaps which of the two it writes first (all lanes but the two whose
        // banks wrap around): every store instruction then touches 32 distinct banks.
        // (d = bank distance between the two half-warps' rows: 514 floats = 2 banks, or 257 = 1 with PAIR2)
        constexpr int d = PAIR2 ? 1 : 2;
        const bool swF = h && (l < 16 - d), swM = h && (l >= d);
        float* f1 = rowA + l + (swF ? 16 : 0);
        float* f2 = rowA + l - (swF ? 16 : 0);
        float* m1 = rowA + 256 - l - (swM ? 16 : 0);
        float* m2 = rowA + 256 - l + (swM ? 16 : 0);
        auto fmt2 = [&](float2 re, float2 im) {
          const float2 s = fma2(re, re, fma2(im, im, make_float2(A.eps, A.eps)));
          if (RF) return s;                                   // formatted after the floor, below
          return make_float2(fmt1<FMT>(s.x), fmt1<FMT>(s.y));
        };
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float2 xr_, xi_, mr, mi;
          bin_pair(2 * u, xr_, xi_, mr, mi);
          const float2 F0 = fmt2(xr_, xi_), M0 = fmt2(mr, mi);
          bin_pair(2 * u + 1, xr_, xi_, mr, mi);
          const float2 F1 = fmt2(xr_, xi_), M1 = fmt2(mr, mi);
          float2 v = swF ? F1 : F0;
          f1[32 * u] = v.x;
          f1[32 * u + kRowB] = v.y;
          v = swF ? F0 : F1;
          f2[32 * u + 16] = v.x;
          f2[32 * u + 16 + kRowB] = v.y;
          v = swM ? M1 : M0;
          m1[-32 * u] = v.x;
          m1[-32 * u + kRowB] = v.y;
          v = swM ? M0 : M1;
          m2[-32 * u - 16] = v.x;
          m2[-32 * u - 16 + kRowB] = v.y;
        }
      } else {
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
          float2 xr_, xi_, mr, mi;
          bin_pair(k1, xr_, xi_, mr, mi);
          const int k = 16 * k1 + l;
          put_bin<FMT, ST>(rowA, rowB, vB, k, xr_, xi_, A.eps);
          put_bin<FMT, ST>(rowA, rowB, vB, 256 - k, mr, mi, A.eps);
        }
      }
      if (l == 0) {  // bin 128 pairs with itself: X[128] = conj(Z[128])
        if constexpr (ST && RF && FMT != DSB200_SPEC_COMPLEX && FMT != kFmtMfcc) {
          const C2 z8 = a[dig(8)];
          const float2 s = fma2(z8.re, z8.re, fma2(z8.im, z8.im, make_float2(A.eps, A.eps)));
          rowA[128] = s.x;
          rowB[128] = s.y;
        } else {
          put_bin<FMT, ST>(rowA, rowB, vB, 128, a[dig(8)].re, make_float2(-a[dig(8)].im.x, -a[dig(8)].im.y), A.eps);
        }
      }
      if constexpr (ST && RF && FMT != DSB200_SPEC_COMPLEX && FMT != kFmtMfcc) {
        // relative floor + formatter on the two staged rows of this half-warp
        __syncwarp();
#pragma unroll 1
        for (int rr = 0; rr < 2; ++rr) {
          float* row = rr ? rowB : rowA;
          float v[17];
          float m = 0.0f;                                      // powers are >= 0
#pragma unroll
          for (int i = 0; i < 17; ++i) {
            const int k = l + 16 * i;
            v[i] = k < 257 ? row[k] : 0.0f;
            m = fmaxf(m, v[i]);
          }
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          const float fl = m * A.rel_floor;
#pragma unroll
          for (int i = 0; i < 17; ++i) {
            const int k = l + 16 * i;
            if (k < 257) row[k] = fmt1<FMT>(fmaxf(v[i], fl));
          }
        }
      }
    };

    if (FMT == kFmtMfcc) {
      // ---- MFCC epilogue (mfcc.py:243-256 on chip).  Amplitudes of the quad are staged as two rows of
      //      float2 = (frame A, frame B); everything below is packed arithmetic on all four frames ----------
      split(std::true_type{});
      if (l < kAmpPitch - 257) reinterpret_cast<float2*>(rowA)[257 + l] = make_float2(0.0f, 0.0f);
      __syncwarp();
      // Table pointers are re-derived here from an opaque offset: hoisted out of the quad loop they would pin
      // ~12 registers across the register-tight FFT (128 per thread at 16 warps) and spill its loop state.
      uint32_t tbl_off = 16u * kWarps + 512u * sizeof(float) + (128u + (TWS ? 256u : 0u)) * sizeof(float2);
      int C = A.mf_C, M = A.mf_M, D = A.mf_D;
      asm volatile("" : "+r"(tbl_off), "+r"(C), "+r"(M), "+r"(D));
      const int M1 = M + 1;
      const float* mfW = reinterpret_cast<const float*>(smem_raw + tbl_off);
      const float* mfL = mfW + C * M1;
      const int* seg_k0 = reinterpret_cast<const int*>(mfL + M1);
      const uint8_t* ch_slots = reinterpret_cast<const uint8_t*>(seg_k0 + kMaxSeg);
      const float* wT = reinterpret_cast<const float*>(seg_k0 + kMaxSeg + mf_slot_words(C));
      const int* mf_info = reinterpret_cast<const int*>(wT + kSegLen * kMaxSeg);
      const float2* amp0 = reinterpret_cast<const float2*>(ostage);
      const float2* amp1 = amp0 + kAmpPitch;
      float4* segsum = reinterpret_cast<float4*>(ostage + 4 * kAmpPitch);   // [kMaxSeg + 1] partial sums, 4 frames
      float4* mel4 = segsum + kMaxSeg + 1;                                   // [C] log filter-bank outputs
      float* orow = ostage + kXchBytesPerWarp / 4;                           // [run][4][D] parked feature rows (behind the planes)
      auto fb_out = [&](int c, float2 u, float2 v) {                         // fbank.py:195-202
        u.x = fmaxf(u.x, A.mf_floor); u.y = fmaxf(u.y, A.mf_floor);
        v.x = fmaxf(v.x, A.mf_floor); v.y = fmaxf(v.y, A.mf_floor);
        if (A.mf_gamma == 0.0f) {
          mel4[c] = make_float4(fast_ln(u.x), fast_ln(u.y), fast_ln(v.x), fast_ln(v.y));   // inputs >= floor > 0
        } else {
          const float g = A.mf_gamma, ig = 1.0f / g;
          mel4[c] = make_float4((powf(u.x, g) - 1.0f) * ig, (powf(u.y, g) - 1.0f) * ig,
                                (powf(v.x, g) - 1.0f) * ig, (powf(v.y, g) - 1.0f) * ig);
        }
      };
      const int ns = mf_info[0];
      if (ns > 0) {
        // lane = segment of <= 8 bins of one filter; 32 segments x 4 frames per round
        if (lane == 0) segsum[kMaxSeg] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);   // "no segment"
        for (int s0 = 0; s0 < ns; s0 += 32) {
          const int sg = s0 + lane;
          const float2* p0 = amp0 + seg_k0[sg];
          const float2* p1 = amp1 + seg_k0[sg];
          float2 u = make_float2(0.0f, 0.0f), v = u;
#pragma unroll
          for (int i = 0; i < kSegLen; ++i) {
            const float w = wT[i * kMaxSeg + sg];
            u = fma2s(p0[i], w, u);
            v = fma2s(p1[i], w, v);
          }
          segsum[sg] = make_float4(u.x, u.y, v.x, v.y);
        }
        __syncwarp();
        for (int c = lane; c < C; c += 32) {
          float2 u = make_float2(0.0f, 0.0f), v = u;
          const uint8_t* sl = ch_slots + c * kSlotsPerCh;
#pragma unroll 2
          for (int t = 0; t < kSlotsPerCh && sl[t] < kMaxSeg; ++t) {
            const float4 q4 = segsum[sl[t]];
            u = add2(u, make_float2(q4.x, q4.y));
            v = add2(v, make_float2(q4.z, q4.w));
          }
          fb_out(c, u, v);
        }
      } else {
        // general filter bank (dense / learnable supports): lane = channel, weights from global memory
        for (int c = lane; c < C; c += 32) {
          float2 u = make_float2(0.0f, 0.0f), v = u;
          const int k0 = A.mf_cb != nullptr ? A.mf_cb[c] : 0, k1 = A.mf_ce != nullptr ? A.mf_ce[c] : 257;
          for (int k = k0; k < k1; ++k) {
            const float w = A.mf_H[k * C + c];
            u = fma2s(amp0[k], w, u);
            v = fma2s(amp1[k], w, v);
          }
          fb_out(c, u, v);
        }
      }
      float2 En = make_float2(0.0f, 0.0f);
      const bool want_e = (A.mf_format == DSB200_MFCC_YE) || (A.mf_format == DSB200_MFCC_YCE);
      if (want_e) {                                  // E = log((2 sum_{0<k<256} x_k + x_0 + x_256) / 512)
        const float2* amp2 = amp0 + h * kAmpPitch;
        float2 e = make_float2(0.0f, 0.0f);
        for (int k = l; k < 257; k += 16) {
          const float2 v = amp2[k];
          const float wgt = (k == 0 || k == 256) ? 1.0f : 2.0f;
          e = fma2(mul2s(v, wgt), v, e);
        }
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
          e.x += __shfl_xor_sync(0xffffffffu, e.x, o);
          e.y += __shfl_xor_sync(0xffffffffu, e.y, o);
        }
        En = make_float2(__logf(e.x * (1.0f / 512.0f)), __logf(e.y * (1.0f / 512.0f)));
      }
      __syncwarp();
      // DCT-II columns 0..M (lane l), the channel range split between the half-warps, then lifter and
      // y | yE | yc | ycE packing into the quad's staged rows; half-warp h owns frames hf, hf + kFB.
      const int64_t off = (A.mf_row_off + row0) * D;
      auto flush = [&]() {
        // the parked rows are contiguous in the output: 128-bit stores, 32 lanes x 16 bytes per instruction, to every
        // destination (the local tensor; with the all-gather fused in, every rank's copy over NVLink, or the
        // multicast address that the switch replicates)
        if (A.mf_vec && ((run_off & 3) == 0) && ((run_n & 3) == 0)) {
          for (int i = lane; i < (run_n >> 2); i += 32) {
            const float4 v4 = reinterpret_cast<const float4*>(orow)[i];
            for (int d = 0; d < A.n_dst; ++d) reinterpret_cast<float4*>(A.y_dst[d] + run_off)[i] = v4;
          }
        } else {
          for (int i = lane; i < run_n; i += 32) {
            const float v1 = orow[i];
            for (int d = 0; d < A.n_dst; ++d) A.y_dst[d][run_off + i] = v1;
          }
        }
        run_n = 0;
        __syncwarp();
      };
      if (run_n > 0 && off != run_off + run_n) flush();     // the run broke (utterance end with a partial quad)
      if (run_n == 0) run_off = off;
      float* outA = orow + run_n + hf * D;
      float* outB = outA + kFB * D;
      const int ch = (C + 1) >> 1, c0 = h * ch, c1 = (c0 + ch < C) ? c0 + ch : C;
      for (int m0 = 0; m0 < M1; m0 += 16) {
        const int m = m0 + l;
        float2 u0 = make_float2(0.0f, 0.0f), v0 = u0, u1 = u0, v1 = u0;
        if (m < M1) {
          const float* wp = mfW + c0 * M1 + m;
          const float4* mp = mel4 + c0;
          int left = c1 - c0;
#pragma unroll 2
          for (; left >= 2; left -= 2, wp += 2 * M1, mp += 2) {
            const float wa = wp[0], wb = wp[M1];
            const float4 ta = mp[0], tb = mp[1];
            u0 = fma2s(make_float2(ta.x, ta.y), wa, u0);
            v0 = fma2s(make_float2(ta.z, ta.w), wa, v0);
            u1 = fma2s(make_float2(tb.x, tb.y), wb, u1);
            v1 = fma2s(make_float2(tb.z, tb.w), wb, v1);
          }
          if (left > 0) {
            const float wa = wp[0];
            const float4 ta = mp[0];
            u0 = fma2s(make_float2(ta.x, ta.y), wa, u0);
            v0 = fma2s(make_float2(ta.z, ta.w), wa, v0);
          }
        }
        float2 u = add2(u0, u1), v = add2(v0, v1);
        u.x += __shfl_xor_sync(0xffffffffu, u.x, 16);
        u.y += __shfl_xor_sync(0xffffffffu, u.y, 16);
        v.x += __shfl_xor_sync(0xffffffffu, v.x, 16);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, 16);
        if (m < M1) {
          const float2 acc = mul2s(h ? v : u, mfL[m]);
          int pos = m - 1;
          if (m == 0) pos = (A.mf_format == DSB200_MFCC_YC || A.mf_format == DSB200_MFCC_YCE) ? M : -1;
          if (pos >= 0) {
            outA[pos] = acc.x;
            outB[pos] = acc.y;
          }
        }
      }
      if (l == 0 && want_e) {
        const int pos = (A.mf_format == DSB200_MFCC_YE) ? M : M + 1;
        outA[pos] = En.x;
        outB[pos] = En.y;
      }
      __syncwarp();
      run_n += rows * D;
      if (rows < 4 || (it % run) == run - 1 || qn >= A.n_quads) flush();
    } else if (staged) {
      split(std::true_type{});
      if (bulk_ok) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) bulk_s2g(A.y + row0 * 257, ostage, kOutFloats * 4u);
        store_pending = true;
      } else {   // RF builds only: partial quad or unaligned rows
        __syncwarp();
        for (int i = lane; i < rows * 257; i += 32) A.y[row0 * 257 + i] = ostage[i];
        __syncwarp();
      }
    } else if (vA) {
      split(std::false_type{});
    }

    q = qn; b = bn; g = gn;
  }
  if (store_pending && lane == 0) bulk_wait_read();
}

template <int NJ, bool MASK_ALL, int W, int V = 0>
int launch_fmt(const Args& A, int fmt, size_t smem, int device, cudaStream_t stream) {
  const int blocks = static_cast<int>(std::min<int64_t>((A.n_quads + W - 1) / W, sm_count(device)));
#define DSB_LAUNCH(F)                                                                                      \
  case F: {                                                                                                \
    DSB_CUDA(cudaFuncSetAttribute(stft512_kernel<NJ, MASK_ALL, F, W, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  static_cast<int>(smem)));                                                \
    stft512_kernel<NJ, MASK_ALL, F, W, V><<<blocks, W * 32, smem, stream>>>(A);                            \
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

// Shared argument set-up of the spectrum and MFCC entry points.  Returns DSB200_E_UNSUPPORTED outside
// the kernel's envelope (the callers then fall back to the generic kernels).
static int setup_args(Args& A, const float* x, const float* window, float* y, int64_t batch, int64_t T_len,
                      const dsb200_stft_params* p, int device, cudaStream_t stream, int* NJ_out) {
  const dsb200_frame_params& f = p->frame;
  const dsb200_spec_params& s = p->spec;
  const int left = f.center ? f.frame_length / 2 : 0;
  if (s.fft_length != 512 || f.frame_length > 512 || (f.frame_period & 1) || T_len > (1 << 30))
    return DSB200_E_UNSUPPORTED;
  const int64_t N = dsb200_num_frames(T_len, f.frame_period);
  const int64_t Q = (N + 3) / 4;
  if (batch * Q > (1LL << 30) || batch * N * 257 > (1LL << 40)) return DSB200_E_UNSUPPORTED;
  const int NJ = ((f.frame_length + 31) / 32 == 13) ? 13 : 16;
  *NJ_out = NJ;
  const void* tw = twiddle_table(device, 512, false, stream);
  if (tw == nullptr) return fail(DSB200_E_CUDA, "could not build the twiddle table for fft_length=512");
  A.x = x;
  A.window = window;
  A.tw512 = static_cast<const float*>(tw);
  A.y = y;
  A.T = static_cast<int>(T_len);
  A.n_frames = static_cast<int>(N);
  A.quads_per_utt = static_cast<int>(Q);
  A.n_quads = static_cast<int>(batch * Q);
  A.L = f.frame_length;
  A.P = f.frame_period;
  A.left = left;
  A.pad_mode = f.pad_mode;
  // The quad's samples: 3 P + L.  The column loads below run up to 32 NJ - L floats past that into the next
  // shared-memory region; those lanes are masked (p0 >= L), so the span is not padded to 32 NJ.
  A.span = (3 * f.frame_period + f.frame_length + 3) & ~3;
  // Frame B of a pair (two hops on) reads its masked tail columns up to 16 floats past the span: the padding keeps
  // those reads inside the warp's own buffer instead of the first floats of the other one, which the next bulk copy
  // is filling at that moment (the values are discarded either way; compute-sanitizer racecheck flags the overlap).
  A.in_floats = A.span + (knob("STFT_INPAD", kInPad) & ~3);
  // bulk copies need 16-byte aligned global addresses and sizes: every span start (4 g P - left) and
  // every utterance start (b T) must be a multiple of 4 floats.
  A.bulk_in = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (T_len % 4 == 0) && (left % 4 == 0) &&
              (f.frame_period % 4 == 0);
  A.eps = static_cast<float>(s.eps);
  A.rel_floor = s.has_relative_floor ? static_cast<float>(s.relative_floor) : 0.0f;
  A.stagger = knob("STFT_STAGGER", kStftStagger);
  return DSB200_OK;
}

// knob STFT_V: kernel variant bits, see stft512_kernel
static int stft_variant_knob() { return knob("STFT_V", kDefaultVariant); }

static size_t smem_bytes(const Args& A, int mf_floats, int kWarps, int variant = 0) {
  const size_t bufs = (variant & kVSingleBuf) ? 1 : 2, tws = (variant & kVTwSmem) ? 256 : 0;
  return 16 * kWarps + 512 * sizeof(float) + (128 + tws) * sizeof(float2) + static_cast<size_t>(mf_floats) * 4 +
         static_cast<size_t>(kWarps) * (bufs * static_cast<size_t>(A.in_floats) * 4 + kXchBytesPerWarp +
                                        static_cast<size_t>(A.mf_tail) * 4);
}

int stft512_try(const float* x, const float* window, float* y, int64_t batch, int64_t T_len,
                const dsb200_stft_params* p, int device, cudaStream_t stream) {
  Args A{};
  int NJ = 16;
  if (int rc = setup_args(A, x, window, y, batch, T_len, p, device, stream, &NJ)) return rc;
  // the BASELINE frame length gets the 16-warp build; other lengths stage longer spans and use 12 warps
  const int kWarps = (NJ == 13) ? kWarpsSpectrum : kWarpsMfcc;
  const size_t smem = smem_bytes(A, 0, kWarps);
  if (smem > static_cast<size_t>(max_dynamic_smem(device))) return DSB200_E_UNSUPPORTED;
  // Output path: rows staged in shared memory + one bulk store per quad, or plain per-lane stores.
  // DSB200_STFT_STORE=direct|bulk overrides the default (tuning knob, read once).
  static const int store_mode = [] {
    const char* e = getenv("DSB200_STFT_STORE");
    if (e != nullptr && e[0] == 'd') return 0;
    if (e != nullptr && e[0] == 'b') return 1;
    return kDefaultBulkStore;
  }();
  A.bulk_out = store_mode && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
  const bool zm = p->frame.zmean != 0;
  const bool rf = p->spec.has_relative_floor != 0 && p->spec.out_format != DSB200_SPEC_COMPLEX;
  if (zm || rf) {
    // zmean / relative-floor builds of the two main shapes: the BASELINE frame geometry (16 warps, frame pairing)
    // and everything else (12 warps, every column masked)
    const int fmt = p->spec.out_format;
    if (NJ == 13 && A.P == 80) {
      if (zm && rf) return launch_fmt<13, false, kWarpsSpectrum, kVPair2 | kVZmean | kVRelFloor>(A, fmt, smem, device, stream);
      if (zm) return launch_fmt<13, false, kWarpsSpectrum, kVPair2 | kVZmean>(A, fmt, smem, device, stream);
      return launch_fmt<13, false, kWarpsSpectrum, kVPair2 | kVRelFloor>(A, fmt, smem, device, stream);
    }
    const size_t sm12 = smem_bytes(A, 0, kWarpsMfcc);
    if (sm12 > static_cast<size_t>(max_dynamic_smem(device))) return DSB200_E_UNSUPPORTED;
    if (zm && rf) return launch_fmt<16, true, kWarpsMfcc, kVZmean | kVRelFloor>(A, fmt, sm12, device, stream);
    if (zm) return launch_fmt<16, true, kWarpsMfcc, kVZmean>(A, fmt, sm12, device, stream);
    return launch_fmt<16, true, kWarpsMfcc, kVRelFloor>(A, fmt, sm12, device, stream);
  }
  if (NJ == 13) {
    const int v = stft_variant_knob();
    if ((v & kVPair2) && A.P == 80) {
      if ((v & 6) == 6) {   // single buffer + shared-memory twiddles: 16 or 20 warps (DSB200_STFT_W)
        if (knob("STFT_W", kDefaultWarpsV7) == 20) {
          const size_t sm20 = smem_bytes(A, 0, 20, 7);
          if (sm20 <= static_cast<size_t>(max_dynamic_smem(device)))
            return launch_fmt<13, false, 20, 7>(A, p->spec.out_format, sm20, device, stream);
        }
        return launch_fmt<13, false, kWarpsSpectrum, 7>(A, p->spec.out_format, smem_bytes(A, 0, kWarpsSpectrum, 7),
                                                        device, stream);
      }
      if (v & kVRegTables)
        return launch_fmt<13, false, kWarpsMfcc, kVPair2 | kVRegTables>(A, p->spec.out_format, smem_bytes(A, 0, kWarpsMfcc),
                                                                       device, stream);
      return launch_fmt<13, false, kWarpsSpectrum, kVPair2>(A, p->spec.out_format, smem, device, stream);
    }
    return launch_fmt<13, false, kWarpsSpectrum, 0>(A, p->spec.out_format, smem, device, stream);
  }
  return launch_fmt<16, true, kWarpsMfcc>(A, p->spec.out_format, smem, device, stream);
}

// Fused waveform -> STFT power -> MFCC (stft.py:237-241 + mfcc.py:243-256) in the same kernel.  `plan` (device,
// optional) is the host-built segment plan of mfcc_plan.cu; the rows go to y_dst[0 .. n_dst) at row `row_off`.
int mfcc_wave_try(const float* x, const float* window, const float* H, const int32_t* cb, const int32_t* ce,
                  const float* W, const float* lifter, const int32_t* plan, float* const* y_dst, int n_dst,
                  int64_t row_off, int64_t batch, int64_t T_len, const dsb200_stft_params* sp,
                  const dsb200_mfcc_params* mp, int device, cudaStream_t stream) {
  if (sp->spec.out_format != DSB200_SPEC_POWER || n_dst < 1 || n_dst > kMaxDst) return DSB200_E_UNSUPPORTED;
  const int C = mp->fbank.n_channel, M = mp->mfcc_order;
  if (C > 128 || M >= C || mp->fbank.fft_length != 512) return DSB200_E_UNSUPPORTED;
  Args A{};
  int NJ = 16;
  if (int rc = setup_args(A, x, window, y_dst[0], batch, T_len, sp, device, stream, &NJ)) return rc;
  const int mf_floats = mf_table_floats(C, M);
  const int D = M + (mp->out_format == DSB200_MFCC_Y ? 0 : (mp->out_format == DSB200_MFCC_YCE ? 2 : 1));
  // amplitude rows, segment sums, mel rows and the quad's feature rows live inside the per-warp exchange region
  if (mf_warp_floats(C) * 4 > kXchBytesPerWarp) return DSB200_E_UNSUPPORTED;
  const int run_knob = knob("MFCC_RUN", 4);   // 1 | 4
  const size_t smem_max = static_cast<size_t>(max_dynamic_smem(device));
  int mf_run = run_knob == 4 ? 4 : 1;
  A.mf_tail = (mf_run * 4 * D + 3) & ~3;
  if (mf_run > 1 && smem_bytes(A, mf_floats, kWarpsMfcc) > smem_max) {   // wide rows: park one quad only
    mf_run = 1;
    A.mf_tail = (4 * D + 3) & ~3;
  }
  // 12 warps (168 registers) by default: with the planned filter bank the epilogue is short enough that the extra
  // registers beat the extra warps (round 2, 1024 x 10 s: 1.39 ms at 12 warps, 1.44 ms at 16).
  const int warps_knob = knob("MFCC_WARPS", 12);   // 12 | 16
  const bool w16 = NJ == 13 && warps_knob == 16 && smem_bytes(A, mf_floats, kWarpsSpectrum) <= smem_max;
  const int kWarps = w16 ? kWarpsSpectrum : kWarpsMfcc;
  const size_t smem = smem_bytes(A, mf_floats, kWarps);
  if (smem > smem_max) return DSB200_E_UNSUPPORTED;
  A.mf_H = H;
  A.mf_cb = cb;
  A.mf_ce = ce;
  A.mf_W = W;
  A.mf_lifter = lifter;
  A.mf_C = C;
  A.mf_M = M;
  A.mf_format = mp->out_format;
  A.mf_D = D;
  A.mf_plan = (cb != nullptr && ce != nullptr) ? plan : nullptr;
  A.n_dst = n_dst;
  A.mf_run = mf_run;
  A.mf_row_off = row_off;
  A.stagger = knob("MFCC_STAGGER", kMfccStagger);
  A.mf_vec = 1;
  for (int d = 0; d < n_dst; ++d) {
    A.y_dst[d] = y_dst[d];
    if (reinterpret_cast<uintptr_t>(y_dst[d]) & 15) A.mf_vec = 0;
  }
  A.mf_floor = static_cast<float>(mp->fbank.floor);
  A.mf_gamma = static_cast<float>(mp->fbank.gamma);
  const int64_t n_runs = (A.n_quads + mf_run - 1) / mf_run;
  const int blocks = static_cast<int>(std::min<int64_t>((n_runs + kWarps - 1) / kWarps, sm_count(device)));
  auto launch = [&](auto kern) -> int {
    DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<blocks, kWarps * 32, smem, stream>>>(A);
    return DSB200_OK;
  };
  const bool pair2 = (stft_variant_knob() & kVPair2) && A.P == 80;
  int rc;
  if (w16) rc = pair2 ? launch(stft512_kernel<13, false, kFmtMfcc, kWarpsSpectrum, kVPair2>)
                      : launch(stft512_kernel<13, false, kFmtMfcc, kWarpsSpectrum, 0>);
  else if (NJ == 13) rc = pair2 ? launch(stft512_kernel<13, false, kFmtMfcc, kWarpsMfcc, kVPair2>)
                                : launch(stft512_kernel<13, false, kFmtMfcc, kWarpsMfcc, 0>);
  else rc = launch(stft512_kernel<16, true, kFmtMfcc, kWarpsMfcc>);
  if (rc != DSB200_OK) return rc;
  return after_launch("stft512_kernel<mfcc>");
}

}  // namespace dsb200
