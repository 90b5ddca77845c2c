// Fused waveform -> LPC kernel (fp32, sm_100a): Frame -> Window -> autocorrelation -> Levinson-Durbin
// in one pass over HBM (BASELINE.json config 3: fl=400, fp=80, M=24; 320 B read + 100 B written per frame).
//
// Reference cascade being replaced: diffsptk/modules/frame.py:120-141, window.py:185-193 (out_length
// None), acorr.py:110-120, levdur.py:113-127 (README.md:198-201 pipeline).
//
// Mapping: persistent CTAs; every warp runs its own pipeline over "units" of 32 consecutive frames of
// one utterance (no CTA barrier in the loop):
//   * the unit's contiguous sample span is staged in shared memory by one bulk async copy (UBLKCP) on a
//     warp-private mbarrier; the next unit's copy is issued before the Levinson phase, which no longer
//     reads the span, so the copy overlaps it;
//   * 8 sub-steps of 4 frames: each half-warp takes a PAIR of frames (float2 = (frame A, frame B),
//     packed FFMA2); lane l owns samples [25 l, 25 l + 25) plus a 24-sample halo in registers and
//     accumulates all 25 lags (625 FFMA2, no memory traffic); the 16 partial sums per lag are reduced
//     through shared memory; the autocorrelations of the 32 frames collect in a 32 x 25 tile;
//   * Levinson-Durbin for the 32 frames runs with one frame per lane, float64 state in registers
//     (reciprocals by MUFU.RCP + two Newton steps), and the 32 x (M+1) result tile leaves with one bulk
//     async store.
// The FP32 pipe bounds this kernel (9.7 k multiply-adds per frame = 76 clk/frame/SM), not HBM.
//
// Envelope: float32, frame_length <= 400, lpc_order <= 24, even frame_period, no zmean.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "bulk.cuh"

namespace dsb200 {
namespace {

// Warps per CTA (template parameter): 8 = 2 per scheduler, 255 registers/thread (round 1: FMA pipe 47 %, every
// load / reduction / Levinson phase of one warp exposed); 12 = 3 per scheduler at 168 registers -- the 25 packed lag
// accumulators + the 25-deep window need ~100, the rest covers the loads in flight.
constexpr int kUnit = 32;     // frames per warp unit (one Levinson frame per lane)
constexpr int kHalfUnit = 16; // frames staged at a time
constexpr int kCh = 25;       // samples per lane (16 lanes x 25 = 400 >= frame_length)
constexpr int kLag = 25;      // lags 0..24
constexpr int kHalo = kLag - 1;
constexpr int kLpcVariant = 16 | 4;  // default of the knob LPC_V (kLv* bits): lag-pair form, order-24 Levinson unguarded
constexpr int kLpcWarps2 = 16;  // default of the knob LPC_W2 (warps per CTA of the lag-pair kernel; 12 when M < 24)
constexpr int kLpcStagger = 0;  // default of the knob LPC_STAGGER (cycles per scheduler slot)

struct LArgs {
  const float* x;
  const float* window;  // [L]
  float* y;             // [batch, n_frames, M + 1]
  int T, n_frames, units_per_utt, n_units;
  int L, P, left, pad_mode, M;
  int span;             // floats staged per half unit: 15 P + 400 + 24, rounded up to 4
  int bulk_in, bulk_out;
  int stagger;          // start-up offset between the warps of one scheduler, cycles (stagger_start)
  double eps;
};

// FULL: frame_length == 400 exactly -- every chunk sample is inside the frame, and only lane 15's halo
// (samples 400..423) must be forced to zero; otherwise every load is compared with the per-lane limit.
//
// V (variant bits, knob LPC_V; FULL builds only): 1 = the last lane reads its halo through the zero tail of the window
// table instead of selecting zeros (48 FSEL per lane and sub-step); 2 = the window product is one packed multiply per
// sample pair; 4 = lpc_order == 24 exactly: the Levinson recursion is unrolled without its per-order guards.
constexpr int kLvHalo = 1, kLvMul2 = 2, kLvM24 = 4, kLvRolled = 8;   // 8: the rolled Levinson recursion (see there)
// Timing diagnostics (WRONG results by construction; compiled only with -DDSB200_LPC_DIAG): 64 = no Levinson phase
// (the lane's autocorrelation row is stored instead), 128 = no cross-lane reduction of the lag sums.
constexpr int kLvNoLev = 64, kLvNoRed = 128;

template <bool FULL, int kLWarps, int V = 0>
__global__ void __launch_bounds__(kLWarps * 32, 1) lpc_wave_kernel(const LArgs A) {
  constexpr int kLThreads = kLWarps * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int l = lane & 15, h = lane >> 4;
  const int D = A.M + 1;
  const bool last_lane = (l == 15);
  const int lim = A.L - kCh * l;                  // valid samples counted from this lane's chunk start

  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw) + warp;
  float* win = reinterpret_cast<float*>(smem_raw + 8 * kLWarps);  // [448], zero padded
  const size_t per_warp = static_cast<size_t>(A.span) * 4 + 2 * 16 * kLag * 8 + kUnit * kLag * 4;
  unsigned char* wbase = reinterpret_cast<unsigned char*>(win + 448) + warp * per_warp;
  float* span = reinterpret_cast<float*>(wbase);
  float2* part = reinterpret_cast<float2*>(wbase + static_cast<size_t>(A.span) * 4) + h * (16 * kLag);  // [16][25]
  float* rbuf = reinterpret_cast<float*>(wbase + static_cast<size_t>(A.span) * 4 + 2 * 16 * kLag * 8);  // [32][25]

  for (int i = tid; i < 448; i += kLThreads) win[i] = i < A.L ? A.window[i] : 0.0f;
  if (lane == 0) {
    mbar_init(mbar, 1);
    mbar_fence_init();
  }
  __syncthreads();  // the only CTA-wide barrier

  const int n_warps = gridDim.x * kLWarps;
  int u = blockIdx.x * kLWarps + warp;
  int b = u / A.units_per_utt, g = u - b * A.units_per_utt;
  const int db = n_warps / A.units_per_utt, dg = n_warps - db * A.units_per_utt;
  uint32_t phase = 0u;
  // every staging completes one phase of the warp's barrier (bulk copy, or a plain arrive on the out-of-line
  // general path), so the wait below is unconditional
  auto stage_half = [&](int bq, int gq, int half) {
    stage_span_fast(A.x + static_cast<int64_t>(bq) * A.T, A.T, (kUnit * gq + kHalfUnit * half) * A.P - A.left,
                    A.span, A.pad_mode, A.bulk_in != 0, span, mbar, lane);
    return true;
  };
  bool cur_bulk = false;
  if (u < A.n_units) cur_bulk = stage_half(b, g, 0);
  bool store_pending = false;
  stagger_start(A.stagger);

  while (u < A.n_units) {
    const int f0 = kUnit * g;                       // first frame of the unit
    const int nvalid = (A.n_frames - f0) < kUnit ? (A.n_frames - f0) : kUnit;
    const int un = u + n_warps;
    int bn = b + db, gn = g + dg;
    if (gn >= A.units_per_utt) { gn -= A.units_per_utt; ++bn; }

#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      if (cur_bulk) {
        mbar_wait(mbar, phase);
        phase ^= 1u;
      }
      __syncwarp();
      if (half == 0 && store_pending) {             // rbuf doubles as the output staging tile
        if (lane == 0) bulk_wait_read();
        store_pending = false;
        __syncwarp();
      }
#pragma unroll 1
      for (int sub = 0; sub < kHalfUnit / 4; ++sub) {
        // frame pair of this half-warp: (fa, fa + 2).  The two half-warps start one frame apart, which
        // shifts their shared-memory banks by P mod 32 (= 16 at P = 80): 32-lane loads are conflict-free.
        const int fa = 4 * sub + h;                 // relative to the staged half unit
        const float* pa = span + fa * A.P + kCh * l;
        const float* pb = pa + 2 * A.P;
        const float* pw = win + kCh * l;
        // kLvHalo: the last lane's halo (samples 400..423) is read from win[400..423] = 0 (0 * 0, never sample * 0)
        const bool redirect = FULL && (V & kLvHalo) && last_lane;
        const float* pah = redirect ? pw : pa;
        const float* pbh = redirect ? pw : pb;
        auto ld = [&](int i) {                      // windowed sample pair i of this lane's chunk (+ halo)
          // samples past the frame end are structural zeros (selected, never multiplied: 0 * inf = nan)
          const float w = pw[i];
          float xa, xb;
          if (FULL && (V & kLvHalo)) {
            xa = (i >= kCh) ? pah[i] : pa[i];
            xb = (i >= kCh) ? pbh[i] : pb[i];
          } else {
            const bool out = FULL ? (last_lane && i >= kCh) : (i >= lim);
            xa = pa[i];
            xb = pb[i];
            if (out) { xa = 0.0f; xb = 0.0f; }
          }
          if (V & kLvMul2) return __fmul2_rn(make_float2(xa, xb), make_float2(w, w));
          return make_float2(xa * w, xb * w);
        };
        float2 x2[kCh + kHalo];
        float2 acc[kLag];
#pragma unroll
        for (int k = 0; k < kLag; ++k) acc[k] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int i = 0; i < kLag; ++i) x2[i] = ld(i);
#pragma unroll
        for (int i = 0; i < kCh; ++i) {             // x2[i .. i+24] live: a 25-deep sliding register window
          if (i + kLag < kCh + kHalo) x2[i + kLag] = ld(i + kLag);
#pragma unroll
          for (int k = 0; k < kLag; ++k) acc[k] = __ffma2_rn(x2[i], x2[i + k], acc[k]);
        }
        // reduce the 16 per-lane partial sums of every lag through shared memory
        float2 s0 = make_float2(0.0f, 0.0f), s1 = make_float2(0.0f, 0.0f);
        if (V & kLvNoRed) {
#pragma unroll
          for (int k = 0; k < kLag; ++k) s0 = __fadd2_rn(s0, acc[k]);
          s1 = s0;
        } else {
#pragma unroll
          for (int k = 0; k < kLag; ++k) part[l * kLag + k] = acc[k];
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            s0 = __fadd2_rn(s0, part[j * kLag + l]);
            if (l < kLag - 16) s1 = __fadd2_rn(s1, part[j * kLag + 16 + l]);
          }
        }
        const int fu = kHalfUnit * half + fa;       // frame within the unit
        rbuf[fu * kLag + l] = s0.x;
        rbuf[(fu + 2) * kLag + l] = s0.y;
        if (l < kLag - 16) {
          rbuf[fu * kLag + 16 + l] = s1.x;
          rbuf[(fu + 2) * kLag + 16 + l] = s1.y;
        }
        __syncwarp();
      }
      // the staged samples are dead: fetch the second half, or the next unit's first half (that copy
      // overlaps the Levinson phase below)
      if (half == 0) cur_bulk = stage_half(b, g, 1);
      else cur_bulk = (un < A.n_units) ? stage_half(bn, gn, 0) : false;
    }

    // Levinson-Durbin, one frame per lane, float64 state (see lpc.cu for the recursion)
    double a[kLag];
    if (V & kLvRolled) {
      // Rolled form: one loop body serves every order, so the phase is ~250 instructions instead of ~3 000 (the
      // unrolled recursion streams 48 KB of code per unit through the instruction cache and evicts the lag loop).
      // No register is indexed by the order: besides a[j] the lane keeps the REVERSED predictor ar[j] = a[i - j]
      // (ar[i] = a[0] = 1, zero beyond), so that
      //     acc_i   = r[i] + sum_{j<i} a[j] r[i-j] = sum_m ar[m] r[m]              (r[m]: float64 column in `part`)
      //     a'[j]   = a[j] + k ar[j]            (j = i gives a'[i] = k, positions beyond i stay 0)
      //     ar'[j]  = ar[j-1] + k a[j-1],  ar'[1] = k                               (the reversal pivot moves with i)
      // are the same statements for every i.  Orders 1..11 touch positions <= 12 only (half-width body).
      double* rd = reinterpret_cast<double*>(wbase + static_cast<size_t>(A.span) * 4) + lane;   // [24][32]
#pragma unroll
      for (int k = 1; k < kLag; ++k) rd[(k - 1) * 32] = static_cast<double>(rbuf[lane * kLag + k]);
      const double r0 = static_cast<double>(rbuf[lane * kLag]);
      double ar[kLag];   // positions 1..24 at indices 1..24 (index 0 unused)
#pragma unroll
      for (int j = 1; j < kLag; ++j) { a[j] = 0.0; ar[j] = 0.0; }
      ar[1] = 1.0;
      double E = r0 + A.eps;
      auto orders = [&](auto width, int i0, int i1) {
        constexpr int W = decltype(width)::value;
#pragma unroll 1
        for (int i = i0; i <= i1; ++i) {
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
          for (int m = 1; m <= W; m += 4) {
            s0 = fma(ar[m], rd[(m - 1) * 32], s0);
            s1 = fma(ar[m + 1], rd[m * 32], s1);
            s2 = fma(ar[m + 2], rd[(m + 1) * 32], s2);
            s3 = fma(ar[m + 3], rd[(m + 2) * 32], s3);
          }
          const double acc = (s0 + s1) + (s2 + s3);
          double inv = static_cast<double>(__frcp_rn(static_cast<float>(E)));
          inv = fma(inv, fma(-E, inv, 1.0), inv);
          inv = fma(inv, fma(-E, inv, 1.0), inv);
          const bool tame = fabs(E) > 1e-30 && fabs(E) < 1e30;
          const double kk = tame ? (-acc * inv) : (-acc / E);
#pragma unroll
          for (int j = W; j >= 2; --j) {   // descending: positions j - 1 are still the old ones when j reads them
            const double ta = fma(kk, ar[j], a[j]);
            ar[j] = fma(kk, a[j - 1], ar[j - 1]);
            a[j] = ta;
          }
          a[1] = fma(kk, ar[1], a[1]);
          ar[1] = kk;
          E *= fma(-kk, kk, 1.0);
        }
      };
      orders(std::integral_constant<int, 12>{}, 1, A.M < 11 ? A.M : 11);
      orders(std::integral_constant<int, 24>{}, 12, A.M);
      double g0 = r0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
#pragma unroll
      for (int m = 1; m < kLag; m += 4) {
        g0 = fma(a[m], rd[(m - 1) * 32], g0);
        g1 = fma(a[m + 1], rd[m * 32], g1);
        g2 = fma(a[m + 2], rd[(m + 1) * 32], g2);
        g3 = fma(a[m + 3], rd[(m + 2) * 32], g3);
      }
      a[0] = sqrt((g0 + g1) + (g2 + g3));
    } else {
    double r[kLag];
#pragma unroll
    for (int k = 0; k < kLag; ++k) r[k] = static_cast<double>(rbuf[lane * kLag + k]);
    double E = r[0] + A.eps;
    if (V & kLvNoLev) {
#pragma unroll
      for (int i = 1; i < kLag; ++i) a[i] = r[i];
    }
#pragma unroll
    for (int i = 1; i < kLag; ++i) {
      if (V & kLvNoLev) break;
      if ((V & kLvM24) || i <= A.M) {
        double acc = r[i];
#pragma unroll
        for (int j = 1; j < i; ++j) acc = fma(a[j], r[i - j], acc);
        // k = -acc / E with a Newton-refined reciprocal (exact division outside the float range)
        double inv = static_cast<double>(__frcp_rn(static_cast<float>(E)));
        inv = fma(inv, fma(-E, inv, 1.0), inv);
        inv = fma(inv, fma(-E, inv, 1.0), inv);
        const bool tame = fabs(E) > 1e-30 && fabs(E) < 1e30;
        const double kk = tame ? (-acc * inv) : (-acc / E);
#pragma unroll
        for (int j = 1; 2 * j <= i; ++j) {  // a_j <- a_j + k a_{i-j}, updated in symmetric pairs
          const double lo = a[j], hi = a[i - j];
          a[j] = fma(kk, hi, lo);
          if (2 * j != i) a[i - j] = fma(kk, lo, hi);
        }
        a[i] = kk;
        E *= fma(-kk, kk, 1.0);
      }
    }
    double gain = r[0];
#pragma unroll
    for (int j = 1; j < kLag; ++j)
      if ((V & kLvM24) || j <= A.M) gain = fma(r[j], a[j], gain);
    a[0] = sqrt(gain);
    }

    const int64_t row0 = static_cast<int64_t>(b) * A.n_frames + f0;
    const bool staged = A.bulk_out && nvalid == kUnit && (((row0 * D) & 3) == 0);
    __syncwarp();  // every lane has read its rbuf row
    if (staged) {
      float* o = rbuf + lane * D;                   // dense [32][D] tile (D <= 25 fits in the row storage)
#pragma unroll
      for (int k = 0; k < kLag; ++k)
        if ((V & kLvM24) || k < D) o[k] = static_cast<float>(a[k]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) bulk_s2g(A.y + row0 * D, rbuf, static_cast<uint32_t>(kUnit * D) * 4u);
      store_pending = true;
    } else if (lane < nvalid) {
      float* o = A.y + (row0 + lane) * D;
#pragma unroll
      for (int k = 0; k < kLag; ++k)
        if (k < D) o[k] = static_cast<float>(a[k]);
    }
    u = un; b = bn; g = gn;
  }
  if (store_pending && lane == 0) bulk_wait_read();
}


// ---- lag-pair form (frame_length = 400 exactly; knob LPC_V bit 16) -------------------------------------------
// Measured on B200 (tools/bench_ffma2.cu): a packed FFMA2 whose three operands are three different register pairs
// issues every 3 cycles per scheduler (register-file read bandwidth); with one operand a broadcast scalar (or a
// reused pair) it issues every 2.  The frame-pair form above, acc[k] += x2[i] * x2[i + k], is the 3-cycle kind.
// Here a half-warp owns ONE frame and the pair is two adjacent lags of it:
//     (r[2m], r[2m+1]) += x[i] * (x[i + 2m], x[i + 2m + 1])            -- scalar x[i] times a pair of samples
// Lane l owns samples [26 l, 26 l + 26) and reads 50 (25 aligned pairs E[j] = (x[2j], x[2j+1]), windowed by one
// FMUL2 each, 64-bit loads).  An even sample i = 2t multiplies E[t + m] into (r[2m], r[2m+1]); an odd sample multiplies
// the same kind of aligned pair, E[t + 1 + m], into a SECOND accumulator set (r[2m+1], r[2m+2]); the two sets are
// merged once per frame (lag 0 of the odd samples is one scalar FFMA each).  13 + 12 accumulator pairs, 13 x 25 =
// 325 FFMA2 per frame and lane.  Samples past the frame
// end (lane 14 from pair 18, lane 15 from pair 5) are read from the zero tail of the window table instead of the
// staged waveform: 0 * 0, never sample * 0.
constexpr int kCh2 = 26, kPairs2 = 25, kAcc2 = 13;
constexpr int kLvLagPair = 16, kLvSplit = 32;   // 32: split accumulation chains in the Levinson phase

template <int kLWarps, int V>
__global__ void __launch_bounds__(kLWarps * 32, 1) lpc_wave2_kernel(const LArgs A) {
  constexpr int kLThreads = kLWarps * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int l = lane & 15, h = lane >> 4;
  const int D = A.M + 1;

  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw) + warp;
  float* win = reinterpret_cast<float*>(smem_raw + 8 * kLWarps);  // [448], zero from 400 on
  constexpr int kPartBytes = 2 * 16 * kAcc2 * 8;
  const size_t per_warp = static_cast<size_t>(A.span) * 4 + kPartBytes + kUnit * kLag * 4;
  unsigned char* wbase = reinterpret_cast<unsigned char*>(win + 448) + warp * per_warp;
  float* span = reinterpret_cast<float*>(wbase);
  float2* part = reinterpret_cast<float2*>(wbase + static_cast<size_t>(A.span) * 4) + h * (16 * kAcc2);  // [16][13]
  float* rbuf = reinterpret_cast<float*>(wbase + static_cast<size_t>(A.span) * 4 + kPartBytes);          // [32][25]

  for (int i = tid; i < 448; i += kLThreads) win[i] = i < A.L ? A.window[i] : 0.0f;
  if (lane == 0) {
    mbar_init(mbar, 1);
    mbar_fence_init();
  }
  __syncthreads();  // the only CTA-wide barrier

  const int n_warps = gridDim.x * kLWarps;
  int u = blockIdx.x * kLWarps + warp;
  int b = u / A.units_per_utt, g = u - b * A.units_per_utt;
  const int db = n_warps / A.units_per_utt, dg = n_warps - db * A.units_per_utt;
  uint32_t phase = 0u;
  auto stage_half = [&](int bq, int gq, int half) {
    stage_span_fast(A.x + static_cast<int64_t>(bq) * A.T, A.T, (kUnit * gq + kHalfUnit * half) * A.P - A.left,
                    A.span, A.pad_mode, A.bulk_in != 0, span, mbar, lane);
  };
  bool cur = false;
  if (u < A.n_units) { stage_half(b, g, 0); cur = true; }
  bool store_pending = false;
  const float2* pw = reinterpret_cast<const float2*>(win + kCh2 * l);
  const float2* zero = reinterpret_cast<const float2*>(win + 400);

  while (u < A.n_units) {
    const int f0 = kUnit * g;
    const int nvalid = (A.n_frames - f0) < kUnit ? (A.n_frames - f0) : kUnit;
    const int un = u + n_warps;
    int bn = b + db, gn = g + dg;
    if (gn >= A.units_per_utt) { gn -= A.units_per_utt; ++bn; }

#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      if (cur) {
        mbar_wait(mbar, phase);
        phase ^= 1u;
      }
      __syncwarp();
      if (half == 0 && store_pending) {             // rbuf doubles as the output staging tile
        if (lane == 0) bulk_wait_read();
        store_pending = false;
        __syncwarp();
      }
#pragma unroll 1
      for (int sub = 0; sub < kHalfUnit / 2; ++sub) {
        const int fa = 2 * sub + h;                 // this half-warp's frame within the staged half unit
        const float2* pa = reinterpret_cast<const float2*>(span + fa * A.P + kCh2 * l);
        const float2* pb = (l == 15) ? zero - 5 : pa;     // pairs 5..17: past the frame end for lane 15
        const float2* pc = (l >= 14) ? zero - 18 : pa;    // pairs 18..25: past it for lanes 14 and 15
        auto ld = [&](int j) {
          const float2 x = j < 5 ? pa[j] : (j < 18 ? pb[j] : pc[j]);
          return __fmul2_rn(x, pw[j]);
        };
        float2 E[kPairs2], ae[kAcc2], ao[kAcc2 - 1];
        float r0o = 0.0f;                           // lag 0 of the odd samples (no aligned pair starts at lag 0 there)
#pragma unroll
        for (int m = 0; m < kAcc2; ++m) ae[m] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int m = 0; m < kAcc2 - 1; ++m) ao[m] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int j = 0; j < kAcc2; ++j) E[j] = ld(j);
#pragma unroll
        for (int t = 0; t < kAcc2; ++t) {           // samples 2t and 2t + 1 of the lane's chunk
          if (t + kAcc2 < kPairs2) E[t + kAcc2] = ld(t + kAcc2);
          const float xe = E[t].x, xo = E[t].y;
          // even sample: lags (2m, 2m+1) from the aligned pair E[t + m]; odd sample: the SAME aligned pairs give
          // lags (2m+1, 2m+2), kept in a second accumulator set -- no odd-aligned pair is ever formed
#pragma unroll
          for (int m = 0; m < kAcc2; ++m) ae[m] = __ffma2_rn(E[t + m], make_float2(xe, xe), ae[m]);
#pragma unroll
          for (int m = 0; m < kAcc2 - 1; ++m) ao[m] = __ffma2_rn(E[t + 1 + m], make_float2(xo, xo), ao[m]);
          r0o = fmaf(xo, xo, r0o);
        }
        float2 acc[kAcc2];                          // (r[2m], r[2m+1]) = (ae[m].x + ao[m-1].y, ae[m].y + ao[m].x)
#pragma unroll
        for (int m = 0; m < kAcc2; ++m) {
          acc[m].x = ae[m].x + (m > 0 ? ao[m - 1].y : r0o);
          acc[m].y = m < kAcc2 - 1 ? ae[m].y + ao[m].x : ae[m].y;
        }
        // reduce the 16 per-lane partial sums of every lag pair through shared memory
        if (V & kLvNoRed) {
          float2 s0 = make_float2(0.0f, 0.0f);
#pragma unroll
          for (int m = 0; m < kAcc2; ++m) s0 = __fadd2_rn(s0, acc[m]);
          if (l < kAcc2) rbuf[(kHalfUnit * half + fa) * kLag + 2 * l] = s0.x + s0.y;
          __syncwarp();
          continue;
        }
#pragma unroll
        for (int m = 0; m < kAcc2; ++m) part[l * kAcc2 + m] = acc[m];
        __syncwarp();
        if (l < kAcc2) {
          float2 s0 = make_float2(0.0f, 0.0f), s1 = make_float2(0.0f, 0.0f);
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            s0 = __fadd2_rn(s0, part[j * kAcc2 + l]);
            s1 = __fadd2_rn(s1, part[(j + 1) * kAcc2 + l]);
          }
          s0 = __fadd2_rn(s0, s1);
          const int fu = kHalfUnit * half + fa;     // frame within the unit
          rbuf[fu * kLag + 2 * l] = s0.x;
          if (2 * l + 1 < kLag) rbuf[fu * kLag + 2 * l + 1] = s0.y;
        }
        __syncwarp();
      }
      if (half == 0) { stage_half(b, g, 1); cur = true; }
      else if (un < A.n_units) { stage_half(bn, gn, 0); cur = true; }
      else cur = false;
    }

    // Levinson-Durbin, one frame per lane, float64 state (as in lpc_wave_kernel)
    double r[kLag], a[kLag];
#pragma unroll
    for (int k = 0; k < kLag; ++k) r[k] = static_cast<double>(rbuf[lane * kLag + k]);
    double E = r[0] + A.eps;
    if (V & kLvNoLev) {
#pragma unroll
      for (int i = 1; i < kLag; ++i) a[i] = r[i];
    }
#pragma unroll
    for (int i = 1; i < kLag; ++i) {
      if (V & kLvNoLev) break;
      if ((V & kLvM24) || i <= A.M) {
        double acc = r[i];
        if (V & kLvSplit) {   // four interleaved partial sums: the dependent chain is i / 4 long instead of i
          double c1 = 0.0, c2 = 0.0, c3 = 0.0;
#pragma unroll
          for (int j = 1; j < i; ++j) {
            if ((j & 3) == 0) acc = fma(a[j], r[i - j], acc);
            if ((j & 3) == 1) c1 = fma(a[j], r[i - j], c1);
            if ((j & 3) == 2) c2 = fma(a[j], r[i - j], c2);
            if ((j & 3) == 3) c3 = fma(a[j], r[i - j], c3);
          }
          if (i > 1) acc = (acc + c1) + (c2 + c3);
        } else {
#pragma unroll
          for (int j = 1; j < i; ++j) acc = fma(a[j], r[i - j], acc);
        }
        double inv = static_cast<double>(__frcp_rn(static_cast<float>(E)));
        inv = fma(inv, fma(-E, inv, 1.0), inv);
        inv = fma(inv, fma(-E, inv, 1.0), inv);
        const bool tame = fabs(E) > 1e-30 && fabs(E) < 1e30;
        const double kk = tame ? (-acc * inv) : (-acc / E);
#pragma unroll
        for (int j = 1; 2 * j <= i; ++j) {
          const double lo = a[j], hi = a[i - j];
          a[j] = fma(kk, hi, lo);
          if (2 * j != i) a[i - j] = fma(kk, lo, hi);
        }
        a[i] = kk;
        E *= fma(-kk, kk, 1.0);
      }
    }
    double gain = r[0];
#pragma unroll
    for (int j = 1; j < kLag; ++j)
      if ((V & kLvM24) || j <= A.M) gain = fma(r[j], a[j], gain);
    a[0] = sqrt(gain);

    const int64_t row0 = static_cast<int64_t>(b) * A.n_frames + f0;
    const bool staged = A.bulk_out && nvalid == kUnit && (((row0 * D) & 3) == 0);
    __syncwarp();  // every lane has read its rbuf row
    if (staged) {
      float* o = rbuf + lane * D;
#pragma unroll
      for (int k = 0; k < kLag; ++k)
        if ((V & kLvM24) || k < D) o[k] = static_cast<float>(a[k]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) bulk_s2g(A.y + row0 * D, rbuf, static_cast<uint32_t>(kUnit * D) * 4u);
      store_pending = true;
    } else if (lane < nvalid) {
      float* o = A.y + (row0 + lane) * D;
#pragma unroll
      for (int k = 0; k < kLag; ++k)
        if (k < D) o[k] = static_cast<float>(a[k]);
    }
    u = un; b = bn; g = gn;
  }
  if (store_pending && lane == 0) bulk_wait_read();
}

}  // namespace

int lpc_wave_fast_try(const float* x, const float* window, float* y, int64_t batch, int64_t T_len,
                      const dsb200_frame_params* fp, int32_t M, double eps, int device, cudaStream_t stream) {
  if (fp->frame_length > 16 * kCh || M >= kLag || (fp->frame_period & 1) || fp->zmean || T_len > (1 << 30))
    return DSB200_E_UNSUPPORTED;
  const int64_t N = dsb200_num_frames(T_len, fp->frame_period);
  const int64_t U = (N + kUnit - 1) / kUnit;
  if (batch * U > (1LL << 30)) return DSB200_E_UNSUPPORTED;
  const int left = fp->center ? fp->frame_length / 2 : 0;
  const int span = ((kHalfUnit - 1) * fp->frame_period + 16 * kCh + kHalo + 3) & ~3;
  const size_t per_warp = static_cast<size_t>(span) * 4 + 2 * 16 * kLag * 8 + kUnit * kLag * 4;
  // knob LPC_W = 8 | 12: warps per CTA
  int kLWarps = (knob("LPC_W", 12) == 8) ? 8 : 12;
  if (8 * kLWarps + 448 * sizeof(float) + kLWarps * per_warp > static_cast<size_t>(max_dynamic_smem(device))) kLWarps = 8;
  const size_t smem = 8 * kLWarps + 448 * sizeof(float) + kLWarps * per_warp;
  if (smem > static_cast<size_t>(max_dynamic_smem(device))) return DSB200_E_UNSUPPORTED;

  LArgs A{};
  A.x = x;
  A.window = window;
  A.y = y;
  A.T = static_cast<int>(T_len);
  A.n_frames = static_cast<int>(N);
  A.units_per_utt = static_cast<int>(U);
  A.n_units = static_cast<int>(batch * U);
  A.L = fp->frame_length;
  A.P = fp->frame_period;
  A.left = left;
  A.pad_mode = fp->pad_mode;
  A.M = M;
  A.span = span;
  A.bulk_in = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (T_len % 4 == 0) && (left % 4 == 0) &&
              ((kHalfUnit * fp->frame_period) % 4 == 0);
  A.bulk_out = ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
  A.eps = eps;
  A.stagger = knob("LPC_STAGGER", kLpcStagger);
  const bool full = fp->frame_length == 16 * kCh;
  const int blocks = static_cast<int>(std::min<int64_t>((A.n_units + kLWarps - 1) / kLWarps, sm_count(device)));
  auto launch = [&](auto kern) -> int {
    DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<blocks, kLWarps * 32, smem, stream>>>(A);
    return DSB200_OK;
  };
  int rc;
  int v = full ? knob("LPC_V", kLpcVariant) : 0;
  if (M != kLag - 1) v &= ~kLvM24;
  if (v & kLvRolled) v &= ~kLvM24;
  if (v & kLvLagPair) {
    // lag-pair form: only samples < 400 of a frame are ever read from the staged span
    A.span = ((kHalfUnit - 1) * fp->frame_period + 16 * kCh + 3) & ~3;
    const size_t per_warp2 = static_cast<size_t>(A.span) * 4 + 2 * 16 * kAcc2 * 8 + kUnit * kLag * 4;
    // 16 warps at 128 registers: the lag loop needs 123, the unguarded order-24 recursion fits with ~100 B of spills; with
    // the per-order guards of M < 24 it would spill 1 KB, so those launches stay at 12 warps
    const int w2 = ((v & kLvM24) && knob("LPC_W2", kLpcWarps2) == 16) ? 16 : 12;
    const size_t smem2 = 8 * w2 + 448 * sizeof(float) + w2 * per_warp2;
    if (smem2 <= static_cast<size_t>(max_dynamic_smem(device))) {
      const int blocks2 = static_cast<int>(std::min<int64_t>((A.n_units + w2 - 1) / w2, sm_count(device)));
      auto launch2 = [&](auto kern) -> int {
        DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2)));
        kern<<<blocks2, w2 * 32, smem2, stream>>>(A);
        return DSB200_OK;
      };
      const bool m24 = (v & kLvM24) != 0;
      const int extra = v & ~(kLvLagPair | kLvM24 | kLvHalo | kLvMul2 | kLvRolled);
      if (m24 && extra == kLvSplit)
        rc = w2 == 16 ? launch2(lpc_wave2_kernel<16, kLvM24 | kLvSplit>) : launch2(lpc_wave2_kernel<12, kLvM24 | kLvSplit>);
#ifdef DSB200_LPC_DIAG
      else if (m24 && extra == kLvNoLev)
        rc = w2 == 16 ? launch2(lpc_wave2_kernel<16, kLvM24 | kLvNoLev>) : launch2(lpc_wave2_kernel<12, kLvM24 | kLvNoLev>);
      else if (m24 && extra == (kLvNoLev | kLvNoRed))
        rc = w2 == 16 ? launch2(lpc_wave2_kernel<16, kLvM24 | kLvNoLev | kLvNoRed>)
                      : launch2(lpc_wave2_kernel<12, kLvM24 | kLvNoLev | kLvNoRed>);
#endif
      else if (w2 == 16) rc = m24 ? launch2(lpc_wave2_kernel<16, kLvM24>) : launch2(lpc_wave2_kernel<16, 0>);
      else rc = m24 ? launch2(lpc_wave2_kernel<12, kLvM24>) : launch2(lpc_wave2_kernel<12, 0>);
      if (rc != DSB200_OK) return rc;
      return after_launch("lpc_wave2_kernel");
    }
    A.span = span;
  }
  if (kLWarps == 8) rc = full ? launch(lpc_wave_kernel<true, 8>) : launch(lpc_wave_kernel<false, 8>);
  else if (!full) rc = launch(lpc_wave_kernel<false, 12>);
  else switch (v) {
    case 3: rc = launch(lpc_wave_kernel<true, 12, 3>); break;
    case 7: rc = launch(lpc_wave_kernel<true, 12, 7>); break;
    case 8: rc = launch(lpc_wave_kernel<true, 12, 8>); break;
    case 11: rc = launch(lpc_wave_kernel<true, 12, 11>); break;
#ifdef DSB200_LPC_DIAG
    case 64: rc = launch(lpc_wave_kernel<true, 12, 64>); break;
    case 128: rc = launch(lpc_wave_kernel<true, 12, 128>); break;
    case 192: rc = launch(lpc_wave_kernel<true, 12, 192>); break;
    case 195: rc = launch(lpc_wave_kernel<true, 12, 195>); break;
#endif
    default: rc = launch(lpc_wave_kernel<true, 12>); break;
  }
  if (rc != DSB200_OK) return rc;
  return after_launch("lpc_wave_kernel");
}

}  // namespace dsb200

namespace dsb200 {
int mfcc_wave_try(const float* x, const float* window, const float* H, const int32_t* cb, const int32_t* ce,
                  const float* W, const float* lifter, const int32_t* plan, float* const* y_dst, int n_dst,
                  int64_t row_off, int64_t batch, int64_t T_len, const dsb200_stft_params* sp,
                  const dsb200_mfcc_params* mp, int device, cudaStream_t stream);

static int mfcc_wave_checked(const void* x, const void* window, const void* H, const int32_t* cb, const int32_t* ce,
                             const void* W, const void* lifter, const int32_t* plan, void* const* y_dst, int n_dst,
                             int64_t row_off, int64_t batch, int64_t T, const dsb200_stft_params* sp,
                             const dsb200_mfcc_params* mp, int device, void* stream) {
  DSB_REQUIRE(sp != nullptr && mp != nullptr, "params are NULL");
  DSB_REQUIRE(sp->frame.frame_length > 0, "frame_length must be positive.");
  DSB_REQUIRE(sp->frame.frame_period > 0, "frame_period must be positive.");
  DSB_REQUIRE(sp->spec.fft_length > 1 && sp->spec.fft_length % 2 == 0, "fft_length must be positive even.");
  DSB_REQUIRE(sp->spec.eps >= 0, "eps must be non-negative.");
  DSB_REQUIRE(mp->fbank.n_channel > 0, "n_channel must be positive.");
  DSB_REQUIRE(mp->fbank.floor > 0, "floor must be positive.");
  DSB_REQUIRE(mp->mfcc_order >= 0 && mp->mfcc_order < mp->fbank.n_channel, "mfcc_order must be less than n_channel.");
  DSB_REQUIRE(mp->out_format >= DSB200_MFCC_Y && mp->out_format <= DSB200_MFCC_YCE, "out_format %d is not supported.", mp->out_format);
  DSB_REQUIRE(mp->fbank.fft_length == sp->spec.fft_length, "the filter bank and the STFT must share fft_length.");
  DSB_REQUIRE(T >= 1 && batch >= 0 && row_off >= 0, "bad batch / waveform length / row offset");
  DSB_REQUIRE(y_dst != nullptr && n_dst >= 1 && n_dst <= 8, "between 1 and 8 destinations");
  if (batch == 0) return DSB200_OK;
  DSB_REQUIRE(x && window && H && W && lifter, "NULL data pointer");
  for (int d = 0; d < n_dst; ++d) DSB_REQUIRE(y_dst[d] != nullptr, "NULL destination pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  const int rc = mfcc_wave_try(static_cast<const float*>(x), static_cast<const float*>(window),
                               static_cast<const float*>(H), cb, ce, static_cast<const float*>(W),
                               static_cast<const float*>(lifter), plan, reinterpret_cast<float* const*>(y_dst), n_dst,
                               row_off, batch, T, sp, mp, device, static_cast<cudaStream_t>(stream));
  if (rc == DSB200_E_UNSUPPORTED)
    return fail(DSB200_E_UNSUPPORTED, "fused waveform->MFCC kernel not available for this configuration");
  return rc;
}
}  // namespace dsb200

extern "C" {
int dsb200_mfcc_wave_f32(const void* x, const void* window, const void* H, const int32_t* cb, const int32_t* ce,
                         const void* W, const void* lifter, void* y, int64_t batch, int64_t T,
                         const dsb200_stft_params* sp, const dsb200_mfcc_params* mp, int device, void* stream) {
  void* dst[1] = {y};
  return dsb200::mfcc_wave_checked(x, window, H, cb, ce, W, lifter, nullptr, dst, 1, 0, batch, T, sp, mp, device, stream);
}
int dsb200_mfcc_wave_f64(const void*, const void*, const void*, const int32_t*, const int32_t*, const void*, const void*,
                         void*, int64_t, int64_t, const dsb200_stft_params*, const dsb200_mfcc_params*, int, void*) {
  return dsb200::fail(DSB200_E_UNSUPPORTED, "fused waveform->MFCC kernel is float32 only");
}
int dsb200_mfcc_wave_ex_f32(const void* x, const void* window, const void* H, const int32_t* cb, const int32_t* ce,
                            const void* W, const void* lifter, const int32_t* plan, void* const* y_dst, int32_t n_dst,
                            int64_t row_offset, int64_t batch, int64_t T, const dsb200_stft_params* sp,
                            const dsb200_mfcc_params* mp, int device, void* stream) {
  return dsb200::mfcc_wave_checked(x, window, H, cb, ce, W, lifter, plan, y_dst, n_dst, row_offset, batch, T, sp, mp,
                                   device, stream);
}
int dsb200_mfcc_wave_ex_f64(const void*, const void*, const void*, const int32_t*, const int32_t*, const void*,
                            const void*, const int32_t*, void* const*, int32_t, int64_t, int64_t, int64_t,
                            const dsb200_stft_params*, const dsb200_mfcc_params*, int, void*) {
  return dsb200::fail(DSB200_E_UNSUPPORTED, "fused waveform->MFCC kernel is float32 only");
}
}
