// Mel-cepstral analysis, fast path (fp32, sm_100a): fft_length = 512 (257 bins), cep_order <= 24.
//
// Reference: diffsptk/modules/mcep.py:189-224.  Per Newton step, with the FFT pairs folded into the
// host-built tables (tables.make_mcep_tables):
//     d  = mc @ G            [25 x 257]       e = exp(log x - 2 d)
//     rt = e @ Hm            [257 x 49]       solve (Toeplitz(rt[:25]) + Hankel(rt)) g = rt[:25] - alpha;  mc += g
// This is ~250 k FP32 multiply-adds per frame: the FP32 pipe bounds it at 1-2 % of the HBM roofline, so
// the job of the mapping is to keep every operand of those multiply-adds on chip and reused:
//   * a warp owns an OCTET of 8 frames; all arithmetic is packed float2 = (frame 2p, frame 2p+1);
//   * lane l owns bins k = l + 32 t (t < 9) of all 8 frames in registers (log x, d / e): each table
//     element fetched from shared memory feeds 8 multiply-adds (the generic kernel: 1);
//   * the 49 (25) dot products over bins are reduced across the 32 lanes with a transposed butterfly
//     (9 shuffles per output instead of 40);
//   * the 25 x 25 Newton system is symmetric, so Gaussian elimination never needs another lane's ROW:
//     the pivot row equals the pivot COLUMN, one entry per lane, published through shared memory in one
//     parallel store; two frames' systems are eliminated per pass (packed), back substitution likewise.
// The three tables (G, Hm, P0: 114 KB) stay resident in shared memory for the whole persistent CTA.
//
// v2 (round 2).  v1 ran 8 warps of 255 registers (two per scheduler) and kept the log spectrum of the octet in
// shared memory (9 KB per warp): FMA pipe 39 %, every dependency stall exposed.  Now
//   * e = exp(log x - 2 d) is evaluated as x * exp(-2 d) with x re-read from global memory in every Newton step
//     (the octet's 8 KB stay in L2 for its ten steps; HBM traffic is unchanged) -- no per-warp spectrum copy, so
//     the CTA has room for 12 (default) or 16 warps;
//   * the Newton systems are eliminated two frame pairs per pass (100 instead of 200 matrix registers), which
//     fits the 168-register budget of 12 warps without spills;
//   * rows past the end of a partial octet are clamped to the last valid row and simply not stored: no load
//     predicates anywhere in the step.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace dsb200 {
namespace {

constexpr int kKT = 9;               // bins per lane: k = lane + 32 t
constexpr int kKS = 32 * kKT;        // 288: padded number of bins
constexpr int kK = 257;
constexpr int kDM = 25;              // max cepstral dimension (M + 1)
constexpr int kJS = 49;              // row stride of Hm in shared memory (odd: conflict-free), = 2 * 24 + 1
constexpr int kPS = 25;              // row stride of P0 in shared memory (odd)
constexpr int kJO = kDM - 1;          // rt is kept symmetric about its origin, rt[-j] = rt[j] (j <= 24): |i - c| needs no abs
constexpr int kJSE = kJO + kJS;      // 73 values per frame pair
constexpr int kWarpFloats = kDM * 8 + 4 * kJSE * 2 + 512 + 256 + 8;   // per-warp scratch: mc, rt, pivot column, solution, rhs

struct MArgs {
  const float* x;    // [rows, 257] power spectrum
  float* y;          // [rows, D]
  const float* P0;   // [257, D]
  const float* G;    // [D, 257]
  const float* Hm;   // [257, J]
  const float* av;   // [D]
  int64_t rows;
  int D, J, n_iter;
  int stagger;       // start-up offset between the warps of one scheduler, cycles (stagger_start)
};

constexpr int kMcepStagger = 0;  // default of the knob MCEP_STAGGER
constexpr int kMcepVariant = 122; // default of the knob MCEP_V

// Compile-time loop: the elimination below indexes register arrays with the loop variable, so it must be
// expanded even when the body is too large for `#pragma unroll` heuristics.
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 shfl_xor2(float2 v, int m) {
  return f2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
__device__ __forceinline__ float2 sel2(bool c, float2 a, float2 b) { return c ? a : b; }

// Sum p[0..3] (frames (0,1), (2,3), (4,5), (6,7)) over the 32 lanes.  Returns, in EVERY lane, the total
// of frame (lane >> 2); 9 shuffles (a plain butterfly would need 40).
__device__ __forceinline__ float reduce8(const float2 (&p)[4], int lane) {
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
  float2 k0 = sel2(h16, p[2], p[0]), k1 = sel2(h16, p[3], p[1]);
  const float2 s0 = sel2(h16, p[0], p[2]), s1 = sel2(h16, p[1], p[3]);
  k0 = __fadd2_rn(k0, shfl_xor2(s0, 16));
  k1 = __fadd2_rn(k1, shfl_xor2(s1, 16));
  float2 k = sel2(h8, k1, k0);
  const float2 s = sel2(h8, k0, k1);
  k = __fadd2_rn(k, shfl_xor2(s, 8));
  float v = h4 ? k.y : k.x;
  const float w = h4 ? k.x : k.y;
  v += __shfl_xor_sync(0xffffffffu, w, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

__device__ __forceinline__ float fast_rcp(float x) {  // MUFU.RCP + one Newton step (~full fp32 accuracy)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.0f), r);
}

// kMW warps per CTA (12: 168 registers per thread, 16: 128); kQ packed frame pairs eliminated per pass.
template <int kMW, int kQ, int SOLVE>
__global__ void __launch_bounds__(kMW * 32, 1) mcep_fast_kernel(const MArgs A) {
  constexpr bool ROLLED = SOLVE != 0;  // SOLVE: 0 unrolled elimination + back substitution, 1 rolled Gauss-Jordan with
                                       // lane = row and two frame pairs per pass, 2 four rows of one system per lane
  constexpr int kMT = kMW * 32;
  constexpr int kH = kQ / 2;           // float4 groups (two pairs each) exchanged through shared memory
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = A.D, J = A.J;

  float* Gs = reinterpret_cast<float*>(smem_raw);        // [kDM][kKS]   G[m][k]
  float* Hs = Gs + kDM * kKS;                            // [kKS][kJS]   Hm[k][j]
  float* Ps = Hs + kKS * kJS;                            // [kKS][kPS]   P0[k][m]
  float* avs = Ps + kKS * kPS;                           // [32]
  float* wbase = avs + 32 + warp * kWarpFloats;
  float* mcs = wbase;                                    // [kDM][8]   mc[m][frame]
  float2* rts = reinterpret_cast<float2*>(mcs + kDM * 8);   // [4][kJSE] rt[pair][kJO + j] = (frame 2p, frame 2p+1), j = -24..48
  float4* col = reinterpret_cast<float4*>(rts + 4 * kJSE);   // [2][64] pivot column (= pivot row, by symmetry), 8 systems; [32, 64) stay zero
  float4* xs = col + 128;                                // [2][32] solution broadcast
  float4* pb = xs + 64;                                  // [2]     pivot right-hand sides

  // tables -> shared memory (zero padded to 288 bins so that the tail lanes contribute nothing)
  for (int i = tid; i < kDM * kKS; i += kMT) {
    const int m = i / kKS, k = i - m * kKS;
    Gs[i] = (m < D && k < kK) ? A.G[m * kK + k] : 0.0f;
  }
  for (int i = tid; i < kKS * kJS; i += kMT) {
    const int k = i / kJS, j = i - k * kJS;
    Hs[i] = (k < kK && j < J) ? A.Hm[k * J + j] : 0.0f;
  }
  for (int i = tid; i < kKS * kPS; i += kMT) {
    const int k = i / kPS, m = i - k * kPS;
    Ps[i] = (k < kK && m < D) ? A.P0[k * D + m] : 0.0f;
  }
  if (tid < 32) avs[tid] = tid < D ? A.av[tid] : 0.0f;
  __syncthreads();

  for (int g = 0; g < 2; ++g) col[64 * g + 32 + lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  __syncwarp();
  stagger_start(A.stagger);
  const int64_t n_oct = (A.rows + 7) / 8;
  for (int64_t oct = static_cast<int64_t>(blockIdx.x) * kMW + warp; oct < n_oct;
       oct += static_cast<int64_t>(gridDim.x) * kMW) {
    // The last octet of a batch that is not a multiple of 8 is moved back to end at the last row: its first rows
    // are then computed twice (bit-identical results, stored twice) and no load or store needs a predicate.
    const int64_t r0 = (oct * 8 + 8 <= A.rows) ? oct * 8 : A.rows - 8;
    constexpr int nf = 8;
    const float* xb = A.x + r0 * kK;                       // row f, bin k: xb[f * kK + k] (immediate offsets)
    // Lane column t = 8 holds bins 256 + lane: only lane 0 has one.  The others re-read bin 256; their table
    // entries are zero, so the (finite) value contributes nothing.
    const int klast = (lane + 32 * (kKT - 1) < kK) ? lane + 32 * (kKT - 1) : kK - 1;

    // ---- initial estimate mc = log x @ P0 (mcep.py:203-207 folded) ------------------------------
    {
      float2 lx[kKT][4];                                   // (log x[2p][k_t], log x[2p+1][k_t])
#pragma unroll
      for (int t = 0; t < kKT; ++t) {
        const int k = (t < kKT - 1) ? lane + 32 * t : klast;
#pragma unroll
        for (int p = 0; p < 4; ++p) lx[t][p] = f2(logf(__ldg(xb + 2 * p * kK + k)), logf(__ldg(xb + (2 * p + 1) * kK + k)));
      }
      for (int m = 0; m < D; ++m) {
        float2 p4[4] = {f2(0, 0), f2(0, 0), f2(0, 0), f2(0, 0)};
#pragma unroll
        for (int t = 0; t < kKT; ++t) {
          const float c = Ps[(lane + 32 * t) * kPS + m];
#pragma unroll
          for (int p = 0; p < 4; ++p) p4[p] = __ffma2_rn(lx[t][p], f2(c, c), p4[p]);
        }
        const float v = reduce8(p4, lane);
        if ((lane & 3) == 0) mcs[m * 8 + (lane >> 2)] = v;
      }
    }
    __syncwarp();

    for (int it = 0; it < A.n_iter; ++it) {
      // ---- d = mc @ G ; e = exp(log x - 2 d) ---------------------------------------------------
      float2 e[kKT][4];
#pragma unroll
      for (int t = 0; t < kKT; ++t)
#pragma unroll
        for (int p = 0; p < 4; ++p) e[t][p] = f2(0, 0);
      for (int m = 0; m < D; ++m) {
        const float4 ma = *reinterpret_cast<const float4*>(mcs + m * 8);
        const float4 mb = *reinterpret_cast<const float4*>(mcs + m * 8 + 4);
        const float2 mc2[4] = {f2(ma.x, ma.y), f2(ma.z, ma.w), f2(mb.x, mb.y), f2(mb.z, mb.w)};
#pragma unroll
        for (int t = 0; t < kKT; ++t) {
          const float gk = Gs[m * kKS + lane + 32 * t];
#pragma unroll
          for (int p = 0; p < 4; ++p) e[t][p] = __ffma2_rn(mc2[p], f2(gk, gk), e[t][p]);
        }
      }
      // e = exp(log x - 2 d) = x * 2^(-2 log2(e) d): the spectrum comes back from L2, one MUFU.EX2 per value
      constexpr float kM2L2E = -2.885390081777927f;        // -2 / ln 2
#pragma unroll
      for (int t = 0; t < kKT; ++t) {
        const int k = (t < kKT - 1) ? lane + 32 * t : klast;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float2 xv = f2(__ldg(xb + 2 * p * kK + k), __ldg(xb + (2 * p + 1) * kK + k));
          const float2 a = __fmul2_rn(e[t][p], f2(kM2L2E, kM2L2E));
          e[t][p] = __fmul2_rn(xv, f2(exp2f(a.x), exp2f(a.y)));
        }
      }
      // ---- rt = e @ Hm, reduced over the lanes; four columns per round so that the shuffle chains of one
      //      column overlap the multiply-adds of the next ---------------------------------------------------
      auto rt_columns = [&](auto nc_c, int j0) {
        constexpr int NC = decltype(nc_c)::value;
        float2 acc[NC][4];
#pragma unroll
        for (int u = 0; u < NC; ++u)
#pragma unroll
          for (int p = 0; p < 4; ++p) acc[u][p] = f2(0, 0);
#pragma unroll
        for (int t = 0; t < kKT; ++t) {
#pragma unroll
          for (int u = 0; u < NC; ++u) {
            const float hk = Hs[(lane + 32 * t) * kJS + j0 + u];
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[u][p] = __ffma2_rn(e[t][p], f2(hk, hk), acc[u][p]);
          }
        }
#pragma unroll
        for (int u = 0; u < NC; ++u) {
          const float v = reduce8(acc[u], lane);
          if ((lane & 3) == 0)
          {
            if constexpr (SOLVE == 2) {      // one scalar row per frame: rtf[frame][kJO + j]
              float* rp = reinterpret_cast<float*>(rts) + (lane >> 2) * kJSE + kJO;
              rp[j0 + u] = v;
              if (j0 + u <= kJO) rp[-(j0 + u)] = v;
            } else {
              float* rp = reinterpret_cast<float*>(rts) + ((lane >> 3) * kJSE + kJO) * 2 + ((lane >> 2) & 1);
              rp[2 * (j0 + u)] = v;
              if (j0 + u <= kJO) rp[-2 * (j0 + u)] = v;     // mirror (j = 0 rewrites itself)
            }
          }
        }
      };
      {
        int j = 0;
        for (; j + 4 <= J; j += 4) rt_columns(std::integral_constant<int, 4>{}, j);
        for (; j < J; ++j) rt_columns(std::integral_constant<int, 1>{}, j);
      }
      __syncwarp();

      if constexpr (SOLVE == 2) {
        // ---- Newton systems, variant 2: FOUR ROWS OF ONE SYSTEM PER LANE, four systems per pass ------------------
        // With lane = row and two packed frame pairs per register (variant 1) every multiply-add needs a pivot-row
        // entry of ITS system: a 16-byte shared-memory broadcast (four systems) per two FFMA2 -- the kernel was bound
        // by the shared-memory pipe (70 %, profiles/r2_mcep_rolled.txt).  Here a float2 holds two ROWS of the same
        // system, so the pivot-row entry is one scalar that FFMA2 broadcasts to both halves: a 4-byte load (four
        // distinct addresses, one wavefront) per two FFMA2.  lane = 8 s + r: system s of the pass, rows r, r + 7,
        // r + 14, r + 21 (r < 7; rows 25..27 are zero rows).  Same rolled Gauss-Jordan, same shifting registers.
        const int s4 = lane >> 3, r = lane & 7;
        const bool lane_on = r < 7;
        float* colf = reinterpret_cast<float*>(col) + s4 * 33;     // [4][33] published pivot column per system (odd pitch:
                                                                   // the four systems' entries sit in different banks)
        float* pbf = reinterpret_cast<float*>(col) + 136 + s4;     // [4] pivot right-hand side per system
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          const int sys = 4 * pass + s4;
          const float* rt = reinterpret_cast<const float*>(rts) + sys * kJSE + kJO;
          float2 lo[kDM], hi[kDM], blo, bhi;
          auto elem = [&](int row, int c) -> float {      // A[row][c]; rows D..24 are identity rows, rows >= 25 zero
            const bool real = lane_on && row < D;
            float v = (lane_on && row == c && row >= D) ? 1.0f : 0.0f;
            if (real && c < D) v = rt[row - c] + rt[row + c];
            return v;
          };
#pragma unroll
          for (int c = 0; c < kDM; ++c) {
            lo[c] = f2(elem(r, c), elem(r + 7, c));
            hi[c] = f2(elem(r + 14, c), elem(r + 21, c));
          }
          auto rhs = [&](int row) -> float { return (lane_on && row < D) ? rt[row] - avs[row] : 0.0f; };
          blo = f2(rhs(r), rhs(r + 7));
          bhi = f2(rhs(r + 14), rhs(r + 21));
          float2 dlo = f2(1, 1), dhi = f2(1, 1);
          int pvr = 0, pvt = 0;                                   // pivot row = 7 pvt + pvr
          auto pivot_step = [&](auto nc_c, int pv) {
            constexpr int NC = decltype(nc_c)::value;
            if (lane_on) {
              colf[r] = lo[0].x;
              colf[r + 7] = lo[0].y;
              colf[r + 14] = hi[0].x;
              colf[r + 21] = hi[0].y;
            }
            const bool mine = lane_on && r == pvr;
            {
              const float2 bsel = pvt < 2 ? blo : bhi;
              const float bval = (pvt & 1) ? bsel.y : bsel.x;
              if (mine) *pbf = bval;
            }
            __syncwarp();
            const float rinv = fast_rcp(colf[pv]);
            const bool p0 = mine && pvt == 0, p1 = mine && pvt == 1, p2 = mine && pvt == 2, p3 = mine && pvt == 3;
            const float2 flo = f2(p0 ? 0.0f : -lo[0].x * rinv, p1 ? 0.0f : -lo[0].y * rinv);
            const float2 fhi = f2(p2 ? 0.0f : -hi[0].x * rinv, p3 ? 0.0f : -hi[0].y * rinv);
            dlo = f2(p0 ? rinv : dlo.x, p1 ? rinv : dlo.y);
            dhi = f2(p2 ? rinv : dhi.x, p3 ? rinv : dhi.y);
            const float* cp = colf + pv + 1;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
              const float rc = cp[k];
              lo[k] = __ffma2_rn(flo, f2(rc, rc), lo[k + 1]);
              hi[k] = __ffma2_rn(fhi, f2(rc, rc), hi[k + 1]);
            }
            const float bp = *pbf;
            blo = __ffma2_rn(flo, f2(bp, bp), blo);
            bhi = __ffma2_rn(fhi, f2(bp, bp), bhi);
            __syncwarp();
            if (++pvr == 7) { pvr = 0; ++pvt; }
          };
          static_for<0, (kDM - 1) / 4>([&](auto ph_c) {
            constexpr int ph = decltype(ph_c)::value;
#pragma unroll 1
            for (int pv = 4 * ph; pv < 4 * ph + 4; ++pv) pivot_step(std::integral_constant<int, kDM - 1 - 4 * ph>{}, pv);
          });
          pivot_step(std::integral_constant<int, 0>{}, kDM - 1);
          // mc += g: x_row = b_row / pivot_row
          const float2 xlo = __fmul2_rn(blo, dlo), xhi = __fmul2_rn(bhi, dhi);
          if (lane_on) {
            if (r < D) mcs[r * 8 + sys] += xlo.x;
            if (r + 7 < D) mcs[(r + 7) * 8 + sys] += xlo.y;
            if (r + 14 < D) mcs[(r + 14) * 8 + sys] += xhi.x;
            if (r + 21 < D) mcs[(r + 21) * 8 + sys] += xhi.y;
          }
          __syncwarp();
        }
      } else {
      // ---- Newton systems, kQ packed frame pairs per pass (default: all 8 frames in one pass): lane = row ----
#pragma unroll 1
      for (int pass = 0; pass < 4 / kQ; ++pass) {
        const int i = lane;
        const float2* rt0 = rts + pass * kQ * kJSE + kJO + i;   // (rt0 + q kJSE)[+-c] = rt[i +- c]
        float2 a[kQ][kDM], b[kQ];
        if (i < D) {
          const float alpha_i = avs[i];
#pragma unroll
          for (int q = 0; q < kQ; ++q) {
            const float2* rt = rt0 + q * kJSE;
#pragma unroll
            for (int c = 0; c < kDM; ++c) a[q][c] = (c < D) ? __fadd2_rn(rt[-c], rt[c]) : f2(0, 0);
            b[q] = __fadd2_rn(rt[0], f2(-alpha_i, -alpha_i));
          }
        } else {
          // Rows D..24 are identity rows with a zero right-hand side (lanes >= 25 are all-zero rows that are
          // never pivots): the elimination can then run all 25 pivots unconditionally -- no run-time
          // `pv < D` tests, which the compiler otherwise keeps as a bit mask of 50 predicates.
#pragma unroll
          for (int q = 0; q < kQ; ++q) {
#pragma unroll
            for (int c = 0; c < kDM; ++c) a[q][c] = f2(c == i ? 1.0f : 0.0f, c == i ? 1.0f : 0.0f);
            b[q] = f2(0, 0);
          }
        }
        if constexpr (ROLLED) {
          // ---- Gauss-Jordan in a ROLLED pivot loop (round 2).  The fully unrolled elimination below is ~1 700
          // instructions per pass, the Newton step ~4 000 = 64 KB of code for 12 warps at different places of it:
          // `stall_no_inst` was 20 % (profiles/r2_mcep_fast_v2).  Here every row shifts its registers left by one
          // column per pivot -- register k always holds column pv + k -- so the same ~150 instructions serve all 25
          // pivots.  Eliminating above the diagonal as well keeps every lane in the loop to the end and removes the
          // back substitution: x_i = b_i / pivot_i.  The trailing block stays symmetric, so the pivot row is still
          // read from the column the lanes publish.
          float2 dinv[kQ];
#pragma unroll
          for (int q = 0; q < kQ; ++q) dinv[q] = f2(1, 1);
          // One pivot.  NC = number of columns right of the pivot that are updated: a compile-time multiple of four
          // (>= 24 - pv; the surplus columns multiply the zeros that lanes 25.. publish), so the pivot loop is six
          // short rolled loops of four pivots each with a fixed trip count -- no exit tests inside the update, and
          // no register moves to merge exit paths (first rolled version: 8 % of the kernel's instructions).
          auto pivot_step = [&](auto nc_c, int pv) {
            constexpr int NC = decltype(nc_c)::value;
#pragma unroll
            for (int g = 0; g < kH; ++g)
              col[64 * g + lane] = make_float4(a[2 * g][0].x, a[2 * g][0].y, a[2 * g + 1][0].x, a[2 * g + 1][0].y);
            if (lane == pv) {
#pragma unroll
              for (int g = 0; g < kH; ++g) pb[g] = make_float4(b[2 * g].x, b[2 * g].y, b[2 * g + 1].x, b[2 * g + 1].y);
            }
            __syncwarp();
            const bool piv = lane == pv;
            float2 f[kQ];
#pragma unroll
            for (int g = 0; g < kH; ++g) {
              const float4 p = col[64 * g + pv];
              const float2 r0 = f2(fast_rcp(p.x), fast_rcp(p.y)), r1 = f2(fast_rcp(p.z), fast_rcp(p.w));
              f[2 * g] = piv ? f2(0, 0) : __fmul2_rn(a[2 * g][0], f2(-r0.x, -r0.y));        // -a_ip / a_pp
              f[2 * g + 1] = piv ? f2(0, 0) : __fmul2_rn(a[2 * g + 1][0], f2(-r1.x, -r1.y));
              dinv[2 * g] = sel2(piv, r0, dinv[2 * g]);
              dinv[2 * g + 1] = sel2(piv, r1, dinv[2 * g + 1]);
            }
            const float4* cp = col + pv + 1;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
#pragma unroll
              for (int g = 0; g < kH; ++g) {
                const float4 cv = cp[64 * g + k];            // entries past column 24 are zero (lanes 25.., padding)
                a[2 * g][k] = __ffma2_rn(f[2 * g], f2(cv.x, cv.y), a[2 * g][k + 1]);
                a[2 * g + 1][k] = __ffma2_rn(f[2 * g + 1], f2(cv.z, cv.w), a[2 * g + 1][k + 1]);
              }
            }
#pragma unroll
            for (int g = 0; g < kH; ++g) {
              const float4 bv = pb[g];
              b[2 * g] = __ffma2_rn(f[2 * g], f2(bv.x, bv.y), b[2 * g]);
              b[2 * g + 1] = __ffma2_rn(f[2 * g + 1], f2(bv.z, bv.w), b[2 * g + 1]);
            }
            __syncwarp();
          };
#if defined(DSB200_MCEP_SINGLE_LOOP)   // A/B knob: one rolled loop over all 25 pivots, every update 24 columns wide
#pragma unroll 1
          for (int pv = 0; pv < kDM; ++pv) pivot_step(std::integral_constant<int, kDM - 1>{}, pv);
#elif !defined(DSB200_MCEP_PHASED)     // default: one rolled loop, warp-uniform exit from the update every four columns
#pragma unroll 1
          for (int pv = 0; pv < kDM; ++pv) {
#pragma unroll
            for (int g = 0; g < kH; ++g)
              col[64 * g + lane] = make_float4(a[2 * g][0].x, a[2 * g][0].y, a[2 * g + 1][0].x, a[2 * g + 1][0].y);
            if (lane == pv) {
#pragma unroll
              for (int g = 0; g < kH; ++g) pb[g] = make_float4(b[2 * g].x, b[2 * g].y, b[2 * g + 1].x, b[2 * g + 1].y);
            }
            __syncwarp();
            const bool piv = lane == pv;
            float2 f[kQ];
#pragma unroll
            for (int g = 0; g < kH; ++g) {
              const float4 p = col[64 * g + pv];
              const float2 r0 = f2(fast_rcp(p.x), fast_rcp(p.y)), r1 = f2(fast_rcp(p.z), fast_rcp(p.w));
              f[2 * g] = piv ? f2(0, 0) : __fmul2_rn(a[2 * g][0], f2(-r0.x, -r0.y));
              f[2 * g + 1] = piv ? f2(0, 0) : __fmul2_rn(a[2 * g + 1][0], f2(-r1.x, -r1.y));
              dinv[2 * g] = sel2(piv, r0, dinv[2 * g]);
              dinv[2 * g + 1] = sel2(piv, r1, dinv[2 * g + 1]);
            }
            const int n_left = kDM - 1 - pv;
#pragma unroll
            for (int k = 0; k < kDM - 1; ++k) {
              if ((k & 3) == 0 && k >= n_left) break;
#pragma unroll
              for (int g = 0; g < kH; ++g) {
                const float4 cv = col[64 * g + pv + 1 + k];
                a[2 * g][k] = __ffma2_rn(f[2 * g], f2(cv.x, cv.y), a[2 * g][k + 1]);
                a[2 * g + 1][k] = __ffma2_rn(f[2 * g + 1], f2(cv.z, cv.w), a[2 * g + 1][k + 1]);
              }
            }
#pragma unroll
            for (int g = 0; g < kH; ++g) {
              const float4 bv = pb[g];
              b[2 * g] = __ffma2_rn(f[2 * g], f2(bv.x, bv.y), b[2 * g]);
              b[2 * g + 1] = __ffma2_rn(f[2 * g + 1], f2(bv.z, bv.w), b[2 * g + 1]);
            }
            __syncwarp();
          }
#else   // A/B knob DSB200_MCEP_PHASED: six rolled loops of four pivots with static update widths (no exit tests; measured
        // slower: the seven instantiations of the step spill 400 bytes at 168 registers)
          static_for<0, (kDM - 1) / 4>([&](auto ph_c) {
            constexpr int ph = decltype(ph_c)::value;
#pragma unroll 1
            for (int pv = 4 * ph; pv < 4 * ph + 4; ++pv) pivot_step(std::integral_constant<int, kDM - 1 - 4 * ph>{}, pv);
          });
          pivot_step(std::integral_constant<int, 0>{}, kDM - 1);
#endif
#pragma unroll
          for (int g = 0; g < kH; ++g) {
            const float2 x0 = __fmul2_rn(b[2 * g], dinv[2 * g]), x1 = __fmul2_rn(b[2 * g + 1], dinv[2 * g + 1]);
            xs[32 * g + lane] = make_float4(x0.x, x0.y, x1.x, x1.y);
          }
          __syncwarp();
        } else {
        // Elimination.  The sub-matrix stays symmetric, so pivot-row entry c == column entry held by lane c:
        // one parallel store publishes the whole pivot row (float4 = two frame pairs).
        float2 dinv[kQ];                                   // 1 / (this lane's own pivot)
#pragma unroll
        for (int q = 0; q < kQ; ++q) dinv[q] = f2(1, 1);
        static_for<0, kDM>([&](auto pv_c) {
          constexpr int pv = decltype(pv_c)::value;
#pragma unroll
          for (int g = 0; g < kH; ++g)
            col[64 * g + lane] = make_float4(a[2 * g][pv].x, a[2 * g][pv].y, a[2 * g + 1][pv].x, a[2 * g + 1][pv].y);
          if (lane == pv) {
#pragma unroll
            for (int g = 0; g < kH; ++g) pb[g] = make_float4(b[2 * g].x, b[2 * g].y, b[2 * g + 1].x, b[2 * g + 1].y);
          }
          __syncwarp();
          const bool act = i > pv;                 // identity / zero rows hold a zero here: factor 0
          float2 f[kQ];
#pragma unroll
          for (int g = 0; g < kH; ++g) {
            const float4 p = col[64 * g + pv];
            const float2 r0 = f2(fast_rcp(p.x), fast_rcp(p.y)), r1 = f2(fast_rcp(p.z), fast_rcp(p.w));
            f[2 * g] = act ? __fmul2_rn(a[2 * g][pv], f2(-r0.x, -r0.y)) : f2(0, 0);       // -a_ip / a_pp
            f[2 * g + 1] = act ? __fmul2_rn(a[2 * g + 1][pv], f2(-r1.x, -r1.y)) : f2(0, 0);
            dinv[2 * g] = sel2(lane == pv, r0, dinv[2 * g]);
            dinv[2 * g + 1] = sel2(lane == pv, r1, dinv[2 * g + 1]);
          }
          if constexpr (pv < kDM - 1) {
#pragma unroll
            for (int c = pv + 1; c < kDM; ++c) {
#pragma unroll
              for (int g = 0; g < kH; ++g) {
                const float4 cv = col[64 * g + c];
                a[2 * g][c] = __ffma2_rn(f[2 * g], f2(cv.x, cv.y), a[2 * g][c]);
                a[2 * g + 1][c] = __ffma2_rn(f[2 * g + 1], f2(cv.z, cv.w), a[2 * g + 1][c]);
              }
            }
#pragma unroll
            for (int g = 0; g < kH; ++g) {
              const float4 bv = pb[g];
              b[2 * g] = __ffma2_rn(f[2 * g], f2(bv.x, bv.y), b[2 * g]);
              b[2 * g + 1] = __ffma2_rn(f[2 * g + 1], f2(bv.z, bv.w), b[2 * g + 1]);
            }
          }
          __syncwarp();
        });
        // Back substitution in place: x_c published by lane c, b -= U[., c] x_c in the rows above it
        // (x_c = 0 for c >= D: identity rows with a zero right-hand side).
        static_for<0, kDM>([&](auto cc) {
          constexpr int c = kDM - 1 - decltype(cc)::value;
          if (lane == c) {
#pragma unroll
            for (int g = 0; g < kH; ++g) {
              const float2 x0 = __fmul2_rn(b[2 * g], dinv[2 * g]), x1 = __fmul2_rn(b[2 * g + 1], dinv[2 * g + 1]);
              xs[32 * g + c] = make_float4(x0.x, x0.y, x1.x, x1.y);
            }
          }
          __syncwarp();
          if (i < c) {
#pragma unroll
            for (int g = 0; g < kH; ++g) {
              const float4 xv = xs[32 * g + c];
              b[2 * g] = __ffma2_rn(a[2 * g][c], f2(-xv.x, -xv.y), b[2 * g]);
              b[2 * g + 1] = __ffma2_rn(a[2 * g + 1][c], f2(-xv.z, -xv.w), b[2 * g + 1]);
            }
          }
        });
        __syncwarp();
        }
        if (i < D) {  // mc += g for the frames of this pass
          float4* mp = reinterpret_cast<float4*>(mcs + i * 8 + pass * kQ * 2);
#pragma unroll
          for (int g = 0; g < kH; ++g) {
            const float4 gv = xs[32 * g + i];
            float4 m = mp[g];
            m.x += gv.x; m.y += gv.y; m.z += gv.z; m.w += gv.w;
            mp[g] = m;
          }
        }
        __syncwarp();
      }
      }
    }
    // ---- store the 8 x D block (contiguous in HBM) ----------------------------------------------
    for (int idx = lane; idx < nf * D; idx += 32) {
      const int f = idx / D, m = idx - f * D;
      A.y[r0 * D + idx] = mcs[m * 8 + f];
    }
    __syncwarp();
  }
}

}  // namespace

template <int kMW, int kQ, int SOLVE>
static int launch_mcep_fast(const MArgs& A, int device, cudaStream_t stream) {
  const size_t smem = (static_cast<size_t>(kDM) * kKS + kKS * kJS + kKS * kPS + 32 +
                       static_cast<size_t>(kMW) * kWarpFloats) * sizeof(float);
  if (smem > static_cast<size_t>(max_dynamic_smem(device))) return DSB200_E_UNSUPPORTED;
  DSB_CUDA(cudaFuncSetAttribute(mcep_fast_kernel<kMW, kQ, SOLVE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  const int64_t n_oct = (A.rows + 7) / 8;
  const int blocks = static_cast<int>(std::min<int64_t>((n_oct + kMW - 1) / kMW, sm_count(device)));
  mcep_fast_kernel<kMW, kQ, SOLVE><<<blocks, kMW * 32, smem, stream>>>(A);
  return after_launch("mcep_fast_kernel");
}

int mcep_fast_try(const float* x, float* y, int64_t rows, const dsb200_mcep_params* p, const float* P0,
                  const float* G, const float* Hm, const float* av, int device, cudaStream_t stream) {
  if (p->fft_length != 512 || p->cep_order > kDM - 1) return DSB200_E_UNSUPPORTED;
  if (rows < 8) return DSB200_E_UNSUPPORTED;   // fewer rows than one octet: the generic kernel
  MArgs A{};
  A.x = x;
  A.y = y;
  A.P0 = P0;
  A.G = G;
  A.Hm = Hm;
  A.av = av;
  A.rows = rows;
  A.D = p->cep_order + 1;
  A.J = 2 * p->cep_order + 1;
  A.n_iter = p->n_iter;
  // knob MCEP_V = 8 | 120 | 12 | 16 | 122 | 162: warps per CTA (x frame pairs per elimination pass:
  // 4 at 8 warps, else 2); 8 and 120 keep the fully unrolled elimination + back substitution for A/B runs
  const int variant = knob("MCEP_V", kMcepVariant);
  A.stagger = knob("MCEP_STAGGER", kMcepStagger);
  if (variant == 8) return launch_mcep_fast<8, 4, 0>(A, device, stream);      // round-1 shape
  if (variant == 120) return launch_mcep_fast<12, 2, 0>(A, device, stream);   // 12 warps, unrolled elimination
  if (variant == 122) return launch_mcep_fast<12, 2, 2>(A, device, stream);   // four rows of one system per lane
  if (variant == 162) return launch_mcep_fast<16, 2, 2>(A, device, stream);
  if (variant == 16) return launch_mcep_fast<16, 2, 1>(A, device, stream);
  return launch_mcep_fast<12, 2, 1>(A, device, stream);                        // rolled Gauss-Jordan, lane = row, 12 warps
}

}  // namespace dsb200
