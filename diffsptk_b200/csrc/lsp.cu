// LPC -> line spectral pairs.  Reference: diffsptk/modules/lpc2lsp.py:159-197 builds
// P(z) = A(z) - z^-(M+1) A(1/z) and Q(z) = A(z) + z^-(M+1) A(1/z), removes their trivial roots at z = +-1
// (deconv1d) and takes the angles of the remaining roots from the EIGENVALUES of two companion matrices
// (root_pol.py:130-146).  All those roots lie on the unit circle, so here they are found where they are: after the
// deflation both polynomials are symmetric of even degree 2n and on z = e^{jw}
//     z^-n C(z) = c_n + 2 sum_{k=1..n} c_{n-k} cos(k w)  =  a Chebyshev series in x = cos w,
// a real function of one variable with n simple zeros in (0, pi).
//
// Mapping: ONE WARP per row, float64 throughout (the zeros of an order-24 polynomial are too ill-conditioned for
// float32 evaluation).  The lanes evaluate the series on a grid of w (Clenshaw recurrence; the abscissae cos(w_i)
// are a per-CTA table, one evaluation per grid point, signs exchanged by warp ballots), mark the sign changes,
// then one lane per bracket refines in x (bisection + Illinois regula falsi, fixed trip count) -- and
// the M angles are rank-sorted into the output row.  A row whose brackets do not add up to n on the first grid is
// searched again on a 16x finer one; zeros still missing after that (a double zero: an unstable or degenerate
// filter) are reported as NaN.
#include <algorithm>

#include "common.cuh"

namespace dsb200 {
namespace {

// f(x) = g[0] + sum_{k=1..n} g[k] T_k(x)
__device__ __forceinline__ double cheb_eval(const double* g, int n, double x) {
  double b1 = 0.0, b2 = 0.0;
  const double x2 = 2.0 * x;
  for (int k = n; k >= 1; --k) {
    const double b0 = fma(x2, b1, g[k] - b2);
    b2 = b1;
    b1 = b0;
  }
  return fma(x, b1, g[0] - b2);
}

template <typename T>
__global__ void __launch_bounds__(256) lpc2lsp_kernel(const T* __restrict__ a, T* __restrict__ w, int64_t rows, int M,
                                                      int log_gain, double scale, int G, int NC, int per_warp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double* xg = reinterpret_cast<double*>(smem_raw);     // [G + 1] cos(pi i / G): the same grid for every row
  for (int i = threadIdx.x; i <= G; i += blockDim.x) xg[i] = cospi(static_cast<double>(i) / G);
  __syncthreads();
  double* base = xg + (G + 2) + static_cast<size_t>(warp) * per_warp;
  double* pq = base;                 // [2][M + 2] p then q (raw, then deflated in place)
  double* gc = pq + 2 * (M + 2);     // [2][NC]    Chebyshev coefficients of the two deflated polynomials
  double* lo = gc + 2 * NC;          // [M]        bracket ends (x = cos w, descending in w) / refined angles
  double* hi = lo + M;               // [M]
  int* cnt = reinterpret_cast<int*>(hi + M);   // [2]
  const int nP = (M % 2 == 0) ? M / 2 : (M - 1) / 2, nQ = (M % 2 == 0) ? M / 2 : (M + 1) / 2;

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + warp; row < rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    const T* ar = a + row * (M + 1);
    T* wr = w + row * (M + 1);
    if (lane == 0) {
      const T K = ar[0];
      wr[0] = log_gain ? dlog(K) : K;
    }
    if (M == 0) continue;
    // ---- p = a1 - flip(a1), q = a1 + flip(a1), a1 = [1, a_1..a_M, 0] ----------------------------------------
    for (int i = lane; i <= M + 1; i += 32) {
      const double u = (i == 0) ? 1.0 : (i <= M ? static_cast<double>(ar[i]) : 0.0);
      const int j = M + 1 - i;
      const double v = (j == 0) ? 1.0 : (j <= M ? static_cast<double>(ar[j]) : 0.0);
      pq[i] = u - v;
      pq[M + 2 + i] = u + v;
    }
    __syncwarp();
    // ---- remove the roots at z = 1 / z = -1 (polynomial division, sequential; two lanes, one polynomial each) --
    if (lane == 0) {
      double* p = pq;
      if (M % 2 == 0) { for (int i = 1; i <= M; ++i) p[i] += p[i - 1]; }        // / (1 - z^-1)
      else { for (int i = 2; i <= M - 1; ++i) p[i] += p[i - 2]; }               // / (1 - z^-2)
      for (int k = 0; k <= nP; ++k) gc[k] = (k == 0 ? 1.0 : 2.0) * p[nP - k];
    } else if (lane == 1) {
      double* q = pq + M + 2;
      if (M % 2 == 0) { for (int i = 1; i <= M; ++i) q[i] -= q[i - 1]; }        // / (1 + z^-1)
      for (int k = 0; k <= nQ; ++k) gc[NC + k] = (k == 0 ? 1.0 : 2.0) * q[nQ - k];
    }
    __syncwarp();
    // ---- zeros of both series ---------------------------------------------------------------------------------
    bool failed = false;
    for (int which = 0; which < 2; ++which) {
      const int n = which ? nQ : nP, off = which ? nP : 0;
      const double* g = gc + which * NC;
      if (n == 0) continue;
      int found = 0;
      for (int fine = 1; fine <= 16; fine *= 16) {      // second round on a 16x finer grid if needed
        const int Gn = G * fine;
        if (lane == 0) cnt[0] = 0;
        __syncwarp();
        auto grid_x = [&](int i) { return fine == 1 ? xg[i] : cospi(static_cast<double>(i) / Gn); };
        unsigned carry = 0;
        // 32 grid points per round, one evaluation each; the signs are exchanged by a ballot.  Lane l < 31 owns the
        // interval (base + l, base + l + 1); the interval that straddles two rounds belongs to lane 31 of the later.
        for (int base = 0; base <= Gn; base += 32) {
          const int i = base + lane;
          const bool valid = i <= Gn;
          const bool neg = valid && cheb_eval(g, n, grid_x(valid ? i : Gn)) < 0.0;
          const unsigned vbits = __ballot_sync(0xffffffffu, valid), bits = __ballot_sync(0xffffffffu, neg);
          bool change;
          int idx;
          if (lane < 31) {
            idx = i;
            change = ((vbits >> (lane + 1)) & 1u) && (((bits >> lane) ^ (bits >> (lane + 1))) & 1u);
          } else {
            idx = base - 1;
            change = base > 0 && ((carry ^ bits) & 1u);
          }
          if (change) {
            const int slot = atomicAdd(&cnt[0], 1);
            if (slot < n) { lo[off + slot] = grid_x(idx); hi[off + slot] = grid_x(idx + 1); }
          }
          carry = bits >> 31;
        }
        __syncwarp();
        found = cnt[0];
        __syncwarp();
        if (found == n) break;
      }
      if (found != n) failed = true;
      const int m = found < n ? found : n;
      for (int r = lane; r < n; r += 32) {
        if (r < m) {
          // f changes sign in [xb, xa] (xa > xb: w ascending).  4 bisections, then 12 regula-falsi steps with the
          // Illinois damping (bracket kept, superlinear): 18 evaluations instead of the 55 of a plain bisection
          // to machine precision, still a fixed trip count (tests/kernel_models.py: lsp_model, 5e-12 vs the
          // reference's eigenvalues).
          double xa = lo[off + r], xb = hi[off + r];
          double fa = cheb_eval(g, n, xa), fb = cheb_eval(g, n, xb);
          for (int it = 0; it < 4; ++it) {
            const double xm = 0.5 * (xa + xb), fm = cheb_eval(g, n, xm);
            if ((fm < 0.0) == (fa < 0.0)) { xa = xm; fa = fm; } else { xb = xm; fb = fm; }
          }
          for (int it = 0; it < 12; ++it) {
            const double d = fb - fa;
            if (d == 0.0) break;
            const double xc = (xa * fb - xb * fa) / d, fc = cheb_eval(g, n, xc);
            if ((fc < 0.0) != (fb < 0.0)) { xa = xb; fa = fb; } else { fa *= 0.5; }
            xb = xc;
            fb = fc;
          }
          lo[off + r] = acos(xb);
        } else {
          lo[off + r] = nan("");
        }
      }
      __syncwarp();
    }
    // ---- rank sort of the M angles (NaN last), scale, store ----------------------------------------------------
    for (int r = lane; r < M; r += 32) {
      const double v = lo[r];
      const double kv = (v == v) ? v : 1e300;                  // NaN (zero not found) sorts last
      int rank = 0;
      for (int j = 0; j < M; ++j) {
        const double u = lo[j];
        const double ku = (u == u) ? u : 1e300;
        rank += (ku < kv || (ku == kv && j < r)) ? 1 : 0;      // ties by index: a permutation in every case
      }
      wr[1 + rank] = static_cast<T>(v * scale);
    }
    (void)failed;
    __syncwarp();
  }
}

template <typename T>
int lpc2lsp_impl(const void* a, void* w, int64_t rows, int32_t M, int32_t log_gain, double scale, int device,
                 void* stream) {
  DSB_REQUIRE(M >= 0, "lpc_order must be non-negative.");
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(a != nullptr && w != nullptr, "NULL data pointer");
  if (M > 256) return fail(DSB200_E_UNSUPPORTED, "lpc_order > 256 is not implemented");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  const int NC = M / 2 + 2;
  const int per_warp = 2 * (M + 2) + 2 * NC + 2 * std::max(M, 1) + 2;   // doubles
  const int wpb = 8;
  const int G = std::min(2048, std::max(256, 32 * M));   // grid intervals on (0, pi): ~ 64 per expected zero
  const size_t smem = (static_cast<size_t>(G) + 2 + static_cast<size_t>(wpb) * per_warp) * sizeof(double);
  if (smem > static_cast<size_t>(max_dynamic_smem(device)))
    return fail(DSB200_E_UNSUPPORTED, "lpc_order %d does not fit in shared memory", M);
  DSB_CUDA(cudaFuncSetAttribute(lpc2lsp_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t need = (rows + wpb - 1) / wpb;
  const int blocks = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(sm_count(device)) * 8));
  lpc2lsp_kernel<T><<<blocks, wpb * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const T*>(a), static_cast<T*>(w), rows, M, log_gain, scale, G, NC, per_warp);
  return after_launch("lpc2lsp_kernel");
}

}  // namespace
}  // namespace dsb200

extern "C" {

int dsb200_lpc2lsp_f32(const void* a, void* w, int64_t rows, int32_t lpc_order, int32_t log_gain, double scale,
                       int device, void* stream) {
  return dsb200::lpc2lsp_impl<float>(a, w, rows, lpc_order, log_gain, scale, device, stream);
}
int dsb200_lpc2lsp_f64(const void* a, void* w, int64_t rows, int32_t lpc_order, int32_t log_gain, double scale,
                       int device, void* stream) {
  return dsb200::lpc2lsp_impl<double>(a, w, rows, lpc_order, log_gain, scale, device, stream);
}

}  // extern "C"
