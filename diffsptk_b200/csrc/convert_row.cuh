// The per-row recursions of csrc/convert.cu, host- and device-callable so that tests/test_convert_host.py can run
// the very same code on the CPU against the oracle (no GPU needed to check the recursion logic).
#pragma once

#include "common.cuh"

namespace dsb200 {

// gnorm.py:101-112, ignorm.py:98-109, norm0.py:88-94: y_0 and the factor that multiplies x_1..x_M.
template <typename T>
__host__ __device__ __forceinline__ void row_scalar(T x0, int op, T g, T* y0, T* mult) {
  const T one = static_cast<T>(1);
  if (op == DSB200_CONV_GNORM) {
    if (g == static_cast<T>(0)) {
      *y0 = dexp(x0);
      *mult = one;
    } else {
      const T z = one + g * x0;
      *y0 = dpow(z, one / g);
      *mult = one / z;
    }
  } else if (op == DSB200_CONV_IGNORM) {
    if (g == static_cast<T>(0)) {
      *y0 = dlog(x0);
      *mult = one;
    } else {
      const T z = dpow(x0, g);
      *y0 = (z - one) / g;
      *mult = z;
    }
  } else {  // DSB200_CONV_NORM0
    *y0 = one / x0;
    *mult = one / x0;
  }
}

template <typename T>
__host__ __device__ __forceinline__ void convert_row(T* a, int D, int op, T g) {
  const int M = D - 1;
  switch (op) {
    case DSB200_CONV_LPC2PAR: {
      // a[1 + i] <- gamma a_i; for m = M-1..1: k_m = a[m]; a[i] <- (a[i] - k_m a[m-1-i]) / (1 - k_m^2), i < m
      T* c = a + 1;
      for (int i = 0; i < M; ++i) c[i] *= g;
      for (int m = M - 1; m >= 1; --m) {
        const T km = c[m];
        const T z = static_cast<T>(1) - km * km;
        for (int i = 0, j = m - 1; i <= j; ++i, --j) {   // true divisions: the recursion amplifies every rounding
          const T ci = c[i], cj = c[j];
          c[i] = (ci - km * cj) / z;
          if (j != i) c[j] = (cj - km * ci) / z;
        }
      }
      break;
    }
    case DSB200_CONV_PAR2LPC: {
      // a <- k / gamma (the gain too, as the reference does); for m = 2..M: a[1..m) += k_m flip(a[1..m))
      // k_m is the UNDIVIDED coefficient, so a[m] is divided only after it was used as k_m.
      a[0] = a[0] / g;
      if (M >= 1) a[1] = a[1] / g;
      for (int m = 2; m <= M; ++m) {
        const T km = a[m];
        for (int i = 1, j = m - 1; i <= j; ++i, --j) {
          const T ai = a[i], aj = a[j];
          a[i] = ai + km * aj;
          if (j != i) a[j] = aj + km * ai;
        }
        a[m] = km / g;
      }
      break;
    }
    default: {  // gnorm / ignorm / norm0: a new zeroth element and one multiplier for the rest of the row
      T mult;
      row_scalar<T>(a[0], op, g, &a[0], &mult);
      if (mult != static_cast<T>(1))
        for (int i = 1; i <= M; ++i) a[i] = a[i] * mult;
      break;
    }
  }
}

// The two O(M^2) recursions with the row in REGISTERS (compile-time length, fully unrolled): the shared-memory
// version above costs ~2 M^2 shared-memory accesses and M^2 / 2 divisions per row, 4-5x the HBM time of the row at
// M = 24.  The state is float64 for float32 rows too (like the Levinson kernels): the step-down recursion
// amplifies each rounding by 1 / (1 - k^2) per order, and with a float64 state one reciprocal per order can
// replace the per-element divisions without leaving the reference's own float32 error band.
template <typename T, int D>
__host__ __device__ __forceinline__ void convert_row_fixed(T* a, int op, T g_) {
  constexpr int M = D - 1;
  const double g = static_cast<double>(g_);
  double c[D];
#pragma unroll
  for (int i = 0; i < D; ++i) c[i] = static_cast<double>(a[i]);
  if (op == DSB200_CONV_LPC2PAR) {
#pragma unroll
    for (int i = 1; i <= M; ++i) c[i] *= g;
#pragma unroll
    for (int m = M - 1; m >= 1; --m) {
      const double km = c[1 + m];
      const double rz = 1.0 / (1.0 - km * km);
#pragma unroll
      for (int i = 0; i <= (m - 1) / 2; ++i) {
        const int j = m - 1 - i;
        const double ci = c[1 + i], cj = c[1 + j];
        c[1 + i] = (ci - km * cj) * rz;
        if (j != i) c[1 + j] = (cj - km * ci) * rz;
      }
    }
  } else {  // DSB200_CONV_PAR2LPC
    const double rg = 1.0 / g;
    c[0] *= rg;
    if (M >= 1) c[M >= 1 ? 1 : 0] *= rg;
#pragma unroll
    for (int m = 2; m <= M; ++m) {
      const double km = c[m];
#pragma unroll
      for (int i = 1; i <= m / 2; ++i) {
        const int j = m - i;
        const double ai = c[i], aj = c[j];
        c[i] = ai + km * aj;
        if (j != i) c[j] = aj + km * ai;
      }
      c[m] = km * rg;
    }
  }
#pragma unroll
  for (int i = 0; i < D; ++i) a[i] = static_cast<T>(c[i]);
}

}  // namespace dsb200
