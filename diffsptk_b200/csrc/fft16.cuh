// Frame-pair packed 16-point FFT building blocks shared by the fused forward (stft512.cu) and inverse
// (istft512.cu) kernels.  Every value is a float2 = (frame A, frame B), so butterflies issue as packed
// FADD2 / FMUL2 / FFMA2 with scalar-broadcast twiddles: the same 128 results per clock per SM as scalar FP32
// with half the issue slots (profiles/ubench_r1.txt).
#pragma once

#include "common.cuh"

namespace dsb200 {
namespace fft16_detail {

constexpr int kXRow = 17;                              // float2 units per transpose row (16 + 1 pad)
constexpr int kPlane = 16 * kXRow;                     // float2 units per plane
constexpr int kXchBytesPerWarp = 2 * 2 * kPlane * 8;   // 2 half-warps x (re, im) planes = 8704 B

// Hint the L2 to fetch [p, p + bytes) of a global tensor spanning [base, base + total): the range is shrunk to
// 16-byte boundaries inside the tensor (cp.async.bulk.prefetch.L2 needs aligned address and size).
__device__ __forceinline__ void prefetch_l2(const void* base, size_t total, const void* p, size_t bytes) {
  const uintptr_t lo0 = reinterpret_cast<uintptr_t>(base), hi0 = lo0 + total;
  uintptr_t lo = reinterpret_cast<uintptr_t>(p), hi = lo + bytes;
  if (lo < lo0) lo = lo0;
  if (hi > hi0) hi = hi0;
  lo = (lo + 15) & ~static_cast<uintptr_t>(15);
  hi &= ~static_cast<uintptr_t>(15);
  if (hi > lo)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"(static_cast<uint32_t>(hi - lo)) : "memory");
}

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
__device__ __forceinline__ C2 csub_i(C2 a, C2 b) { return {add2(a.re, b.im), sub2(a.im, b.re)}; }  // a - i b
__device__ __forceinline__ C2 cadd_i(C2 a, C2 b) { return {sub2(a.re, b.im), add2(a.im, b.re)}; }  // a + i b
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

// 16-point DFT of a[0..15] (inputs a[j], j >= NJ, structurally zero).  With j = 4 s + c and
// k = r + 4 t:  W16^(jk) = W4^(s r) W16^(c r) W4^(c t).  The result is left in the "digit-swapped"
// register order a[4 r + t] = A[r + 4 t]; callers index through dig().
__host__ __device__ constexpr int dig(int k) { return 4 * (k & 3) + (k >> 2); }


#ifdef DSB200_FFT16_LEGACY
// Round-1 version: the W16^(c r) twiddles as separate multiplies (96 packed operations in the second layer).
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
  radix4<false>(a[0], a[1], a[2], a[3]);      // r = 0 -> k = 0, 4, 8, 12
  radix4<false>(a[4], a[5], a[6], a[7]);      // r = 1 -> k = 1, 5, 9, 13
  radix4<false>(a[8], a[9], a[10], a[11]);    // r = 2 -> k = 2, 6, 10, 14
  radix4<false>(a[12], a[13], a[14], a[15]);  // r = 3 -> k = 3, 7, 11, 15
}
#else
constexpr float kT8 = 0.41421356237309504880f;  // tan(pi/8)

// out = (x + s y, x - s y) with a real scalar s: two packed multiply-adds per complex value
__device__ __forceinline__ void pm_s(C2 x, C2 y, float s, C2& plus, C2& minus) {
  plus = {fma2s(y.re, s, x.re), fma2s(y.im, s, x.im)};
  minus = {fma2s(y.re, -s, x.re), fma2s(y.im, -s, x.im)};
}
// out = (x - i s y, x + i s y)
__device__ __forceinline__ void pm_is(C2 x, C2 y, float s, C2& minus_i, C2& plus_i) {
  minus_i = {fma2s(y.im, s, x.re), fma2s(y.re, -s, x.im)};
  plus_i = {fma2s(y.im, -s, x.re), fma2s(y.re, s, x.im)};
}

// Second layer for one residue r: X_t = sum_c W4^(c t) W16^(c r) b_c, with the W16 twiddles folded into the
// butterflies' multiply-adds (round 2): a twiddle cos(1 - i tan) costs two multiply-adds for the rotation-by-tan
// and its cosine rides on the following +/- as a multiply-add; R (1 -+ i) costs two additions and its R rides
// likewise.  80 packed operations for the four residues instead of 96, none of them a bare multiply.
template <int R>
__device__ __forceinline__ void layer2(C2& x0, C2& x1, C2& x2, C2& x3) {
  if constexpr (R == 0) {
    radix4<false>(x0, x1, x2, x3);
  } else if constexpr (R == 2) {
    // W16^2 = R (1 - i), W16^4 = -i, W16^6 = -R (1 + i)
    const C2 t0 = csub_i(x0, x2), t1 = cadd_i(x0, x2);          // x0 -+ i b2
    const C2 s = cadd(x1, x3), d = csub(x1, x3);
    const C2 p = csub_i(d, s);                                   // (1 - i) b1 - (1 + i) b3 = d - i s
    const C2 m = csub_i(s, d);                                   // (1 - i) b1 + (1 + i) b3 = s - i d
    pm_s(t0, p, kR, x0, x2);
    pm_is(t1, m, kR, x1, x3);
  } else {
    // R == 1: W16^1 = c (1 - i t), W16^2 = R (1 - i), W16^3 = c (t - i)
    // R == 3: W16^3 = c (t - i),  W16^6 = -R (1 + i), W16^9 = -c (1 - i t)       (c = cos pi/8, t = tan pi/8)
    C2 t0, t1, u, v;
    if constexpr (R == 1) {
      const C2 q = {add2(x2.re, x2.im), sub2(x2.im, x2.re)};     // (1 - i) b2
      pm_s(x0, q, kR, t0, t1);
      u = {fma2s(x1.im, kT8, x1.re), fma2s(x1.re, -kT8, x1.im)};             // (1 - i t) b1
      v = {fma2s(x3.re, kT8, x3.im), fma2s(x3.im, kT8, make_float2(-x3.re.x, -x3.re.y))};   // (t - i) b3
      const C2 p = cadd(u, v), m = csub(u, v);
      pm_s(t0, p, kC8, x0, x2);
      pm_is(t1, m, kC8, x1, x3);
    } else {
      const C2 q = {sub2(x2.re, x2.im), add2(x2.im, x2.re)};     // (1 + i) b2
      pm_s(x0, q, kR, t1, t0);                                   // t0 = x0 - R q, t1 = x0 + R q
      u = {fma2s(x1.re, kT8, x1.im), fma2s(x1.im, kT8, make_float2(-x1.re.x, -x1.re.y))};   // (t - i) b1
      v = {fma2s(x3.im, kT8, x3.re), fma2s(x3.re, -kT8, x3.im)};             // (1 - i t) b3
      const C2 p = csub(u, v), m = cadd(u, v);
      pm_s(t0, p, kC8, x0, x2);
      pm_is(t1, m, kC8, x1, x3);
    }
  }
}

template <int NJ>
__device__ __forceinline__ void fft16(C2 (&a)[16]) {
  radix4<(12 >= NJ)>(a[0], a[4], a[8], a[12]);
  radix4<(13 >= NJ)>(a[1], a[5], a[9], a[13]);
  radix4<(14 >= NJ)>(a[2], a[6], a[10], a[14]);
  radix4<(15 >= NJ)>(a[3], a[7], a[11], a[15]);
  // a[c + 4 r] now holds b_c[r]
  layer2<0>(a[0], a[1], a[2], a[3]);      // r = 0 -> k = 0, 4, 8, 12
  layer2<1>(a[4], a[5], a[6], a[7]);      // r = 1 -> k = 1, 5, 9, 13
  layer2<2>(a[8], a[9], a[10], a[11]);    // r = 2 -> k = 2, 6, 10, 14
  layer2<3>(a[12], a[13], a[14], a[15]);  // r = 3 -> k = 3, 7, 11, 15
}
#endif

}  // namespace fft16_detail
}  // namespace dsb200
