// Warp-private staging of waveform spans with bulk async copies (sm_90+/sm_100a PTX).
//
// cp.async.bulk (SASS: UBLKCP) moves a contiguous, 16-byte aligned span global -> shared and signals
// an mbarrier with the byte count; the store direction (shared -> global) uses bulk groups.  One
// elected lane issues the copy, the whole warp waits on the warp's own mbarrier -- no CTA barrier.
#pragma once

#include "common.cuh"

namespace dsb200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Stage samples [s0, s0 + span) of one utterance (xb, T samples) into dst[0..span), applying the
// reference's padding (frame.py:130-137) outside [0, T).  Called by all 32 lanes of a warp.
// With `bulk_ok` (16-byte alignment of every span start and of every utterance) the in-range part is
// one bulk copy completing on `bar` and the out-of-range part is zero-filled (constant padding); other
// pad modes take the guarded-load path on the (rare) spans that touch an utterance end.
// Returns true when a bulk copy is in flight (the caller must mbar_wait before reading).
// With ALWAYS_ARRIVE the guarded path completes the barrier phase too (plain arrive by lane 0), so that the
// caller can wait unconditionally and derive the phase parity from its iteration count instead of carrying
// per-buffer phase bits and "was it a bulk copy" flags through its loop.
template <bool ALWAYS_ARRIVE = false>
__device__ __forceinline__ bool stage_span(const float* xb, int T, int s0, int span, int pad_mode, bool bulk_ok,
                                           float* dst, uint64_t* bar, int lane) {
  const int lo = s0 < 0 ? 0 : s0;
  const int hi = (s0 + span) > T ? T : (s0 + span);
  const bool interior = (lo == s0) && (hi == s0 + span);
  if (bulk_ok && (interior || pad_mode == DSB200_PAD_CONSTANT) && hi > lo) {
    for (int i = lane; i < lo - s0; i += 32) dst[i] = 0.0f;
    for (int i = hi - s0 + lane; i < span; i += 32) dst[i] = 0.0f;
    if (lane == 0) {
      fence_async_smem();
      const uint32_t bytes = static_cast<uint32_t>(hi - lo) * 4u;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(dst + (lo - s0), xb + lo, bytes, bar);
    }
    return true;
  }
  for (int i = lane; i < span; i += 32) {
    // spans may run past the last real frame, i.e. further out than any valid padding: clamp to zero
    const int64_t p = pad_index(static_cast<int64_t>(s0) + i, T, pad_mode);
    dst[i] = (p < 0 || p >= T) ? 0.0f : xb[p];
  }
  if (ALWAYS_ARRIVE && lane == 0) mbar_arrive(bar);
  return false;
}

// Out-of-line copy of stage_span<true> for the pipelines' hot loops: interior spans (all but the first and last
// few of an utterance) take a ten-instruction bulk-copy path inline, everything else -- zero fill, the four pad
// modes, unaligned waveforms -- lives behind one call instead of being inlined into the loop (round 1: the inlined
// general path was ~20 % of stft512's static code and most of its per-quad branch and integer overhead).
static __device__ __noinline__ void stage_span_slow(const float* xb, int T, int s0, int span, int pad_mode, bool bulk_ok,
                                             float* dst, uint64_t* bar, int lane) {
  stage_span<true>(xb, T, s0, span, pad_mode, bulk_ok, dst, bar, lane);
}

__device__ __forceinline__ void stage_span_fast(const float* xb, int T, int s0, int span, int pad_mode, bool bulk_ok,
                                                float* dst, uint64_t* bar, int lane) {
#ifdef DSB200_STAGE_INLINE   // A/B knob: the round-1 code shape (general path inlined into the loop)
  stage_span<true>(xb, T, s0, span, pad_mode, bulk_ok, dst, bar, lane);
  return;
#endif
  if (bulk_ok && s0 >= 0 && s0 + span <= T) {
    if (lane == 0) {
      fence_async_smem();
      const uint32_t bytes = static_cast<uint32_t>(span) * 4u;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(dst, xb + s0, bytes, bar);
    }
  } else {
    stage_span_slow(xb, T, s0, span, pad_mode, bulk_ok, dst, bar, lane);
  }
}

}  // namespace dsb200
