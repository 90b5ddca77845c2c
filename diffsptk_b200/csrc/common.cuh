// Shared device/host helpers for the diffsptk_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "diffsptk_b200.h"

namespace dsb200 {

// ---- host-side error plumbing (api_util.cu) --------------------------------------------
int fail(int code, const char* fmt, ...);          // records the message, returns `code`
int cuda_fail(cudaError_t e, const char* what);    // DSB200_E_CUDA with cudaGetErrorString
void count_launch(int n = 1);
void note_kernel(const char* name);               // name of the kernel this thread launched last (string literal)

#define DSB_CUDA(call)                                            \
  do {                                                            \
    cudaError_t e__ = (call);                                     \
    if (e__ != cudaSuccess) return ::dsb200::cuda_fail(e__, #call); \
  } while (0)

#define DSB_REQUIRE(cond, ...)                                                   \
  do {                                                                           \
    if (!(cond)) return ::dsb200::fail(DSB200_E_BAD_PARAM, __VA_ARGS__);         \
  } while (0)

// Checks the launch that was just issued (cudaPeekAtLastError keeps sticky errors visible).
static inline int after_launch(const char* what) {
  count_launch();
  note_kernel(what);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  return DSB200_OK;
}

struct DeviceScope {  // selects `device` for the duration of one ABI call
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceScope(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
  }
  ~DeviceScope() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

int sm_count(int device);               // cached multiProcessorCount
int max_dynamic_smem(int device);       // cached sharedMemPerBlockOptin
int knob(const char* name, int dflt);   // tuning knob: dsb200_set_knob > environment DSB200_<name> > dflt

// Device twiddle table W_n^k = exp(-2*pi*i*k/n), k = 0..n-1, interleaved (re, im), built in
// float64 with sincospi and cached per (device, n, dtype).  Returns nullptr on failure.
const void* twiddle_table(int device, int n, bool is_f64, cudaStream_t stream);

// ---- element-type traits -------------------------------------------------------------
template <typename T> struct Cx;
template <> struct Cx<float> { using type = float2; };
template <> struct Cx<double> { using type = double2; };
template <typename T> using cx_t = typename Cx<T>::type;

template <typename T> __host__ __device__ __forceinline__ cx_t<T> mk(T a, T b) {
  cx_t<T> r; r.x = a; r.y = b; return r;
}
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
  C r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}

__host__ __device__ __forceinline__ float dsqrt(float x) { return sqrtf(x); }
__host__ __device__ __forceinline__ double dsqrt(double x) { return sqrt(x); }
__host__ __device__ __forceinline__ float dlog(float x) { return logf(x); }
__host__ __device__ __forceinline__ double dlog(double x) { return log(x); }
__host__ __device__ __forceinline__ float dlog10(float x) { return log10f(x); }
__host__ __device__ __forceinline__ double dlog10(double x) { return log10(x); }
__host__ __device__ __forceinline__ float dexp(float x) { return expf(x); }
__host__ __device__ __forceinline__ double dexp(double x) { return exp(x); }
__host__ __device__ __forceinline__ float dpow(float x, float y) { return powf(x, y); }
__host__ __device__ __forceinline__ double dpow(double x, double y) { return pow(x, y); }
__host__ __device__ __forceinline__ float dmax(float a, float b) { return fmaxf(a, b); }
__host__ __device__ __forceinline__ double dmax(double a, double b) { return fmax(a, b); }
__host__ __device__ __forceinline__ float dfma(float a, float b, float c) { return fmaf(a, b, c); }
__host__ __device__ __forceinline__ double dfma(double a, double b, double c) { return fma(a, b, c); }

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = dmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Source index of padded position p (may be < 0 or >= T); returns -1 for "zero".
// torch F.pad semantics used by diffsptk/modules/frame.py:134-137.
__device__ __forceinline__ int64_t pad_index(int64_t p, int64_t T, int mode) {
  if (p >= 0 && p < T) return p;
  switch (mode) {
    case DSB200_PAD_REFLECT:
      return p < 0 ? -p : 2 * (T - 1) - p;
    case DSB200_PAD_REPLICATE:
      return p < 0 ? 0 : T - 1;
    case DSB200_PAD_CIRCULAR:
      return p < 0 ? p + T : p - T;
    default:
      return -1;
  }
}

// Persistent kernels whose warps all run the same alternation of phases (a multiply-add phase, then a latency-bound
// one) start in lockstep and stay there: the warps that share a scheduler then fight over the FP32 pipe in one phase
// and leave it idle in the other.  Delaying the k-th warp of a scheduler (warp index / 4) by k * cycles at start-up
// offsets the phases once; nothing re-synchronises them afterwards (no CTA barrier in the main loops).
__device__ __forceinline__ void stagger_start(int cycles) {
  const int slot = static_cast<int>(threadIdx.x >> 7);
  if (cycles > 0 && slot > 0) {
    const long long t0 = clock64(), span = static_cast<long long>(cycles) * slot;
    while (clock64() - t0 < span) __nanosleep(64);
  }
}

static inline bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

}  // namespace dsb200
