// Pipe-throughput micro-benchmarks for sm_100a (B200).
//
// Purpose: the fused frame+window+rFFT kernel sits at the HBM/ALU ridge
// (SURVEY.md §7 "hard parts"), so its design depends on measured issue rates of
// scalar FP32, packed FP32x2 (FFMA2/FADD2/FMUL2, new on sm_100), shared-memory
// loads, shuffles, MUFU and FP64.  Each kernel runs a dependent-chain loop with
// enough independent chains to cover latency and reports lane-ops / clk / SM,
// using clock64() inside the kernel (so DVFS does not matter).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
// Run  : ./ubench > gpurun_out/ubench.txt
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
constexpr int NCH = 16;  // independent chains per thread

__global__ void k_ffma(float* out, long long* cyc, float b, float c) {
  float a[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) a[i] = threadIdx.x * 1e-3f + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) a[i] = fmaf(a[i], b, c);
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_ffma2(float* out, long long* cyc, float b, float c) {
  float2 a[NCH / 2];
  float2 bb = make_float2(b, b * 1.0001f), cc = make_float2(c, c * 0.999f);
#pragma unroll
  for (int i = 0; i < NCH / 2; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH / 2; ++i) a[i] = __ffma2_rn(a[i], bb, cc);
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH / 2; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_fadd(float* out, long long* cyc, float b, float c) {
  float a[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) a[i] = threadIdx.x * 1e-3f + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) a[i] = a[i] + b;
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_fadd2(float* out, long long* cyc, float b, float c) {
  float2 a[NCH / 2];
  float2 bb = make_float2(b, c);
#pragma unroll
  for (int i = 0; i < NCH / 2; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH / 2; ++i) a[i] = __fadd2_rn(a[i], bb);
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH / 2; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Mixed: 8 scalar FFMA + 8 FADD (alternating pipes?) per iteration.
__global__ void k_ffma_fadd(float* out, long long* cyc, float b, float c) {
  float a[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) a[i] = threadIdx.x * 1e-3f + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; i += 2) { a[i] = fmaf(a[i], b, c); a[i + 1] = a[i + 1] + c; }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Mixed FFMA2 + integer ALU (IADD3/LOP3) to see whether they co-issue.
__global__ void k_ffma2_alu(float* out, long long* cyc, float b, float c) {
  float2 a[NCH / 2];
  unsigned u[NCH / 2];
  float2 bb = make_float2(b, b * 1.0001f), cc = make_float2(c, c * 0.999f);
#pragma unroll
  for (int i = 0; i < NCH / 2; ++i) { a[i] = make_float2(threadIdx.x * 1e-3f + i, i); u[i] = threadIdx.x + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH / 2; ++i) { a[i] = __ffma2_rn(a[i], bb, cc); u[i] = (u[i] ^ 0x5bd1e995u) + it; }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH / 2; ++i) s += a[i].x + a[i].y + u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int W>  // W = words per load: 1, 2, 4
__global__ void k_lds(float* out, long long* cyc) {
  __shared__ __align__(16) float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int base = (threadIdx.x % 32) * W + (threadIdx.x / 32) * 128;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int idx = (base + j * 32 * W + (it & 7) * 4) & 4095 & ~(W - 1);
      if (W == 1) s[j] += sm[idx];
      if (W == 2) { float2 v = *reinterpret_cast<float2*>(&sm[idx]); s[j] += v.x + v.y; }
      if (W == 4) { float4 v = *reinterpret_cast<float4*>(&sm[idx]); s[j] += v.x + v.y + v.z + v.w; }
    }
  }
  long long t1 = clock64();
  float t = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) t += s[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_shfl(float* out, long long* cyc) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 3));
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>  // 0 sqrt.approx, 1 ex2.approx, 2 lg2.approx, 3 rsqrt.approx, 4 rcp.approx
__global__ void k_mufu(float* out, long long* cyc) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0f + threadIdx.x * 1e-3f + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_dfma(float* out, long long* cyc, double b, double c) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Mixed FFMA2 + LDS.64 (the FFT inner-loop mix): 8 FFMA2 + 2 LDS.64 per iteration.
__global__ void k_ffma2_lds(float* out, long long* cyc, float b, float c) {
  __shared__ __align__(16) float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-6f;
  float2 a[NCH / 2];
  float2 bb = make_float2(b, b * 1.0001f);
#pragma unroll
  for (int i = 0; i < NCH / 2; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i);
  __syncthreads();
  int base = (threadIdx.x % 32) * 2 + (threadIdx.x / 32) * 128;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    float2 v0 = *reinterpret_cast<float2*>(&sm[(base + (it & 15) * 64) & 4094]);
    float2 v1 = *reinterpret_cast<float2*>(&sm[(base + 2048 + (it & 15) * 64) & 4094]);
#pragma unroll
    for (int i = 0; i < NCH / 2; ++i) a[i] = __ffma2_rn(a[i], bb, (i & 1) ? v1 : v0);
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH / 2; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
void run(const char* name, F launch, int blocks_per_sm, int threads, double lane_ops_per_thread_iter, int nsm) {
  int nblk = nsm * blocks_per_sm;
  float* out; long long* cyc;
  CK(cudaMalloc(&out, sizeof(float) * nblk * threads));
  CK(cudaMalloc(&cyc, sizeof(long long) * nblk));
  launch(nblk, threads, out, cyc);  // warm-up
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  launch(nblk, threads, out, cyc);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = (long long*)malloc(sizeof(long long) * nblk);
  CK(cudaMemcpy(h, cyc, sizeof(long long) * nblk, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < nblk; ++i) avg += h[i]; avg /= nblk;
  double ops_per_sm = (double)blocks_per_sm * threads * ITERS * lane_ops_per_thread_iter;
  printf("%-14s blocks/SM=%d thr=%4d  cycles=%9.0f  lane-ops/clk/SM=%7.2f  warp-instr/clk/SM=%5.2f  time=%.3f ms\n",
         name, blocks_per_sm, threads, avg, ops_per_sm / avg, ops_per_sm / avg / 32.0, ms);
  free(h); cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("device %s  SMs=%d  clockRate=%d kHz  smem/SM=%zu\n", p.name, nsm, p.clockRate, p.sharedMemPerMultiprocessor);
  for (int thr : {256, 512, 1024}) {
    int bps = 1024 / thr;  // keep 32 warps per SM
    run("ffma", [&](int g, int t, float* o, long long* c) { k_ffma<<<g, t>>>(o, c, 1.0001f, 0.5f); }, bps, thr, NCH, nsm);
    run("ffma2(x2)", [&](int g, int t, float* o, long long* c) { k_ffma2<<<g, t>>>(o, c, 1.0001f, 0.5f); }, bps, thr, NCH, nsm);
    run("fadd", [&](int g, int t, float* o, long long* c) { k_fadd<<<g, t>>>(o, c, 1.0001f, 0.5f); }, bps, thr, NCH, nsm);
    run("fadd2(x2)", [&](int g, int t, float* o, long long* c) { k_fadd2<<<g, t>>>(o, c, 1.0001f, 0.5f); }, bps, thr, NCH, nsm);
    run("ffma+fadd", [&](int g, int t, float* o, long long* c) { k_ffma_fadd<<<g, t>>>(o, c, 1.0001f, 0.5f); }, bps, thr, NCH, nsm);
    run("ffma2+alu", [&](int g, int t, float* o, long long* c) { k_ffma2_alu<<<g, t>>>(o, c, 1.0001f, 0.5f); }, bps, thr, NCH + 2 * (NCH / 2), nsm);
    run("ffma2+lds64", [&](int g, int t, float* o, long long* c) { k_ffma2_lds<<<g, t>>>(o, c, 1.0001f, 0.5f); }, bps, thr, NCH + 2, nsm);
    run("lds32", [&](int g, int t, float* o, long long* c) { k_lds<1><<<g, t>>>(o, c); }, bps, thr, 8, nsm);
    run("lds64", [&](int g, int t, float* o, long long* c) { k_lds<2><<<g, t>>>(o, c); }, bps, thr, 8, nsm);
    run("lds128", [&](int g, int t, float* o, long long* c) { k_lds<4><<<g, t>>>(o, c); }, bps, thr, 8, nsm);
    run("shfl", [&](int g, int t, float* o, long long* c) { k_shfl<<<g, t>>>(o, c); }, bps, thr, 8, nsm);
    run("mufu.sqrt", [&](int g, int t, float* o, long long* c) { k_mufu<0><<<g, t>>>(o, c); }, bps, thr, 8, nsm);
    run("mufu.ex2", [&](int g, int t, float* o, long long* c) { k_mufu<1><<<g, t>>>(o, c); }, bps, thr, 8, nsm);
    run("mufu.lg2", [&](int g, int t, float* o, long long* c) { k_mufu<2><<<g, t>>>(o, c); }, bps, thr, 8, nsm);
    run("mufu.rsqrt", [&](int g, int t, float* o, long long* c) { k_mufu<3><<<g, t>>>(o, c); }, bps, thr, 8, nsm);
    run("dfma", [&](int g, int t, float* o, long long* c) { k_dfma<<<g, t>>>(o, c, 1.0001, 0.5); }, bps, thr, 8, nsm);
  }
  return 0;
}
