// Micro-benchmark: issue rate of packed FFMA2 on sm_100a as a function of operand pattern and warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bench_ffma2 tools/bench_ffma2.cu && /tmp/bench_ffma2
// Patterns (all: 16 independent accumulators, 64 multiply-adds per warp instruction for the packed forms):
//   0  acc[k] += a[i] * b[i+k]   -- the LPC lag loop: operand A is shared by 16 consecutive instructions (reuse cache)
//   1  acc[k] += a[k] * b[i+k]   -- three distinct register pairs per instruction, no reuse
//   2  acc[k] += a[i] * s[k]     -- packed x scalar broadcast (twiddle form of the FFT butterflies)
//   3  scalar FFMA, same arithmetic as 0 (two instructions per packed one)
//   4  acc[k] += acc2[k] * b[k]  -- FMUL2/FADD2-like two-operand traffic: d = a * b + d with a, b distinct per k
// Prints cycles per warp instruction per scheduler (SMSP) and multiply-adds per clock per SM.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

constexpr int kAcc = 16, kWin = 31, kIters = 128;   // 32 + 62 registers of live state: fits 128 per thread

template <int PAT, int W>
__global__ void __launch_bounds__(W * 32, 1) k(float2* out, const float2* in, long long* cyc) {
  float2 acc[kAcc], x[kWin];
  float s[kAcc];
#pragma unroll
  for (int i = 0; i < kWin; ++i) x[i] = in[(threadIdx.x + 37 * i) & 1023];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) {
    acc[i] = make_float2(0.f, 0.f);
    s[i] = in[(threadIdx.x + 11 * i) & 1023].x;
  }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kAcc; ++i) {
#pragma unroll
      for (int kk = 0; kk < kAcc; ++kk) {
        if (PAT == 0) acc[kk] = __ffma2_rn(x[i], x[i + kk], acc[kk]);
        if (PAT == 1) acc[kk] = __ffma2_rn(x[kk], x[(i + kk) % kWin], acc[kk]);
        if (PAT == 2) acc[kk] = __ffma2_rn(x[i], make_float2(s[kk], s[kk]), acc[kk]);
        if (PAT == 3) {
          acc[kk].x = fmaf(x[i].x, x[i + kk].x, acc[kk].x);
          acc[kk].y = fmaf(x[i].y, x[i + kk].y, acc[kk].y);
        }
        if (PAT == 4) acc[kk] = __ffma2_rn(x[15 + kk], x[kk], acc[kk]);
      }
    }
  }
  const long long t1 = clock64();
  float2 r = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < kAcc; ++i) { r.x += acc[i].x; r.y += acc[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int PAT, int W>
void run(float2* out, const float2* in, long long* cyc, int sms) {
  constexpr int warps = W;
  k<PAT, W><<<sms, warps * 32>>>(out, in, cyc);
  k<PAT, W><<<sms, warps * 32>>>(out, in, cyc);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < sms; ++i) mean += static_cast<double>(h[i]);
  mean /= sms;
  const double per_sched = (warps + 3) / 4;                        // warps on the busiest scheduler
  const double inst = static_cast<double>(kIters) * kAcc * kAcc * (PAT == 3 ? 2 : 1);
  const double cyc_per_inst = mean / (inst * per_sched);
  const double fma_per_clk_sm = static_cast<double>(kIters) * kAcc * kAcc * 64.0 * warps / mean;
  printf("{\"pattern\": %d, \"warps_per_sm\": %d, \"cycles\": %.0f, \"cycles_per_warp_inst_per_smsp\": %.3f, "
         "\"fma_per_clk_per_sm\": %.1f}\n", PAT, warps, mean, cyc_per_inst, fma_per_clk_sm);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float2 *out, *in;
  long long* cyc;
  cudaMalloc(&out, sizeof(float2) * 256 * 1024);
  cudaMalloc(&in, sizeof(float2) * 1024);
  cudaMalloc(&cyc, sizeof(long long) * 256);
  cudaMemset(in, 0, sizeof(float2) * 1024);
#define ALL(W) run<0, W>(out, in, cyc, sms); run<1, W>(out, in, cyc, sms); run<2, W>(out, in, cyc, sms); \
               run<3, W>(out, in, cyc, sms); run<4, W>(out, in, cyc, sms);
  ALL(4) ALL(8) ALL(12) ALL(16)
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
