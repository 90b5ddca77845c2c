// Per-row solve of a symmetric-Toeplitz-plus-Hankel system  (T + H) x = r,  T[i][j] = t[|i - j|],  H[i][j] = h[i + j]
// -- the Newton step of the (mel-)generalized cepstral analysis (diffsptk/modules/mgcep.py:219-222:
// symmetric_toeplitz + hankel + torch.linalg.solve; utils/private.py:291-302).  One warp per row: the augmented
// system lives in the warp's shared memory (odd row stride, conflict-free columns), Gaussian elimination without
// pivoting (the Newton matrix is symmetric positive definite) with the rows below the pivot spread over the lanes,
// then back substitution.  Same elimination as the mcep kernels (cepstral.cu), as a stand-alone entry point.
#include <algorithm>

#include "common.cuh"

namespace dsb200 {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) thsolve_kernel(const T* __restrict__ t, const T* __restrict__ h,
                                                      const T* __restrict__ r, T* __restrict__ x, int64_t rows,
                                                      int M, int ldm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int J = 2 * M - 1;
  T* base = reinterpret_cast<T*>(smem_raw) + static_cast<size_t>(warp) * (M + J + static_cast<size_t>(M) * ldm);
  T* ts = base;            // [M]
  T* hs = ts + M;          // [2M - 1]
  T* Aug = hs + J;         // [M][ldm], column M = right-hand side
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * wpb + warp; row < rows;
       row += static_cast<int64_t>(gridDim.x) * wpb) {
    for (int i = lane; i < M; i += 32) ts[i] = t[row * M + i];
    for (int i = lane; i < J; i += 32) hs[i] = h[row * J + i];
    __syncwarp();
    for (int idx = lane; idx < M * (M + 1); idx += 32) {
      const int i = idx / (M + 1), j = idx - i * (M + 1);
      Aug[i * ldm + j] = (j < M) ? ts[i > j ? i - j : j - i] + hs[i + j] : r[row * M + i];
    }
    __syncwarp();
    for (int pc = 0; pc < M - 1; ++pc) {
      const T piv = Aug[pc * ldm + pc];
      for (int i = pc + 1 + lane; i < M; i += 32) {
        const T f = Aug[i * ldm + pc] / piv;
        for (int c = pc + 1; c <= M; ++c) Aug[i * ldm + c] = dfma(-f, Aug[pc * ldm + c], Aug[i * ldm + c]);
      }
      __syncwarp();
    }
    for (int i = M - 1; i >= 0; --i) {
      T s = 0;
      for (int c = i + 1 + lane; c < M; c += 32) s = dfma(Aug[i * ldm + c], Aug[c * ldm + M], s);
      s = warp_sum(s);
      if (lane == 0) Aug[i * ldm + M] = (Aug[i * ldm + M] - s) / Aug[i * ldm + i];
      __syncwarp();
    }
    for (int i = lane; i < M; i += 32) x[row * M + i] = Aug[i * ldm + M];
    __syncwarp();
  }
}

template <typename T>
int thsolve_impl(const void* t, const void* h, const void* r, void* x, int64_t rows, int32_t M, int device,
                 void* stream) {
  DSB_REQUIRE(M >= 1, "order must be positive");
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(t != nullptr && h != nullptr && r != nullptr && x != nullptr, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  const int ldm = (M + 1) | 1;
  const size_t per_warp = (static_cast<size_t>(M) + 2 * M - 1 + static_cast<size_t>(M) * ldm) * sizeof(T);
  const size_t cap = static_cast<size_t>(max_dynamic_smem(device));
  int wpb = 8;
  while (wpb > 1 && wpb * per_warp > cap / 2) --wpb;
  if (wpb * per_warp > cap) return fail(DSB200_E_UNSUPPORTED, "order %d does not fit in shared memory", M);
  DSB_CUDA(cudaFuncSetAttribute(thsolve_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cap)));
  const int64_t need = (rows + wpb - 1) / wpb;
  const int blocks = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(sm_count(device)) * 4));
  thsolve_kernel<T><<<blocks, wpb * 32, wpb * per_warp, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const T*>(t), static_cast<const T*>(h), static_cast<const T*>(r), static_cast<T*>(x), rows, M, ldm);
  return after_launch("thsolve_kernel");
}

}  // namespace
}  // namespace dsb200

extern "C" {

int dsb200_thsolve_f32(const void* t, const void* h, const void* r, void* x, int64_t rows, int32_t order, int device,
                       void* stream) {
  return dsb200::thsolve_impl<float>(t, h, r, x, rows, order, device, stream);
}
int dsb200_thsolve_f64(const void* t, const void* h, const void* r, void* x, int64_t rows, int32_t order, int device,
                       void* stream) {
  return dsb200::thsolve_impl<double>(t, h, r, x, rows, order, device, stream);
}

}  // extern "C"
