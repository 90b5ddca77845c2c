// Per-frame coefficient converters on LPC / cepstral rows: one launch reads every
// row once and writes it once -- HBM-bound at 2 * dim * sizeof(T) bytes per row.
//
//   DSB200_CONV_LPC2PAR  lpc2par.py:104-120   [K, a_1..a_M] -> [K, k_1..k_M]   (step-down recursion, gamma)
//   DSB200_CONV_PAR2LPC  par2lpc.py:100-107   [K, k_1..k_M] -> [K, a_1..a_M] / gamma (step-up recursion)
//   DSB200_CONV_GNORM    gnorm.py:101-112     gain normalisation of a generalized cepstrum (gamma)
//   DSB200_CONV_IGNORM   ignorm.py:98-109     its inverse
//   DSB200_CONV_NORM0    norm0.py:88-94       [K, a_1..a_M] -> [1/K, a_1/K..a_M/K]
//
// Mapping: a CTA stages a tile of 256 rows in shared memory with coalesced loads (row pitch odd, so that the
// per-thread row walks below are bank-conflict free), ONE THREAD owns one row and runs the recursion in place --
// the recursions are sequential in the order m and only O(M^2) flops on M + 1 values, far below the HBM time of
// the row -- then the tile leaves with coalesced stores.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "convert_row.cuh"

namespace dsb200 {
namespace {

constexpr int kThreads = 256;

// DF > 0: the row length is the compile-time DF and the O(M^2) recursions run in registers.
template <typename T, int DF>
__global__ void __launch_bounds__(kThreads) rowconv_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows,
                                                           int D, int pitch, int op, T g, int RB) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);   // [RB][pitch], RB <= kThreads rows per tile
  T* scale = tile + static_cast<size_t>(RB) * pitch;   // [RB] multiplier of x_1..x_M (gnorm / ignorm / norm0)
  const bool scalar_op = op == DSB200_CONV_GNORM || op == DSB200_CONV_IGNORM || op == DSB200_CONV_NORM0;
  const int64_t n_tiles = (rows + RB - 1) / RB;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t base = t * RB;
    const int nr = static_cast<int>(rows - base < RB ? rows - base : RB);
    const T* src = x + base * D;
    T* dst = y + base * D;
    // coalesced load: element e of the tile -> (row e / D, column e % D), walked incrementally
    {
      int r = threadIdx.x / D, c = threadIdx.x - r * D;
      const int dr = kThreads / D, dc = kThreads - dr * D;
      for (int e = threadIdx.x; e < nr * D; e += kThreads) {
        tile[r * pitch + c] = src[e];
        r += dr;
        c += dc;
        if (c >= D) { c -= D; ++r; }
      }
    }
    __syncthreads();
    if (threadIdx.x < nr) {
      T* row = tile + threadIdx.x * pitch;
      if (scalar_op) row_scalar<T>(row[0], op, g, &row[0], &scale[threadIdx.x]);   // the row is scaled on the way out
      else if (DF > 0) convert_row_fixed<T, (DF > 0 ? DF : 1)>(row, op, g);
      else convert_row<T>(row, D, op, g);
    }
    __syncthreads();
    {
      int r = threadIdx.x / D, c = threadIdx.x - r * D;
      const int dr = kThreads / D, dc = kThreads - dr * D;
      for (int e = threadIdx.x; e < nr * D; e += kThreads) {
        T v = tile[r * pitch + c];
        if (scalar_op && c > 0) v *= scale[r];
        dst[e] = v;
        r += dr;
        c += dc;
        if (c >= D) { c -= D; ++r; }
      }
    }
    __syncthreads();
  }
}

template <typename T>
int rowconv_impl(const void* x, void* y, int64_t rows, int32_t dim, int32_t op, double param, int device, void* stream) {
  DSB_REQUIRE(dim >= 1, "dim must be positive");
  DSB_REQUIRE(rows >= 0, "rows must be non-negative");
  DSB_REQUIRE(op >= DSB200_CONV_LPC2PAR && op <= DSB200_CONV_NORM0, "converter %d is not supported.", op);
  if ((op == DSB200_CONV_LPC2PAR || op == DSB200_CONV_PAR2LPC || op == DSB200_CONV_GNORM || op == DSB200_CONV_IGNORM) &&
      !(param >= -1.0 && param <= 1.0))
    return fail(DSB200_E_BAD_PARAM, "gamma must be in [-1, 1].");
  if (rows == 0) return DSB200_OK;
  DSB_REQUIRE(x != nullptr && y != nullptr, "NULL data pointer");
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  const int pitch = dim | 1;
  // rows per tile: one per thread while a tile stays below ~64 KB, fewer for long rows
  const size_t row_bytes = static_cast<size_t>(pitch) * sizeof(T);
  const int RB = static_cast<int>(std::max<size_t>(1, std::min<size_t>(kThreads, (64 * 1024) / row_bytes)));
  const size_t smem = static_cast<size_t>(RB) * (row_bytes + sizeof(T));
  if (smem > static_cast<size_t>(max_dynamic_smem(device)))
    return fail(DSB200_E_UNSUPPORTED, "row length %d does not fit in shared memory", dim);
  const int64_t n_tiles = (rows + RB - 1) / RB;
  const int per_sm = std::max<int>(1, std::min<int>(8, static_cast<int>((200 * 1024) / std::max<size_t>(smem, 1))));
  const int blocks = static_cast<int>(std::min<int64_t>(n_tiles, static_cast<int64_t>(sm_count(device)) * per_sm));
  auto launch = [&](auto kern) -> int {
    DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<blocks, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const T*>(x), static_cast<T*>(y), rows, dim, pitch, op, static_cast<T>(param), RB);
    return DSB200_OK;
  };
  // register-resident recursions for the usual LPC orders 12 / 16 / 20 / 24 (DSB200_ROWCONV_GENERIC=1: A/B knob)
  static const bool generic_only = getenv("DSB200_ROWCONV_GENERIC") != nullptr;
  const bool quad = (op == DSB200_CONV_LPC2PAR || op == DSB200_CONV_PAR2LPC) && !generic_only;
  int rc;
  if (quad && dim == 13) rc = launch(rowconv_kernel<T, 13>);
  else if (quad && dim == 17) rc = launch(rowconv_kernel<T, 17>);
  else if (quad && dim == 21) rc = launch(rowconv_kernel<T, 21>);
  else if (quad && dim == 25) rc = launch(rowconv_kernel<T, 25>);
  else rc = launch(rowconv_kernel<T, 0>);
  if (rc != DSB200_OK) return rc;
  return after_launch("rowconv_kernel");
}

}  // namespace
}  // namespace dsb200

extern "C" {

int dsb200_rowconv_f32(const void* x, void* y, int64_t rows, int32_t dim, int32_t op, double param, int device,
                       void* stream) {
  return dsb200::rowconv_impl<float>(x, y, rows, dim, op, param, device, stream);
}
int dsb200_rowconv_f64(const void* x, void* y, int64_t rows, int32_t dim, int32_t op, double param, int device,
                       void* stream) {
  return dsb200::rowconv_impl<double>(x, y, rows, dim, op, param, device, stream);
}

}  // extern "C"
