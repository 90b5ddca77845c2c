// Placeholder until the specialised kernel lands: always defers to the generic path.
#include "common.cuh"
namespace dsb200 {
int stft512_try(const float*, const float*, float*, int64_t, int64_t, const dsb200_stft_params*, int, cudaStream_t) {
  return DSB200_E_UNSUPPORTED;
}
}  // namespace dsb200
