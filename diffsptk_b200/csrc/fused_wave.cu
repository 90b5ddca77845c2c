// Fused waveform -> {LPC, MFCC} fast paths.  Placeholder: defers to the generic kernels.
#include "common.cuh"
namespace dsb200 {
int lpc_wave_fast_try(const float*, const float*, float*, int64_t, int64_t, const dsb200_frame_params*, int32_t, double,
                      int, cudaStream_t) {
  return DSB200_E_UNSUPPORTED;
}
}  // namespace dsb200

extern "C" {
int dsb200_mfcc_wave_f32(const void*, const void*, const void*, const int32_t*, const int32_t*, const void*, const void*,
                         void*, int64_t, int64_t, const dsb200_stft_params*, const dsb200_mfcc_params*, int, void*) {
  return dsb200::fail(DSB200_E_UNSUPPORTED, "fused waveform->MFCC kernel not available for this configuration");
}
int dsb200_mfcc_wave_f64(const void*, const void*, const void*, const int32_t*, const int32_t*, const void*, const void*,
                         void*, int64_t, int64_t, const dsb200_stft_params*, const dsb200_mfcc_params*, int, void*) {
  return dsb200::fail(DSB200_E_UNSUPPORTED, "fused waveform->MFCC kernel not available for this configuration");
}
}
