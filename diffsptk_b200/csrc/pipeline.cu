// Host-buffer STFT pipeline: the end-to-end path of the C ABI.
//
// Owns two device staging slots and three streams so that, for utterance chunk c,
// H2D(c+1), compute(c) and D2H(c-1) overlap.  Host buffers should be page-locked
// (cudaHostAlloc / torch pin_memory) for the copies to be truly asynchronous.
#include <new>

#include "common.cuh"

struct dsb200_pipeline {
  int device = 0;
  int64_t chunk = 0, T = 0, N = 0, K = 0;
  int is_f64 = 0;
  int out_mult = 1;  // 2 for complex output
  dsb200_stft_params p{};
  void* dx[2] = {nullptr, nullptr};
  void* dy[2] = {nullptr, nullptr};
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  cudaEvent_t e_in[2] = {nullptr, nullptr}, e_run[2] = {nullptr, nullptr}, e_out[2] = {nullptr, nullptr};
};

using namespace dsb200;

extern "C" {

int dsb200_pipeline_destroy(dsb200_pipeline* pl) {
  if (pl == nullptr) return DSB200_OK;
  DeviceScope ds(pl->device);
  for (int i = 0; i < 2; ++i) {
    if (pl->dx[i]) cudaFree(pl->dx[i]);
    if (pl->dy[i]) cudaFree(pl->dy[i]);
    if (pl->e_in[i]) cudaEventDestroy(pl->e_in[i]);
    if (pl->e_run[i]) cudaEventDestroy(pl->e_run[i]);
    if (pl->e_out[i]) cudaEventDestroy(pl->e_out[i]);
  }
  if (pl->s_in) cudaStreamDestroy(pl->s_in);
  if (pl->s_run) cudaStreamDestroy(pl->s_run);
  if (pl->s_out) cudaStreamDestroy(pl->s_out);
  delete pl;
  return DSB200_OK;
}

int dsb200_pipeline_create(dsb200_pipeline** out, int device, int64_t chunk_utterances, int64_t T,
                           const dsb200_stft_params* p, int is_f64) {
  DSB_REQUIRE(out != nullptr && p != nullptr, "NULL argument");
  DSB_REQUIRE(chunk_utterances > 0 && T > 0, "chunk size and waveform length must be positive");
  DSB_REQUIRE(p->frame.frame_period > 0 && p->spec.fft_length > 1, "bad stft params");
  *out = nullptr;
  DeviceScope ds(device);
  DSB_CUDA(ds.err);
  dsb200_pipeline* pl = new (std::nothrow) dsb200_pipeline();
  if (pl == nullptr) return fail(DSB200_E_CUDA, "out of host memory");
  pl->device = device;
  pl->chunk = chunk_utterances;
  pl->T = T;
  pl->N = dsb200_num_frames(T, p->frame.frame_period);
  pl->K = p->spec.fft_length / 2 + 1;
  pl->is_f64 = is_f64;
  pl->out_mult = p->spec.out_format == DSB200_SPEC_COMPLEX ? 2 : 1;
  pl->p = *p;
  const size_t es = is_f64 ? 8 : 4;
  const size_t xb = static_cast<size_t>(pl->chunk) * T * es;
  const size_t yb = static_cast<size_t>(pl->chunk) * pl->N * pl->K * pl->out_mult * es;
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaMalloc(&pl->dx[i], xb);
    if (e == cudaSuccess) e = cudaMalloc(&pl->dy[i], yb);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->e_in[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->e_run[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->e_out[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&pl->s_in, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&pl->s_run, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&pl->s_out, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    dsb200_pipeline_destroy(pl);
    return cuda_fail(e, "dsb200_pipeline_create");
  }
  *out = pl;
  return DSB200_OK;
}

int dsb200_pipeline_stft_host(dsb200_pipeline* pl, const void* x_host, const void* window_dev, void* y_host,
                              int64_t batch) {
  DSB_REQUIRE(pl != nullptr, "pipeline is NULL");
  DSB_REQUIRE(batch >= 0, "batch must be non-negative");
  if (batch == 0) return DSB200_OK;
  DSB_REQUIRE(x_host != nullptr && y_host != nullptr && window_dev != nullptr, "NULL data pointer");
  DeviceScope ds(pl->device);
  DSB_CUDA(ds.err);
  const size_t es = pl->is_f64 ? 8 : 4;
  const size_t x_row = static_cast<size_t>(pl->T) * es;
  const size_t y_row = static_cast<size_t>(pl->N) * pl->K * pl->out_mult * es;
  const char* xh = static_cast<const char*>(x_host);
  char* yh = static_cast<char*>(y_host);
  // Every exit, error exits included, first drains the three streams: the copies read and write caller-owned
  // host buffers, which must not be touched after this call returns.
  auto drain = [&](int rc) {
    cudaStreamSynchronize(pl->s_out);
    cudaStreamSynchronize(pl->s_run);
    cudaStreamSynchronize(pl->s_in);
    return rc;
  };
#define DSB_PIPE(call)                                                    \
  do {                                                                    \
    cudaError_t e__ = (call);                                             \
    if (e__ != cudaSuccess) return drain(::dsb200::cuda_fail(e__, #call)); \
  } while (0)
  int64_t c = 0;
  for (int64_t b0 = 0; b0 < batch; b0 += pl->chunk, ++c) {
    const int s = static_cast<int>(c & 1);
    const int64_t nb = batch - b0 < pl->chunk ? batch - b0 : pl->chunk;
    if (c >= 2) DSB_PIPE(cudaStreamWaitEvent(pl->s_in, pl->e_run[s], 0));   // dx[s] no longer read
    DSB_PIPE(cudaMemcpyAsync(pl->dx[s], xh + b0 * x_row, nb * x_row, cudaMemcpyHostToDevice, pl->s_in));
    DSB_PIPE(cudaEventRecord(pl->e_in[s], pl->s_in));
    DSB_PIPE(cudaStreamWaitEvent(pl->s_run, pl->e_in[s], 0));
    if (c >= 2) DSB_PIPE(cudaStreamWaitEvent(pl->s_run, pl->e_out[s], 0));  // dy[s] drained
    const int rc = pl->is_f64 ? dsb200_stft_f64(pl->dx[s], window_dev, pl->dy[s], nb, pl->T, &pl->p, pl->device, pl->s_run)
                              : dsb200_stft_f32(pl->dx[s], window_dev, pl->dy[s], nb, pl->T, &pl->p, pl->device, pl->s_run);
    if (rc != DSB200_OK) return drain(rc);
    DSB_PIPE(cudaEventRecord(pl->e_run[s], pl->s_run));
    DSB_PIPE(cudaStreamWaitEvent(pl->s_out, pl->e_run[s], 0));
    DSB_PIPE(cudaMemcpyAsync(yh + b0 * y_row, pl->dy[s], nb * y_row, cudaMemcpyDeviceToHost, pl->s_out));
    DSB_PIPE(cudaEventRecord(pl->e_out[s], pl->s_out));
  }
#undef DSB_PIPE
  DSB_CUDA(cudaStreamSynchronize(pl->s_out));
  DSB_CUDA(cudaStreamSynchronize(pl->s_run));
  DSB_CUDA(cudaStreamSynchronize(pl->s_in));
  return DSB200_OK;
}

}  // extern "C"
