// Error reporting, device-property cache and the twiddle-table cache of the C ABI.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <map>
#include <cstdlib>
#include <mutex>
#include <string>
#include <tuple>

#include "common.cuh"

namespace dsb200 {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(DSB200_E_CUDA, "CUDA error '%s' in %s", cudaGetErrorString(e), what);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static thread_local const char* g_last_kernel = "";
void note_kernel(const char* name) { g_last_kernel = name; }

static std::mutex g_mu;
static std::map<int, cudaDeviceProp> g_props;

static const cudaDeviceProp* props(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_props.find(device);
  if (it == g_props.end()) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return nullptr;
    it = g_props.emplace(device, p).first;
  }
  return &it->second;
}

static std::atomic<int> g_sm_margin{0};

int sm_count(int device) {
  const cudaDeviceProp* p = props(device);
  const int n = (p ? p->multiProcessorCount : 148) - g_sm_margin.load(std::memory_order_relaxed);
  return n > 1 ? n : 1;
}

// Tuning knobs: a value set through dsb200_set_knob wins over the environment variable DSB200_<name>, which wins over
// the built-in default.  Looked up on every launch (one map probe + one getenv), so one process can sweep a knob.
static std::map<std::string, int> g_knobs;

int knob(const char* name, int dflt) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_knobs.find(name);
    if (it != g_knobs.end()) return it->second;
  }
  const std::string env = std::string("DSB200_") + name;
  const char* e = getenv(env.c_str());
  return (e != nullptr && *e != '\0') ? atoi(e) : dflt;
}

int max_dynamic_smem(int device) {
  const cudaDeviceProp* p = props(device);
  return p ? static_cast<int>(p->sharedMemPerBlockOptin) : 227 * 1024;
}

template <typename T>
__global__ void fill_twiddles(T* tw, int n) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double s, c;
  sincospi(-2.0 * static_cast<double>(k) / static_cast<double>(n), &s, &c);
  tw[2 * k] = static_cast<T>(c);
  tw[2 * k + 1] = static_cast<T>(s);
}

static std::map<std::tuple<int, int, bool>, void*> g_tw;

const void* twiddle_table(int device, int n, bool is_f64, cudaStream_t stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_tuple(device, n, is_f64);
  auto it = g_tw.find(key);
  if (it != g_tw.end()) return it->second;
  void* p = nullptr;
  size_t bytes = static_cast<size_t>(n) * 2 * (is_f64 ? sizeof(double) : sizeof(float));
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  int threads = 128, blocks = (n + threads - 1) / threads;
  if (is_f64)
    fill_twiddles<double><<<blocks, threads, 0, stream>>>(static_cast<double*>(p), n);
  else
    fill_twiddles<float><<<blocks, threads, 0, stream>>>(static_cast<float*>(p), n);
  count_launch();
  // First use only: make the table visible to every stream before it is cached.
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) {
    cudaFree(p);
    return nullptr;
  }
  g_tw[key] = p;
  return p;
}

}  // namespace dsb200

extern "C" {

int dsb200_version(void) { return DSB200_VERSION; }

const char* dsb200_last_error(void) { return dsb200::g_err; }

int64_t dsb200_launch_count(void) { return dsb200::g_launches.load(std::memory_order_relaxed); }

const char* dsb200_last_kernel(void) { return dsb200::g_last_kernel; }

int dsb200_set_sm_margin(int32_t n_sms) {
  if (n_sms < 0) return dsb200::fail(DSB200_E_BAD_PARAM, "sm margin must be non-negative");
  return dsb200::g_sm_margin.exchange(n_sms, std::memory_order_relaxed);
}

int dsb200_set_knob(const char* name, int32_t value) {
  if (name == nullptr || *name == '\0') return dsb200::fail(DSB200_E_BAD_PARAM, "knob name is empty");
  std::lock_guard<std::mutex> lk(dsb200::g_mu);
  dsb200::g_knobs[name] = value;
  return DSB200_OK;
}

int dsb200_clear_knobs(void) {
  std::lock_guard<std::mutex> lk(dsb200::g_mu);
  dsb200::g_knobs.clear();
  return DSB200_OK;
}

int64_t dsb200_num_frames(int64_t T, int32_t frame_period) {
  if (T <= 0 || frame_period <= 0) return 0;
  return (T - 1) / frame_period + 1;
}

}  // extern "C"
