// Library bring-up, error reporting, cached FFT tables.
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace ac {

static thread_local std::string t_error;
std::atomic<long long> g_launches{0};
static int g_sm_count = 0;
static int g_bound_device = -1;  // the library keeps per-device tables (FFT plans, window caches, abort flag): one device per process
static std::mutex g_plan_mu;
static std::map<int, FftPlan*> g_plans;

void set_error(const std::string& msg) { t_error = msg; }

bool g_prof_on = false;
struct ProfSpan { cudaEvent_t a, b; int cls; double flops, bytes; };
static std::vector<ProfSpan> g_spans;
static std::vector<cudaEvent_t> g_event_pool;
static std::mutex g_prof_mu;
static cudaEvent_t prof_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
int prof_start(int cls, double flops, double bytes, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfSpan s{prof_event(), prof_event(), cls, flops, bytes};
  cudaEventRecord(s.a, st);
  g_spans.push_back(s);
  return (int)g_spans.size() - 1;
}
void prof_stop(int idx, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (idx >= 0 && idx < (int)g_spans.size()) cudaEventRecord(g_spans[idx].b, st);
}
int device_sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }

const FftPlan* get_fft_plan(int n) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  auto it = g_plans.find(n);
  if (it != g_plans.end()) return it->second;
  FftPlan* p = new FftPlan();
  p->n = n;
  int m = n, c3 = 0, c5 = 0, c2 = 0;
  while (m % 3 == 0) { m /= 3; ++c3; }
  while (m % 5 == 0) { m /= 5; ++c5; }
  while (m % 2 == 0) { m /= 2; ++c2; }
  if (m != 1 || n < 2) {
    set_error("fft length " + std::to_string(n) + " is not of the form 2^a 3^b 5^c");
    delete p;
    return nullptr;
  }
  // odd radices first: their Ns=1 pass writes with an odd stride (no smem bank conflicts)
  int k = 0;
  for (int i = 0; i < c3; ++i) p->radix[k++] = 3;
  for (int i = 0; i < c5; ++i) p->radix[k++] = 5;
  while (c2 >= 3) { p->radix[k++] = 8; c2 -= 3; }
  if (c2 == 2) p->radix[k++] = 4;
  if (c2 == 1) p->radix[k++] = 2;
  p->n_radix = k;
  if (k > 12) {
    set_error("fft length too large");
    delete p;
    return nullptr;
  }
  std::vector<float2> tw(n);
  std::vector<float> hann(n);
  const double two_pi = 6.283185307179586476925286766559;
  for (int i = 0; i < n; ++i) {
    double a = -two_pi * (double)i / (double)n;
    tw[i] = make_float2((float)std::cos(a), (float)std::sin(a));
    hann[i] = (float)(0.5 - 0.5 * std::cos(two_pi * (double)i / (double)n));
  }
  const int n_hi = (n + 63) / 64;
  std::vector<float2> tw2(64 + n_hi);
  for (int i = 0; i < 64; ++i) {
    double a = -two_pi * (double)i / (double)n;
    tw2[i] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  for (int i = 0; i < n_hi; ++i) {
    double a = -two_pi * (double)(64 * i) / (double)n;
    tw2[64 + i] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  float2* d_tw2 = nullptr;
  if (cudaMalloc(&d_tw2, sizeof(float2) * tw2.size()) != cudaSuccess ||
      cudaMemcpy(d_tw2, tw2.data(), sizeof(float2) * tw2.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error(std::string("fft plan upload: ") + cudaGetErrorString(cudaGetLastError()));
    delete p;
    return nullptr;
  }
  p->d_tw_lo = d_tw2;
  p->d_tw_hi = d_tw2 + 64;
  float2* d_tw = nullptr;
  float* d_h = nullptr;
  if (cudaMalloc(&d_tw, sizeof(float2) * n) != cudaSuccess || cudaMalloc(&d_h, sizeof(float) * n) != cudaSuccess ||
      cudaMemcpy(d_tw, tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(d_h, hann.data(), sizeof(float) * n, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error(std::string("fft plan upload: ") + cudaGetErrorString(cudaGetLastError()));
    delete p;
    return nullptr;
  }
  p->d_twiddle = d_tw;
  p->d_hann = d_h;
  if (n == 7680 || n == 6144) {
    // tables of the three-pass FFT (fft3.cuh): radices 16 x 20 x 24 (7680) / 16 x 16 x 24 (6144)
    const int r1 = 16, r2 = n == 7680 ? 20 : 16, r3 = 24;
    std::vector<float2> t2((size_t)r1 * r2), t3((size_t)r1 * r2 * r3);
    for (int k = 0; k < r1; ++k)
      for (int r = 0; r < r2; ++r) {
        double a = -two_pi * (double)((k * r) % (r1 * r2)) / (double)(r1 * r2);
        t2[(size_t)k * r2 + r] = make_float2((float)std::cos(a), (float)std::sin(a));
      }
    for (int k = 0; k < r1 * r2; ++k)
      for (int r = 0; r < r3; ++r) {
        double a = -two_pi * (double)((k * r) % n) / (double)n;
        t3[(size_t)k * r3 + r] = make_float2((float)std::cos(a), (float)std::sin(a));
      }
    float2 *d2 = nullptr, *d3 = nullptr;
    if (cudaMalloc(&d2, sizeof(float2) * t2.size()) != cudaSuccess || cudaMalloc(&d3, sizeof(float2) * t3.size()) != cudaSuccess ||
        cudaMemcpy(d2, t2.data(), sizeof(float2) * t2.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d3, t3.data(), sizeof(float2) * t3.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error(std::string("fft plan upload: ") + cudaGetErrorString(cudaGetLastError()));
      delete p;
      return nullptr;
    }
    p->d_tw3_p2 = d2;
    p->d_tw3_p3 = d3;
  }
  g_plans[n] = p;
  return p;
}

}  // namespace ac

extern "C" {

int ac_profile_begin(void) {
  std::lock_guard<std::mutex> lk(ac::g_prof_mu);
  for (auto& s : ac::g_spans) { ac::g_event_pool.push_back(s.a); ac::g_event_pool.push_back(s.b); }
  ac::g_spans.clear();
  ac::g_prof_on = true;
  return AC_OK;
}

int ac_profile_collect(ac_kernel_stat* out, int max_classes) {
  ac::g_prof_on = false;
  AC_REQUIRE(out && max_classes >= ac::KC_COUNT, "ac_profile_collect needs room for every kernel class");
  AC_CHECK_CUDA(cudaDeviceSynchronize());
  static const char* names[ac::KC_COUNT] = {"stft_mdx", "istft_ola_stems", "conv3x3_tcgen05", "conv3x3_simt", "tdf_simt",
                                            "resample_simt", "conv1x1", "frame_rms", "stft2048_features", "onset_flux",
                                            "misc", "tdf_tcgen05", "resample_tcgen05"};
  for (int c = 0; c < ac::KC_COUNT; ++c) {
    out[c].name = names[c];
    out[c].launches = 0;
    out[c].total_ms = out[c].flops = out[c].bytes = 0.0;
  }
  std::lock_guard<std::mutex> lk(ac::g_prof_mu);
  for (auto& s : ac::g_spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
      out[s.cls].launches += 1;
      out[s.cls].total_ms += ms;
      out[s.cls].flops += s.flops;
      out[s.cls].bytes += s.bytes;
    }
    ac::g_event_pool.push_back(s.a);
    ac::g_event_pool.push_back(s.b);
  }
  ac::g_spans.clear();
  return ac::KC_COUNT;
}

int ac_abi_version(void) { return 1; }

int ac_host_is_pinned(const void* h_ptr, size_t bytes) {
  if (!h_ptr || bytes == 0) return 0;
  cudaPointerAttributes a0, a1;
  if (cudaPointerGetAttributes(&a0, h_ptr) != cudaSuccess ||
      cudaPointerGetAttributes(&a1, static_cast<const char*>(h_ptr) + bytes - 1) != cudaSuccess) {
    cudaGetLastError();  // unregistered host memory reports an error on old drivers: not an error for the caller
    return 0;
  }
  return a0.type == cudaMemoryTypeHost && a1.type == cudaMemoryTypeHost ? 1 : 0;
}

int ac_copy_h2d_async(void* d_dst, const void* h_src, size_t bytes, void* stream) {
  AC_REQUIRE(d_dst && h_src, "null pointer");
  AC_CHECK_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return AC_OK;
}
const char* ac_last_error(void) { return ac::t_error.c_str(); }
long long ac_launch_count(void) { return ac::g_launches.load(); }

int ac_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    ac::set_error(std::string("no CUDA device: ") + cudaGetErrorString(e));
    return AC_E_NODEVICE;
  }
  AC_REQUIRE(device >= 0 && device < n, "device index out of range");
  if (ac::g_bound_device >= 0 && ac::g_bound_device != device) {
    ac::set_error("libaudiocut_b200 is bound to cuda:" + std::to_string(ac::g_bound_device) + " in this process (its cached device tables "
                  "live there); use one process per GPU (torchrun / CUDA_VISIBLE_DEVICES), as bench.py --gpus N does");
    return AC_E_INVALID;
  }
  AC_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  AC_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    ac::set_error(std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) +
                  std::to_string(prop.minor) + "; this library is built for sm_100a only");
    return AC_E_NODEVICE;
  }
  ac::g_sm_count = prop.multiProcessorCount;
  ac::g_bound_device = device;
  return AC_OK;
}

long long ac_frame_count(long long n, int frame, int hop, int center) {
  if (frame <= 0 || hop <= 0 || n < 0) return 0;
  long long padded = center ? n + 2LL * (frame / 2) : n;
  if (padded < frame) return 0;
  return 1 + (padded - frame) / hop;
}

}  // extern "C"
