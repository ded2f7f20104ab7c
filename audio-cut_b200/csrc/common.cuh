// Shared helpers for libaudiocut_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/audiocut_b200.h"

namespace ac {

void set_error(const std::string& msg);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define AC_CHECK_CUDA(expr)                                                                             \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      ac::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" +        \
                    std::to_string(__LINE__) + ")");                                                    \
      return AC_E_CUDA;                                                                                 \
    }                                                                                                   \
  } while (0)

#define AC_REQUIRE(cond, msg)                   \
  do {                                          \
    if (!(cond)) {                              \
      ac::set_error(std::string("invalid: ") + (msg)); \
      return AC_E_INVALID;                      \
    }                                           \
  } while (0)

#define AC_LAUNCH_CHECK()                  \
  do {                                     \
    ac::count_launch();                    \
    AC_CHECK_CUDA(cudaPeekAtLastError());  \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- optional per-kernel-class timing (ac_profile_begin / ac_profile_collect) ---------------
enum KernelClass {
  KC_STFT = 0, KC_ISTFT, KC_CONV_TC, KC_CONV_SIMT, KC_TDF_SIMT, KC_RESAMPLE_SIMT, KC_CONV1X1, KC_RMS, KC_FEAT_STFT,
  KC_FEAT_FLUX, KC_MISC, KC_TDF_TC, KC_RESAMPLE_TC, KC_COUNT
};
int prof_start(int cls, double flops, double bytes, cudaStream_t st);
void prof_stop(int idx, cudaStream_t st);
extern bool g_prof_on;
struct ProfScope {  // brackets the launches issued during its lifetime with two events on `st`
  int idx; cudaStream_t st;
  ProfScope(int cls, double flops, double bytes, cudaStream_t s) : idx(-1), st(s) {
    if (g_prof_on) idx = prof_start(cls, flops, bytes, st);
  }
  ~ProfScope() { if (idx >= 0) prof_stop(idx, st); }
};

// storage type helpers: activations are float, __half or __nv_bfloat16, math is always fp32
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// ---- 16-bit storage of the tensor-core path ---------------------------------------------------
// The tcgen05 kernels are templated on the operand format FMT of tcgen05.mma kind::f16: IEEE half (11-bit
// significand; the production format: its rounding noise is 18 dB below bfloat16's, which is what the >= 40 dB
// stem-SDR gate of the 16-bit path needs) or bfloat16 (fp32 range, 8-bit significand).  Both run at the same
// tensor-core rate and move the same bytes.  `h16` is the opaque storage element; pack2 / unpack2 convert
// a pair through one 32-bit register.
constexpr int kFmtF16 = 0, kFmtBF16 = 1;
struct h16 { uint16_t bits; };
inline bool is_h16_dtype(int dtype) { return dtype == AC_BF16 || dtype == AC_F16; }
inline int fmt_of_dtype(int dtype) { return dtype == AC_BF16 ? kFmtBF16 : kFmtF16; }
inline h16 h16_rn(float v, int fmt) {  // host: round to nearest even
  h16 r;
  if (fmt == kFmtBF16) { __nv_bfloat16 b = __float2bfloat16_rn(v); r.bits = *reinterpret_cast<uint16_t*>(&b); }
  else { __half h = __float2half_rn(v); r.bits = *reinterpret_cast<uint16_t*>(&h); }
  return r;
}
template <int FMT>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  if constexpr (FMT == kFmtBF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // |x| > 65504 saturates instead of inf
  return r;
}
template <int FMT>
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  if constexpr (FMT == kFmtBF16) return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
  else return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
template <int FMT>
__device__ __forceinline__ float unpack1(h16 v) {
  if constexpr (FMT == kFmtBF16) return __uint_as_float((uint32_t)v.bits << 16);
  else return __half2float(*reinterpret_cast<const __half*>(&v));
}

// streaming 128-bit global load (read once: keep it out of L1)
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// streaming 128-bit global load of raw bits (bf16 x 8)
__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

int device_sm_count();

// ---- per-geometry FFT tables (device resident, cached) ------------------------------------
struct FftPlan {
  int n = 0;
  int n_radix = 0;
  int radix[16];
  const float2* d_twiddle = nullptr;  // exp(-2*pi*i*k/n), k in [0,n)
  // two-level twiddles: exp(-2*pi*i*m/n) = d_tw_lo[m & 63] * d_tw_hi[m >> 6]; 64 + ceil(n/64) entries that stay in L1
  // (gathering from the n-entry table cost one L2 sector per butterfly input: 4 GB of L2 traffic per 16-window STFT launch)
  const float2* d_tw_lo = nullptr;
  const float2* d_tw_hi = nullptr;
  const float* d_hann = nullptr;      // periodic hann, n values
  // three-pass FFT (fft3.cuh; n = 7680 / 6144 only): W_{R1 R2}^{k r} as [R1][R2] and W_n^{k r} as [R1 R2][R3]
  const float2* d_tw3_p2 = nullptr;
  const float2* d_tw3_p3 = nullptr;
};
// returns nullptr (and sets the error) when n has a prime factor other than 2,3,5
const FftPlan* get_fft_plan(int n);

}  // namespace ac
