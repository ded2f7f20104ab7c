// CUDA-core (FFMA) kernels of the TFC-TDF U-Net: the fp32 parity path (>= 60 dB vs the CPU
// oracle) and the layer-by-layer cross-check for the tcgen05 kernels (bf16 storage, same math).
//
// Everything the network does between its 1x1 convs is one implicit GEMM
//     C[M][N] = A[M][K] * B[K][N],  A rows K-contiguous (gathered), B rows N-contiguous
//   conv3x3      M = (b,t,f)        N = c_out      K = 9*c_in   A = im2col gather (zero padded)
//   down 2x2/s2  M = (b,t/2,f/2)    N = c_out      K = 4*c_in   A = gather
//   up   2x2/s2  M = (b,t,f)        N = 4*c_out    K = c_in     A = activations, scatter epilogue * skip
//   TDF linear   M = f_out          N = c          K = f_in     A = weights, B = activations of one (b,t)
// with channels-last activations [B][T][F][C].  128 x BN x 16 tiles, 256 threads, 8 x BN/16
// outputs per thread, register-staged double buffering.
#include "unet_kernels.cuh"

namespace ac {

template <typename T>
__device__ __forceinline__ void load8(const T* p, float* r);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float* r) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* r) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    r[2 * i] = f.x;
    r[2 * i + 1] = f.y;
  }
}

template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float* r) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    r[2 * i] = f.x;
    r[2 * i + 1] = f.y;
  }
}

template <typename T, int BN>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs a) {
  constexpr int BM = 128, BK = 16, LDA = BM + 4, NB = BN / 16, NBL = (BK * BN + 255) / 256;
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int tm = tid & 15, tn = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const long long bz = blockIdx.z;
  const T* Ab = reinterpret_cast<const T*>(a.A) + (a.a_mode == A_PLAIN ? bz * a.a_batch_stride : 0);
  const T* Bb = reinterpret_cast<const T*>(a.Bm) + bz * a.b_batch_stride;
  const int lrow = tid >> 1, lhalf = tid & 1;
  const int gm = m0 + lrow;
  const bool row_ok = gm < a.M;
  int rb = 0, rt = 0, rf = 0;
  if (a.a_mode != A_PLAIN && row_ok) {
    rf = gm % a.F;
    const int q = gm / a.F;
    rt = q % a.T;
    rb = q / a.T;
  }
  float acc[8][NB];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int i = 0; i < NB; ++i) acc[r][i] = 0.f;
  float ra[8], rbv[NBL];
  const int nk = (a.K + BK - 1) / BK;

  auto load_tiles = [&](int kt) {
    const int k0 = kt * BK + lhalf * 8;
    const T* p;
    bool ok = row_ok;
    if (a.a_mode == A_CONV3) {
      const int tap = k0 / a.C, c0 = k0 - tap * a.C;
      const int tt = rt + tap / 3 - 1, ff = rf + tap % 3 - 1;
      ok = ok && tt >= 0 && tt < a.T && ff >= 0 && ff < a.F;
      p = Ab + (((long long)rb * a.T + tt) * a.F + ff) * a.C + c0;
    } else if (a.a_mode == A_DOWN2) {
      const int tap = k0 / a.C, c0 = k0 - tap * a.C;
      const int tt = 2 * rt + tap / 2, ff = 2 * rf + tap % 2;
      p = Ab + (((long long)rb * (2 * a.T) + tt) * (2 * a.F) + ff) * a.C + c0;
    } else {
      p = Ab + (long long)gm * a.K + k0;
    }
    if (ok && k0 + 8 <= a.K && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      load8<T>(p, ra);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) ra[e] = (ok && k0 + e < a.K) ? to_f32(p[e]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NBL; ++i) {
      const int idx = tid + i * 256;
      const int k = idx / BN, n = idx - k * BN;
      const int gk = kt * BK + k;
      rbv[i] = (idx < BK * BN && gk < a.K) ? to_f32(Bb[(long long)gk * a.N + n0 + n]) : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 8; ++e) As[buf][lhalf * 8 + e][lrow] = ra[e];
#pragma unroll
    for (int i = 0; i < NBL; ++i) {
      const int idx = tid + i * 256;
      if (idx < BK * BN) Bs[buf][idx / BN][idx % BN] = rbv[i];
    }
  };

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) load_tiles(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][tm * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) bv[i] = Bs[cur][k][tn + 16 * i];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < NB; ++i) acc[r][i] = fmaf(av[r], bv[i], acc[r][i]);
    }
    if (kt + 1 < nk) {
      store_tiles(cur ^ 1);
      __syncthreads();
    }
  }

  T* out = reinterpret_cast<T*>(a.out);
  const T* extra = reinterpret_cast<const T*>(a.extra);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int m = m0 + tm * 8 + r;
    if (m >= a.M) continue;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int n = n0 + tn + 16 * i;
      const int c = n % a.cmod;
      float y = fmaxf(fmaf(acc[r][i], __ldg(a.scale + c), __ldg(a.shift + c)), 0.f);
      if (a.epi == EPI_UP_SKIP) {
        const int cout = a.N >> 2;
        const int tap = n / cout;
        const int f = m % a.up_F;
        const int q = m / a.up_F;
        const int t = q % a.up_T, b = q / a.up_T;
        const long long idx =
            (((long long)b * (2 * a.up_T) + 2 * t + (tap >> 1)) * (2 * a.up_F) + 2 * f + (tap & 1)) * cout + c;
        y *= to_f32(extra[idx]);
        out[idx] = from_f32<T>(y);
      } else {
        const long long idx = bz * a.c_batch_stride + (long long)m * a.N + n;
        if (a.epi == EPI_RESIDUAL) y += to_f32(extra[idx]);
        out[idx] = from_f32<T>(y);
      }
    }
  }
}

template <typename T>
static int launch_gemm_t(const GemmArgs& a, cudaStream_t st) {
  AC_REQUIRE(a.N % 16 == 0, "gemm: N must be a multiple of 16");
  int bn = (a.N % 48 == 0) ? 48 : (a.N % 32 == 0 ? 32 : 16);
  dim3 grid((a.M + 127) / 128, a.N / bn, a.batch);
  AC_REQUIRE(grid.z <= 65535 && grid.y <= 65535, "gemm: grid too large");
  ProfScope ps(a.kclass, 2.0 * a.M * (double)a.N * a.K * a.batch, 0.0, st);
  if (bn == 48) gemm_simt_kernel<T, 48><<<grid, 256, 0, st>>>(a);
  else if (bn == 32) gemm_simt_kernel<T, 32><<<grid, 256, 0, st>>>(a);
  else gemm_simt_kernel<T, 16><<<grid, 256, 0, st>>>(a);
  AC_LAUNCH_CHECK();
  return AC_OK;
}

int launch_gemm_simt(const GemmArgs& a, int dtype, cudaStream_t st) {
  if (a.a_mode != A_PLAIN) AC_REQUIRE(a.C % 16 == 0, "gemm: channels must be a multiple of 16");
  if (dtype == AC_F32) return launch_gemm_t<float>(a, st);
  return dtype == AC_F16 ? launch_gemm_t<__half>(a, st) : launch_gemm_t<__nv_bfloat16>(a, st);
}

// ---- 1x1 convs -------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) first_conv_kernel(const T* __restrict__ in, T* __restrict__ out, long long P,
                                                         int g, const float* __restrict__ w,
                                                         const float* __restrict__ scale,
                                                         const float* __restrict__ shift) {
  // one thread per (position, 8 output channels)
  const int groups = g >> 3;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= P * groups) return;
  const long long pos = gid / groups;
  const int c0 = (int)(gid - pos * groups) * 8;
  float x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = to_f32(in[pos * 4 + i]);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s = fmaf(__ldg(w + c * 4 + i), x[i], s);
    out[pos * g + c] = from_f32<T>(fmaxf(fmaf(s, __ldg(scale + c), __ldg(shift + c)), 0.f));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) final_conv_kernel(const T* __restrict__ in, T* __restrict__ out, long long P,
                                                         int g, const float* __restrict__ w,
                                                         const float* __restrict__ bias) {
  // one warp per 32 positions; lanes stride channels so loads are coalesced
  extern __shared__ float sw[];  // [4][g] + bias[4]
  for (int i = threadIdx.x; i < 4 * g + 4; i += blockDim.x) sw[i] = i < 4 * g ? w[i] : bias[i - 4 * g];
  __syncthreads();
  const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= P) return;
  float s0 = sw[4 * g], s1 = sw[4 * g + 1], s2 = sw[4 * g + 2], s3 = sw[4 * g + 3];
  const T* p = in + pos * g;
  for (int c = 0; c < g; c += 8) {
    float r[8];
    load8<T>(p + c, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s0 = fmaf(sw[c + j], r[j], s0);
      s1 = fmaf(sw[g + c + j], r[j], s1);
      s2 = fmaf(sw[2 * g + c + j], r[j], s2);
      s3 = fmaf(sw[3 * g + c + j], r[j], s3);
    }
  }
  out[pos * 4 + 0] = from_f32<T>(s0);
  out[pos * 4 + 1] = from_f32<T>(s1);
  out[pos * 4 + 2] = from_f32<T>(s2);
  out[pos * 4 + 3] = from_f32<T>(s3);
}

int launch_first_conv(const void* in, void* out, long long P, int g, const float* w, const float* scale,
                      const float* shift, int dtype, cudaStream_t st) {
  const long long total = P * (g / 8);
  const unsigned grid = (unsigned)((total + 255) / 256);
  ProfScope ps(KC_CONV1X1, 2.0 * P * g * 4, (double)P * (4 + g) * (dtype == AC_F32 ? 4 : 2), st);
  if (dtype == AC_F32)
    first_conv_kernel<float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, P, g, w, scale, shift);
  else if (dtype == AC_F16)
    first_conv_kernel<__half><<<grid, 256, 0, st>>>((const __half*)in, (__half*)out, P, g, w, scale, shift);
  else
    first_conv_kernel<__nv_bfloat16>
        <<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, P, g, w, scale, shift);
  AC_LAUNCH_CHECK();
  return AC_OK;
}

int launch_final_conv(const void* in, void* out, long long P, int g, const float* w, const float* bias, int dtype,
                      cudaStream_t st) {
  const unsigned grid = (unsigned)((P + 255) / 256);
  const size_t smem = sizeof(float) * (4 * g + 4);
  ProfScope ps(KC_CONV1X1, 2.0 * P * g * 4, (double)P * (4 + g) * (dtype == AC_F32 ? 4 : 2), st);
  if (dtype == AC_F32)
    final_conv_kernel<float><<<grid, 256, smem, st>>>((const float*)in, (float*)out, P, g, w, bias);
  else if (dtype == AC_F16)
    final_conv_kernel<__half><<<grid, 256, smem, st>>>((const __half*)in, (__half*)out, P, g, w, bias);
  else
    final_conv_kernel<__nv_bfloat16>
        <<<grid, 256, smem, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, P, g, w, bias);
  AC_LAUNCH_CHECK();
  return AC_OK;
}

}  // namespace ac
