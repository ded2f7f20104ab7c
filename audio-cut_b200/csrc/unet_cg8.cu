// CUDA-core kernels of the bf16 tensor-core path that work on the CG8 activation layout
// [B][T][C/8][F][8] (tc_common.cuh): the two 1x1 convolutions at the ends of the network (4 <-> g
// channels; HBM-bound streaming) and the TDF layers of the deepest levels, whose GEMMs are too
// small for a 128-row UMMA tile (M or K below 64).
#include "tc_common.cuh"
#include "unet_kernels.cuh"

namespace ac {

// ---- first 1x1 conv: spectrogram [P][4] bf16 -> CG8 [B][T][g/8][F][8], folded BN + ReLU ---------
// One thread per position; lanes walk f, so every 16-byte store of a warp is 512 contiguous bytes.
template <int G, int FMT>
__global__ void __launch_bounds__(256) first_conv_cg8_kernel(const h16* __restrict__ in, h16* __restrict__ out,
                                                             long long rows /* B*T */, int F, const float* __restrict__ w,
                                                             const float* __restrict__ scale, const float* __restrict__ shift) {
  __shared__ float sw[G * 4], ssc[G], ssh[G];
  for (int i = threadIdx.x; i < G * 4; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < G; i += blockDim.x) { ssc[i] = scale[i]; ssh[i] = shift[i]; }
  __syncthreads();
  const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= rows * F) return;
  const long long row = pos / F;
  const int f = (int)(pos - row * F);
  const uint2 raw = *reinterpret_cast<const uint2*>(in + pos * 4);
  const float2 x01 = unpack2<FMT>(raw.x);
  const float2 x23 = unpack2<FMT>(raw.y);
  h16* dst = out + ((size_t)row * (G / 8) * F + f) * 8;
#pragma unroll
  for (int cg = 0; cg < G / 8; ++cg) {
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = cg * 8 + 2 * e + h;
        float s = sw[c * 4] * x01.x;
        s = fmaf(sw[c * 4 + 1], x01.y, s);
        s = fmaf(sw[c * 4 + 2], x23.x, s);
        s = fmaf(sw[c * 4 + 3], x23.y, s);
        v[h] = fmaxf(fmaf(s, ssc[c], ssh[c]), 0.f);
      }
      pk[e] = pack2<FMT>(v[0], v[1]);
    }
    *reinterpret_cast<uint4*>(dst + (size_t)cg * F * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// ---- final 1x1 conv: CG8 [B][T][g/8][F][8] -> [P][4] bf16 (+ bias) -----------------------------
template <int G, int FMT>
__global__ void __launch_bounds__(256) final_conv_cg8_kernel(const h16* __restrict__ in, h16* __restrict__ out,
                                                             long long rows, int F, const float* __restrict__ w,
                                                             const float* __restrict__ bias) {
  __shared__ float sw[4 * G + 4];
  for (int i = threadIdx.x; i < 4 * G + 4; i += blockDim.x) sw[i] = i < 4 * G ? w[i] : bias[i - 4 * G];
  __syncthreads();
  const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= rows * F) return;
  const long long row = pos / F;
  const int f = (int)(pos - row * F);
  const h16* src = in + ((size_t)row * (G / 8) * F + f) * 8;
  float s0 = sw[4 * G], s1 = sw[4 * G + 1], s2 = sw[4 * G + 2], s3 = sw[4 * G + 3];
#pragma unroll
  for (int cg = 0; cg < G / 8; ++cg) {
    const uint4 q = *reinterpret_cast<const uint4*>(src + (size_t)cg * F * 8);
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 v = unpack2<FMT>(u[e]);
      const int c = cg * 8 + 2 * e;
      s0 = fmaf(sw[c], v.x, s0); s0 = fmaf(sw[c + 1], v.y, s0);
      s1 = fmaf(sw[G + c], v.x, s1); s1 = fmaf(sw[G + c + 1], v.y, s1);
      s2 = fmaf(sw[2 * G + c], v.x, s2); s2 = fmaf(sw[2 * G + c + 1], v.y, s2);
      s3 = fmaf(sw[3 * G + c], v.x, s3); s3 = fmaf(sw[3 * G + c + 1], v.y, s3);
    }
  }
  uint2 o;
  o.x = pack2<FMT>(s0, s1);
  o.y = pack2<FMT>(s2, s3);
  *reinterpret_cast<uint2*>(out + pos * 4) = o;
}

template <int G, int FMT>
static void launch_ends(bool first, const void* in, void* out, long long rows, int F, const float* w, const float* a,
                        const float* b, cudaStream_t st) {
  const long long P = rows * F;
  const unsigned grid = (unsigned)((P + 255) / 256);
  if (first)
    first_conv_cg8_kernel<G, FMT><<<grid, 256, 0, st>>>((const h16*)in, (h16*)out, rows, F, w, a, b);
  else
    final_conv_cg8_kernel<G, FMT><<<grid, 256, 0, st>>>((const h16*)in, (h16*)out, rows, F, w, a);
}
template <int G>
static void launch_ends_fmt(int fmt, bool first, const void* in, void* out, long long rows, int F, const float* w, const float* a,
                            const float* b, cudaStream_t st) {
  if (fmt == kFmtBF16) launch_ends<G, kFmtBF16>(first, in, out, rows, F, w, a, b, st);
  else launch_ends<G, kFmtF16>(first, in, out, rows, F, w, a, b, st);
}

int cg8_ends_supported(int g) { return (g == 16 || g == 32 || g == 48 || g == 64) ? AC_OK : AC_E_INVALID; }

static int launch_end_conv(bool first, const void* in, void* out, long long rows, int F, int g, const float* w,
                           const float* a, const float* b, int fmt, cudaStream_t st) {
  AC_REQUIRE(cg8_ends_supported(g) == AC_OK, "cg8 1x1 conv: unsupported channel count");
  ProfScope ps(KC_CONV1X1, 2.0 * rows * F * g * 4, (double)rows * F * (4 + g) * 2, st);
  switch (g) {
    case 16: launch_ends_fmt<16>(fmt, first, in, out, rows, F, w, a, b, st); break;
    case 32: launch_ends_fmt<32>(fmt, first, in, out, rows, F, w, a, b, st); break;
    case 48: launch_ends_fmt<48>(fmt, first, in, out, rows, F, w, a, b, st); break;
    default: launch_ends_fmt<64>(fmt, first, in, out, rows, F, w, a, b, st); break;
  }
  AC_LAUNCH_CHECK();
  return AC_OK;
}
int launch_first_conv_cg8(const void* in, void* out, long long rows, int F, int g, const float* w, const float* scale,
                          const float* shift, int fmt, cudaStream_t st) {
  return launch_end_conv(true, in, out, rows, F, g, w, scale, shift, fmt, st);
}
int launch_final_conv_cg8(const void* in, void* out, long long rows, int F, int g, const float* w, const float* bias,
                          int fmt, cudaStream_t st) {
  return launch_end_conv(false, in, out, rows, F, g, w, bias, nullptr, fmt, st);
}

// ---- small TDF layers on CUDA cores, CG8 in and out --------------------------------------------
//   out[b][t][cg][m][e] = relu(scale[c] * sum_k W[m][k] * in[b][t][cg][k][e] + shift[c]) (+ res),  c = cg*8 + e
// The whole weight matrix (<= 36 KB as fp32, transposed to [K][M] so that lanes read consecutive m) and
// PL (b, t, cg) planes ([K][8] each) are staged in shared memory; a thread owns one output row m of one
// plane for all 8 channels: per k one conflict-free weight read + two broadcast 16-byte plane reads.
constexpr int kTdfSmallThreads = 256;
template <int FMT>
__global__ void __launch_bounds__(kTdfSmallThreads) tdf_small_cg8_kernel(const h16* __restrict__ in,
                                                                         const h16* __restrict__ w /*[M][K]*/,
                                                                         const h16* __restrict__ residual,
                                                                         h16* __restrict__ out, long long n_planes,
                                                                         int cgs, int M, int K, int PL, int planes_per_cta,
                                                                         const float* __restrict__ scale,
                                                                         const float* __restrict__ shift) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  float* sw = reinterpret_cast<float*>(sm_raw);  // [K][M]
  float* sx = sw + (size_t)K * M;                // [PL][K][8]
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
    const int m = i / K, k = i - m * K;
    sw[(size_t)k * M + m] = unpack1<FMT>(w[i]);
  }
  const long long first = (long long)blockIdx.x * planes_per_cta;
  for (long long plane0 = first; plane0 < first + planes_per_cta && plane0 < n_planes; plane0 += PL) {
    const int npl = (int)((n_planes - plane0) < PL ? (n_planes - plane0) : PL);
    __syncthreads();  // weights staged / previous group consumed
    for (int i = threadIdx.x; i < npl * K; i += blockDim.x) {
      const uint4 q = *reinterpret_cast<const uint4*>(in + ((size_t)plane0 * K + i) * 8);
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
      float4 lo, hi;
      float2 v;
      v = unpack2<FMT>(u[0]); lo.x = v.x; lo.y = v.y;
      v = unpack2<FMT>(u[1]); lo.z = v.x; lo.w = v.y;
      v = unpack2<FMT>(u[2]); hi.x = v.x; hi.y = v.y;
      v = unpack2<FMT>(u[3]); hi.z = v.x; hi.w = v.y;
      *reinterpret_cast<float4*>(sx + (size_t)i * 8) = lo;
      *reinterpret_cast<float4*>(sx + (size_t)i * 8 + 4) = hi;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < npl * M; o += blockDim.x) {
      const int pl = o / M, m = o - pl * M;
      const float* x = sx + (size_t)pl * K * 8;
      const float* wc = sw + m;
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        const float wv = wc[(size_t)k * M];
        const float4 a = *reinterpret_cast<const float4*>(x + (size_t)k * 8);
        const float4 b = *reinterpret_cast<const float4*>(x + (size_t)k * 8 + 4);
        acc[0] = fmaf(wv, a.x, acc[0]); acc[1] = fmaf(wv, a.y, acc[1]); acc[2] = fmaf(wv, a.z, acc[2]); acc[3] = fmaf(wv, a.w, acc[3]);
        acc[4] = fmaf(wv, b.x, acc[4]); acc[5] = fmaf(wv, b.y, acc[5]); acc[6] = fmaf(wv, b.z, acc[6]); acc[7] = fmaf(wv, b.w, acc[7]);
      }
      const long long plane = plane0 + pl;
      const int cg = (int)(plane % cgs);
      const size_t oidx = ((size_t)plane * M + m) * 8;
      float res[8];
      if (residual) {
        const uint4 q = *reinterpret_cast<const uint4*>(residual + oidx);
        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 v = unpack2<FMT>(u[e]);
          res[2 * e] = v.x;
          res[2 * e + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) res[e] = 0.f;
      }
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = cg * 8 + 2 * e;
        const float v0 = fmaxf(fmaf(acc[2 * e], __ldg(scale + c), __ldg(shift + c)), 0.f) + res[2 * e];
        const float v1 = fmaxf(fmaf(acc[2 * e + 1], __ldg(scale + c + 1), __ldg(shift + c + 1)), 0.f) + res[2 * e + 1];
        pk[e] = pack2<FMT>(v0, v1);
      }
      *reinterpret_cast<uint4*>(out + oidx) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// in CG8 [nB][T][C/8][K][8]; w bf16 [M][K]; residual/out CG8 [nB][T][C/8][M][8]
int launch_tdf_small_cg8(const h16* in, const h16* w, const h16* residual, h16* out,
                         int nB, int T, int C, int M, int K, const float* scale, const float* shift, int fmt, cudaStream_t st) {
  AC_REQUIRE(C % 8 == 0 && M > 0 && K > 0, "tdf small: shape");
  const long long n_planes = (long long)nB * T * (C / 8);
  const size_t w_bytes = (size_t)M * K * 4;
  AC_REQUIRE(w_bytes <= 64 * 1024, "tdf small: weight matrix too large for the CUDA-core kernel");
  // planes per group: enough outputs for every thread, within ~96 KB of shared memory
  int PL = 1;
  while (PL < 32 && w_bytes + (size_t)(2 * PL) * K * 32 <= 96 * 1024 && PL * M < 2 * kTdfSmallThreads) PL *= 2;
  const size_t smem = w_bytes + (size_t)PL * K * 32;
  AC_REQUIRE(smem <= 160 * 1024, "tdf small: K too large for the CUDA-core kernel");
  static size_t attr_set = 0;
  if (smem > attr_set) {
    AC_CHECK_CUDA(cudaFuncSetAttribute(tdf_small_cg8_kernel<kFmtF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    AC_CHECK_CUDA(cudaFuncSetAttribute(tdf_small_cg8_kernel<kFmtBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = 160 * 1024;
  }
  // each CTA stages the weights once and walks several plane groups: ~4 CTAs per SM in total
  const long long groups = (n_planes + PL - 1) / PL;
  long long ctas = 4LL * device_sm_count();
  if (ctas > groups) ctas = groups;
  const int planes_per_cta = (int)(((groups + ctas - 1) / ctas) * PL);
  const unsigned grid = (unsigned)((n_planes + planes_per_cta - 1) / planes_per_cta);
  ProfScope ps(KC_TDF_SIMT, 2.0 * M * (double)K * C * T * nB, 2.0 * nB * (double)T * C * (K + M * (residual ? 2 : 1)), st);
  if (fmt == kFmtBF16)
    tdf_small_cg8_kernel<kFmtBF16><<<grid, kTdfSmallThreads, smem, st>>>(in, w, residual, out, n_planes, C / 8, M, K, PL, planes_per_cta,
                                                                        scale, shift);
  else
    tdf_small_cg8_kernel<kFmtF16><<<grid, kTdfSmallThreads, smem, st>>>(in, w, residual, out, n_planes, C / 8, M, K, PL, planes_per_cta,
                                                                       scale, shift);
  AC_LAUNCH_CHECK();
  return AC_OK;
}

}  // namespace ac
