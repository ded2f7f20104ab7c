// MDX23 window STFT and the fused inverse path (iFFT + window + overlap-add + envelope
// normalise + trim + stem arithmetic + accumulate).
//
// Replaces Conv_TDF_net_trim_model.stft/.istft (external MVSEP-MDX23 inference.py; called at
// backends.py:355, :376) and, in stems mode, backends.py:377 (trim/concat), :389-406 (crop, mix
// subtraction, mono mean) plus enhanced_vocal_separator.py:423-437 (effective-region accumulate).
//
// STFT: one CTA per (frame, window).  Left and right are packed as re/im of ONE complex FFT of
// length n_fft (two real transforms for the price of one); the Hermitian split yields
// {L_re,L_im,R_re,R_im}, written as one 16-byte (f32) / 8-byte (bf16) store per bin into the
// [win][t][f][4] layout, i.e. every frame is a contiguous 48 KB run in HBM.
//
// iSTFT: one CTA per (strip of consecutive frames, window).  The CTA walks its frames in order,
// keeps the ceil(n_fft/hop) hop-blocks that are still receiving contributions in a shared-memory
// ring, and emits a hop-block as soon as its last frame has been added - the overlap-add never
// touches HBM.  Strips re-compute nb-1 warm-up frames.
#include <stdlib.h>

#include <map>
#include <mutex>
#include <vector>

#include "fft.cuh"
#include "fft3.cuh"
#include "stft_mdx.cuh"

namespace ac {

constexpr int kFftThreads = 512;
// The fused iSTFT holds one CTA per SM (two FFT buffers + the overlap-add ring = 187 KB of shared memory): 32 warps instead of
// 16 double the loads in flight of its latency-bound passes (ncu, 16 warps: 25 % warps active, issue 53 %).
constexpr int kIstftThreads = 1024;

static std::mutex g_mdx_mu;
static std::map<std::vector<int>, MdxPlan*> g_mdx_plans;

const MdxPlan* get_mdx_plan(const ac_mdx_geom& g) {
  std::lock_guard<std::mutex> lk(g_mdx_mu);
  std::vector<int> key = {g.n_fft, g.hop, g.dim_f, g.dim_t};
  auto it = g_mdx_plans.find(key);
  if (it != g_mdx_plans.end()) return it->second;
  if (g.n_fft < 16 || g.hop <= 0 || g.dim_t < 2 || g.dim_f <= 0 || g.dim_f > g.n_fft / 2 || (g.n_fft & 1)) {
    set_error("bad mdx geometry");
    return nullptr;
  }
  const int W = g.hop * (g.dim_t - 1);
  if (W <= g.n_fft) {
    set_error("mdx geometry: window shorter than n_fft");
    return nullptr;
  }
  const FftPlan* fft = get_fft_plan(g.n_fft);
  if (!fft) return nullptr;
  const int plen = W + g.n_fft;
  std::vector<double> env(plen, 0.0);
  std::vector<float> w2(g.n_fft);
  for (int j = 0; j < g.n_fft; ++j) {
    float w = (float)(0.5 - 0.5 * std::cos(6.283185307179586476925286766559 * (double)j / (double)g.n_fft));
    w2[j] = w * w;
  }
  for (int t = 0; t < g.dim_t; ++t)
    for (int j = 0; j < g.n_fft; ++j) env[(size_t)t * g.hop + j] += (double)w2[j];
  std::vector<float> envf(plen);
  for (int i = 0; i < plen; ++i) envf[i] = (float)env[i];
  float* d_env = nullptr;
  if (cudaMalloc(&d_env, sizeof(float) * plen) != cudaSuccess ||
      cudaMemcpy(d_env, envf.data(), sizeof(float) * plen, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("mdx plan upload failed");
    return nullptr;
  }
  MdxPlan* p = new MdxPlan{g, W, fft, d_env};
  g_mdx_plans[key] = p;
  return p;
}

template <typename T>
__device__ __forceinline__ void store_spec4(T* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store_spec4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store_spec4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 v;
  v.x = *reinterpret_cast<uint32_t*>(&lo);
  v.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = v;
}
template <>
__device__ __forceinline__ void store_spec4<__half>(__half* p, float a, float b, float c, float d) {
  uint2 v;
  v.x = pack2<kFmtF16>(a, b);
  v.y = pack2<kFmtF16>(c, d);
  *reinterpret_cast<uint2*>(p) = v;
}
__device__ __forceinline__ float4 load_spec4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load_spec4(const __half* p) {
  uint2 v = *reinterpret_cast<const uint2*>(p);
  float2 a = unpack2<kFmtF16>(v.x), b = unpack2<kFmtF16>(v.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 load_spec4(const __nv_bfloat16* p) {
  uint2 v = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162*>(&v.x), hi = *reinterpret_cast<__nv_bfloat162*>(&v.y);
  float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename T, bool kInplace>
__global__ void __launch_bounds__(kFftThreads, kInplace ? 2 : 1) stft_mdx_kernel(const float* __restrict__ src, long long ch_stride,
                                                               int n_ch, const WinDesc* __restrict__ wins, FftDev fft,
                                                               const float* __restrict__ hann, int hop, int dim_f,
                                                               int dim_t, int W, T* __restrict__ spec) {
  extern __shared__ float2 smem_f2[];
  const int N = fft.n;
  float2* buf0 = smem_f2;
  float2* buf1 = smem_f2 + fpad(N) + 1;
  const int t = blockIdx.x;
  const WinDesc wd = wins[blockIdx.y];
  const float* s0 = src + wd.base;
  const float* s1 = src + (n_ch > 1 ? ch_stride : 0) + wd.base;
  const int p0 = t * hop - N / 2;
  for (int j = threadIdx.x; j < N; j += kFftThreads) {
    int p = p0 + j;
    p = p < 0 ? -p : p;
    p = p >= W ? 2 * (W - 1) - p : p;  // torch.stft center=True, pad_mode="reflect"
    float l = 0.f, r = 0.f;
    if (p >= wd.p_lo && p < wd.p_hi) {
      l = __ldg(s0 + p);
      r = __ldg(s1 + p);
    }
    const float w = __ldg(hann + j);
    buf0[fpad(j)] = make_float2(l * w, r * w);
  }
  __syncthreads();
  const float2* Z;
  if constexpr (kInplace) {
    fft_smem_inplace<false>(buf0, fft);
    Z = buf0;
  } else {
    Z = fft_smem<false>(buf0, buf1, fft);
  }
  T* out = spec + ((size_t)blockIdx.y * dim_t + t) * (size_t)dim_f * 4;
  for (int k = threadIdx.x; k < dim_f; k += kFftThreads) {
    const float2 a = Z[fpad(k)];
    const float2 b = Z[fpad(k == 0 ? 0 : N - k)];
    store_spec4<T>(out + (size_t)k * 4, 0.5f * (a.x + b.x), 0.5f * (a.y - b.y), 0.5f * (a.y + b.y),
                   -0.5f * (a.x - b.x));
  }
}

struct IstftArgs {
  const WinDesc* wins;
  FftDev fft;
  const float* hann;
  const float* env;
  int hop, dim_f, dim_t, W;
  int nb;              // ring blocks = ceil(n_fft/hop)
  int blk_lo, blk_hi;  // hop-blocks (padded coordinates) to emit
  int strip;           // blocks per CTA
  int mode;
  float* wave;  // mode 0
  const float* mix;  // mode 1
  long long mix_stride;
  int n_ch;
  int output_is_vocal;
  float* vocal;
  float* instr;
  float* weight;
  float* side_vocal;  // nullable: per-chunk vocal before halo trimming
};

template <typename T>
__global__ void __launch_bounds__(kIstftThreads) istft_mdx_kernel(const T* __restrict__ spec, IstftArgs a) {
  extern __shared__ float2 smem_f2[];
  const int N = a.fft.n;
  float2* buf0 = smem_f2;
  float2* buf1 = buf0 + fpad(N) + 1;
  float2* ring = buf1 + fpad(N) + 1;  // [nb][hop]
  const int hop = a.hop;
  const int e0 = a.blk_lo + blockIdx.x * a.strip;
  const int e1 = min(a.blk_hi, e0 + a.strip);
  if (e0 >= e1) return;
  const WinDesc wd = a.wins[blockIdx.y];
  const int half = N / 2;
  const float inv_n = 1.0f / (float)N;
  for (int i = threadIdx.x; i < a.nb * hop; i += kIstftThreads) ring[i] = make_float2(0.f, 0.f);
  const int t_first = max(0, e0 - a.nb + 1);
  for (int t = t_first; t < e1; ++t) {
    if (t < a.dim_t) {
      // ---- build the full-length spectrum of z = L + iR from the kept bins (Hermitian extension)
      const T* in = spec + ((size_t)blockIdx.y * a.dim_t + t) * (size_t)a.dim_f * 4;
      for (int k = a.dim_f + threadIdx.x; k <= N - a.dim_f; k += kIstftThreads) buf0[fpad(k)] = make_float2(0.f, 0.f);
      for (int k = threadIdx.x; k < a.dim_f; k += kIstftThreads) {
        const float4 v = load_spec4(in + (size_t)k * 4);  // L_re, L_im, R_re, R_im
        if (k == 0) {
          buf0[0] = make_float2(v.x, v.z);  // c2r ignores the imaginary part of DC
        } else {
          buf0[fpad(k)] = make_float2(v.x - v.w, v.y + v.z);
          buf0[fpad(N - k)] = make_float2(v.x + v.w, v.z - v.y);
        }
      }
      __syncthreads();
      const float2* z = fft_smem<true>(buf0, buf1, a.fft);
      // ---- windowed overlap-add into the ring (frame t covers padded positions [t*hop, t*hop+N))
      for (int j = threadIdx.x; j < N; j += kIstftThreads) {
        const float w = __ldg(a.hann + j) * inv_n;
        const float2 v = z[fpad(j)];
        const int blk = t + j / hop;
        float2* dst = ring + (blk % a.nb) * hop + (j % hop);
        float2 acc = *dst;
        acc.x = fmaf(v.x, w, acc.x);
        acc.y = fmaf(v.y, w, acc.y);
        *dst = acc;
      }
      __syncthreads();
    }
    // ---- hop-block t is complete: emit (or discard during warm-up) and recycle its slot
    float2* blk = ring + (t % a.nb) * hop;
    if (t >= e0) {
      for (int i = threadIdx.x; i < hop; i += kIstftThreads) {
        const int pos = t * hop + i;
        const int n = pos - half;  // sample index in the torch.istft output
        if (n < 0 || n >= a.W) continue;
        const float e = __ldg(a.env + pos);
        const float2 acc = blk[i];
        const float y0 = acc.x / e, y1 = acc.y / e;
        if (a.mode == 0) {
          float* w0 = a.wave + (size_t)blockIdx.y * 2 * a.W;
          w0[n] = y0;
          w0[a.W + n] = y1;
        } else {
          const int o = n - half;  // trim = n_fft/2 on both sides (backends.py:377)
          if (o < 0 || o >= wd.out_len || n >= a.W - half) continue;
          const long long tp = wd.out_base + o;
          const bool in_eff = tp >= wd.eff_start && tp < wd.eff_end;
          if (!in_eff && !a.side_vocal) continue;
          const float m0 = __ldg(a.mix + tp);
          const float m1 = __ldg(a.mix + (a.n_ch > 1 ? a.mix_stride : 0) + tp);
          float v, ins;
          if (a.output_is_vocal) {
            v = (y0 + y1) * 0.5f;
            ins = ((m0 - y0) + (m1 - y1)) * 0.5f;
          } else {
            ins = (y0 + y1) * 0.5f;
            v = ((m0 - y0) + (m1 - y1)) * 0.5f;
          }
          if (a.side_vocal) a.side_vocal[wd.side_base + o] = v;
          if (!in_eff) continue;
          atomicAdd(a.vocal + tp, v);
          atomicAdd(a.instr + tp, ins);
          atomicAdd(a.weight + tp, 1.0f);
        }
      }
    }
    for (int i = threadIdx.x; i < hop; i += kIstftThreads) blk[i] = make_float2(0.f, 0.f);
    __syncthreads();
  }
}

// ---- three-pass kernels for n_fft = 7680 / 6144 (fft3.cuh) -----------------------------------------------------------
// Same arithmetic as the kernels above around a different FFT: the frame is loaded from the track (STFT) / built from the
// kept bins (iSTFT) by pass 1 itself, and the iSTFT's windowed overlap-add is pass 3's sink.  One 61 KB buffer.
template <typename T, int N>
__global__ void __launch_bounds__(kFft3Threads, 2) stft_mdx3_kernel(const float* __restrict__ src, long long ch_stride, int n_ch,
                                                                    const WinDesc* __restrict__ wins, Fft3Tw tw,
                                                                    const float* __restrict__ hann, int hop, int dim_f, int dim_t,
                                                                    int W, T* __restrict__ spec) {
  extern __shared__ float2 smem_f2[];
  float2* buf = smem_f2;
  constexpr int nb3 = N / Fft3Geom<N>::R3;
  const int t = blockIdx.x;
  const WinDesc wd = wins[blockIdx.y];
  const float* s0 = src + wd.base;
  const float* s1 = src + (n_ch > 1 ? ch_stride : 0) + wd.base;
  const int p0 = t * hop - N / 2;
  auto load = [&](int m) {
    int p = p0 + m;
    p = p < 0 ? -p : p;
    p = p >= W ? 2 * (W - 1) - p : p;  // torch.stft center=True, pad_mode="reflect"
    float l = 0.f, r = 0.f;
    if (p >= wd.p_lo && p < wd.p_hi) {
      l = __ldg(s0 + p);
      r = __ldg(s1 + p);
    }
    const float w = __ldg(hann + m);
    return make_float2(l * w, r * w);
  };
  float2* zrow = buf + threadIdx.x + (threadIdx.x >> 5);
  auto sink = [&](int, int r, float2 v) { zrow[r * (nb3 + nb3 / 32)] = v; };
  fft3_run<N, false>(buf, tw, load, sink);
  __syncthreads();
  T* out = spec + ((size_t)blockIdx.y * dim_t + t) * (size_t)dim_f * 4;
  for (int k = threadIdx.x; k < dim_f; k += kFft3Threads) {
    const float2 a = buf[fpad(k)];
    const float2 b = buf[fpad(k == 0 ? 0 : N - k)];
    store_spec4<T>(out + (size_t)k * 4, 0.5f * (a.x + b.x), 0.5f * (a.y - b.y), 0.5f * (a.y + b.y),
                   -0.5f * (a.x - b.x));
  }
}

// STFT + the network's first 1x1 conv: the frame's 4 spectrogram values per bin never go to HBM as a [f][4] tensor; each
// thread turns its bins into the 48-channel CG8 row the level-0 conv chain reads (6 x 16-byte stores per bin, lanes walk f).
template <int FMT, int N>
__global__ void __launch_bounds__(kFft3Threads, 2) stft_mdx3_first_kernel(const float* __restrict__ src, long long ch_stride, int n_ch,
                                                                          const WinDesc* __restrict__ wins, Fft3Tw tw,
                                                                          const float* __restrict__ hann, int hop, int dim_f, int dim_t,
                                                                          int W, h16* __restrict__ out, const float* __restrict__ cw,
                                                                          const float* __restrict__ cscale,
                                                                          const float* __restrict__ cshift) {
  constexpr int G = 48;
  extern __shared__ float2 smem_f2[];
  float2* buf = smem_f2;
  float* sw = reinterpret_cast<float*>(buf + ((fpad(N) + 2) & ~1));  // 16-byte aligned: [G][4] weights, [G] scale, [G] shift
  float* ssc = sw + G * 4;
  float* ssh = ssc + G;
  constexpr int nb3 = N / Fft3Geom<N>::R3;
  for (int i = threadIdx.x; i < G * 4; i += kFft3Threads) sw[i] = cw[i];
  for (int i = threadIdx.x; i < G; i += kFft3Threads) {
    ssc[i] = cscale[i];
    ssh[i] = cshift[i];
  }
  const int t = blockIdx.x;
  const WinDesc wd = wins[blockIdx.y];
  const float* s0 = src + wd.base;
  const float* s1 = src + (n_ch > 1 ? ch_stride : 0) + wd.base;
  const int p0 = t * hop - N / 2;
  auto load = [&](int m) {
    int p = p0 + m;
    p = p < 0 ? -p : p;
    p = p >= W ? 2 * (W - 1) - p : p;  // torch.stft center=True, pad_mode="reflect"
    float l = 0.f, r = 0.f;
    if (p >= wd.p_lo && p < wd.p_hi) {
      l = __ldg(s0 + p);
      r = __ldg(s1 + p);
    }
    const float w = __ldg(hann + m);
    return make_float2(l * w, r * w);
  };
  float2* zrow = buf + threadIdx.x + (threadIdx.x >> 5);
  auto sink = [&](int, int r, float2 v) { zrow[r * (nb3 + nb3 / 32)] = v; };
  fft3_run<N, false>(buf, tw, load, sink);
  __syncthreads();
  // row = (window, t); CG8: out[((row * G/8 + cg) * F + f) * 8 + c%8].  Three bins per thread and round share every weight /
  // scale / shift load (the conv is ~2600 instructions per thread and frame, more than the FFT itself).
  h16* orow = out + ((size_t)blockIdx.y * dim_t + t) * (size_t)(G / 8) * dim_f * 8;
  for (int k0 = threadIdx.x; k0 < dim_f; k0 += 3 * kFft3Threads) {
    float2 x01[3], x23[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int k = k0 + j * kFft3Threads;
      const int kk = k < dim_f ? k : 0;
      const float2 a = buf[fpad(kk)];
      const float2 b = buf[fpad(kk == 0 ? 0 : N - kk)];
      // the spectrogram values exactly as the [f][4] store would round them
      x01[j] = unpack2<FMT>(pack2<FMT>(0.5f * (a.x + b.x), 0.5f * (a.y - b.y)));
      x23[j] = unpack2<FMT>(pack2<FMT>(0.5f * (a.y + b.y), -0.5f * (a.x - b.x)));
    }
#pragma unroll 1  // (unrolled, the 48-channel body spills under the 64-register budget of two CTAs per SM)
    for (int cg = 0; cg < G / 8; ++cg) {
      uint32_t pk[3][4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v[3][2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = cg * 8 + 2 * e + h;
          const float4 wv = *reinterpret_cast<const float4*>(sw + c * 4);
          const float sc = ssc[c], sh = ssh[c];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            float s = wv.x * x01[j].x;
            s = fmaf(wv.y, x01[j].y, s);
            s = fmaf(wv.z, x23[j].x, s);
            s = fmaf(wv.w, x23[j].y, s);
            v[j][h] = fmaxf(fmaf(s, sc, sh), 0.f);
          }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) pk[j][e] = pack2<FMT>(v[j][0], v[j][1]);
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int k = k0 + j * kFft3Threads;
        if (k < dim_f)
          *reinterpret_cast<uint4*>(orow + ((size_t)cg * dim_f + k) * 8) = make_uint4(pk[j][0], pk[j][1], pk[j][2], pk[j][3]);
      }
    }
  }
}

template <typename T, int N>
__global__ void __launch_bounds__(kFft3Threads, 1) istft_mdx3_kernel(const T* __restrict__ spec, IstftArgs a, Fft3Tw tw,
                                                                     unsigned hop_magic) {
  extern __shared__ float2 smem_f2[];
  float2* buf = smem_f2;
  float2* ring = buf + fpad(N) + 1;  // [nb][hop]
  constexpr int nb3 = N / Fft3Geom<N>::R3;
  const int hop = a.hop;
  const int e0 = a.blk_lo + blockIdx.x * a.strip;
  const int e1 = min(a.blk_hi, e0 + a.strip);
  if (e0 >= e1) return;
  const WinDesc wd = a.wins[blockIdx.y];
  const int half = N / 2;
  const float inv_n = 1.0f / (float)N;
  for (int i = threadIdx.x; i < a.nb * hop; i += kFft3Threads) ring[i] = make_float2(0.f, 0.f);
  __syncthreads();
  const int t_first = max(0, e0 - a.nb + 1);
  int tmod = t_first % a.nb;  // t mod nb
  for (int t = t_first; t < e1; ++t) {
    if (t < a.dim_t) {
      // ---- pass 1 builds the full-length spectrum of z = L + iR from the kept bins (Hermitian extension) on the fly
      const T* in = spec + ((size_t)blockIdx.y * a.dim_t + t) * (size_t)a.dim_f * 4;
      auto load = [&](int m) {
        const int k = m < half ? m : N - m;
        if (k >= a.dim_f) return make_float2(0.f, 0.f);
        const float4 v = load_spec4(in + (size_t)k * 4);  // L_re, L_im, R_re, R_im
        if (m == 0) return make_float2(v.x, v.z);          // c2r ignores the imaginary part of DC
        return m < half ? make_float2(v.x - v.w, v.y + v.z) : make_float2(v.x + v.w, v.z - v.y);
      };
      // ---- pass 3 hands every output sample to the windowed overlap-add (frame t covers padded positions [t*hop, t*hop+N))
      auto sink = [&](int j, int r, float2 v) {
        const int m = j + r * nb3;
        const float w = __ldg(a.hann + m) * inv_n;
        const int q = (int)__umulhi((unsigned)m, hop_magic);  // m / hop
        int blk = tmod + q;
        if (blk >= a.nb) blk -= a.nb;
        float2* dst = ring + blk * hop + (m - q * hop);
        float2 acc = *dst;
        acc.x = fmaf(v.x, w, acc.x);
        acc.y = fmaf(v.y, w, acc.y);
        *dst = acc;
      };
      fft3_run<N, true>(buf, tw, load, sink);
      __syncthreads();
    }
    // ---- hop-block t is complete: emit (or discard during warm-up) and recycle its slot
    float2* blk = ring + tmod * hop;
    if (t >= e0) {
      for (int i = threadIdx.x; i < hop; i += kFft3Threads) {
        const int pos = t * hop + i;
        const int n = pos - half;  // sample index in the torch.istft output
        if (n < 0 || n >= a.W) continue;
        const float e = __ldg(a.env + pos);
        const float2 acc = blk[i];
        const float y0 = acc.x / e, y1 = acc.y / e;
        if (a.mode == 0) {
          float* w0 = a.wave + (size_t)blockIdx.y * 2 * a.W;
          w0[n] = y0;
          w0[a.W + n] = y1;
        } else {
          const int o = n - half;  // trim = n_fft/2 on both sides (backends.py:377)
          if (o < 0 || o >= wd.out_len || n >= a.W - half) continue;
          const long long tp = wd.out_base + o;
          const bool in_eff = tp >= wd.eff_start && tp < wd.eff_end;
          if (!in_eff && !a.side_vocal) continue;
          const float m0 = __ldg(a.mix + tp);
          const float m1 = __ldg(a.mix + (a.n_ch > 1 ? a.mix_stride : 0) + tp);
          float v, ins;
          if (a.output_is_vocal) {
            v = (y0 + y1) * 0.5f;
            ins = ((m0 - y0) + (m1 - y1)) * 0.5f;
          } else {
            ins = (y0 + y1) * 0.5f;
            v = ((m0 - y0) + (m1 - y1)) * 0.5f;
          }
          if (a.side_vocal) a.side_vocal[wd.side_base + o] = v;
          if (!in_eff) continue;
          atomicAdd(a.vocal + tp, v);
          atomicAdd(a.instr + tp, ins);
          atomicAdd(a.weight + tp, 1.0f);
        }
      }
    }
    for (int i = threadIdx.x; i < hop; i += kFft3Threads) blk[i] = make_float2(0.f, 0.f);
    __syncthreads();
    if (++tmod == a.nb) tmod = 0;
  }
}

// dev hook: AC_NO_FFT3=1 keeps the generic mixed-radix kernels for every size (A/B timing)
static bool use_fft3(const MdxPlan* plan) {
  static const bool off = getenv("AC_NO_FFT3") && atoi(getenv("AC_NO_FFT3")) != 0;
  return !off && fft3_supported(plan->g.n_fft) && plan->fft->d_tw3_p2 && plan->fft->d_tw3_p3 && plan->g.hop <= 32768;
}

bool stft_first_conv_supported(const MdxPlan* plan, int g, int dtype) {
  return use_fft3(plan) && g == 48 && (dtype == AC_F16 || dtype == AC_BF16);
}

int launch_stft_first_conv(const MdxPlan* plan, const float* d_src, long long ch_stride, int n_ch, const WinDesc* d_wins, int n_win,
                           void* d_cg8, int g, const float* d_w, const float* d_scale, const float* d_shift, int dtype,
                           cudaStream_t st) {
  if (n_win <= 0) return AC_OK;
  if (!stft_first_conv_supported(plan, g, dtype)) return AC_E_INVALID;
  const ac_mdx_geom& gm = plan->g;
  const Fft3Tw tw{plan->fft->d_tw3_p2, plan->fft->d_tw3_p3};
  const size_t smem = sizeof(float2) * fft_smem_floats2(gm.n_fft) + sizeof(float) * (48 * 4 + 2 * 48) + 16;
  dim3 grid(gm.dim_t, n_win);
  // algorithmic bytes: the window's samples in, the 48-channel 16-bit tensor out
  ProfScope ps(KC_STFT, 2.0 * n_win * (double)gm.dim_t * gm.dim_f * g * 4,
               n_win * (2.0 * plan->W * 4 + (double)gm.dim_t * gm.dim_f * g * 2.0), st);
#define AC_STFT3F_LAUNCH(FMT, NN)                                                                                                 \
  do {                                                                                                                            \
    AC_CHECK_CUDA(cudaFuncSetAttribute(stft_mdx3_first_kernel<FMT, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    stft_mdx3_first_kernel<FMT, NN><<<grid, kFft3Threads, smem, st>>>(d_src, ch_stride, n_ch, d_wins, tw, plan->fft->d_hann, gm.hop,  \
                                                                     gm.dim_f, gm.dim_t, plan->W, (h16*)d_cg8, d_w, d_scale, d_shift); \
  } while (0)
  if (dtype == AC_F16) {
    if (gm.n_fft == 7680) AC_STFT3F_LAUNCH(kFmtF16, 7680); else AC_STFT3F_LAUNCH(kFmtF16, 6144);
  } else {
    if (gm.n_fft == 7680) AC_STFT3F_LAUNCH(kFmtBF16, 7680); else AC_STFT3F_LAUNCH(kFmtBF16, 6144);
  }
#undef AC_STFT3F_LAUNCH
  AC_LAUNCH_CHECK();
  return AC_OK;
}

static size_t stft_smem_bytes(int n_fft) { return sizeof(float2) * 2 * (fft_smem_floats2(n_fft)); }

int launch_stft(const MdxPlan* plan, const float* d_src, long long ch_stride, int n_ch, const WinDesc* d_wins,
                int n_win, void* d_spec, int dtype, cudaStream_t st) {
  if (n_win <= 0) return AC_OK;
  const ac_mdx_geom& g = plan->g;
  const bool inplace = fft_inplace_ok(plan->fft, kFftThreads);
  const size_t smem = inplace ? stft_smem_bytes(g.n_fft) / 2 : stft_smem_bytes(g.n_fft);
  AC_REQUIRE(smem <= 227 * 1024, "n_fft too large for shared memory");
  FftDev fd = make_fft_dev(plan->fft);
  dim3 grid(g.dim_t, n_win);
  const double es = dtype == AC_F32 ? 4.0 : 2.0;
  ProfScope ps(KC_STFT, 0.0, n_win * (2.0 * plan->W * 4 + (double)g.dim_t * g.dim_f * 4 * es), st);
  if (use_fft3(plan)) {
    const Fft3Tw tw{plan->fft->d_tw3_p2, plan->fft->d_tw3_p3};
    const size_t smem3 = sizeof(float2) * fft_smem_floats2(g.n_fft);
#define AC_STFT3_LAUNCH(TYPE, NN)                                                                                             \
  do {                                                                                                                        \
    AC_CHECK_CUDA(cudaFuncSetAttribute(stft_mdx3_kernel<TYPE, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3)); \
    stft_mdx3_kernel<TYPE, NN><<<grid, kFft3Threads, smem3, st>>>(d_src, ch_stride, n_ch, d_wins, tw, plan->fft->d_hann, g.hop, \
                                                                  g.dim_f, g.dim_t, plan->W, (TYPE*)d_spec);                   \
  } while (0)
#define AC_STFT3_TYPE(TYPE)                                  \
  do {                                                       \
    if (g.n_fft == 7680) AC_STFT3_LAUNCH(TYPE, 7680);        \
    else AC_STFT3_LAUNCH(TYPE, 6144);                        \
  } while (0)
    if (dtype == AC_F32) AC_STFT3_TYPE(float);
    else if (dtype == AC_F16) AC_STFT3_TYPE(__half);
    else AC_STFT3_TYPE(__nv_bfloat16);
#undef AC_STFT3_TYPE
#undef AC_STFT3_LAUNCH
    AC_LAUNCH_CHECK();
    return AC_OK;
  }
#define AC_STFT_LAUNCH(TYPE, INPLACE)                                                                                        \
  do {                                                                                                                       \
    AC_CHECK_CUDA(cudaFuncSetAttribute(stft_mdx_kernel<TYPE, INPLACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    stft_mdx_kernel<TYPE, INPLACE><<<grid, kFftThreads, smem, st>>>(d_src, ch_stride, n_ch, d_wins, fd, plan->fft->d_hann, g.hop, \
                                                                    g.dim_f, g.dim_t, plan->W, (TYPE*)d_spec);                 \
  } while (0)
  if (dtype == AC_F32) {
    if (inplace) AC_STFT_LAUNCH(float, true); else AC_STFT_LAUNCH(float, false);
  } else if (dtype == AC_F16) {
    if (inplace) AC_STFT_LAUNCH(__half, true); else AC_STFT_LAUNCH(__half, false);
  } else {
    if (inplace) AC_STFT_LAUNCH(__nv_bfloat16, true); else AC_STFT_LAUNCH(__nv_bfloat16, false);
  }
#undef AC_STFT_LAUNCH
  AC_LAUNCH_CHECK();
  return AC_OK;
}

int launch_istft(const MdxPlan* plan, const void* d_spec, int dtype, const WinDesc* d_wins, int n_win, int mode,
                 float* d_wave, const float* d_mix, long long mix_stride, int n_ch, int output_is_vocal,
                 float* d_vocal, float* d_instr, float* d_weight, cudaStream_t st, float* d_chunk_vocal) {
  if (n_win <= 0) return AC_OK;
  const ac_mdx_geom& g = plan->g;
  IstftArgs a;
  a.wins = d_wins;
  a.fft = make_fft_dev(plan->fft);
  a.hann = plan->fft->d_hann;
  a.env = plan->d_env;
  a.hop = g.hop;
  a.dim_f = g.dim_f;
  a.dim_t = g.dim_t;
  a.W = plan->W;
  a.nb = (g.n_fft + g.hop - 1) / g.hop;
  const int half = g.n_fft / 2;
  const int pos_lo = mode == 0 ? half : g.n_fft;        // first padded position needed
  const int pos_hi = mode == 0 ? plan->W + half : plan->W;  // one past the last
  a.blk_lo = pos_lo / g.hop;
  a.blk_hi = (pos_hi + g.hop - 1) / g.hop;
  const int n_blk = a.blk_hi - a.blk_lo;
  // One CTA per SM (the kernel needs > 113 KB of shared memory): fill the machine with exactly one wave,
  // strips no shorter than 2x the nb-1 warm-up frames every strip recomputes.
  int strips = device_sm_count() / n_win;
  int max_strips = n_blk / (2 * a.nb);
  if (max_strips < 1) max_strips = 1;
  if (strips > max_strips) strips = max_strips;
  if (strips < 1) strips = 1;
  a.strip = (n_blk + strips - 1) / strips;
  strips = (n_blk + a.strip - 1) / a.strip;
  a.mode = mode;
  a.wave = d_wave;
  a.mix = d_mix;
  a.mix_stride = mix_stride;
  a.n_ch = n_ch;
  a.output_is_vocal = output_is_vocal;
  a.vocal = d_vocal;
  a.instr = d_instr;
  a.weight = d_weight;
  a.side_vocal = mode == 1 ? d_chunk_vocal : nullptr;
  const size_t smem = stft_smem_bytes(g.n_fft) + sizeof(float2) * (size_t)a.nb * g.hop;
  AC_REQUIRE(smem <= 227 * 1024, "n_fft too large for shared memory");
  dim3 grid(strips, n_win);
  const double es = dtype == AC_F32 ? 4.0 : 2.0;
  const double gen = plan->W - g.n_fft;
  // read spec + (stems: read mix 2ch, read-modify-write 3 accumulators | raw: write 2ch wave)
  ProfScope ps(KC_ISTFT, 0.0, n_win * ((double)g.dim_t * g.dim_f * 4 * es + (mode == 1 ? gen * (8 + 24) : plan->W * 8.0)), st);
  if (use_fft3(plan)) {
    const Fft3Tw tw{plan->fft->d_tw3_p2, plan->fft->d_tw3_p3};
    const size_t smem3 = sizeof(float2) * (fft_smem_floats2(g.n_fft) + (size_t)a.nb * g.hop);
    AC_REQUIRE(smem3 <= 227 * 1024, "n_fft too large for shared memory");
    const unsigned hop_magic = (unsigned)((0x100000000ull + (unsigned)g.hop - 1) / (unsigned)g.hop);  // m / hop = umulhi(m, magic), m < 2^16
#define AC_ISTFT3_LAUNCH(TYPE, NN)                                                                                             \
  do {                                                                                                                         \
    AC_CHECK_CUDA(cudaFuncSetAttribute(istft_mdx3_kernel<TYPE, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3)); \
    istft_mdx3_kernel<TYPE, NN><<<grid, kFft3Threads, smem3, st>>>((const TYPE*)d_spec, a, tw, hop_magic);                     \
  } while (0)
#define AC_ISTFT3_TYPE(TYPE)                                  \
  do {                                                        \
    if (g.n_fft == 7680) AC_ISTFT3_LAUNCH(TYPE, 7680);        \
    else AC_ISTFT3_LAUNCH(TYPE, 6144);                        \
  } while (0)
    if (dtype == AC_F32) AC_ISTFT3_TYPE(float);
    else if (dtype == AC_F16) AC_ISTFT3_TYPE(__half);
    else AC_ISTFT3_TYPE(__nv_bfloat16);
#undef AC_ISTFT3_TYPE
#undef AC_ISTFT3_LAUNCH
    AC_LAUNCH_CHECK();
    return AC_OK;
  }
  if (dtype == AC_F32) {
    AC_CHECK_CUDA(cudaFuncSetAttribute(istft_mdx_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    istft_mdx_kernel<float><<<grid, kIstftThreads, smem, st>>>((const float*)d_spec, a);
  } else if (dtype == AC_F16) {
    AC_CHECK_CUDA(cudaFuncSetAttribute(istft_mdx_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    istft_mdx_kernel<__half><<<grid, kIstftThreads, smem, st>>>((const __half*)d_spec, a);
  } else {
    AC_CHECK_CUDA(
        cudaFuncSetAttribute(istft_mdx_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    istft_mdx_kernel<__nv_bfloat16><<<grid, kIstftThreads, smem, st>>>((const __nv_bfloat16*)d_spec, a);
  }
  AC_LAUNCH_CHECK();
  return AC_OK;
}

}  // namespace ac

// ---- C ABI: standalone batch forms ([B][2][W] waves) ------------------------------------------
namespace {
struct WinCache {  // identity window descriptors for the [B][2][W] layout, grown on demand
  ac::WinDesc* d = nullptr;
  int cap = 0;
  int W = 0;
};
WinCache g_wc;
std::mutex g_wc_mu;

const ac::WinDesc* batch_windows(int B, int W) {
  std::lock_guard<std::mutex> lk(g_wc_mu);
  if (g_wc.cap < B || g_wc.W != W) {
    if (g_wc.d) cudaFree(g_wc.d);
    int cap = B < 64 ? 64 : B;
    std::vector<ac::WinDesc> h(cap);
    for (int b = 0; b < cap; ++b) {
      h[b] = ac::WinDesc{(long long)b * 2 * W, 0, 0, 0, 0, W, 0, 0};
    }
    if (cudaMalloc(&g_wc.d, sizeof(ac::WinDesc) * cap) != cudaSuccess ||
        cudaMemcpy(g_wc.d, h.data(), sizeof(ac::WinDesc) * cap, cudaMemcpyHostToDevice) != cudaSuccess) {
      g_wc = WinCache();
      ac::set_error("window descriptor upload failed");
      return nullptr;
    }
    g_wc.cap = cap;
    g_wc.W = W;
  }
  return g_wc.d;
}
}  // namespace

extern "C" int ac_stft_mdx(const float* d_wave, void* d_spec, int B, const ac_mdx_geom* g, int dtype, void* stream) {
  AC_REQUIRE(d_wave && d_spec && g && B >= 0, "null pointer");
  AC_REQUIRE(dtype == AC_F32 || dtype == AC_BF16 || dtype == AC_F16, "dtype");
  const ac::MdxPlan* plan = ac::get_mdx_plan(*g);
  if (!plan) return AC_E_INVALID;
  const ac::WinDesc* w = batch_windows(B, plan->W);
  if (!w) return AC_E_CUDA;
  return ac::launch_stft(plan, d_wave, plan->W, 2, w, B, d_spec, dtype, (cudaStream_t)stream);
}

extern "C" int ac_istft_mdx(const void* d_spec, float* d_wave, int B, const ac_mdx_geom* g, int dtype, void* stream) {
  AC_REQUIRE(d_wave && d_spec && g && B >= 0, "null pointer");
  AC_REQUIRE(dtype == AC_F32 || dtype == AC_BF16 || dtype == AC_F16, "dtype");
  const ac::MdxPlan* plan = ac::get_mdx_plan(*g);
  if (!plan) return AC_E_INVALID;
  const ac::WinDesc* w = batch_windows(B, plan->W);
  if (!w) return AC_E_CUDA;
  return ac::launch_istft(plan, d_spec, dtype, w, B, 0, d_wave, nullptr, 0, 2, 1, nullptr, nullptr, nullptr,
                          (cudaStream_t)stream);
}
