// Framewise RMS and zero-crossing rate: one HBM pass, frames cut out of a shared-memory tile.
//
// librosa.feature.rms(y, frame_length, hop_length, center=True, pad_mode="constant"):
//   features_cache.py:182 (4410/2205), pure_vocal_pause_detector.py:1111-1113 (1102/441), :1397 (2048/441),
//   seamless_splitter.py:1714, :1848 (2048/441), vocal_separator.py:483 (2205/882).
// librosa.feature.zero_crossing_rate(y, frame_length, hop_length, center=True): pure_vocal_pause_detector.py:444.
//
// A CTA stages the samples of FR consecutive frames (one coalesced 128-bit sweep; each sample
// is read from HBM ~once even though frames overlap frame/hop times) and every warp reduces
// whole frames out of shared memory with lane-strided, conflict-free reads + a shuffle tree.
#include "common.cuh"

namespace ac {

constexpr int kRmsThreads = 256;
constexpr int kRmsMaxTileFloats = 24064;  // 94 KB -> two CTAs per SM

// One signal segment of a launch: frames are cut out of x[lo, hi) exactly as if that slice were the whole signal
// (zero / edge padding at ITS ends), results go to out[out_off ...].  Up to kRmsMaxSegs segments ride in the kernel
// parameters, so the per-chunk RMS of a whole track (features_cache.py:182, one librosa call per pipeline chunk) is one launch.
struct RmsSeg {
  long long lo, hi, out_off, n_frames;
};
constexpr int kRmsMaxSegs = 96;
struct RmsSegs {
  RmsSeg s[kRmsMaxSegs];
};

// MODE 0: rms (zero padding)   MODE 1: zero-crossing rate (edge padding, |y|<=1e-10 -> 0)
template <int MODE>
__global__ void __launch_bounds__(kRmsThreads) frame_reduce_kernel(const float* __restrict__ x, const __grid_constant__ RmsSegs segs,
                                                                   int frame, int hop, int pad, int frames_per_cta,
                                                                   float* __restrict__ out_base) {
  extern __shared__ float tile[];
  const RmsSeg sg = segs.s[blockIdx.y];
  const long long t0 = (long long)blockIdx.x * frames_per_cta;
  if (t0 >= sg.n_frames) return;
  const int nf = (int)min((long long)frames_per_cta, sg.n_frames - t0);
  float* __restrict__ out = out_base + sg.out_off;
  // absolute sample indices: the tile is [a0, a0 + 4*len4), a0 a multiple of 4 so that every 128-bit load is aligned
  const long long start = sg.lo + t0 * hop - pad;                  // first sample of the tile (may precede the segment)
  const long long end = start + (long long)(nf - 1) * hop + frame;  // one past the last
  const long long a0 = (start >= 0 ? start / 4 : -((-start + 3) / 4)) * 4;  // floor to a multiple of 4
  const int len4 = (int)((end - a0 + 3) / 4);
  const bool aligned = (((uintptr_t)x) & 15) == 0;
  if (aligned && a0 >= sg.lo && a0 + 4LL * len4 <= sg.hi) {
    // interior tile: 8 independent 128-bit loads in flight per thread (a one-load-at-a-time loop keeps only
    // ~8 KB per SM in flight, which caps the kernel near 1.5 TB/s)
    const float4* src = reinterpret_cast<const float4*>(x + a0);
    float4* dst = reinterpret_cast<float4*>(tile);
    for (int i = threadIdx.x; i < len4; i += kRmsThreads * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = i + u * kRmsThreads;
        if (k < len4) v[u] = ldg_stream_f4(src + k);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = i + u * kRmsThreads;
        if (k < len4) dst[k] = v[u];
      }
    }
  } else {
    for (int i = threadIdx.x; i < len4; i += kRmsThreads) {
      const long long idx = a0 + 4LL * i;
      float4 v;
      if (aligned && idx >= sg.lo && idx + 3 < sg.hi) {
        v = ldg_stream_f4(reinterpret_cast<const float4*>(x + idx));
      } else {
        float e[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          long long j = idx + k;
          if (MODE == 1) {
            j = j < sg.lo ? sg.lo : (j >= sg.hi ? sg.hi - 1 : j);
            e[k] = x[j];
          } else {
            e[k] = (j >= sg.lo && j < sg.hi) ? x[j] : 0.f;
          }
        }
        v = make_float4(e[0], e[1], e[2], e[3]);
      }
      reinterpret_cast<float4*>(tile)[i] = v;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int off0 = (int)(start - a0);
  for (int f = warp; f < nf; f += kRmsThreads / 32) {
    const float* fr = tile + off0 + f * hop;
    float acc = 0.f;
    if (MODE == 0) {
      for (int j = lane; j < frame; j += 32) {
        float v = fr[j];
        acc = fmaf(v, v, acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) out[t0 + f] = sqrtf(acc / (float)frame);
    } else {
      for (int j = lane + 1; j < frame; j += 32) {
        float a = fr[j - 1], b = fr[j];
        a = fabsf(a) <= 1e-10f ? 0.f : a;
        b = fabsf(b) <= 1e-10f ? 0.f : b;
        acc += (signbit(a) != signbit(b)) ? 1.f : 0.f;
      }
      acc = warp_sum(acc);
      if (lane == 0) out[t0 + f] = acc / (float)frame;
    }
  }
}

// segs: host array; every segment must lie inside [0, n)
template <int MODE>
static int launch_frame_reduce(const float* d_x, long long n, const ac_feat_segment* segs, int n_seg, int frame, int hop,
                               int center, float* d_out, cudaStream_t st) {
  AC_REQUIRE(d_x && d_out && segs, "null pointer");
  AC_REQUIRE(frame > 0 && hop > 0 && n > 0 && n_seg >= 0, "frame, hop and n must be positive");
  AC_REQUIRE(frame + 8 <= kRmsMaxTileFloats, "frame too long");
  long long total_frames = 0, max_frames = 0;
  double total_samples = 0.0;
  for (int i = 0; i < n_seg; ++i) {
    AC_REQUIRE(segs[i].start >= 0 && segs[i].len >= 0 && segs[i].start + segs[i].len <= n && segs[i].frame_off >= 0,
               "segment outside the signal");
    const long long f = ac_frame_count(segs[i].len, frame, hop, center);
    total_frames += f;
    total_samples += (double)segs[i].len;
    if (f > max_frames) max_frames = f;
  }
  if (total_frames <= 0) return AC_OK;
  // Tile size: small tiles put 3-4 CTAs on an SM, so one CTA's load phase overlaps the others' reduce phase (the
  // 94 KB tile, 2 CTAs per SM, ran at 2.8-3.4 TB/s; 32-46 KB tiles reach 3.9-4.9 TB/s on one hour of audio).  Frames
  // overlap by frame - hop samples, re-read once per tile: take the smallest tile that keeps that below ~26 %.
  const int caps[4] = {8192, 11776, 15600, kRmsMaxTileFloats};
  int fr = 1;
  for (int i = 0; i < 4; ++i) {
    if (caps[i] - 8 < frame) continue;
    fr = 1 + (caps[i] - 8 - frame) / hop;
    if ((double)(frame > hop ? frame - hop : 0) <= 0.26 * (double)fr * hop) break;
  }
  if (fr > 64) fr = 64;
  // keep at least ~4 CTAs per SM worth of work when the signal is short
  const long long want = 4LL * device_sm_count();
  while (fr > 8 && (total_frames + fr - 1) / fr < want) fr /= 2;
  const size_t smem = sizeof(float) * (size_t)((fr - 1) * hop + frame + 8);
  static bool attr_done[2] = {false, false};
  if (!attr_done[MODE]) {
    AC_CHECK_CUDA(cudaFuncSetAttribute(frame_reduce_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(sizeof(float) * kRmsMaxTileFloats)));
    attr_done[MODE] = true;
  }
  ProfScope ps(KC_RMS, 0.0, 4.0 * total_samples + 4.0 * (double)total_frames, st);
  for (int s0 = 0; s0 < n_seg; s0 += kRmsMaxSegs) {
    RmsSegs rs;
    const int cnt = n_seg - s0 < kRmsMaxSegs ? n_seg - s0 : kRmsMaxSegs;
    long long mx = 0;
    for (int i = 0; i < cnt; ++i) {
      const ac_feat_segment& g = segs[s0 + i];
      rs.s[i] = RmsSeg{g.start, g.start + g.len, g.frame_off, ac_frame_count(g.len, frame, hop, center)};
      if (rs.s[i].n_frames > mx) mx = rs.s[i].n_frames;
    }
    if (mx <= 0) continue;
    const dim3 grid((unsigned)((mx + fr - 1) / fr), (unsigned)cnt);
    frame_reduce_kernel<MODE><<<grid, kRmsThreads, smem, st>>>(d_x, rs, frame, hop, center ? frame / 2 : 0, fr, d_out);
    AC_LAUNCH_CHECK();
  }
  return AC_OK;
}

template <int MODE>
static int launch_frame_reduce_whole(const float* d_x, long long n, int frame, int hop, int center, float* d_out, cudaStream_t st) {
  AC_REQUIRE(n > 0, "frame, hop and n must be positive");
  const ac_feat_segment one{0, n, 0};
  return launch_frame_reduce<MODE>(d_x, n, &one, 1, frame, hop, center, d_out, st);
}

}  // namespace ac

extern "C" int ac_frame_rms(const float* d_x, long long n, int frame, int hop, int center, float* d_out, void* stream) {
  return ac::launch_frame_reduce_whole<0>(d_x, n, frame, hop, center, d_out, (cudaStream_t)stream);
}

extern "C" int ac_frame_rms_segments(const float* d_x, long long n, const ac_feat_segment* h_segs, int n_segs, int frame, int hop,
                                     int center, float* d_out, void* stream) {
  return ac::launch_frame_reduce<0>(d_x, n, h_segs, n_segs, frame, hop, center, d_out, (cudaStream_t)stream);
}

extern "C" int ac_zero_crossing_rate(const float* d_x, long long n, int frame, int hop, float* d_out, void* stream) {
  return ac::launch_frame_reduce_whole<1>(d_x, n, frame, hop, 1, d_out, (cudaStream_t)stream);
}
