// Data formats either side of the hot path: integer PCM <-> float32 and the
// 44.1 kHz -> 16 kHz polyphase resampler in front of the per-chunk VAD.  All of it is byte / sample streaming:
// HBM-bound, one pass, coalesced 128-bit accesses where the layout allows.
//
//   in   AudioProcessor.load_audio      src/vocal_smart_splitter/utils/audio_processor.py:32-60
//        librosa.load(path, sr, mono=True) = decode (libsndfile) -> channel mean -> [resample] -> float32, then
//        audio / max|audio| (:54-56).  libsndfile's sf_read_float normalisation: PCM_16 / 2^15, PCM_24 / 2^23.
//   out  export_audio / _write_wav       src/vocal_smart_splitter/utils/audio_export.py:70-112
//        sf.write(path, audio, sr, subtype="PCM_24"): libsndfile pcm.c f2let_array, value = lrintf(x * 0x7FFFFF), the low
//        three bytes little endian (no clipping unless SFC_SET_CLIPPING: f2let_clip_array, scale 2^31, saturate, top
//        three bytes).  _write_mp3 (:113-133): clip to [-1,1], np.round(x * 32767) -> int16.
//   vad  VocalPauseDetectorV2._detect_speech_timestamps   src/vocal_smart_splitter/core/vocal_pause_detector.py:175-296
//        librosa.resample(audio, 44100 -> 16000) per chunk, zero-pad to a multiple of 4096.  librosa's default filter is
//        libsoxr's (absent from this image: parity unpinned); the kernel is librosa's res_type="polyphase" =
//        scipy.signal.resample_poly (zero-phase Kaiser-5.0 FIR, upfirdn), which the tests compare against scipy itself.
#include "common.cuh"

namespace ac {

// ---- PCM -> float -------------------------------------------------------------------------------
// bytes: interleaved little-endian frames [n][ch] of `bits` (16 / 24); out: planar [ch][n], or the channel mean [n] when
// mono != 0 (np.mean over channels in float32, as librosa.to_mono).
template <int BITS>
__global__ void __launch_bounds__(256) pcm_decode_kernel(const uint8_t* __restrict__ bytes, long long n, int ch, int mono,
                                                         float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int B = BITS / 8;
  const uint8_t* p = bytes + (size_t)i * ch * B;
  float acc = 0.f;
  for (int c = 0; c < ch; ++c) {
    int v;
    if (BITS == 16) v = (int)(int16_t)((uint16_t)p[0] | ((uint16_t)p[1] << 8));
    else v = ((int)(int8_t)p[2] << 16) | ((int)p[1] << 8) | (int)p[0];
    const float f = (float)v * (BITS == 16 ? (1.0f / 32768.0f) : (1.0f / 8388608.0f));
    if (mono) acc += f;
    else out[(size_t)c * n + i] = f;
    p += B;
  }
  if (mono) out[i] = ch == 1 ? acc : acc / (float)ch;
}

// ---- peak normalisation: x / max|x| ---------------------------------------------------------------
__global__ void __launch_bounds__(256) abs_max_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, sh[w]);
    atomicMax(out, __float_as_uint(m));  // non-negative floats order like their bit patterns
  }
}
__global__ void __launch_bounds__(256) scale_by_inv_max_kernel(float* __restrict__ x, long long n, const unsigned int* __restrict__ mx) {
  const float m = __uint_as_float(*mx);
  if (!(m > 0.f)) return;  // audio_processor.py:55: only when np.max(np.abs(audio)) > 0
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = x[i] / m;       // a true division, as numpy does (x * (1/m) differs in the last bit)
}

// ---- float -> PCM ---------------------------------------------------------------------------------
// in: planar [ch][n]; out: interleaved frames.  MODE 0: PCM_24 libsndfile default, 1: PCM_24 with clipping, 2: int16 (mp3 path)
template <int MODE>
__global__ void __launch_bounds__(256) pcm_pack_kernel(const float* __restrict__ x, long long n, int ch, uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int B = MODE == 2 ? 2 : 3;
  uint8_t* p = out + (size_t)i * ch * B;
  for (int c = 0; c < ch; ++c) {
    const float s = x[(size_t)c * n + i];
    if (MODE == 0) {
      // out-of-range products are undefined in C (lrintf of a value beyond int); x86's cvtss2si gives INT_MIN, kept here
      const float prod = s * 8388607.0f;
      const int v = (prod >= 2147483648.0f || prod < -2147483648.0f || prod != prod) ? (int)0x80000000 : __float2int_rn(prod);
      p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16);
    } else if (MODE == 1) {
      const float sc = s * 2147483648.0f;
      if (sc >= 2147483648.0f) { p[0] = 0xFF; p[1] = 0xFF; p[2] = 0x7F; }
      else if (sc <= -2147483648.0f) { p[0] = 0x00; p[1] = 0x00; p[2] = 0x80; }
      else { const int v = __float2int_rn(sc); p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 24); }
    } else {
      const float cl = fminf(fmaxf(s, -1.0f), 1.0f);
      const int v = __float2int_rn(cl * 32767.0f);  // np.round: half to even
      p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8);
    }
    p += B;
  }
}

// ---- polyphase resampler ----------------------------------------------------------------------------
// y[m] = sum_j h[(m + n_pre_remove) * down - n_pre_pad - j * up] * x[j]   (scipy.signal.resample_poly with zero padding);
// the segments of x (chunks, back to back) are independent signals; row s of the output holds segment s, zero-filled to
// row_len.  One thread per output sample: its ~h_len/up taps read x at consecutive j (coalesced across the warp, since
// neighbouring m advance j by down/up) and h at stride `up` (through the read-only cache; the taps are 35 KB).
struct ResampleSeg {
  long long in_off, in_len, out_len;
};
constexpr int kResampleMaxSegs = 128;
struct ResampleSegs {
  ResampleSeg s[kResampleMaxSegs];
};
__global__ void __launch_bounds__(256) resample_poly_kernel(const float* __restrict__ x, const __grid_constant__ ResampleSegs segs,
                                                            const float* __restrict__ h, int h_len, int up, int down, int n_pre_pad,
                                                            int n_pre_remove, long long row_len, float* __restrict__ out) {
  const ResampleSeg sg = segs.s[blockIdx.y];
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= row_len) return;
  float* dst = out + (size_t)blockIdx.y * row_len;
  if (m >= sg.out_len) { dst[m] = 0.f; return; }
  const long long T = (m + n_pre_remove) * (long long)down - n_pre_pad;  // h index for j = 0
  // taps: 0 <= T - j*up < h_len  ->  j in [ceil((T - h_len + 1)/up), floor(T/up)]
  long long j_hi = T >= 0 ? T / up : -1;
  long long lo_num = T - h_len + 1;
  long long j_lo = lo_num <= 0 ? 0 : (lo_num + up - 1) / up;
  if (j_hi >= sg.in_len) j_hi = sg.in_len - 1;
  const float* xs = x + sg.in_off;
  float acc = 0.f;
  for (long long j = j_lo; j <= j_hi; ++j) acc = fmaf(__ldg(h + (T - j * up)), xs[j], acc);
  dst[m] = acc;
}

}  // namespace ac

extern "C" int ac_pcm_decode(const void* d_bytes, long long n_frames, int channels, int bits, int mono, float* d_out, void* stream) {
  using namespace ac;
  AC_REQUIRE(d_bytes && d_out, "null pointer");
  AC_REQUIRE(n_frames >= 0 && channels >= 1 && channels <= 8, "bad sizes");
  AC_REQUIRE(bits == 16 || bits == 24, "bits must be 16 or 24");
  if (n_frames == 0) return AC_OK;
  const unsigned grid = (unsigned)((n_frames + 255) / 256);
  ProfScope ps(KC_MISC, 0.0, (double)n_frames * channels * (bits / 8) + 4.0 * n_frames * (mono ? 1 : channels), (cudaStream_t)stream);
  if (bits == 16) pcm_decode_kernel<16><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)d_bytes, n_frames, channels, mono, d_out);
  else pcm_decode_kernel<24><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)d_bytes, n_frames, channels, mono, d_out);
  AC_LAUNCH_CHECK();
  return AC_OK;
}

extern "C" int ac_peak_normalize(float* d_x, long long n, unsigned int* d_scratch, void* stream) {
  using namespace ac;
  AC_REQUIRE(d_x && d_scratch && n >= 0, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  AC_CHECK_CUDA(cudaMemsetAsync(d_scratch, 0, sizeof(unsigned int), st));
  if (n == 0) return AC_OK;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  ProfScope ps(KC_MISC, 0.0, 12.0 * (double)n, st);
  abs_max_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_x, n, d_scratch);
  AC_LAUNCH_CHECK();
  scale_by_inv_max_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_x, n, d_scratch);
  AC_LAUNCH_CHECK();
  return AC_OK;
}

extern "C" int ac_pcm_pack(const float* d_x, long long n_frames, int channels, int format, void* d_out, void* stream) {
  using namespace ac;
  AC_REQUIRE(d_x && d_out, "null pointer");
  AC_REQUIRE(n_frames >= 0 && channels >= 1 && channels <= 8, "bad sizes");
  AC_REQUIRE(format >= 0 && format <= 2, "format: 0 = PCM_24 (libsndfile default), 1 = PCM_24 clipped, 2 = int16");
  if (n_frames == 0) return AC_OK;
  const unsigned grid = (unsigned)((n_frames + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(KC_MISC, 0.0, (double)n_frames * channels * (4 + (format == 2 ? 2 : 3)), st);
  if (format == 0) pcm_pack_kernel<0><<<grid, 256, 0, st>>>(d_x, n_frames, channels, (uint8_t*)d_out);
  else if (format == 1) pcm_pack_kernel<1><<<grid, 256, 0, st>>>(d_x, n_frames, channels, (uint8_t*)d_out);
  else pcm_pack_kernel<2><<<grid, 256, 0, st>>>(d_x, n_frames, channels, (uint8_t*)d_out);
  AC_LAUNCH_CHECK();
  return AC_OK;
}

extern "C" long long ac_resample_out_len(long long n_in, int up, int down) {
  if (n_in < 0 || up <= 0 || down <= 0) return 0;
  const long long t = n_in * (long long)up;
  return t / down + (t % down ? 1 : 0);
}

extern "C" int ac_resample_poly(const float* d_x, const long long* h_seg_off, const long long* h_seg_len, int n_segs, int up, int down,
                                const float* d_taps, int n_taps, int n_pre_pad, int n_pre_remove, long long row_len, float* d_out,
                                void* stream) {
  using namespace ac;
  AC_REQUIRE(d_x && h_seg_off && h_seg_len && d_taps && d_out, "null pointer");
  AC_REQUIRE(n_segs >= 0 && up > 0 && down > 0 && n_taps > 0 && row_len > 0 && n_pre_pad >= 0 && n_pre_remove >= 0, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  double bytes = 0.0;
  for (int s = 0; s < n_segs; ++s) {
    AC_REQUIRE(h_seg_off[s] >= 0 && h_seg_len[s] >= 0, "bad segment");
    AC_REQUIRE(ac_resample_out_len(h_seg_len[s], up, down) <= row_len, "row_len shorter than a resampled segment");
    bytes += 4.0 * (double)h_seg_len[s];
  }
  ProfScope ps(KC_MISC, 2.0 * (double)n_segs * (double)row_len * ((double)n_taps / up), bytes + 4.0 * (double)n_segs * (double)row_len, st);
  for (int s0 = 0; s0 < n_segs; s0 += kResampleMaxSegs) {
    ResampleSegs rs;
    const int cnt = n_segs - s0 < kResampleMaxSegs ? n_segs - s0 : kResampleMaxSegs;
    for (int i = 0; i < cnt; ++i)
      rs.s[i] = ResampleSeg{h_seg_off[s0 + i], h_seg_len[s0 + i], ac_resample_out_len(h_seg_len[s0 + i], up, down)};
    const dim3 grid((unsigned)((row_len + 255) / 256), (unsigned)cnt);
    resample_poly_kernel<<<grid, 256, 0, st>>>(d_x, rs, d_taps, n_taps, up, down, n_pre_pad, n_pre_remove, row_len,
                                               d_out + (size_t)s0 * row_len);
    AC_LAUNCH_CHECK();
  }
  return AC_OK;
}
