// STFT-2048 framewise features: spectral flatness, mel-dB spectral flux (onset envelope, mean and
// median aggregates), spectral centroid and the low-third band ratio, from ONE pass over the signal.
//
//   librosa.feature.spectral_flatness(y, hop_length)   features_cache.py:183; pure_vocal_pause_detector.py:1117
//   librosa.onset.onset_strength(y, sr, hop_length)     features_cache.py:184; adaptive_vad_enhancer.py:61-67,143-148
//   librosa.feature.spectral_centroid                   pure_vocal_pause_detector.py:434
//   _calculate_harmonic_ratio_direct                    pure_vocal_pause_detector.py:937-959
//
// Pass 1 (one CTA per PAIR of frames): the two real frames are packed as re/im of one complex
// 2048-point shared-memory FFT, split, and reduced to flatness / centroid / band ratio; the
// 128-band slaney mel projection is a sparse gather over the power spectrum; mel dB rows go to
// the workspace and the segment maximum (the power_to_db top_db reference, which librosa takes
// over the whole call = one chunk) is folded in with an ordered-int atomicMax.
// Pass 2 (one warp per frame): clip at max-80 dB, rectified difference to the previous frame,
// mean and exact median (rank counting through shuffles) over the 128 bands.
#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "fft.cuh"

namespace ac {

constexpr int kNfft = 2048;
constexpr int kBins = kNfft / 2 + 1;
constexpr int kMels = 128;
constexpr int kFeatThreads = 256;

struct MelTable {
  int sr = 0;
  int* d_start = nullptr;   // [128] first bin of each filter
  int* d_count = nullptr;   // [128] bins in each filter
  int* d_woff = nullptr;    // [128] offset into d_w
  float* d_w = nullptr;     // concatenated non-zero weights
};
static MelTable g_mel;
static std::mutex g_mel_mu;

static double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// librosa.filters.mel(sr, n_fft=2048, n_mels=128, fmin=0, fmax=sr/2, htk=False, norm="slaney")
static const MelTable* get_mel_table(int sr) {
  std::lock_guard<std::mutex> lk(g_mel_mu);
  if (g_mel.sr == sr) return &g_mel;
  if (g_mel.sr != 0) {
    cudaFree(g_mel.d_start); cudaFree(g_mel.d_count); cudaFree(g_mel.d_woff); cudaFree(g_mel.d_w);
    g_mel = MelTable();
  }
  std::vector<double> mel_f(kMels + 2);
  const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(sr / 2.0);
  for (int i = 0; i < kMels + 2; ++i) mel_f[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (kMels + 1));
  std::vector<int> start(kMels), count(kMels), woff(kMels);
  std::vector<float> w;
  for (int m = 0; m < kMels; ++m) {
    const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
    int lo = -1, hi = -1;
    std::vector<float> row(kBins, 0.f);
    for (int k = 0; k < kBins; ++k) {
      const double fk = (sr / 2.0) * k / (kBins - 1);
      const double lower = (fk - mel_f[m]) / (mel_f[m + 1] - mel_f[m]);
      const double upper = (mel_f[m + 2] - fk) / (mel_f[m + 2] - mel_f[m + 1]);
      const double v = std::max(0.0, std::min(lower, upper)) * enorm;
      row[k] = (float)v;
      if (row[k] > 0.f) {
        if (lo < 0) lo = k;
        hi = k;
      }
    }
    start[m] = lo < 0 ? 0 : lo;
    count[m] = lo < 0 ? 0 : hi - lo + 1;
    woff[m] = (int)w.size();
    for (int k = 0; k < count[m]; ++k) w.push_back(row[start[m] + k]);
  }
  if (w.empty()) w.push_back(0.f);
  bool ok = cudaMalloc(&g_mel.d_start, sizeof(int) * kMels) == cudaSuccess &&
            cudaMalloc(&g_mel.d_count, sizeof(int) * kMels) == cudaSuccess &&
            cudaMalloc(&g_mel.d_woff, sizeof(int) * kMels) == cudaSuccess &&
            cudaMalloc(&g_mel.d_w, sizeof(float) * w.size()) == cudaSuccess &&
            cudaMemcpy(g_mel.d_start, start.data(), sizeof(int) * kMels, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(g_mel.d_count, count.data(), sizeof(int) * kMels, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(g_mel.d_woff, woff.data(), sizeof(int) * kMels, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(g_mel.d_w, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) {
    set_error("mel table upload failed");
    return nullptr;
  }
  g_mel.sr = sr;
  return &g_mel;
}

struct SegDev {
  long long start, len, frame_off;
  long long pair_off;  // first pair-CTA index of the segment
  int n_frames;
  int pad_;
};

__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void init_segmax_kernel(int* segmax, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) segmax[i] = float_to_ordered(-INFINITY);
}

struct FeatArgs {
  const float* x;
  const SegDev* segs;
  int n_segs;
  int hop;
  float sr;
  FftDev fft;
  const float* hann;
  const int* mel_start;
  const int* mel_count;
  const int* mel_woff;
  const float* mel_w;
  float* flatness;
  float* centroid;
  float* low_ratio;
  float* mel_db;  // workspace [total_frames][128] (nullptr when no onset output is wanted)
  int* segmax;    // workspace [n_segs] ordered-int max of mel_db
};

__global__ void __launch_bounds__(kFeatThreads) stft_feat_kernel(FeatArgs a) {
  extern __shared__ float2 smem_f2[];
  float2* buf0 = smem_f2;
  float2* buf1 = buf0 + fpad(kNfft) + 1;
  float* pw = reinterpret_cast<float*>(buf1 + fpad(kNfft) + 1);  // [2][kBins+3] power spectra
  __shared__ float red[2][5][kFeatThreads / 64];
  __shared__ int s_seg;
  // ---- which segment does this pair belong to (binary search over pair offsets)
  if (threadIdx.x == 0) {
    int lo = 0, hi = a.n_segs - 1;
    const long long me = blockIdx.x;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (a.segs[mid].pair_off <= me) lo = mid; else hi = mid - 1;
    }
    s_seg = lo;
  }
  __syncthreads();
  const SegDev sg = a.segs[s_seg];
  const int t0 = (int)(blockIdx.x - sg.pair_off) * 2;
  const bool has_b = t0 + 1 < sg.n_frames;
  const float* x = a.x + sg.start;
  const long long pa = (long long)t0 * a.hop - kNfft / 2;
  const long long pb = pa + a.hop;
  for (int j = threadIdx.x; j < kNfft; j += kFeatThreads) {
    const long long ia = pa + j, ib = pb + j;
    const float w = __ldg(a.hann + j);
    const float va = (ia >= 0 && ia < sg.len) ? __ldg(x + ia) : 0.f;
    const float vb = (has_b && ib >= 0 && ib < sg.len) ? __ldg(x + ib) : 0.f;
    buf0[fpad(j)] = make_float2(va * w, vb * w);
  }
  __syncthreads();
  const float2* Z = fft_smem<false>(buf0, buf1, a.fft);
  const int PS = kBins + 3;
  for (int k = threadIdx.x; k < kBins; k += kFeatThreads) {
    const float2 u = Z[fpad(k)];
    const float2 v = Z[fpad(k == 0 ? 0 : kNfft - k)];
    const float are = 0.5f * (u.x + v.x), aim = 0.5f * (u.y - v.y);
    const float bre = 0.5f * (u.y + v.y), bim = -0.5f * (u.x - v.x);
    pw[k] = are * are + aim * aim;
    pw[PS + k] = bre * bre + bim * bim;
  }
  __syncthreads();
  // ---- per-frame reductions: threads 0..127 -> frame A, 128..255 -> frame B
  {
    const int fr = threadIdx.x >> 7, lt = threadIdx.x & 127;
    const float* p = pw + fr * PS;
    float s_log = 0.f, s_pow = 0.f, s_mag = 0.f, s_fmag = 0.f, s_low = 0.f;
    const float df = a.sr / (float)kNfft;
    for (int k = lt; k < kBins; k += 128) {
      const float P = p[k];
      const float St = fmaxf(1e-10f, P);
      s_log += logf(St);
      s_pow += St;
      const float S = sqrtf(P);
      s_mag += S;
      s_fmag = fmaf(S, df * (float)k, s_fmag);
      if (k < kBins / 3) s_low += S;
    }
    float vals[5] = {s_log, s_pow, s_mag, s_fmag, s_low};
    const int w_in = lt >> 5;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const float r = warp_sum(vals[q]);
      if ((threadIdx.x & 31) == 0) red[fr][q][w_in] = r;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    const int fr = threadIdx.x;
    if (fr == 0 || has_b) {
      float v[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) v[q] = red[fr][q][0] + red[fr][q][1] + red[fr][q][2] + red[fr][q][3];
      const long long o = sg.frame_off + t0 + fr;
      if (a.flatness) a.flatness[o] = expf(v[0] / (float)kBins) / (v[1] / (float)kBins);
      if (a.centroid) a.centroid[o] = v[3] / fmaxf(v[2], 1.17549435e-38f);
      if (a.low_ratio) a.low_ratio[o] = v[4] / (v[2] + 1e-10f);
    }
  }
  // ---- mel projection + dB: thread (frame = tid/128, band = tid%128)
  if (a.mel_db) {
    const int fr = threadIdx.x >> 7, m = threadIdx.x & 127;
    float db = -INFINITY;
    if (fr == 0 || has_b) {
      const float* p = pw + fr * PS + a.mel_start[m];
      const float* w = a.mel_w + a.mel_woff[m];
      const int cnt = a.mel_count[m];
      float acc = 0.f;
      for (int k = 0; k < cnt; ++k) acc = fmaf(__ldg(w + k), p[k], acc);
      db = 10.0f * log10f(fmaxf(1e-10f, acc));
      a.mel_db[(sg.frame_off + t0 + fr) * kMels + m] = db;
    }
    const float mx = warp_max(db);
    if ((threadIdx.x & 31) == 0 && mx > -INFINITY) atomicMax(a.segmax + s_seg, float_to_ordered(mx));
  }
}

struct FluxArgs {
  const SegDev* segs;
  int n_segs;
  long long total_frames;
  const float* mel_db;
  const int* segmax;
  int shift;  // n_fft // (2*hop): librosa centers the envelope by padding lag + shift zeros in front
  float* onset_mean;
  float* onset_median;
};

__global__ void __launch_bounds__(256) onset_flux_kernel(FluxArgs a) {
  const long long gf = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);  // global frame = output slot
  const int lane = threadIdx.x & 31;
  if (gf >= a.total_frames) return;
  int lo = 0, hi = a.n_segs - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (a.segs[mid].frame_off <= gf) lo = mid; else hi = mid - 1;
  }
  const SegDev sg = a.segs[lo];
  const int i = (int)(gf - sg.frame_off);  // envelope index within the segment
  const int t = i - a.shift;               // flux between mel frames t and t-1 lands at env[t + shift]
  float mean = 0.f, med = 0.f;
  if (i >= sg.n_frames) return;  // slot in a gap between segments: not ours
  if (t >= 1 && t < sg.n_frames) {
    const float floor_db = ordered_to_float(a.segmax[lo]) - 80.0f;
    const float4 c = *reinterpret_cast<const float4*>(a.mel_db + (sg.frame_off + t) * kMels + lane * 4);
    const float4 p = *reinterpret_cast<const float4*>(a.mel_db + (sg.frame_off + t - 1) * kMels + lane * 4);
    float d[4];
    d[0] = fmaxf(0.f, fmaxf(c.x, floor_db) - fmaxf(p.x, floor_db));
    d[1] = fmaxf(0.f, fmaxf(c.y, floor_db) - fmaxf(p.y, floor_db));
    d[2] = fmaxf(0.f, fmaxf(c.z, floor_db) - fmaxf(p.z, floor_db));
    d[3] = fmaxf(0.f, fmaxf(c.w, floor_db) - fmaxf(p.w, floor_db));
    mean = warp_sum(d[0] + d[1] + d[2] + d[3]) * (1.0f / kMels);
    if (a.onset_median) {
      // exact median of 128 values: ranks by counting (ties broken by index), average of ranks 63 and 64
      int rank[4] = {0, 0, 0, 0};
      for (int src = 0; src < 32; ++src) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = __shfl_sync(0xffffffffu, d[e], src);
          const int vi = src * 4 + e;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int mi = lane * 4 + q;
            rank[q] += (v < d[q] || (v == d[q] && vi < mi)) ? 1 : 0;
          }
        }
      }
      float part = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (rank[q] == kMels / 2 - 1 || rank[q] == kMels / 2) part += d[q];
      med = warp_sum(part) * 0.5f;
    }
  }
  if (lane == 0) {
    if (a.onset_mean) a.onset_mean[gf] = mean;
    if (a.onset_median) a.onset_median[gf] = med;
  }
}

static void layout_segments(const ac_feat_segment* h, int n, int hop, std::vector<SegDev>& out, long long& pairs,
                            long long& frames_hi) {
  out.resize(n);
  pairs = 0;
  frames_hi = 0;
  for (int i = 0; i < n; ++i) {
    SegDev s;
    s.start = h[i].start;
    s.len = h[i].len;
    s.frame_off = h[i].frame_off;
    s.n_frames = (int)(1 + h[i].len / hop);
    s.pair_off = pairs;
    s.pad_ = 0;
    pairs += (s.n_frames + 1) / 2;
    frames_hi = std::max(frames_hi, s.frame_off + s.n_frames);
    out[i] = s;
  }
}

}  // namespace ac

extern "C" size_t ac_stft_features_workspace_bytes(const ac_feat_segment* h_segs, int n_segs, int hop) {
  if (!h_segs || n_segs <= 0 || hop <= 0) return 0;
  std::vector<ac::SegDev> segs;
  long long pairs, frames;
  ac::layout_segments(h_segs, n_segs, hop, segs, pairs, frames);
  return ac::align_up(sizeof(ac::SegDev) * n_segs, 256) + ac::align_up(sizeof(int) * n_segs, 256) +
         (size_t)frames * ac::kMels * sizeof(float) + 512;
}

extern "C" int ac_stft_features(const float* d_x, const ac_feat_segment* h_segs, int n_segs, int hop, int sr,
                                float* d_flatness, float* d_onset_mean, float* d_onset_median, float* d_centroid,
                                float* d_low_ratio, void* d_ws, size_t ws_bytes, void* stream) {
  using namespace ac;
  AC_REQUIRE(d_x && h_segs && d_ws, "null pointer");
  AC_REQUIRE(n_segs > 0 && hop > 0 && sr > 0, "bad arguments");
  if (ws_bytes < ac_stft_features_workspace_bytes(h_segs, n_segs, hop)) {
    set_error("feature workspace too small");
    return AC_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<SegDev> segs;
  long long pairs, frames;
  layout_segments(h_segs, n_segs, hop, segs, pairs, frames);
  // segments must be laid out with increasing, non-overlapping frame ranges (pass 2 searches them)
  for (int i = 1; i < n_segs; ++i)
    AC_REQUIRE(segs[i].frame_off >= segs[i - 1].frame_off + segs[i - 1].n_frames, "segments must have increasing frame_off");
  for (int i = 0; i < n_segs; ++i) AC_REQUIRE(segs[i].len > 0 && segs[i].start >= 0, "empty segment");
  const FftPlan* fp = get_fft_plan(kNfft);
  if (!fp) return AC_E_CUDA;
  const MelTable* mt = get_mel_table(sr);
  if (!mt) return AC_E_CUDA;
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(d_ws) + 255) & ~uintptr_t(255));
  SegDev* d_segs = reinterpret_cast<SegDev*>(base);
  int* d_segmax = reinterpret_cast<int*>(base + align_up(sizeof(SegDev) * n_segs, 256));
  float* d_mel = reinterpret_cast<float*>(reinterpret_cast<char*>(d_segmax) + align_up(sizeof(int) * n_segs, 256));
  AC_CHECK_CUDA(cudaMemcpyAsync(d_segs, segs.data(), sizeof(SegDev) * n_segs, cudaMemcpyHostToDevice, st));
  const bool want_onset = d_onset_mean || d_onset_median;
  if (want_onset) {
    init_segmax_kernel<<<(n_segs + 255) / 256, 256, 0, st>>>(d_segmax, n_segs);
    AC_LAUNCH_CHECK();
  }
  FeatArgs fa;
  fa.x = d_x; fa.segs = d_segs; fa.n_segs = n_segs; fa.hop = hop; fa.sr = (float)sr;
  fa.fft = make_fft_dev(fp); fa.hann = fp->d_hann;
  fa.mel_start = mt->d_start; fa.mel_count = mt->d_count; fa.mel_woff = mt->d_woff; fa.mel_w = mt->d_w;
  fa.flatness = d_flatness; fa.centroid = d_centroid; fa.low_ratio = d_low_ratio;
  fa.mel_db = want_onset ? d_mel : nullptr; fa.segmax = d_segmax;
  const size_t smem = sizeof(float2) * 2 * fft_smem_floats2(kNfft) + sizeof(float) * 2 * (kBins + 3);
  AC_CHECK_CUDA(cudaFuncSetAttribute(stft_feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  AC_REQUIRE(pairs < 0x7fffffffLL, "too many frames");
  {
    double in_bytes = 0.0, n_fr = 0.0;
    for (auto& sgm : segs) { in_bytes += 4.0 * (double)sgm.len; n_fr += sgm.n_frames; }
    const int n_out = (d_flatness != nullptr) + (d_centroid != nullptr) + (d_low_ratio != nullptr);
    ProfScope ps(KC_FEAT_STFT, 0.0, in_bytes + n_fr * (4.0 * n_out + (want_onset ? 512.0 : 0.0)), st);
    stft_feat_kernel<<<(unsigned)pairs, kFeatThreads, smem, st>>>(fa);
    AC_LAUNCH_CHECK();
  }
  if (want_onset) {
    // the mel rows of a segment live at [frame_off, frame_off + n_frames): same slots as the outputs
    FluxArgs xa;
    xa.segs = d_segs; xa.n_segs = n_segs; xa.total_frames = frames; xa.mel_db = d_mel; xa.segmax = d_segmax;
    xa.shift = kNfft / (2 * hop); xa.onset_mean = d_onset_mean; xa.onset_median = d_onset_median;
    ProfScope ps(KC_FEAT_FLUX, 0.0, (double)frames * (1024.0 + 8.0), st);
    onset_flux_kernel<<<(unsigned)((frames + 7) / 8), 256, 0, st>>>(xa);
    AC_LAUNCH_CHECK();
  }
  return AC_OK;
}

// =================================================================================================
// Autocorrelation tempogram statistics (librosa.feature.tempogram + rhythm.tempo front end):
//   features_cache.py:283-288 (aggregate=None), adaptive_vad_enhancer.py:61-67 (beat_track -> mean
//   aggregate), :151-156 (aggregate=None).  SURVEY.md section 8(f) row N2.
// For every frame t: window = hann(win) * padded_env[t : t+win] (linear-ramp centring), autocorrelation
// by FFT (two frames per complex transform, forward and inverse), max-normalised.  Emitted per frame:
// argmax_lag(log1p(1e6*tg) + logprior); accumulated over frames: sum of the normalised tempogram.
// =================================================================================================
namespace ac {

struct TgArgs {
  const float* env;
  long long n;
  int win, half;
  FftDev fft;
  const float* window;    // hann(win), periodic
  const float* logprior;  // [win], -inf where excluded
  float* tg_sum;          // [win], pre-zeroed
  int* best;              // [n]
  int pairs_per_cta;
};

__device__ __forceinline__ float tg_padded(const TgArgs& a, long long p) {
  // np.pad(env, half, mode="linear_ramp", end_values=0)
  if (p < a.half) return a.env[0] * ((float)p / (float)a.half);
  const long long q = p - a.half;
  if (q < a.n) return a.env[q];
  const long long i = q - a.n;  // 0 .. half-1
  return a.env[a.n - 1] * (1.0f - (float)(i + 1) / (float)a.half);
}

__global__ void __launch_bounds__(256) tempogram_kernel(TgArgs a) {
  extern __shared__ float2 smem_f2[];
  const int N = a.fft.n;
  float2* buf0 = smem_f2;
  float2* buf1 = buf0 + fpad(N) + 1;
  float* acc = reinterpret_cast<float*>(buf1 + fpad(N) + 1);  // [win]
  __shared__ float s_max[2][8];
  __shared__ float s_val[2][8];
  __shared__ int s_idx[2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = threadIdx.x; j < a.win; j += 256) acc[j] = 0.f;
  __syncthreads();
  for (int pp = 0; pp < a.pairs_per_cta; ++pp) {
    const long long t0 = 2 * ((long long)blockIdx.x * a.pairs_per_cta + pp);
    if (t0 >= a.n) break;
    const bool has_b = t0 + 1 < a.n;
    for (int j = threadIdx.x; j < N; j += 256) {
      float va = 0.f, vb = 0.f;
      if (j < a.win) {
        const float w = __ldg(a.window + j);
        va = tg_padded(a, t0 + j) * w;
        vb = has_b ? tg_padded(a, t0 + 1 + j) * w : 0.f;
      }
      buf0[fpad(j)] = make_float2(va, vb);
    }
    __syncthreads();
    float2* Z = fft_smem<false>(buf0, buf1, a.fft);
    float2* Q = (Z == buf0) ? buf1 : buf0;
    for (int k = threadIdx.x; k < N; k += 256) {
      const float2 u = Z[fpad(k)];
      const float2 v = Z[fpad(k == 0 ? 0 : N - k)];
      const float are = 0.5f * (u.x + v.x), aim = 0.5f * (u.y - v.y);
      const float bre = 0.5f * (u.y + v.y), bim = -0.5f * (u.x - v.x);
      Q[fpad(k)] = make_float2(are * are + aim * aim, bre * bre + bim * bim);
    }
    __syncthreads();
    float2* R = fft_smem<true>(Q, Z, a.fft);  // R[j] = (autocorr_a[j], autocorr_b[j]) * N
    // ---- per-frame max |ac| over the first `win` lags
    float ma = 0.f, mb = 0.f;
    for (int j = threadIdx.x; j < a.win; j += 256) {
      const float2 r = R[fpad(j)];
      ma = fmaxf(ma, fabsf(r.x));
      mb = fmaxf(mb, fabsf(r.y));
    }
    ma = warp_max(ma);
    mb = warp_max(mb);
    if (lane == 0) { s_max[0][warp] = ma; s_max[1][warp] = mb; }
    __syncthreads();
    ma = mb = 0.f;
    for (int w = 0; w < 8; ++w) { ma = fmaxf(ma, s_max[0][w]); mb = fmaxf(mb, s_max[1][w]); }
    const float ia = ma < 1.17549435e-38f ? 1.f : 1.f / ma, ib = mb < 1.17549435e-38f ? 1.f : 1.f / mb;
    // ---- normalise, accumulate, argmax of log1p(1e6*tg) + logprior (first maximum wins)
    float best_a = -INFINITY, best_b = -INFINITY;
    int idx_a = 0x7fffffff, idx_b = 0x7fffffff;
    for (int j = threadIdx.x; j < a.win; j += 256) {
      const float2 r = R[fpad(j)];
      const float ta = r.x * ia, tb = r.y * ib;
      acc[j] += ta + (has_b ? tb : 0.f);
      const float lp = __ldg(a.logprior + j);
      const float sa = log1pf(1e6f * ta) + lp, sb = log1pf(1e6f * tb) + lp;
      if (sa > best_a) { best_a = sa; idx_a = j; }
      if (sb > best_b) { best_b = sb; idx_b = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best_a, o);
      int oi = __shfl_xor_sync(0xffffffffu, idx_a, o);
      if (ov > best_a || (ov == best_a && oi < idx_a)) { best_a = ov; idx_a = oi; }
      ov = __shfl_xor_sync(0xffffffffu, best_b, o);
      oi = __shfl_xor_sync(0xffffffffu, idx_b, o);
      if (ov > best_b || (ov == best_b && oi < idx_b)) { best_b = ov; idx_b = oi; }
    }
    if (lane == 0) { s_val[0][warp] = best_a; s_idx[0][warp] = idx_a; s_val[1][warp] = best_b; s_idx[1][warp] = idx_b; }
    __syncthreads();
    if (threadIdx.x < 2) {
      const int f = threadIdx.x;
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int w = 0; w < 8; ++w)
        if (s_val[f][w] > bv || (s_val[f][w] == bv && s_idx[f][w] < bi)) { bv = s_val[f][w]; bi = s_idx[f][w]; }
      if (bi == 0x7fffffff) bi = 0;  // every candidate was -inf / NaN: numpy's argmax returns 0
      if (f == 0 || has_b) a.best[t0 + f] = bi;
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j < a.win; j += 256) atomicAdd(a.tg_sum + j, acc[j]);
}

}  // namespace ac

extern "C" int ac_tempogram_stats(const float* d_env, long long n, int win, const float* d_logprior, float* d_tg_sum,
                                  int* d_best, void* stream) {
  using namespace ac;
  AC_REQUIRE(d_env && d_logprior && d_tg_sum && d_best, "null pointer");
  AC_REQUIRE(n > 0 && win >= 2, "bad sizes");
  int nfft = 16;
  while (nfft < 2 * win - 1) nfft *= 2;
  AC_REQUIRE(nfft <= 8192, "tempogram window too long");
  cudaStream_t st = (cudaStream_t)stream;
  const FftPlan* fp = get_fft_plan(nfft);
  if (!fp) return AC_E_CUDA;
  // periodic hann(win): cached per win
  static std::mutex mu;
  static std::map<int, float*> cache;
  float* d_win = nullptr;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(win);
    if (it == cache.end()) {
      std::vector<float> h(win);
      for (int i = 0; i < win; ++i) h[i] = (float)(0.5 - 0.5 * std::cos(6.283185307179586476925286766559 * i / win));
      AC_CHECK_CUDA(cudaMalloc(&d_win, sizeof(float) * win));
      AC_CHECK_CUDA(cudaMemcpy(d_win, h.data(), sizeof(float) * win, cudaMemcpyHostToDevice));
      cache[win] = d_win;
    } else {
      d_win = it->second;
    }
  }
  AC_CHECK_CUDA(cudaMemsetAsync(d_tg_sum, 0, sizeof(float) * win, st));
  TgArgs a;
  a.env = d_env; a.n = n; a.win = win; a.half = win / 2; a.fft = make_fft_dev(fp); a.window = d_win;
  a.logprior = d_logprior; a.tg_sum = d_tg_sum; a.best = d_best;
  const long long pairs = (n + 1) / 2;
  a.pairs_per_cta = (int)std::max<long long>(1, std::min<long long>(16, pairs / (4LL * device_sm_count())));
  const long long grid = (pairs + a.pairs_per_cta - 1) / a.pairs_per_cta;
  const size_t smem = sizeof(float2) * 2 * fft_smem_floats2(nfft) + sizeof(float) * win;
  AC_CHECK_CUDA(cudaFuncSetAttribute(tempogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(KC_MISC, 0.0, 4.0 * n * 3, st);
  tempogram_kernel<<<(unsigned)grid, 256, smem, st>>>(a);
  AC_LAUNCH_CHECK();
  return AC_OK;
}

// Ellis dynamic-programming beat tracker inner loop (librosa.beat.__beat_track_dp) - a strictly
// sequential scan over frames, so it runs on the host; C++ instead of a Python loop.
extern "C" int ac_host_beat_dp(const float* localscore, int n, int period, float tightness, long long* backlink,
                               float* cumscore) {
  AC_REQUIRE(localscore && backlink && cumscore && n >= 0 && period >= 1 && tightness > 0, "bad arguments");
  const int w_lo = -2 * period, w_hi = -(int)std::nearbyint(period / 2.0);  // np.round: half to even  // window = arange(w_lo, w_hi + 1)
  const int nw = w_hi - w_lo + 1;
  if (nw <= 0) return AC_E_INVALID;
  std::vector<double> txwt(nw);
  for (int k = 0; k < nw; ++k) {
    const double lg = std::log(-(double)(w_lo + k) / (double)period);
    txwt[k] = -(double)tightness * lg * lg;
  }
  float mx = -INFINITY;
  for (int i = 0; i < n; ++i) mx = std::max(mx, localscore[i]);
  const double thresh = 0.01 * (double)mx;
  bool first = true;
  for (int i = 0; i < n; ++i) {
    // candidates[k] = txwt[k] (+ cumscore[i + w_lo + k] when that index is >= 0); first maximum wins
    double best = -INFINITY;
    int loc = 0;
    for (int k = 0; k < nw; ++k) {
      const int j = i + w_lo + k;
      const double c = txwt[k] + (j >= 0 ? (double)cumscore[j] : 0.0);
      if (c > best) { best = c; loc = k; }
    }
    cumscore[i] = (float)((double)localscore[i] + best);
    if (first && (double)localscore[i] < thresh) {
      backlink[i] = -1;
    } else {
      backlink[i] = i + w_lo + loc;
      first = false;
    }
  }
  return AC_OK;
}
