// Cut-point refinement on the device: SURVEY.md section 8(f) row N1.
//
// Reference: src/audio_cut/cutting/refine.py
//   align_to_zero_cross       :72-110   nearest zero crossing within +-win_ms (float32 arithmetic of numpy scalars)
//   apply_quiet_guard         :113-158  segment-local boxcar RMS-dB (fp32 squares, edge padding), argmin, centre of window
//   _prepare_quiet_lookup     :161-181  whole-track boxcar RMS-dB in fp64 (np.convolve mode="same") + a Python reverse scan
//   _apply_quiet_guard_fast   :184-214  argmin of the lookup over [idx, idx+max_shift)
//   finalize_cut_points       :318-371  per point: zero cross -> vocal guard (fast, else slow) -> zero cross -> mix guard
//
// The reference builds the lookup for the WHOLE track (an O(N) Python loop whose `next_quiet` output nobody
// reads) although only [idx, idx+max_shift) of it is consulted per point.  Here nothing is precomputed: one CTA
// per cut point evaluates the boxcar sums it needs (<= search+win-1 samples) from the stems that are already
// resident in HBM after the separation, and walks the whole per-point chain in one launch.  Points are
// independent, so the launch is P CTAs wide.  All decisions are made in fp64 on the same quantities as the
// reference; float32 is used exactly where numpy (>= 2.0 scalar promotion) computes in float32.
#include "common.cuh"

namespace ac {

constexpr int kRefThreads = 256;
constexpr int kRefR = 8;  // consecutive outputs per thread in the boxcar pass
constexpr double kRefEps = 1e-12;

struct RefineArgs {
  const float* mix;
  const float* vocal;  // may be null
  long long n;
  int sr;
  int zc_half;  // max(1, round(zero_cross_win_ms/1000*sr))
  int search;   // max(1, round(search_right_ms/1000*sr))
  int win;      // max(1, round(guard_win_ms/1000*sr))
  double guard_db, floor_db;
  int vocal_first, vocal_guard, mix_guard;
};

struct RefineShared {
  double red_v[kRefThreads / 32];
  double red_z[kRefThreads / 32];
  int red_i[kRefThreads / 32];
  double out_v, out_z, db0;
  int out_i;
};

__host__ __device__ __forceinline__ int pad8(int x) { return x + (x >> 3); }  // one pad double per 8: stride-8 reads hit 16 bank pairs

// np.argmin order: the first NaN wins, otherwise the smallest value, the lowest index on ties
__device__ __forceinline__ bool arg_better(double a, int ia, double b, int ib) {
  const bool na = isnan(a), nb = isnan(b);
  if (na || nb) return (na && nb) ? ia < ib : na;
  if (a < b) return true;
  if (a > b) return false;
  return ia < ib;
}

// CTA-wide argmin of (v, i) carrying a payload z; the result lands in sh->out_v / out_i / out_z
__device__ void cta_argmin(double v, int i, double z, RefineShared* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    const double oz = __shfl_xor_sync(0xffffffffu, z, o);
    if (arg_better(ov, oi, v, i)) { v = ov; i = oi; z = oz; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // previous readers of the scratch are done
  if (lane == 0) { sh->red_v[warp] = v; sh->red_i[warp] = i; sh->red_z[warp] = z; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kRefThreads / 32; ++w)
      if (arg_better(sh->red_v[w], sh->red_i[w], v, i)) { v = sh->red_v[w]; i = sh->red_i[w]; z = sh->red_z[w]; }
    sh->out_v = v; sh->out_i = i; sh->out_z = z;
  }
  __syncthreads();
}

// refine.py:72-110.  numpy scalar semantics (NEP 50): left/right are float32 scalars, so left*right, |l|+|r|,
// |l|/denom, (pos-1)+frac and zero_pos-idx are all float32 operations (python ints convert to float32).
__device__ double zero_cross(const float* __restrict__ w, const RefineArgs& a, double t, RefineShared* sh) {
  const long long idx = (long long)rint(t * (double)a.sr);
  if (idx <= 0 || idx >= a.n) return t;
  const long long start = max(1LL, idx - a.zc_half), end = min(a.n - 1, idx + a.zc_half);
  if (end <= start) return t;
  const int inf_i = 0x7fffffff;
  double best_d = INFINITY, best_z = 0.0;
  int best_i = inf_i;
  for (long long pos = start + threadIdx.x; pos <= end; pos += kRefThreads) {
    const float left = w[pos - 1], right = w[pos];
    double z;
    float d;
    if (left == 0.0f) {
      z = (double)(pos - 1);
      d = (float)llabs(pos - 1 - idx);
    } else if (right == 0.0f) {
      z = (double)pos;
      d = (float)llabs(pos - idx);
    } else if (__fmul_rn(left, right) < 0.0f) {
      const float den = __fadd_rn(fabsf(left), fabsf(right));
      const float frac = den > 1e-12f ? __fdiv_rn(fabsf(left), den) : 0.5f;
      const float zf = __fadd_rn((float)(pos - 1), frac);
      z = (double)zf;
      d = fabsf(__fsub_rn(zf, (float)idx));
    } else {
      continue;
    }
    const int i = (int)(pos - start);
    if (arg_better((double)d, i, best_d, best_i)) { best_d = (double)d; best_i = i; best_z = z; }
  }
  cta_argmin(best_d, best_i, best_z, sh);
  if (sh->out_i == inf_i) return t;  // no crossing in the window
  return sh->out_z / (double)a.sr;
}

// Boxcar mean of squares over `win` samples for L consecutive outputs, in dB, and its argmin.
// FAST: lookup semantics (refine.py:168-172): fp64 squares, window [i - win/2, i + (win-1)/2] cut at the track ends.
// !FAST: apply_quiet_guard semantics (:140-146): float32 squares, window [i, i + win) with the segment's last
//        sample repeated.  On return sh->out_i = argmin (relative to idx), sh->out_v = its dB, sh->db0 = dB of output 0.
template <bool FAST>
__device__ void window_db_argmin(const float* __restrict__ w, const RefineArgs& a, long long idx, int L, double* S,
                                 RefineShared* sh) {
  const int win = a.win;
  const double inv_win = 1.0 / (double)win;
  const int Lr = (L + kRefR - 1) / kRefR * kRefR;
  const int n_stage = Lr + win - 1;
  __syncthreads();  // S may still be read by a previous pass
  for (int k = threadIdx.x; k < n_stage; k += kRefThreads) {
    double s = 0.0;
    if (k < L + win - 1) {
      if (FAST) {
        const long long p = idx - win / 2 + k;
        if (p >= 0 && p < a.n) {
          const double v = (double)w[p];
          s = v * v * inv_win;
        }
      } else {
        const long long p = min(idx + (long long)k, idx + (long long)L - 1);
        const float v = w[p];
        s = (double)__fmul_rn(v, v) * inv_win;
      }
    }
    S[pad8(k)] = s;
  }
  __syncthreads();
  double best = INFINITY;
  int best_i = 0x7fffffff;
  for (int g = threadIdx.x; g * kRefR < L; g += kRefThreads) {
    const int base = g * kRefR;
    double acc[kRefR];
#pragma unroll
    for (int r = 0; r < kRefR; ++r) acc[r] = 0.0;
    if (win >= kRefR) {
      // every staged value is read once and added to the (up to 8) windows that contain it, in window order
#pragma unroll
      for (int k = 0; k < kRefR - 1; ++k) {
        const double v = S[pad8(base + k)];
#pragma unroll
        for (int r = 0; r < kRefR; ++r)
          if (r <= k) acc[r] += v;
      }
      for (int k = kRefR - 1; k < win; ++k) {
        const double v = S[pad8(base + k)];
#pragma unroll
        for (int r = 0; r < kRefR; ++r) acc[r] += v;
      }
#pragma unroll
      for (int e = 0; e < kRefR - 1; ++e) {
        const double v = S[pad8(base + win + e)];
#pragma unroll
        for (int r = 0; r < kRefR; ++r)
          if (r > e) acc[r] += v;
      }
    } else {
#pragma unroll
      for (int r = 0; r < kRefR; ++r)
        for (int j = 0; j < win; ++j) acc[r] += S[pad8(base + r + j)];
    }
#pragma unroll
    for (int r = 0; r < kRefR; ++r) {
      const int o = base + r;
      if (o < L) {
        const double db = 20.0 * log10(sqrt(acc[r] + kRefEps) + kRefEps);
        if (o == 0) sh->db0 = db;
        if (arg_better(db, o, best, best_i)) { best = db; best_i = o; }
      }
    }
  }
  cta_argmin(best, best_i, 0.0, sh);
}

// refine.py:184-214
__device__ double quiet_guard_fast(const float* __restrict__ w, const RefineArgs& a, double t, double* S, RefineShared* sh) {
  long long idx = (long long)rint(t * (double)a.sr);
  idx = min(max(idx, 0LL), a.n - 1);
  const long long end = min(a.n, idx + a.search);
  if (end <= idx) return t;
  window_db_argmin<true>(w, a, idx, (int)(end - idx), S, sh);
  const int k = sh->out_i;
  const double target = sh->out_v, original = sh->db0;
  if ((original - target) < a.guard_db || target > a.floor_db || k == 0) return t;
  return (double)(idx + k) / (double)a.sr;
}

// refine.py:113-158
__device__ double quiet_guard_slow(const float* __restrict__ w, const RefineArgs& a, double t, double* S, RefineShared* sh) {
  long long idx = (long long)rint(t * (double)a.sr);
  if (idx < 0) idx = 0;
  const long long end = min(a.n, idx + a.search);
  if (end <= idx + 1) return t;
  const int L = (int)(end - idx);
  bool keep;
  if (L <= a.win) {
    // the reference takes the raw float32 samples as "rms" here (end of track only): float32 log10, NaN for negatives
    double best = INFINITY;
    int best_i = 0x7fffffff;
    __syncthreads();  // sh->db0 of the preceding fast pass has been read by everybody
    for (int j = threadIdx.x; j < L; j += kRefThreads) {
      const float db = 20.0f * log10f(__fadd_rn(w[idx + j], 1e-12f));
      if (j == 0) sh->db0 = (double)db;
      if (arg_better((double)db, j, best, best_i)) { best = (double)db; best_i = j; }
    }
    cta_argmin(best, best_i, 0.0, sh);
    const float diff = __fsub_rn((float)sh->db0, (float)sh->out_v);
    keep = diff < (float)a.guard_db || (float)sh->out_v > (float)a.floor_db;  // NaN compares false, as in numpy
  } else {
    window_db_argmin<false>(w, a, idx, L, S, sh);
    keep = (sh->db0 - sh->out_v) < a.guard_db || sh->out_v > a.floor_db;
  }
  if (keep) return t;
  long long center = idx + sh->out_i + a.win / 2;
  center = min(a.n - 1, max(0LL, center));
  return (double)center / (double)a.sr;
}

__global__ void __launch_bounds__(kRefThreads) refine_cuts_kernel(RefineArgs a, const double* __restrict__ t_in,
                                                                  double* __restrict__ guard_out,
                                                                  double* __restrict__ final_out) {
  extern __shared__ double S[];
  __shared__ RefineShared sh;
  const double raw = t_in[blockIdx.x];
  double g = raw;
  if (a.vocal_first && a.vocal) {
    g = zero_cross(a.vocal, a, g, &sh);
    if (a.vocal_guard) {
      const double f = quiet_guard_fast(a.vocal, a, g, S, &sh);
      g = (f != g) ? f : quiet_guard_slow(a.vocal, a, g, S, &sh);
    }
  }
  double m = zero_cross(a.mix, a, g, &sh);
  if (a.mix_guard) {
    const double f = quiet_guard_fast(a.mix, a, m, S, &sh);
    m = (f != m) ? f : quiet_guard_slow(a.mix, a, m, S, &sh);
  }
  const double dur = (double)a.n / (double)a.sr;
  m = fmin(fmax(m, 0.0), fmax(dur, 0.0));
  if (threadIdx.x == 0) {
    guard_out[blockIdx.x] = g;
    final_out[blockIdx.x] = m;
  }
}

// The whole-track lookup itself (refine.py:161-173), for callers that want the array: rms_db[i] in fp64.
__global__ void __launch_bounds__(kRefThreads) quiet_lookup_kernel(const float* __restrict__ w, long long n, int win,
                                                                   int per_cta, double* __restrict__ out) {
  extern __shared__ double S[];
  const long long idx = (long long)blockIdx.x * per_cta;
  const int L = (int)min((long long)per_cta, n - idx);
  // same staging and accumulation order as window_db_argmin<true>, but every output is written
  const double inv_win = 1.0 / (double)win;
  const int Lr = (L + kRefR - 1) / kRefR * kRefR;
  const int n_stage = Lr + win - 1;
  for (int k = threadIdx.x; k < n_stage; k += kRefThreads) {
    double s = 0.0;
    const long long p = idx - win / 2 + k;
    if (k < L + win - 1 && p >= 0 && p < n) {
      const double v = (double)w[p];
      s = v * v * inv_win;
    }
    S[pad8(k)] = s;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < L; o += kRefThreads) {
    double acc = 0.0;
    for (int j = 0; j < win; ++j) acc += S[pad8(o + j)];
    out[idx + o] = 20.0 * log10(sqrt(acc + kRefEps) + kRefEps);
  }
}

static size_t refine_smem_bytes(int L, int win) {
  const int Lr = (L + kRefR - 1) / kRefR * kRefR;
  return sizeof(double) * (size_t)(pad8(Lr + win - 1 + kRefR) + 8);
}

}  // namespace ac

extern "C" int ac_refine_cut_points(const float* d_mix, const float* d_vocal, long long n, int sr, const double* d_times,
                                    int n_points, int zero_cross_half, int search, int win, double guard_db,
                                    double floor_db, int use_vocal_guard_first, int enable_vocal_guard,
                                    int enable_mix_guard, double* d_guard_times, double* d_final_times, void* stream) {
  using namespace ac;
  AC_REQUIRE(d_mix && d_times && d_guard_times && d_final_times, "null pointer");
  AC_REQUIRE(n > 0 && sr > 0 && n_points >= 0, "n and sr must be positive");
  AC_REQUIRE(zero_cross_half >= 1 && search >= 1 && win >= 1, "window sizes must be >= 1 sample");
  if (n_points == 0) return AC_OK;
  const size_t smem = refine_smem_bytes(search, win);
  // 227 KB of dynamic shared memory per CTA minus the kernel's small static arrays: search + win <= ~25.5k samples
  // (the reference's shipped config, search_right_ms 450 + win_ms 80 at 44.1 kHz = 23 373 samples, needs 205.5 KB)
  AC_REQUIRE(smem <= 225 * 1024, "search + guard window does not fit one CTA's shared memory (max ~25.5k samples)");
  static size_t attr = 0;
  if (smem > attr) {
    AC_CHECK_CUDA(cudaFuncSetAttribute(refine_cuts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  RefineArgs a;
  a.mix = d_mix;
  a.vocal = d_vocal;
  a.n = n;
  a.sr = sr;
  a.zc_half = zero_cross_half;
  a.search = search;
  a.win = win;
  a.guard_db = guard_db;
  a.floor_db = floor_db;
  a.vocal_first = use_vocal_guard_first;
  a.vocal_guard = enable_vocal_guard;
  a.mix_guard = enable_mix_guard;
  refine_cuts_kernel<<<n_points, kRefThreads, smem, (cudaStream_t)stream>>>(a, d_times, d_guard_times, d_final_times);
  AC_LAUNCH_CHECK();
  return AC_OK;
}

extern "C" int ac_quiet_lookup_db(const float* d_wave, long long n, int win, double* d_rms_db, void* stream) {
  using namespace ac;
  AC_REQUIRE(d_wave && d_rms_db, "null pointer");
  AC_REQUIRE(n > 0 && win >= 1, "n and win must be positive");
  const int per_cta = 4096;
  const size_t smem = refine_smem_bytes(per_cta, win);
  AC_REQUIRE(smem <= 200 * 1024, "guard window too long");
  static size_t attr = 0;
  if (smem > attr) {
    AC_CHECK_CUDA(cudaFuncSetAttribute(quiet_lookup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const long long grid = (n + per_cta - 1) / per_cta;
  quiet_lookup_kernel<<<(unsigned)grid, kRefThreads, smem, (cudaStream_t)stream>>>(d_wave, n, win, per_cta, d_rms_db);
  AC_LAUNCH_CHECK();
  return AC_OK;
}
