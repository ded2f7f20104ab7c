// F0 and formant features of the legacy pause-detector branch:
//   ac_pyin          librosa.pyin(y, fmin=C2, fmax=C7, sr, hop_length=441)      pure_vocal_pause_detector.py:422-428
//   ac_lpc_formants  _extract_formants (pre-emphasis, Burg LPC(12), |1/A| peaks) pure_vocal_pause_detector.py:961-1018
// librosa is not part of the reference tree: the algorithms follow the published librosa 0.10
// implementation (core/pitch.py, sequence.py, core/audio.py) as restated in oracle/features.py.
//
// pYIN runs as three kernels:
//   1. yin_probs_kernel    one CTA per frame: difference function (direct form, fp32 products, the
//                          reference's FFT autocorrelation evaluates the same sums), cumulative mean
//                          normalisation, parabolic refinement, troughs, per-threshold Boltzmann prior x
//                          beta weights  ->  a short (pitch bin, probability) candidate list per frame;
//   2. pyin_viterbi_kernel one CTA per sequence, fp64: the 2 x 601-state HMM; the transition matrix is
//                          kron(t_switch, 41-wide triangle band), every zero entry is log(tiny) exactly as
//                          librosa's log(transition + tiny), so one block-wide max covers the out-of-band
//                          predecessors; back pointers go to global memory;
//   3. pyin_backtrack_kernel  pointer chase from the best final state.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace ac {

constexpr int kYinFrame = 2048;
constexpr int kYinWin = 1024;
constexpr int kYinThreads = 128;
constexpr int kYinMaxLags = 1024;   // max_period + 1 <= frame - win
constexpr int kYinMaxCand = 328;    // a 655-point CMND curve has at most 328 troughs
constexpr int kNThresholds = 100;
constexpr double kLogTiny = -708.3964185322641;  // log(np.finfo(float64).tiny)

struct YinParams {
  const float* x;
  long long n;
  int hop, sr;
  int min_period, max_period;  // 21, 675
  int n_pitch_bins, bins_per_semitone;
  float fmin;
  float boltzmann, no_trough_prob;
  const float* beta_probs;  // [100]
  long long n_frames;
  uint2* cand;              // [n_frames][kYinMaxCand] (bin, float bits of the probability)
  int* n_cand;              // [n_frames]
  float* voiced_prob;       // [n_frames]
};

__device__ __forceinline__ float block_sum_f(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(kYinThreads) yin_probs_kernel(const YinParams p) {
  __shared__ float y[kYinFrame];
  __shared__ float cs[kYinFrame + 1];     // cs[i] = sum_{j<i} y[j]^2
  __shared__ float yin[kYinMaxLags];      // difference function, then CMND (indexed by tau - min_period)
  __shared__ float probs[kYinMaxLags];    // per CMND index: probability mass (0 = not a candidate)
  __shared__ int trough_list[kYinMaxCand];
  __shared__ float red[8];
  __shared__ float wsum[kYinThreads / 32 + 1];
  __shared__ int n_troughs_s, gmin_idx_s;

  const long long frame = blockIdx.x;
  const int tid = threadIdx.x;
  const long long start = frame * p.hop - kYinFrame / 2;  // centred, zero padded
  for (int i = tid; i < kYinFrame; i += kYinThreads) {
    const long long s = start + i;
    y[i] = (s >= 0 && s < p.n) ? __ldg(p.x + s) : 0.f;
  }
  __syncthreads();
  // ---- energy prefix sums (block scan of y^2; each thread owns 16 consecutive samples)
  {
    constexpr int PER = kYinFrame / kYinThreads;
    float loc[PER];
    float run = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const float v = y[tid * PER + i];
      run += v * v;
      loc[i] = run;
    }
    // exclusive scan of the per-thread totals
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float u = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += u;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = incl;
    __syncthreads();
    float base = 0.f;
    for (int w = 0; w < (tid >> 5); ++w) base += wsum[w];
    const float excl = base + incl - run;
    if (tid == 0) cs[0] = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) cs[tid * PER + i + 1] = excl + loc[i];
  }
  __syncthreads();
  // ---- difference function d[tau] = E[0] + E[tau] - 2 acf[tau], tau = 0..max_period
  const int n_lags = p.max_period + 1;
  const float e0_raw = cs[kYinWin] - cs[0];
  const float e0 = fabsf(e0_raw) < 1e-6f ? 0.f : e0_raw;
  for (int tau = tid; tau < n_lags; tau += kYinThreads) {
    // acf for the reference's |.| < 1e-6 zeroing rules; the difference itself is summed directly as
    // sum (y[j] - y[j+tau])^2, which is the same quantity without the cancellation of E0 + E - 2 acf
    float a0 = 0.f, a1 = 0.f, d0 = 0.f, d1 = 0.f;
    const float* ya = y + 1;
    const float* yb = y + 1 + tau;
#pragma unroll 4
    for (int j = 0; j < kYinWin; j += 2) {
      const float u0 = ya[j], v0 = yb[j], u1 = ya[j + 1], v1 = yb[j + 1];
      a0 = fmaf(u0, v0, a0);
      a1 = fmaf(u1, v1, a1);
      const float q0 = u0 - v0, q1 = u1 - v1;
      d0 = fmaf(q0, q0, d0);
      d1 = fmaf(q1, q1, d1);
    }
    const float acf = a0 + a1;
    const float e = cs[tau + kYinWin] - cs[tau];
    const bool tiny_terms = fabsf(acf) < 1e-6f || fabsf(e) < 1e-6f || fabsf(e0_raw) < 1e-6f;
    yin[tau] = tiny_terms ? e0 + (fabsf(e) < 1e-6f ? 0.f : e) - 2.f * (fabsf(acf) < 1e-6f ? 0.f : acf) : d0 + d1;
  }
  __syncthreads();
  // ---- cumulative mean normalisation: cmnd[tau] = d[tau] / (mean_{1..tau} d + tiny); done by one warp
  //      (sequential prefix over <= 675 values in 32-wide steps), result stored at index tau - min_period
  if (tid < 32) {
    float carry = 0.f;
    for (int b = 1; b < n_lags; b += 32) {
      const int tau = b + tid;
      float v = tau < n_lags ? yin[tau] : 0.f;
      float incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += u;
      }
      const float csum = carry + incl;
      carry = __shfl_sync(0xffffffffu, csum, 31);
      if (tau < n_lags) probs[tau] = v / (csum / (float)tau + 1.17549435e-38f);  // staging: CMND by tau
    }
  }
  __syncthreads();
  const int n_c = p.max_period - p.min_period + 1;
  for (int i = tid; i < n_c; i += kYinThreads) yin[i] = probs[i + p.min_period];
  __syncthreads();
  for (int i = tid; i < kYinMaxLags; i += kYinThreads) probs[i] = 0.f;
  if (tid == 0) { n_troughs_s = 0; gmin_idx_s = 0; }
  __syncthreads();
  // ---- troughs (util.localmin + the first-sample rule), compacted in index order by one warp
  if (tid < 32) {
    int count = 0;
    for (int b = 0; b < n_c; b += 32) {
      const int i = b + tid;
      bool tr = false;
      if (i < n_c) {
        const float v = yin[i];
        if (i == 0) tr = v < yin[1];
        else if (i == n_c - 1) tr = v < yin[i - 1];
        else tr = (v < yin[i - 1]) && (v <= yin[i + 1]);
      }
      const unsigned m = __ballot_sync(0xffffffffu, tr);
      if (tr) {
        const int pos = count + __popc(m & ((1u << tid) - 1u));
        if (pos < kYinMaxCand) trough_list[pos] = i;
      }
      count += __popc(m);
    }
    if (tid == 0) n_troughs_s = count < kYinMaxCand ? count : kYinMaxCand;
  }
  __syncthreads();
  const int nt = n_troughs_s;
  float voiced = 0.f;
  int n_out = 0;
  if (nt > 0) {
    // global minimum among the troughs (first one on ties, as np.argmin)
    if (tid < 32) {
      float best = INFINITY;
      int bi = 0x7fffffff;
      for (int k = tid; k < nt; k += 32) {
        const float h = yin[trough_list[k]];
        if (h < best || (h == best && k < bi)) { best = h; bi = k; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (tid == 0) gmin_idx_s = bi;
    }
    __syncthreads();
    // Per trough k (in index order): first threshold index it is below, k0 = #{thr[1..100] <= h}.
    // For threshold t (1-based, thr = t/100) the troughs below it are those with k0 < t; trough k's rank among
    // them is pos_t(k) = #{k' < k : k0(k') < t}.  One thread per trough walks the thresholds with a running
    // count of earlier troughs that have entered (entry times are read from shared memory).
    int* k0s = reinterpret_cast<int*>(cs);  // reuse: [kYinMaxCand]
    for (int k = tid; k < nt; k += kYinThreads) {
      const float h = yin[trough_list[k]];
      // thresholds are np.linspace(0,1,101)[1:], compared as float64 against the float32 h promoted exactly
      int c = 0;
      for (int t = 1; t <= kNThresholds; ++t)
        if (!((double)h < (double)t / 100.0)) c = t;
      k0s[k] = c;  // below thresholds t > c
    }
    __syncthreads();
    // n_troughs(t) = #{k : k0 < t}: histogram of entry times, prefix summed (100 entries, one thread)
    int* cnt_t = reinterpret_cast<int*>(cs) + kYinMaxCand;  // [101]
    if (tid == 0) {
      for (int t = 0; t <= kNThresholds; ++t) cnt_t[t] = 0;
      for (int k = 0; k < nt; ++k)
        if (k0s[k] < kNThresholds) cnt_t[k0s[k] + 1] += 1;  // enters at threshold k0+1
      for (int t = 1; t <= kNThresholds; ++t) cnt_t[t] += cnt_t[t - 1];
    }
    __syncthreads();
    const float lam = p.boltzmann;
    const float one_m = 1.f - expf(-lam);
    for (int k = tid; k < nt; k += kYinThreads) {
      const int myk0 = k0s[k];
      float acc = 0.f;
      if (myk0 < kNThresholds) {
        // entries of earlier troughs, as a per-threshold running rank
        int rank_at[kNThresholds + 1];
#pragma unroll 1
        for (int t = 0; t <= kNThresholds; ++t) rank_at[t] = 0;
        for (int k2 = 0; k2 < k; ++k2) {
          const int e = k0s[k2] + 1;
          if (e <= kNThresholds) rank_at[e] += 1;
        }
        int run = 0;
        for (int t = 1; t <= kNThresholds; ++t) {
          run += rank_at[t];
          if (t > myk0) {
            const int n_t = cnt_t[t];
            const float prior = one_m * expf(-lam * (float)run) / (1.f - expf(-lam * (float)n_t));
            acc = fmaf(prior, __ldg(p.beta_probs + t - 1), acc);
          }
        }
      }
      if (k == gmin_idx_s) {
        // thresholds the global minimum is NOT below: t <= k0  ->  no-trough mass goes to it
        float s = 0.f;
        for (int t = 0; t < myk0; ++t) s += __ldg(p.beta_probs + t);
        acc = fmaf(p.no_trough_prob, s, acc);
      }
      probs[trough_list[k]] = acc;
    }
    __syncthreads();
    // ---- candidates: refine the period, map to a pitch bin, emit in increasing-period order (one warp)
    if (tid < 32) {
      uint2* out = p.cand + frame * (long long)kYinMaxCand;
      int count = 0;
      for (int b = 0; b < nt; b += 32) {
        const int k = b + tid;
        bool on = false;
        int bin = 0;
        float pr = 0.f;
        if (k < nt) {
          const int i = trough_list[k];
          pr = probs[i];
          on = pr != 0.f;
          if (on) {
            float shift = 0.f;
            if (i > 0 && i < n_c - 1) {
              const float a = yin[i + 1] + yin[i - 1] - 2.f * yin[i];
              const float bb = (yin[i + 1] - yin[i - 1]) * 0.5f;
              shift = fabsf(bb) >= fabsf(a) ? 0.f : -bb / a;
            }
            const float period = (float)(p.min_period + i) + shift;
            const double bidx = 12.0 * p.bins_per_semitone * log2(((double)p.sr / (double)period) / (double)p.fmin);
            double rb = rint(bidx);
            if (rb < 0.0) rb = 0.0;
            if (rb > (double)p.n_pitch_bins) rb = (double)p.n_pitch_bins;
            bin = (int)rb;
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (on) out[count + __popc(m & ((1u << tid) - 1u))] = make_uint2((unsigned)bin, __float_as_uint(pr));
        count += __popc(m);
      }
      n_out = count;
      // voiced probability = sum over DISTINCT bins < n_pitch_bins of the LAST candidate written to the bin
      // (bins decrease with the period, so equal bins are adjacent: the last of a run wins)
      __syncwarp();
      float s = 0.f;
      for (int c = tid; c < count; c += 32) {
        const uint2 me = out[c];
        const bool last = (c == count - 1) || (out[c + 1].x != me.x);
        if (last && (int)me.x < p.n_pitch_bins) s += __uint_as_float(me.y);
      }
      s = warp_sum(s);
      voiced = fminf(fmaxf(s, 0.f), 1.f);
    }
  }
  if (tid == 0) {
    p.n_cand[frame] = n_out;
    p.voiced_prob[frame] = voiced;
  }
  (void)red;
  (void)block_sum_f;
}

// ---- Viterbi ---------------------------------------------------------------------------------------
// log T[(v,k) -> (v',j)] = log t_switch[v][v'] + log tri[j-k] - log rowsum[k]  inside the band |j-k| <= half,
// log(tiny) outside (librosa: log(transition + tiny)).  Per step a thread owns pitch bin j (both voicings):
//   m_v = max_k (val_v[k] - lognorm[k]) + logtri[j-k],   m_u likewise over the unvoiced states,
//   into voiced j:   max(m_v + log stay, m_u + log switch, gmax + log tiny)   (first maximum on ties)
// Two block barriers per step; the next step's observations are prefetched and scattered while this
// step's maxima are being taken.
constexpr int kVitThreads = 640;
constexpr int kVitMaxHalf = 32;
struct VitParams {
  const uint2* cand;
  const int* n_cand;
  const float* voiced_prob;
  long long n_steps;
  int nb;    // pitch bins (601)
  int half;  // band half width (20)
  const double* logtri;   // [2*half+1] log of the un-normalised triangle
  const double* lognorm;  // [nb] log of the row sums (rows are truncated at the edges)
  double log_stay, log_switch, log_init;
  unsigned short* ptr;  // [n_steps][2*nb]
  int* final_state;
};

// adj is stored with kVitPad cells of -inf on either side of each [nb] row, so the fixed-width fast path can read its
// whole band without clipping (an out-of-range predecessor is -inf + logtri = -inf and never wins a strict >).
constexpr int kVitPad = kVitMaxHalf + 2;
__host__ __device__ inline int vit_row(int nb) { return nb + 2 * kVitPad; }

// HALF > 0: compile-time band half width (librosa's default geometry gives 20): the triangle lives in registers
// (it is symmetric: HALF + 1 values), a thread owns TWO adjacent pitch bins and slides one window of
// 2*HALF + 2 predecessors over both, i.e. per step 84 shared loads for 164 candidates instead of 246 -
// the loop was shared-memory bound.  HALF == 0: run-time width (any hop / sr), one bin per thread.
template <int HALF>
__global__ void __launch_bounds__(kVitThreads, 1) pyin_viterbi_kernel(const VitParams p) {
  extern __shared__ double vsm[];
  const int nb = p.nb, S = 2 * nb, half = p.half;
  const int row = vit_row(nb);
  double* val = vsm;                 // [2][S] raw values (double buffered)
  double* adj = val + 2 * S;         // [2][2][row] value - lognorm[k], padded rows (voiced, unvoiced)
  double* obs = adj + 4 * row;       // [2][nb] log observation of the voiced states (double buffered)
  double* tri = obs + 2 * nb;        // [2*half+1]
  double* lnorm = tri + (2 * kVitMaxHalf + 1);  // [nb]
  double* redv = lnorm + nb;         // [32]
  int* redi = reinterpret_cast<int*>(redv + 32);  // [32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kVitThreads / 32;
  const double tiny = 2.2250738585072014e-308;

  for (int i = tid; i < 2 * half + 1; i += kVitThreads) tri[i] = p.logtri[i];
  for (int i = tid; i < nb; i += kVitThreads) { lnorm[i] = p.lognorm[i]; obs[i] = kLogTiny; obs[nb + i] = kLogTiny; }
  for (int i = tid; i < 4 * row; i += kVitThreads) adj[i] = -INFINITY;
  __syncthreads();
  double T[HALF + 1];  // T[d] = logtri[half + d] = logtri[half - d]
#pragma unroll
  for (int d = 0; d <= HALF; ++d) T[d] = tri[half + d];
  // Observation inputs are software-pipelined through registers two steps ahead, so no global-memory
  // latency sits on the per-step critical path: thread i owns candidate i of a frame (n_cand <= 328 < 640).
  struct ObsIn {
    int nc;
    float vp;
    uint2 me, nxt;
  };
  auto fetch = [&](long long t) {
    ObsIn o;
    o.nc = 0; o.vp = 0.f; o.me = make_uint2(0, 0); o.nxt = make_uint2(0, 0);
    if (t < p.n_steps) {
      o.nc = p.n_cand[t];
      o.vp = p.voiced_prob[t];
      if (tid < kYinMaxCand) {
        const uint2* c = p.cand + t * (long long)kYinMaxCand;
        o.me = c[tid];
        o.nxt = tid + 1 < kYinMaxCand ? c[tid + 1] : make_uint2(0xffffffffu, 0);
      }
    }
    return o;
  };
  // returns the bin this thread wrote (or -1): last candidate of a run of equal bins wins
  auto scatter = [&](const ObsIn& o, int ob) -> int {
    if (tid < o.nc) {
      const bool last = (tid == o.nc - 1) || (o.nxt.x != o.me.x);
      if (last && (int)o.me.x < nb) {
        obs[ob * nb + o.me.x] = log((double)__uint_as_float(o.me.y) + tiny);
        return (int)o.me.x;
      }
    }
    return -1;
  };
  ObsIn in0 = fetch(0), in1 = fetch(1), in2 = fetch(2);
  int wrote_cur = scatter(in0, 0);
  __syncthreads();
  {
    const double lu0 = log((1.0 - (double)in0.vp) / (double)nb + tiny);
    for (int s = tid; s < S; s += kVitThreads) {
      const int k = s < nb ? s : s - nb;
      const double v = (s < nb ? obs[s] : lu0) + p.log_init;
      val[s] = v;
      adj[(s < nb ? 0 : row) + kVitPad + k] = v - lnorm[k];
    }
  }
  __syncthreads();
  if (wrote_cur >= 0) obs[wrote_cur] = kLogTiny;
  wrote_cur = scatter(in1, 1);  // observations of step 1 live in buffer 1
  __syncthreads();
  int cur = 0;
  // in1 = inputs of step t (already scattered), in2 = inputs of step t+1
  for (long long t = 1; t < p.n_steps; ++t) {
    const int ob = (int)(t & 1);
    const double* cv = val + cur * S;
    const double* ca = adj + cur * 2 * row + kVitPad;  // ca[k] voiced, ca[row + k] unvoiced
    const ObsIn in3 = fetch(t + 2);  // issued now, consumed two steps later
    // ---- P1: per-warp (max, first argmax) of the previous values + the in-band maxima
    {
      double bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int s = tid; s < S; s += kVitThreads) {
        const double v = cv[s];
        if (v > bv || (v == bv && s < bi)) { bv = v; bi = s; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { redv[warp] = bv; redi[warp] = bi; }
      if (tid == kVitThreads - 1) redv[31] = log((1.0 - (double)in1.vp) / (double)nb + tiny);
    }
    // bins owned by this thread: HALF > 0 -> {2 tid, 2 tid + 1}, else {tid}
    constexpr int kOwn = HALF > 0 ? 2 : 1;
    const int jbase = kOwn * tid;
    double mv[kOwn], mu[kOwn];
    int av[kOwn], au[kOwn];
#pragma unroll
    for (int o = 0; o < kOwn; ++o) { mv[o] = -INFINITY; mu[o] = -INFINITY; av[o] = 0; au[o] = 0; }
    if (jbase < nb) {
      if constexpr (HALF > 0) {
        // window k = jbase - HALF + i, i = 0 .. 2 HALF + 1; bin jbase sees d = i - HALF (i <= 2 HALF), bin jbase + 1
        // sees d = i - HALF - 1 (i >= 1); candidates are visited in increasing k for either bin, as before
        const double* wv = ca + (jbase - HALF);
        const double* wu = wv + row;
#pragma unroll
        for (int i = 0; i <= 2 * HALF + 1; ++i) {
          const double cvk = wv[i], cuk = wu[i];
          const int k = jbase - HALF + i;
          if (i <= 2 * HALF) {
            const double l = T[i >= HALF ? i - HALF : HALF - i];
            const double a = cvk + l, b = cuk + l;
            if (a > mv[0]) { mv[0] = a; av[0] = k; }
            if (b > mu[0]) { mu[0] = b; au[0] = k; }
          }
          if (i >= 1) {
            const double l = T[i - 1 >= HALF ? i - 1 - HALF : HALF - (i - 1)];
            const double a = cvk + l, b = cuk + l;
            if (a > mv[kOwn - 1]) { mv[kOwn - 1] = a; av[kOwn - 1] = k; }
            if (b > mu[kOwn - 1]) { mu[kOwn - 1] = b; au[kOwn - 1] = k; }
          }
        }
      } else {
        const int j = jbase;
        const int k_lo = j - half < 0 ? 0 : j - half;
        const int k_hi = j + half > nb - 1 ? nb - 1 : j + half;
        const double* tr = tri + (half - j);  // tr[k] = logtri[k - j + half]: the triangle is symmetric
        for (int k = k_lo; k <= k_hi; ++k) {
          const double l = tr[k];
          const double a = ca[k] + l, b = ca[row + k] + l;
          if (a > mv[0]) { mv[0] = a; av[0] = k; }
          if (b > mu[0]) { mu[0] = b; au[0] = k; }
        }
      }
    }
    __syncthreads();  // S1: warp partials visible
    // ---- P2: finish, write the new values, scatter the next step's observations into the other buffer
    int wrote_next;
    {
      // block maximum from the warp partials: one partial per lane, shuffle tree (every warp does it for itself)
      double gv = lane < kWarps ? redv[lane] : -INFINITY;
      int gi = lane < kWarps ? redi[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, gv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, gi, o);
        if (ov > gv || (ov == gv && oi < gi)) { gv = ov; gi = oi; }
      }
#pragma unroll
      for (int o = 0; o < kOwn; ++o) {
        const int j = jbase + o;
        if (j < nb) {
          const double oob = gv + kLogTiny;
          // candidates in increasing state index: voiced k, then unvoiced k; replace only on strict improvement
          double best_v = mv[o] + p.log_stay, best_u = mv[o] + p.log_switch;
          int arg_v = av[o], arg_u = av[o];
          const double c2v = mu[o] + p.log_switch, c2u = mu[o] + p.log_stay;
          if (c2v > best_v) { best_v = c2v; arg_v = nb + au[o]; }
          if (c2u > best_u) { best_u = c2u; arg_u = nb + au[o]; }
          if (oob > best_v || (oob == best_v && gi < arg_v)) { best_v = oob; arg_v = gi; }
          if (oob > best_u || (oob == best_u && gi < arg_u)) { best_u = oob; arg_u = gi; }
          const double lu = redv[31];  // log observation of the unvoiced states, published in P1
          const double nv = obs[ob * nb + j] + best_v, nu = lu + best_u;
          double* wv = val + (cur ^ 1) * S;
          double* wa = adj + (cur ^ 1) * 2 * row + kVitPad;
          const double ln = lnorm[j];
          wv[j] = nv; wv[nb + j] = nu;
          wa[j] = nv - ln; wa[row + j] = nu - ln;
          unsigned short* pr = p.ptr + t * (long long)S;
          pr[j] = (unsigned short)arg_v;
          pr[nb + j] = (unsigned short)arg_u;
        }
      }
      wrote_next = scatter(in2, ob ^ 1);
    }
    __syncthreads();  // S2: new values visible, everyone is done reading obs[ob]
    if (wrote_cur >= 0) obs[ob * nb + wrote_cur] = kLogTiny;  // un-scatter this step's observations
    wrote_cur = wrote_next;
    in1 = in2;
    in2 = in3;
    cur ^= 1;
  }
  __syncthreads();
  // final state = first argmax
  if (tid < 32) {
    const double* cv = val + cur * S;
    double bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int s = tid; s < S; s += 32) {
      const double v = cv[s];
      if (v > bv || (v == bv && s < bi)) { bv = v; bi = s; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (tid == 0) *p.final_state = bi;
  }
}

// ---- fast Viterbi: ONE barrier per step ---------------------------------------------------------------------
// Same recurrence, same arithmetic and visiting order as pyin_viterbi_kernel (outputs are bit-identical; the
// generic kernel stays as the fallback for other band widths and as the cross-check), restructured so that the
// only things left on the sequential critical path are the 2 x 42-predecessor window and one block barrier:
//   * warps 0-9 own two adjacent pitch bins per thread (triangle in registers, see above); the block maximum the
//     out-of-band transitions need is reduced from the NEW values while they are still in registers (one partial
//     per warp, published with the step's barrier) instead of a separate pass over shared memory;
//   * warps 10-20 own the candidates: they prefetch frame t+3, take the fp64 log of frame t+2's probabilities,
//     scatter frame t+1's observations into a 3-deep ring and un-scatter frame t-1's - all concurrently with the
//     compute warps, never between two barriers of the same step; thread 320 does the same for the unvoiced term.
constexpr int kVitFastThreads = 672;
constexpr int kVitFastComp = 320;  // threads of the compute warps
template <int HALF>
__global__ void __launch_bounds__(kVitFastThreads, 1) pyin_viterbi_fast_kernel(const VitParams p) {
  extern __shared__ double vsm[];
  const int nb = p.nb, S = 2 * nb;
  const int row = vit_row(nb);
  double* adj = vsm;               // [2][2][row] value - lognorm[k], padded rows (voiced, unvoiced), double buffered
  double* obs = adj + 4 * row;     // [3][nb] log observation of the voiced states, ring over steps
  double* lnorm = obs + 3 * nb;    // [nb]
  double* redv = lnorm + nb;       // [2][16] per-warp maxima of the values of a step
  double* lu_s = redv + 32;        // [2] log observation of the unvoiced states
  int* redi = reinterpret_cast<int*>(lu_s + 2);  // [2][16]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double tiny = 2.2250738585072014e-308;
  constexpr int kCompWarps = kVitFastComp / 32;
  const int jbase = 2 * tid;
  const bool is_comp = tid < kVitFastComp;
  const int cidx = kVitFastThreads - 1 - tid;  // candidate slot (threads >= 320: slots 351 .. 0; >= kYinMaxCand unused)
  const bool is_cand = !is_comp && cidx < kYinMaxCand;
  const bool is_lu = tid == kVitFastComp;

  for (int i = tid; i < nb; i += kVitFastThreads) lnorm[i] = p.lognorm[i];
  for (int i = tid; i < 3 * nb; i += kVitFastThreads) obs[i] = kLogTiny;
  for (int i = tid; i < 4 * row; i += kVitFastThreads) adj[i] = -INFINITY;
  double T[HALF + 1];  // T[d] = logtri[half + d] = logtri[half - d]
#pragma unroll
  for (int d = 0; d <= HALF; ++d) T[d] = p.logtri[HALF + d];
  __syncthreads();

  struct Stage {
    int nc;
    unsigned bin, nbin;
    float prob;
  };
  auto fetch = [&](long long t) {
    Stage st;
    st.nc = 0; st.bin = 0; st.nbin = 0xffffffffu; st.prob = 0.f;
    if (is_cand && t < p.n_steps) {
      st.nc = p.n_cand[t];
      const uint2* c = p.cand + t * (long long)kYinMaxCand;
      const uint2 me = c[cidx];
      st.bin = me.x;
      st.prob = __uint_as_float(me.y);
      st.nbin = cidx + 1 < kYinMaxCand ? c[cidx + 1].x : 0xffffffffu;
    }
    return st;
  };
  // the last candidate of a run of equal bins wins (numpy fancy assignment)
  auto writes = [&](const Stage& st) { return cidx < st.nc && (cidx == st.nc - 1 || st.nbin != st.bin) && (int)st.bin < nb; };
  auto lu_of = [&](long long t) { return log((1.0 - (double)p.voiced_prob[t]) / (double)nb + tiny); };
  // warp partial (max, first index) of this thread's new values -> red[buf][warp]
  auto publish = [&](double bv, int bi, int buf) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { redv[buf * 16 + warp] = bv; redi[buf * 16 + warp] = bi; }
  };

  // ---- step 0 ----
  Stage s0 = fetch(0), s1 = fetch(1), sA = fetch(2), sB = fetch(3);
  int wr0 = -1, wr1 = -1;  // bins this thread scattered for the previous / the current step
  if (is_cand) {
    if (writes(s0)) { obs[s0.bin] = log((double)s0.prob + tiny); wr0 = (int)s0.bin; }
    if (writes(s1)) { obs[nb + s1.bin] = log((double)s1.prob + tiny); wr1 = (int)s1.bin; }
  }
  double lgA = (is_cand && cidx < sA.nc) ? log((double)sA.prob + tiny) : 0.0;
  float vp_next = 0.f;  // voiced_prob of the NEXT step, loaded one step before its logarithm is taken
  if (is_lu) {
    lu_s[0] = lu_of(0);
    if (p.n_steps > 1) lu_s[1] = lu_of(1);
    if (p.n_steps > 2) vp_next = p.voiced_prob[2];
  }
  __syncthreads();
  if (is_comp) {
    double bv = -INFINITY;
    int bi = 0x7fffffff;
    const double lu0 = lu_s[0];
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      const int j = jbase + o;
      if (j < nb) {
        const double vv = obs[j] + p.log_init, vu = lu0 + p.log_init;
        adj[kVitPad + j] = vv - lnorm[j];
        adj[row + kVitPad + j] = vu - lnorm[j];
        if (vv > bv || (vv == bv && j < bi)) { bv = vv; bi = j; }
      }
    }
    // unvoiced states have higher indices than every voiced one: they win only on a strictly larger value here,
    // but a LOWER-indexed unvoiced state must beat a higher-indexed one on ties -> second pass in index order
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      const int j = jbase + o;
      if (j < nb) {
        const double vu = lu0 + p.log_init;
        if (vu > bv || (vu == bv && nb + j < bi)) { bv = vu; bi = nb + j; }
      }
    }
    publish(bv, bi, 0);
  }
  __syncthreads();

  for (long long t = 1; t < p.n_steps; ++t) {
    const int rb = (int)((t - 1) & 1), wb = (int)(t & 1);  // adj / partial buffers read and written by this step
    if (is_comp) {
      // block maximum (and its first index) of the previous step's values, from the per-warp partials
      double gv = -INFINITY;
      int gi = 0x7fffffff;
#pragma unroll
      for (int w = 0; w < kCompWarps; ++w) {
        const double ov = redv[rb * 16 + w];
        const int oi = redi[rb * 16 + w];
        if (ov > gv || (ov == gv && oi < gi)) { gv = ov; gi = oi; }
      }
      double mv[2], mu[2];
      int av[2], au[2];
#pragma unroll
      for (int o = 0; o < 2; ++o) { mv[o] = -INFINITY; mu[o] = -INFINITY; av[o] = 0; au[o] = 0; }
      double bv = -INFINITY;
      int bi = 0x7fffffff;
      if (jbase < nb) {
        const double* wv = adj + rb * 2 * row + kVitPad + (jbase - HALF);
        const double* wu = wv + row;
#pragma unroll
        for (int i = 0; i <= 2 * HALF + 1; ++i) {
          const double cvk = wv[i], cuk = wu[i];
          const int k = jbase - HALF + i;
          if (i <= 2 * HALF) {
            const double l = T[i >= HALF ? i - HALF : HALF - i];
            const double a = cvk + l, b = cuk + l;
            if (a > mv[0]) { mv[0] = a; av[0] = k; }
            if (b > mu[0]) { mu[0] = b; au[0] = k; }
          }
          if (i >= 1) {
            const double l = T[i - 1 >= HALF ? i - 1 - HALF : HALF - (i - 1)];
            const double a = cvk + l, b = cuk + l;
            if (a > mv[1]) { mv[1] = a; av[1] = k; }
            if (b > mu[1]) { mu[1] = b; au[1] = k; }
          }
        }
        const double oob = gv + kLogTiny;
        const double lu = lu_s[wb];
        const double* ob = obs + (int)(t % 3) * nb;
        double* wa = adj + wb * 2 * row + kVitPad;
        unsigned short* pr = p.ptr + t * (long long)S;
        double nvv[2], nuu[2];
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int j = jbase + o;
          nvv[o] = -INFINITY; nuu[o] = -INFINITY;
          if (j < nb) {
            // candidates in increasing state index: voiced k, then unvoiced k; replace only on strict improvement
            double best_v = mv[o] + p.log_stay, best_u = mv[o] + p.log_switch;
            int arg_v = av[o], arg_u = av[o];
            const double c2v = mu[o] + p.log_switch, c2u = mu[o] + p.log_stay;
            if (c2v > best_v) { best_v = c2v; arg_v = nb + au[o]; }
            if (c2u > best_u) { best_u = c2u; arg_u = nb + au[o]; }
            if (oob > best_v || (oob == best_v && gi < arg_v)) { best_v = oob; arg_v = gi; }
            if (oob > best_u || (oob == best_u && gi < arg_u)) { best_u = oob; arg_u = gi; }
            const double nv = ob[j] + best_v, nu = lu + best_u;
            const double ln = lnorm[j];
            wa[j] = nv - ln;
            wa[row + j] = nu - ln;
            pr[j] = (unsigned short)arg_v;
            pr[nb + j] = (unsigned short)arg_u;
            nvv[o] = nv; nuu[o] = nu;
          }
        }
        // first maximum in state-index order: voiced j, j+1, then unvoiced nb+j, nb+j+1
#pragma unroll
        for (int o = 0; o < 2; ++o)
          if (jbase + o < nb && (nvv[o] > bv || (nvv[o] == bv && jbase + o < bi))) { bv = nvv[o]; bi = jbase + o; }
#pragma unroll
        for (int o = 0; o < 2; ++o)
          if (jbase + o < nb && (nuu[o] > bv || (nuu[o] == bv && nb + jbase + o < bi))) { bv = nuu[o]; bi = nb + jbase + o; }
      }
      publish(bv, bi, wb);
    } else {
      // ---- candidate / unvoiced-term warps: everything here targets buffers no compute warp touches in this step ----
      if (is_cand) {
        if (wr0 >= 0) obs[(int)((t - 1) % 3) * nb + wr0] = kLogTiny;  // un-scatter step t-1 (read before the last barrier)
        int wrn = -1;
        if (writes(sA)) { obs[(int)((t + 1) % 3) * nb + sA.bin] = lgA; wrn = (int)sA.bin; }  // step t+1's observations
        wr0 = wr1;
        wr1 = wrn;
        const Stage sC = fetch(t + 3);
        lgA = cidx < sB.nc ? log((double)sB.prob + tiny) : 0.0;  // step t+2's, scattered during step t+1
        sA = sB;
        sB = sC;
      }
      if (is_lu && t + 1 < p.n_steps) {  // invariant: vp_next = voiced_prob[t + 1] on entry to step t
        lu_s[(t + 1) & 1] = log((1.0 - (double)vp_next) / (double)nb + tiny);
        vp_next = t + 2 < p.n_steps ? p.voiced_prob[t + 2] : 0.f;
      }
    }
    __syncthreads();
  }
  // final state = first argmax of the last step's values
  if (tid == 0) {
    const int rb = (int)((p.n_steps - 1) & 1);
    double gv = -INFINITY;
    int gi = 0x7fffffff;
    for (int w = 0; w < kCompWarps; ++w) {
      const double ov = redv[rb * 16 + w];
      const int oi = redi[rb * 16 + w];
      if (ov > gv || (ov == gv && oi < gi)) { gv = ov; gi = oi; }
    }
    *p.final_state = gi;
  }
}

// ---- cluster Viterbi: the pitch bins split over a 4-CTA cluster ------------------------------------------------
// One SM issues the ~7.7 k warp-instructions of a step (601 bins x 82 predecessors x {add, compare, 3 selects}) in
// ~6 k cycles; four SMs each take a quarter of the bins.  A CTA keeps the adj rows of its own bins plus a HALF-wide
// halo on either side; after a step it stores its new values locally, pushes its outermost HALF bins into the two
// neighbours' halos and its per-warp maxima into every CTA's partial table through distributed shared memory
// (st.shared::cluster), then all CTAs meet at ONE cluster barrier (arrive.release / wait.acquire).  Arithmetic,
// visiting order and tie rules are those of the single-CTA kernels: the outputs are bit-identical.
constexpr int kVcN = 4;
constexpr int kVcComp = 160;   // compute threads per CTA (one pitch bin each; Q = ceil(nb / 4) <= 160)
constexpr int kVcCand = 96;    // candidate threads per CTA; each owns kVcSlots candidate slots (slot = ct + 96 k).  Few
constexpr int kVcSlots = 4;    // warps on purpose: the cluster barrier of a step gets slower with every warp that joins it
constexpr int kVcThreads = kVcComp + kVcCand;
static_assert(kVcCand * kVcSlots >= kYinMaxCand, "candidate slots");
__device__ __forceinline__ uint32_t vc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t vc_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void vc_st_f64(uint32_t cluster_addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(cluster_addr), "d"(v) : "memory");
}
__device__ __forceinline__ void vc_st_s32(uint32_t cluster_addr, int v) {
  asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void vc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int HALF>
__global__ void __cluster_dims__(kVcN, 1, 1) __launch_bounds__(kVcThreads, 1) pyin_viterbi_cluster_kernel(const VitParams p) {
  extern __shared__ double vsm[];
  const int nb = p.nb, S = 2 * nb;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int Q = (nb + kVcN - 1) / kVcN;           // bins per CTA (the last CTA may own fewer)
  const int j0 = (int)rank * Q;
  const int nq = max(0, min(Q, nb - j0));
  const int rowlen = Q + 2 * HALF + 2;            // local index of bin j: j - j0 + HALF
  constexpr int kCompWarps = kVcComp / 32;
  constexpr int kParts = kVcN * kCompWarps;       // warp partials of the whole cluster
  double* adj = vsm;                              // [2][2][rowlen]
  double* obs = adj + 4 * rowlen;                 // [3][Q]
  double* lnorm = obs + 3 * Q;                    // [Q]
  double* partv = lnorm + Q;                      // [2][kParts]
  double* lu_s = partv + 2 * kParts;              // [2]
  int* parti = reinterpret_cast<int*>(lu_s + 2);  // [2][kParts]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double tiny = 2.2250738585072014e-308;
  const bool is_comp = tid < kVcComp;
  const int c = tid;                              // local bin of a compute thread
  const int j = j0 + c;
  const bool owns = is_comp && c < nq;
  const int ct = tid - kVcComp;                   // candidate thread index (0 .. 95) of the non-compute warps
  const bool is_cand = !is_comp;
  const bool is_lu = tid == kVcComp;

  for (int i = tid; i < Q; i += kVcThreads) lnorm[i] = j0 + i < nb ? p.lognorm[j0 + i] : 0.0;
  for (int i = tid; i < 3 * Q; i += kVcThreads) obs[i] = kLogTiny;
  for (int i = tid; i < 4 * rowlen; i += kVcThreads) adj[i] = -INFINITY;
  for (int i = tid; i < 2 * kParts; i += kVcThreads) { partv[i] = -INFINITY; parti[i] = 0x7fffffff; }
  double T[HALF + 1];
#pragma unroll
  for (int d = 0; d <= HALF; ++d) T[d] = p.logtri[HALF + d];
  __syncthreads();
  vc_cluster_sync();  // every CTA's shared memory is initialised before anybody stores into it

  struct Stage {
    int nc;
    unsigned bin[kVcSlots], nbin[kVcSlots];
    float prob[kVcSlots];
  };
  auto fetch = [&](long long t) {
    Stage st;
    st.nc = 0;
#pragma unroll
    for (int k = 0; k < kVcSlots; ++k) { st.bin[k] = 0; st.nbin[k] = 0xffffffffu; st.prob[k] = 0.f; }
    if (is_cand && t < p.n_steps) {
      st.nc = p.n_cand[t];
      const uint2* cd = p.cand + t * (long long)kYinMaxCand;
#pragma unroll
      for (int k = 0; k < kVcSlots; ++k) {
        const int idx = ct + kVcCand * k;
        if (idx < st.nc) {
          const uint2 me = cd[idx];
          st.bin[k] = me.x;
          st.prob[k] = __uint_as_float(me.y);
          st.nbin[k] = idx + 1 < kYinMaxCand ? cd[idx + 1].x : 0xffffffffu;
        }
      }
    }
    return st;
  };
  // last candidate of a run of equal bins wins; only bins of this CTA are scattered (local index returned, else -1)
  auto target = [&](const Stage& st, int k) -> int {
    const int idx = ct + kVcCand * k;
    if (idx < st.nc && (idx == st.nc - 1 || st.nbin[k] != st.bin[k]) && (int)st.bin[k] < nb) {
      const int li = (int)st.bin[k] - j0;
      if (li >= 0 && li < nq) return li;
    }
    return -1;
  };
  auto lu_of = [&](long long t) { return log((1.0 - (double)p.voiced_prob[t]) / (double)nb + tiny); };
  // new values -> local adj row, the neighbours' halos, the cluster-wide partial table
  auto publish = [&](double nv, double nu, int wb) {
    double bv = -INFINITY;
    int bi = 0x7fffffff;
    if (owns) {
      const double ln = lnorm[c];
      const double av_ = nv - ln, au_ = nu - ln;
      double* row_v = adj + (wb * 2 + 0) * rowlen;
      double* row_u = adj + (wb * 2 + 1) * rowlen;
      row_v[c + HALF] = av_;
      row_u[c + HALF] = au_;
      if (c < HALF && rank > 0) {  // left neighbour's right halo: its local index of bin j is c + Q + HALF
        vc_st_f64(vc_mapa(vc_smem_u32(row_v + c + Q + HALF), rank - 1), av_);
        vc_st_f64(vc_mapa(vc_smem_u32(row_u + c + Q + HALF), rank - 1), au_);
      }
      if (c >= Q - HALF && rank + 1 < kVcN) {  // right neighbour's left halo: local index c - Q + HALF
        vc_st_f64(vc_mapa(vc_smem_u32(row_v + c - Q + HALF), rank + 1), av_);
        vc_st_f64(vc_mapa(vc_smem_u32(row_u + c - Q + HALF), rank + 1), au_);
      }
      bv = nv; bi = j;
      if (nu > bv) { bv = nu; bi = nb + j; }  // equal values: the voiced state has the lower index
    }
    if (is_comp) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane < kVcN) {  // lane q stores this warp's partial into CTA q's table
        const int slot = wb * kParts + (int)rank * kCompWarps + warp;
        vc_st_f64(vc_mapa(vc_smem_u32(partv + slot), (uint32_t)lane), bv);
        vc_st_s32(vc_mapa(vc_smem_u32(parti + slot), (uint32_t)lane), bi);
      }
    }
  };

  // ---- step 0 ----
  Stage s0 = fetch(0), s1 = fetch(1), sA = fetch(2), sB = fetch(3);
  int wr0[kVcSlots], wr1[kVcSlots];
  double lgA[kVcSlots];
#pragma unroll
  for (int k = 0; k < kVcSlots; ++k) {
    wr0[k] = -1; wr1[k] = -1; lgA[k] = 0.0;
    if (is_cand) {
      const int t0 = target(s0, k), t1 = target(s1, k);
      if (t0 >= 0) { obs[t0] = log((double)s0.prob[k] + tiny); wr0[k] = t0; }
      if (t1 >= 0) { obs[Q + t1] = log((double)s1.prob[k] + tiny); wr1[k] = t1; }
      if (ct + kVcCand * k < sA.nc) lgA[k] = log((double)sA.prob[k] + tiny);
    }
  }
  float vp_next = 0.f;  // voiced_prob of the NEXT step, loaded one step before its logarithm is taken
  if (is_lu) {
    lu_s[0] = lu_of(0);
    if (p.n_steps > 1) lu_s[1] = lu_of(1);
    if (p.n_steps > 2) vp_next = p.voiced_prob[2];
  }
  __syncthreads();
  {
    double nv = -INFINITY, nu = -INFINITY;
    if (owns) {
      nv = obs[c] + p.log_init;
      nu = lu_s[0] + p.log_init;
    }
    publish(nv, nu, 0);
  }
  vc_cluster_sync();

  for (long long t = 1; t < p.n_steps; ++t) {
    const int rb = (int)((t - 1) & 1), wb = (int)(t & 1);
    if (is_comp) {
      double gv = -INFINITY;
      int gi = 0x7fffffff;
#pragma unroll
      for (int w = 0; w < kParts; ++w) {
        const double ov = partv[rb * kParts + w];
        const int oi = parti[rb * kParts + w];
        if (ov > gv || (ov == gv && oi < gi)) { gv = ov; gi = oi; }
      }
      double nv = -INFINITY, nu = -INFINITY;
      if (owns) {
        double mv = -INFINITY, mu = -INFINITY;
        int av = 0, au = 0;
        const double* wv = adj + (rb * 2 + 0) * rowlen + c;  // local index of bin j - HALF
        const double* wu = adj + (rb * 2 + 1) * rowlen + c;
#pragma unroll
        for (int i = 0; i <= 2 * HALF; ++i) {
          const double l = T[i >= HALF ? i - HALF : HALF - i];
          const double a = wv[i] + l, b = wu[i] + l;
          const int k = j - HALF + i;
          if (a > mv) { mv = a; av = k; }
          if (b > mu) { mu = b; au = k; }
        }
        const double oob = gv + kLogTiny;
        double best_v = mv + p.log_stay, best_u = mv + p.log_switch;
        int arg_v = av, arg_u = av;
        const double c2v = mu + p.log_switch, c2u = mu + p.log_stay;
        if (c2v > best_v) { best_v = c2v; arg_v = nb + au; }
        if (c2u > best_u) { best_u = c2u; arg_u = nb + au; }
        if (oob > best_v || (oob == best_v && gi < arg_v)) { best_v = oob; arg_v = gi; }
        if (oob > best_u || (oob == best_u && gi < arg_u)) { best_u = oob; arg_u = gi; }
        nv = obs[(int)(t % 3) * Q + c] + best_v;
        nu = lu_s[wb] + best_u;
        unsigned short* pr = p.ptr + t * (long long)S;
        pr[j] = (unsigned short)arg_v;
        pr[nb + j] = (unsigned short)arg_u;
      }
      publish(nv, nu, wb);
    } else {
      if (is_cand) {
        const Stage sC = fetch(t + 3);
#pragma unroll
        for (int k = 0; k < kVcSlots; ++k) {
          if (wr0[k] >= 0) obs[(int)((t - 1) % 3) * Q + wr0[k]] = kLogTiny;
          const int tn = target(sA, k);
          if (tn >= 0) obs[(int)((t + 1) % 3) * Q + tn] = lgA[k];
          wr0[k] = wr1[k];
          wr1[k] = tn;
          lgA[k] = ct + kVcCand * k < sB.nc ? log((double)sB.prob[k] + tiny) : 0.0;
        }
        sA = sB;
        sB = sC;
      }
      if (is_lu && t + 1 < p.n_steps) {  // invariant: vp_next = voiced_prob[t + 1] on entry to step t
        lu_s[(t + 1) & 1] = log((1.0 - (double)vp_next) / (double)nb + tiny);
        vp_next = t + 2 < p.n_steps ? p.voiced_prob[t + 2] : 0.f;
      }
    }
    vc_cluster_sync();
  }
  if (rank == 0 && tid == 0) {
    const int rb = (int)((p.n_steps - 1) & 1);
    double gv = -INFINITY;
    int gi = 0x7fffffff;
    for (int w = 0; w < kParts; ++w) {
      const double ov = partv[rb * kParts + w];
      const int oi = parti[rb * kParts + w];
      if (ov > gv || (ov == gv && oi < gi)) { gv = ov; gi = oi; }
    }
    *p.final_state = gi;
  }
}

// Back pointers are consumed newest-first in blocks of kBtSteps steps staged through shared memory, so the
// pointer chase itself runs out of smem and the global reads are coalesced.
constexpr int kBtSteps = 64;
__global__ void __launch_bounds__(256) pyin_backtrack_kernel(const unsigned short* __restrict__ ptr, const int* __restrict__ final_state,
                                                             long long n_steps, int nb, float fmin, int bins_per_semitone,
                                                             float* __restrict__ f0, unsigned char* __restrict__ voiced_flag) {
  extern __shared__ unsigned short bsm[];  // [kBtSteps][2*nb]
  __shared__ int states[kBtSteps];
  __shared__ int carry;
  const int S = 2 * nb;
  if (threadIdx.x == 0) carry = *final_state;
  __syncthreads();
  for (long long hi = n_steps - 1; hi >= 0; hi -= kBtSteps) {
    const long long lo = hi - kBtSteps + 1 < 0 ? 0 : hi - kBtSteps + 1;  // steps lo..hi
    const int cnt = (int)(hi - lo + 1);
    // stage ptr[lo+1 .. hi] (ptr[t] maps the state at t to the state at t-1)
    const long long words = (long long)cnt * S;
    const unsigned short* src = ptr + lo * (long long)S;
    for (long long i = threadIdx.x; i < words; i += blockDim.x) bsm[i] = src[i];
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = carry;
      for (int r = cnt - 1; r >= 0; --r) {
        states[r] = s;
        if (lo + r > 0) s = bsm[(size_t)r * S + s];
      }
      carry = s;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < cnt; r += blockDim.x) {
      const int s = states[r];
      const bool v = s < nb;
      const int k = v ? s : s - nb;
      if (f0) f0[lo + r] = v ? fmin * exp2f((float)k / (12.f * (float)bins_per_semitone)) : __int_as_float(0x7fc00000);
      if (voiced_flag) voiced_flag[lo + r] = v ? 1 : 0;
    }
    __syncthreads();
  }
}

// ---- parallel backtrack -----------------------------------------------------------------------------------
// The serial kernel above stages ALL 2*nb back pointers of every step through one CTA to follow ONE of them
// (2.1 us per step: as long as the forward pass).  Back pointers are maps state(t) -> state(t-1), and maps compose:
//   A  (one CTA per 64-step block, whole GPU): compose the block's maps -> G[b][s] = state at the step below the
//      block for every possible state s at its top (64 shared-memory lookups per state);
//   B  (one CTA): chase the block tops through G, 64 tables at a time in shared memory;
//   C  (one CTA per block): walk the block's 64 steps from its known top state, emit f0 / voiced flags.
// The pointers are read once at full bandwidth instead of once through a single SM.
constexpr int kBpSteps = 64;
constexpr int kBpThreads = 640;
__device__ __forceinline__ void bp_stage(unsigned short* dst, const unsigned short* src, long long n_elems) {
  // src is 16-byte aligned (block starts are multiples of 64 rows of 4*nb bytes); bulk as uint4, tail as ushort
  const long long n16 = n_elems / 8;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (long long i = threadIdx.x; i < n16; i += blockDim.x) d4[i] = __ldg(s4 + i);
  for (long long i = n16 * 8 + threadIdx.x; i < n_elems; i += blockDim.x) dst[i] = src[i];
}
__global__ void __launch_bounds__(kBpThreads) pyin_bp_compose_kernel(const unsigned short* __restrict__ ptr, long long n_steps, int nb,
                                                                     unsigned short* __restrict__ G) {
  extern __shared__ __align__(16) unsigned short bsm[];  // [cnt][2*nb]
  const int S = 2 * nb;
  const long long b = blockIdx.x;
  if (b == 0) return;  // nothing lies below block 0
  const long long lo = b * kBpSteps;
  const int cnt = (int)min((long long)kBpSteps, n_steps - lo);
  bp_stage(bsm, ptr + lo * (long long)S, (long long)cnt * S);
  __syncthreads();
  for (int s0 = threadIdx.x; s0 < S; s0 += kBpThreads) {
    int x = s0;
    for (int r = cnt - 1; r >= 0; --r) x = bsm[(size_t)r * S + x];
    G[b * (long long)S + s0] = (unsigned short)x;
  }
}
__global__ void __launch_bounds__(256) pyin_bp_tops_kernel(const unsigned short* __restrict__ G, const int* __restrict__ final_state,
                                                           long long n_blk, int nb, int* __restrict__ top) {
  extern __shared__ __align__(16) unsigned short bsm[];  // [<=64][2*nb]
  __shared__ int carry;
  const int S = 2 * nb;
  if (threadIdx.x == 0) { carry = *final_state; top[n_blk - 1] = carry; }
  __syncthreads();
  // batches of 64 tables, aligned at multiples of 64 blocks so that every batch starts on a 16-byte boundary
  for (long long q = (n_blk - 1) / kBpSteps; q >= 0; --q) {
    const long long b_lo = q * kBpSteps, b_hi = min(n_blk - 1, b_lo + kBpSteps - 1);
    const int cnt = (int)(b_hi - b_lo + 1);
    bp_stage(bsm, G + b_lo * (long long)S, (long long)cnt * S);
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = carry;  // state at the top of block b_hi
      for (long long b = b_hi; b >= max(b_lo, 1LL); --b) {
        s = bsm[(size_t)(b - b_lo) * S + s];  // state at the top of block b - 1
        top[b - 1] = s;
      }
      carry = s;
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kBpSteps) pyin_bp_emit_kernel(const unsigned short* __restrict__ ptr, const int* __restrict__ top,
                                                                long long n_steps, int nb, float fmin, int bins_per_semitone,
                                                                float* __restrict__ f0, unsigned char* __restrict__ voiced_flag) {
  __shared__ int states[kBpSteps];
  const int S = 2 * nb;
  const long long lo = (long long)blockIdx.x * kBpSteps;
  const int cnt = (int)min((long long)kBpSteps, n_steps - lo);
  if (threadIdx.x == 0) {
    int s = top[blockIdx.x];
    for (int r = cnt - 1; r >= 0; --r) {
      states[r] = s;
      if (r > 0) s = ptr[(lo + r) * (long long)S + s];
    }
  }
  __syncthreads();
  const int r = threadIdx.x;
  if (r < cnt) {
    const int s = states[r];
    const bool v = s < nb;
    const int k = v ? s : s - nb;
    if (f0) f0[lo + r] = v ? fmin * exp2f((float)k / (12.f * (float)bins_per_semitone)) : __int_as_float(0x7fc00000);
    if (voiced_flag) voiced_flag[lo + r] = v ? 1 : 0;
  }
}

// ---- host side ---------------------------------------------------------------------------------
static double beta_cdf_int(double x, int a, int b) {  // regularised incomplete beta for integer a, b
  const int n = a + b - 1;
  double s = 0.0;
  for (int j = a; j <= n; ++j) {
    double c = 1.0;
    for (int i = 0; i < j; ++i) c = c * (double)(n - i) / (double)(i + 1);
    s += c * pow(x, j) * pow(1.0 - x, n - j);
  }
  return s;
}

struct PyinGeom {
  int win, min_period, max_period, bps, nb, half;
};
static PyinGeom pyin_geom(int sr, double fmin, double fmax, int frame_length, int hop) {
  PyinGeom g;
  g.win = frame_length / 2;
  g.min_period = (int)floor((double)sr / fmax);
  g.max_period = (int)ceil((double)sr / fmin);
  if (g.max_period > frame_length - g.win - 1) g.max_period = frame_length - g.win - 1;
  g.bps = 10;
  g.nb = (int)floor(12.0 * g.bps * log2(fmax / fmin)) + 1;
  const int semis = (int)llround(35.92 * 12.0 * hop / sr);
  g.half = (semis * g.bps + 1) / 2;
  return g;
}

}  // namespace ac

extern "C" long long ac_pyin_frame_count(long long n, int hop) { return (n < 0 || hop <= 0) ? 0 : 1 + n / hop; }

extern "C" size_t ac_pyin_workspace_bytes(long long n, int hop, int sr, float fmin, float fmax) {
  using namespace ac;
  if (n <= 0 || hop <= 0) return 0;
  const long long nf = 1 + n / hop;
  const PyinGeom g = pyin_geom(sr, fmin, fmax, kYinFrame, hop);
  size_t b = 0;
  b += align_up((size_t)nf * kYinMaxCand * sizeof(uint2), 256);
  b += align_up((size_t)nf * sizeof(int), 256);
  b += align_up((size_t)nf * 2 * g.nb * sizeof(unsigned short), 256);
  b += align_up((size_t)(g.nb + 2 * g.half + 1) * sizeof(double), 256);
  b += align_up((size_t)kNThresholds * sizeof(float), 256);
  const size_t n_blk = (size_t)((nf + kBpSteps - 1) / kBpSteps);
  b += align_up(n_blk * 2 * g.nb * sizeof(unsigned short), 256);  // composed back-pointer maps, one per 64 steps
  b += align_up(n_blk * sizeof(int), 256);                         // state at the top of every block
  b += 256;
  return b;
}

extern "C" int ac_pyin(const float* d_x, long long n, int sr, int hop, float fmin, float fmax, float* d_f0,
                       unsigned char* d_voiced_flag, float* d_voiced_prob, void* d_ws, size_t ws_bytes, void* stream) {
  using namespace ac;
  AC_REQUIRE(d_x && d_voiced_prob && d_ws && n > 0 && hop > 0 && sr > 0 && fmin > 0 && fmax > fmin, "ac_pyin: arguments");
  AC_REQUIRE(ws_bytes >= ac_pyin_workspace_bytes(n, hop, sr, fmin, fmax), "ac_pyin: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const PyinGeom g = pyin_geom(sr, fmin, fmax, kYinFrame, hop);
  AC_REQUIRE(g.max_period + 1 <= kYinMaxLags && g.min_period >= 1 && g.max_period > g.min_period + 2, "ac_pyin: period range");
  AC_REQUIRE((g.max_period - g.min_period + 2) / 2 <= kYinMaxCand, "ac_pyin: candidate capacity");
  AC_REQUIRE(2 * g.nb < 65536 && g.nb <= 1024, "ac_pyin: too many pitch bins");
  const long long nf = 1 + n / hop;
  char* w = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(d_ws) + 255) & ~uintptr_t(255));
  uint2* cand = reinterpret_cast<uint2*>(w); w += align_up((size_t)nf * kYinMaxCand * sizeof(uint2), 256);
  int* n_cand = reinterpret_cast<int*>(w); w += align_up((size_t)nf * sizeof(int), 256);
  unsigned short* ptr = reinterpret_cast<unsigned short*>(w); w += align_up((size_t)nf * 2 * g.nb * sizeof(unsigned short), 256);
  double* logtab = reinterpret_cast<double*>(w); w += align_up((size_t)(g.nb + 2 * g.half + 1) * sizeof(double), 256);
  float* beta = reinterpret_cast<float*>(w); w += align_up((size_t)kNThresholds * sizeof(float), 256);
  const long long n_blk = (nf + kBpSteps - 1) / kBpSteps;
  unsigned short* bp_maps = reinterpret_cast<unsigned short*>(w); w += align_up((size_t)n_blk * 2 * g.nb * sizeof(unsigned short), 256);
  int* bp_top = reinterpret_cast<int*>(w); w += align_up((size_t)n_blk * sizeof(int), 256);
  int* final_state = reinterpret_cast<int*>(w);

  // beta(2, 18) weights of the 100 thresholds and the banded log transition (row-normalised triangle)
  {
    std::vector<float> hb(kNThresholds);
    double prev = 0.0;
    for (int t = 1; t <= kNThresholds; ++t) {
      const double c = beta_cdf_int((double)t / kNThresholds, 2, 18);
      hb[t - 1] = (float)(c - prev);
      prev = c;
    }
    const int W = 2 * g.half + 1;
    // [0, W): log of scipy.signal.windows.triang(W) = 1 - |d| / ((W + 1) / 2);  [W, W + nb): log of the row sums
    std::vector<double> hl((size_t)W + g.nb);
    for (int d = -g.half; d <= g.half; ++d) hl[d + g.half] = log(1.0 - fabs((double)d) / ((W + 1) / 2.0));
    for (int k = 0; k < g.nb; ++k) {
      double sum = 0.0;
      for (int d = -g.half; d <= g.half; ++d)
        if (k + d >= 0 && k + d < g.nb) sum += 1.0 - fabs((double)d) / ((W + 1) / 2.0);
      hl[W + k] = log(sum);
    }
    AC_CHECK_CUDA(cudaMemcpyAsync(beta, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    AC_CHECK_CUDA(cudaMemcpyAsync(logtab, hl.data(), hl.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    AC_CHECK_CUDA(cudaStreamSynchronize(st));  // the host vectors die at the end of this scope
  }
  YinParams yp;
  yp.x = d_x; yp.n = n; yp.hop = hop; yp.sr = sr;
  yp.min_period = g.min_period; yp.max_period = g.max_period;
  yp.n_pitch_bins = g.nb; yp.bins_per_semitone = g.bps;
  yp.fmin = fmin; yp.boltzmann = 2.f; yp.no_trough_prob = 0.01f;
  yp.beta_probs = beta;
  yp.n_frames = nf;
  yp.cand = cand; yp.n_cand = n_cand; yp.voiced_prob = d_voiced_prob;
  {
    ProfScope ps(KC_MISC, 2.0 * nf * (double)(g.max_period + 1) * kYinWin, 4.0 * n, st);
    yin_probs_kernel<<<(unsigned)nf, kYinThreads, 0, st>>>(yp);
    AC_LAUNCH_CHECK();
  }
  if (d_f0 || d_voiced_flag) {
    VitParams vp;
    vp.cand = cand; vp.n_cand = n_cand; vp.voiced_prob = d_voiced_prob;
    vp.n_steps = nf; vp.nb = g.nb; vp.half = g.half; vp.logtri = logtab; vp.lognorm = logtab + (2 * g.half + 1);
    vp.log_stay = log(0.99); vp.log_switch = log(0.01);
    vp.log_init = log(1.0 / (2.0 * g.nb) + 2.2250738585072014e-308);
    vp.ptr = ptr; vp.final_state = final_state;
    AC_REQUIRE(g.half <= kVitMaxHalf, "ac_pyin: transition band too wide");
    const size_t smem = (size_t)(2 * 2 * g.nb + 4 * vit_row(g.nb) + 2 * g.nb + (2 * kVitMaxHalf + 1) + g.nb + 32) * sizeof(double) +
                        32 * sizeof(int);
    AC_CHECK_CUDA(cudaFuncSetAttribute(pyin_viterbi_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AC_CHECK_CUDA(cudaFuncSetAttribute(pyin_viterbi_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(KC_MISC, 0.0, (double)nf * 2 * g.nb * 2, st);
    const char* mode = getenv("AC_PYIN_VITERBI");  // test hook: "generic" / "tiled" force the older kernels
    const bool want_generic = mode && mode[0] == 'g', want_tiled = mode && mode[0] == 't';
    // default: the 4-CTA cluster kernel; "fast" = single-CTA one-barrier kernel, "tiled" / "generic" = the older ones
    const bool want_fast = mode && mode[0] == 'f';
    if (g.half == 20 && !want_generic && !want_tiled && !want_fast && (g.nb + kVcN - 1) / kVcN <= kVcComp &&
        (g.nb + kVcN - 1) / kVcN > 20) {
      const int Q = (g.nb + kVcN - 1) / kVcN;
      const size_t csmem = (size_t)(4 * (Q + 2 * 20 + 2) + 3 * Q + Q + 2 * kVcN * (kVcComp / 32) + 2) * sizeof(double) +
                           (size_t)2 * kVcN * (kVcComp / 32) * sizeof(int);
      pyin_viterbi_cluster_kernel<20><<<kVcN, kVcThreads, csmem, st>>>(vp);
    } else if (g.half == 20 && !want_generic && !want_tiled && (g.nb + 1) / 2 <= kVitFastComp) {
      const size_t fsmem = (size_t)(4 * vit_row(g.nb) + 3 * g.nb + g.nb + 32 + 2) * sizeof(double) + 32 * sizeof(int);
      AC_CHECK_CUDA(cudaFuncSetAttribute(pyin_viterbi_fast_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      pyin_viterbi_fast_kernel<20><<<1, kVitFastThreads, fsmem, st>>>(vp);
    } else if (g.half == 20 && !want_generic) {
      pyin_viterbi_kernel<20><<<1, kVitThreads, smem, st>>>(vp);
    } else {
      pyin_viterbi_kernel<0><<<1, kVitThreads, smem, st>>>(vp);
    }
    AC_LAUNCH_CHECK();
    const size_t bt_smem = (size_t)kBtSteps * 2 * g.nb * sizeof(unsigned short);
    AC_CHECK_CUDA(cudaFuncSetAttribute(pyin_backtrack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bt_smem));
    const char* bt_mode = getenv("AC_PYIN_BACKTRACK");  // test hook: "serial" runs the single-CTA kernel
    if (bt_mode && bt_mode[0] == 's') {
      pyin_backtrack_kernel<<<1, 256, bt_smem, st>>>(ptr, final_state, nf, g.nb, fmin, g.bps, d_f0, d_voiced_flag);
    } else {
      AC_CHECK_CUDA(cudaFuncSetAttribute(pyin_bp_compose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bt_smem));
      AC_CHECK_CUDA(cudaFuncSetAttribute(pyin_bp_tops_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bt_smem));
      pyin_bp_compose_kernel<<<(unsigned)n_blk, kBpThreads, bt_smem, st>>>(ptr, nf, g.nb, bp_maps);
      count_launch();
      pyin_bp_tops_kernel<<<1, 256, bt_smem, st>>>(bp_maps, final_state, n_blk, g.nb, bp_top);
      count_launch();
      pyin_bp_emit_kernel<<<(unsigned)n_blk, kBpSteps, 0, st>>>(ptr, bp_top, nf, g.nb, fmin, g.bps, d_f0, d_voiced_flag);
    }
    AC_LAUNCH_CHECK();
  }
  return AC_OK;
}

// ---- LPC formants ----------------------------------------------------------------------------------
namespace ac {
constexpr int kLpcThreads = 128;
constexpr int kLpcMaxFrame = 2048;
constexpr int kLpcMaxOrder = 16;
constexpr int kLpcFreq = 512;

__device__ __forceinline__ double block_sum_d(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// one CTA per frame: pre-emphasis, Burg recursion in fp64 (librosa.lpc), |1/A(e^jw)| on 512 points of
// [0, pi), peaks (strict local maxima, flat tops at their midpoint as scipy.signal.find_peaks) above
// 0.1 * max, magnitudes of the first three
__global__ void __launch_bounds__(kLpcThreads) lpc_formant_kernel(const float* __restrict__ x, long long n, int frame, int hop,
                                                                  int order, long long n_frames, float* __restrict__ mags,
                                                                  int* __restrict__ counts) {
  __shared__ double fwd[kLpcMaxFrame], bwd[kLpcMaxFrame];
  __shared__ double ar[kLpcMaxOrder + 1], ar_prev[kLpcMaxOrder + 1];
  __shared__ double red[8];
  __shared__ double mag[kLpcFreq];  // fp64 like scipy.signal.freqz: the peak rule compares neighbours, float32 storage created ties
  __shared__ int peak_idx[4];
  __shared__ int n_peaks_s;
  const long long fr = blockIdx.x;
  const int tid = threadIdx.x;
  const long long s0 = fr * hop;
  // pre-emphasised frame p[0] = x[0], p[i] = x[i] - 0.95 x[i-1]  (float32 arithmetic as np.append on float32)
  bool any = false;
  for (int i = tid; i < frame; i += kLpcThreads) {
    const float cur = x[s0 + i];
    const float pe = i == 0 ? cur : cur - 0.95f * x[s0 + i - 1];
    any = any || (pe != 0.f);
    if (i >= 1) fwd[i - 1] = (double)pe;        // fwd = p[1:]
    if (i < frame - 1) bwd[i] = (double)pe;     // bwd = p[:-1]
  }
  const int any_all = __syncthreads_or(any ? 1 : 0);
  if (tid <= order) { ar[tid] = tid == 0 ? 1.0 : 0.0; ar_prev[tid] = ar[tid]; }
  __syncthreads();
  bool ok = any_all != 0;
  int len = frame - 1;
  double den = 0.0;
  {
    double part = 0.0;
    for (int i = tid; i < len; i += kLpcThreads) part += fwd[i] * fwd[i] + bwd[i] * bwd[i];
    den = block_sum_d(part, red);
  }
  double* a_cur = ar;
  double* a_old = ar_prev;
  for (int it = 0; it < order && ok; ++it) {
    double part = 0.0;
    for (int i = tid; i < len; i += kLpcThreads) part += bwd[i] * fwd[i];
    const double dot = block_sum_d(part, red);
    const double rc = -2.0 * dot / (den + 2.2250738585072014e-308);
    // ar_prev, ar = ar, ar_prev ; ar[j] = ar_prev[j] + rc * ar_prev[i - j + 1]
    double* t = a_old; a_old = a_cur; a_cur = t;
    __syncthreads();
    if (tid >= 1 && tid <= it + 1) a_cur[tid] = a_old[tid] + rc * a_old[it - tid + 1];
    if (tid == 0) a_cur[0] = 1.0;
    // fwd' = fwd + rc*bwd ; bwd' = bwd + rc*fwd   (both from the old values), then drop fwd'[0] and bwd'[-1]
    const double b_last_new = bwd[len - 1] + rc * fwd[len - 1];
    const double f_first_new = fwd[0] + rc * bwd[0];
    __syncthreads();
    // new fwd[i] = old fwd[i+1] + rc*old bwd[i+1] (shifted by one), new bwd[i] = old bwd[i] + rc*old fwd[i]
    double nf_[ (kLpcMaxFrame + kLpcThreads - 1) / kLpcThreads ], nb_[ (kLpcMaxFrame + kLpcThreads - 1) / kLpcThreads ];
    int c = 0;
    for (int i = tid; i < len - 1; i += kLpcThreads, ++c) {
      nf_[c] = fwd[i + 1] + rc * bwd[i + 1];
      nb_[c] = bwd[i] + rc * fwd[i];
    }
    __syncthreads();
    c = 0;
    for (int i = tid; i < len - 1; i += kLpcThreads, ++c) {
      fwd[i] = nf_[c];
      bwd[i] = nb_[c];
    }
    const double q = 1.0 - rc * rc;
    den = q * den - b_last_new * b_last_new - f_first_new * f_first_new;
    len -= 1;
    if (!isfinite(rc)) ok = false;
    __syncthreads();
  }
  // |1 / A(e^{jw})|, w_k = pi * k / 512
  double local_max = 0.0;
  for (int k = tid; k < kLpcFreq; k += kLpcThreads) {
    const double w = 3.14159265358979323846 * (double)k / (double)kLpcFreq;
    double re = 0.0, im = 0.0;
    for (int m = 0; m <= order; ++m) {
      double s, c;
      sincos(w * m, &s, &c);
      re += a_cur[m] * c;
      im -= a_cur[m] * s;
    }
    const double v = 1.0 / sqrt(re * re + im * im);
    mag[k] = v;
    local_max = fmax(local_max, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local_max = fmax(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  __shared__ double mx_s[8];
  if ((tid & 31) == 0) mx_s[tid >> 5] = local_max;
  __syncthreads();
  double mx = 0.0;
  for (int i = 0; i < kLpcThreads / 32; ++i) mx = fmax(mx, mx_s[i]);
  if (tid == 0) {
    int np = 0;
    if (ok && isfinite(mx)) {
      const double height = mx * 0.1;
      int i = 1;
      while (i < kLpcFreq - 1 && np < 3) {
        if (mag[i - 1] < mag[i]) {
          int ahead = i + 1;
          while (ahead < kLpcFreq - 1 && mag[ahead] == mag[i]) ++ahead;  // flat top
          if (mag[ahead] < mag[i]) {
            const int mid = (i + ahead - 1) / 2;
            if (mag[mid] >= height) peak_idx[np++] = mid;
            i = ahead;
            continue;
          }
        }
        ++i;
      }
    }
    n_peaks_s = np;
  }
  __syncthreads();
  if (tid < 3) mags[fr * 3 + tid] = tid < n_peaks_s ? (float)mag[peak_idx[tid]] : 0.f;
  if (tid == 0) counts[fr] = n_peaks_s;
}
}  // namespace ac

extern "C" long long ac_lpc_frame_count(long long n, int frame, int hop) {
  if (n - frame <= 0 || hop <= 0) return 0;
  return (n - frame + hop - 1) / hop;  // len(range(0, n - frame, hop))
}

extern "C" int ac_lpc_formants(const float* d_x, long long n, int frame, int hop, int order, float* d_mags, int* d_counts,
                               void* stream) {
  using namespace ac;
  AC_REQUIRE(d_x && d_mags && d_counts, "ac_lpc_formants: null pointer");
  AC_REQUIRE(frame >= order + 2 && frame <= kLpcMaxFrame && order >= 1 && order <= kLpcMaxOrder && hop > 0, "ac_lpc_formants: geometry");
  const long long nf = ac_lpc_frame_count(n, frame, hop);
  if (nf == 0) return AC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(KC_MISC, 8.0 * nf * order * frame, 4.0 * n, st);
  lpc_formant_kernel<<<(unsigned)nf, kLpcThreads, 0, st>>>(d_x, n, frame, hop, order, nf, d_mags, d_counts);
  AC_LAUNCH_CHECK();
  return AC_OK;
}
