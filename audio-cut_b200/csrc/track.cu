// Whole-track chunked separation: the device side of EnhancedVocalSeparator._separate_with_pipeline
// (enhanced_vocal_separator.py:366-458) + MDX23OnnxBackend.infer_chunk (backends.py:299-406).
//
// The track is uploaded once; every model window of every chunk is described by a WinDesc that
// the STFT kernel reads straight out of the track buffer (window build, align_hop / gen / trim
// zero padding are index arithmetic, not copies) and that the fused iSTFT kernel uses to map its
// output back to track samples (trim, concat, crop, halo trim).  Windows are independent, so they
// are simply batched through STFT -> U-Net -> iSTFT max_batch at a time.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "stft_mdx.cuh"
#include "unet_kernels.cuh"

namespace ac {

// hop*(dim_t-1) must exceed n_fft (gen > 0), everything positive: otherwise the window arithmetic divides by zero
static bool track_geom_ok(const ac_track_params& p) {
  return p.mdx.hop > 0 && p.mdx.dim_t > 1 && p.mdx.n_fft > 0 && p.mdx.dim_f > 0 && p.align_hop >= 0 &&
         (long long)p.mdx.hop * (p.mdx.dim_t - 1) > (long long)p.mdx.n_fft;
}

static int chunk_windows(int chunk_len, const ac_track_params& p) {
  const int hop = p.align_hop > 0 ? p.align_hop : 1;
  const long long L = (long long)chunk_len + ((hop - chunk_len % hop) % hop);
  const int W = p.mdx.hop * (p.mdx.dim_t - 1);
  const int gen = W - p.mdx.n_fft;
  const long long pad = (gen - L % gen) % gen;
  return (int)((L + pad) / gen);
}

static void build_windows(const ac_chunk_desc* ch, int n_chunks, const ac_track_params& p, std::vector<WinDesc>& out) {
  const int W = p.mdx.hop * (p.mdx.dim_t - 1);
  const int trim = p.mdx.n_fft / 2;
  const int gen = W - p.mdx.n_fft;
  out.clear();
  long long side = 0;  // chunks lie back to back in the per-chunk side buffer
  for (int c = 0; c < n_chunks; side += ch[c].chunk_len > 0 ? ch[c].chunk_len : 0, ++c) {
    if (ch[c].chunk_len <= 0) continue;
    const int nw = chunk_windows(ch[c].chunk_len, p);
    for (int w = 0; w < nw; ++w) {
      const long long q0 = (long long)w * gen - trim;  // chunk-local index of window position 0
      WinDesc d;
      d.base = ch[c].chunk_start + q0;
      long long lo = -q0, hi = (long long)ch[c].chunk_len - q0;
      if (lo < 0) lo = 0;
      if (hi > W) hi = W;
      if (hi < lo) hi = lo;
      d.p_lo = (int)lo;
      d.p_hi = (int)hi;
      d.out_base = ch[c].chunk_start + (long long)w * gen;
      long long ol = (long long)ch[c].chunk_len - (long long)w * gen;
      if (ol > gen) ol = gen;
      if (ol <= 0) continue;  // window made only of alignment padding: its output is cropped away
      d.out_len = (int)ol;
      d.eff_start = ch[c].eff_start;
      d.eff_end = ch[c].eff_end;
      d.pad_ = 0;
      d.side_base = side + (long long)w * gen;
      out.push_back(d);
    }
  }
}

// divides samples [lo, lo + n) of both stems by their window count
__global__ void finalize_stems_kernel(float* __restrict__ vocal, float* __restrict__ instr,
                                      const float* __restrict__ weight, long long lo, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  i += lo;
  float w = weight[i];
  w = w == 0.f ? 1.f : w;
  vocal[i] = vocal[i] / w;
  instr[i] = instr[i] / w;
}

// windows per STFT -> U-Net -> iSTFT batch.  A single short forward is ~3 % faster per window at 8 than at 16 (909 vs 883 TFLOP/s),
// but inside the sustained, power-capped step 16 wins (60.5-61.3 vs 63.4-63.6 ms per 4-min track: fewer, longer launches and
// 24 % instead of 33 % strip warm-up in the fused iSTFT).  Dev hook: AC_BATCH.
static int default_batch(int dtype) {
  static const int env = getenv("AC_BATCH") ? atoi(getenv("AC_BATCH")) : 0;
  if (env > 0) return env;
  return dtype == AC_F32 ? 8 : 16;
}

// Pinned staging for the window descriptors, so that ac_separate_track never blocks the host:
// a ring of slots, each guarded by the event recorded after its H2D copy was enqueued.
struct WinStage {
  static constexpr int kSlots = 4;
  WinDesc* h[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  size_t cap[kSlots] = {0, 0, 0, 0};
  cudaEvent_t ev[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  int next = 0;
  // returns a pinned buffer of >= n descriptors whose previous use has completed
  WinDesc* acquire(size_t n, int& slot) {
    slot = next;
    next = (next + 1) % kSlots;
    if (ev[slot]) cudaEventSynchronize(ev[slot]);
    else if (cudaEventCreateWithFlags(&ev[slot], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cap[slot] < n) {
      if (h[slot]) cudaFreeHost(h[slot]);
      h[slot] = nullptr;
      cap[slot] = 0;
      const size_t want = n < 256 ? 256 : n * 2;
      if (cudaMallocHost(reinterpret_cast<void**>(&h[slot]), want * sizeof(WinDesc)) != cudaSuccess) return nullptr;
      cap[slot] = want;
    }
    return h[slot];
  }
};
static thread_local WinStage g_win_stage;

__global__ void downmix_kernel(const float* __restrict__ mix, int n_ch, long long n, float* __restrict__ out) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n && n_ch == 2) {
    const float4 a = ldg_stream_f4(reinterpret_cast<const float4*>(mix + i));
    float4 b;
    if ((n & 3) == 0) b = ldg_stream_f4(reinterpret_cast<const float4*>(mix + n + i));
    else b = make_float4(mix[n + i], mix[n + i + 1], mix[n + i + 2], mix[n + i + 3]);
    // numpy/torch mean over 2 channels: (a + b) / 2 in float32
    *reinterpret_cast<float4*>(out + i) = make_float4((a.x + b.x) * 0.5f, (a.y + b.y) * 0.5f, (a.z + b.z) * 0.5f, (a.w + b.w) * 0.5f);
    return;
  }
  for (long long j = i; j < n && j < i + 4; ++j) out[j] = n_ch == 2 ? (mix[j] + mix[n + j]) * 0.5f : mix[j];
}

// sum of squares of up to three arrays + count of non-zero samples of the second one (fp64 accumulators)
__global__ void track_stats_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                   long long n, double* __restrict__ out, int* __restrict__ abort_flag) {
  if (abort_flag && blockIdx.x == 0 && threadIdx.x == 0) {  // hand the tcgen05 watchdog flag to the host and re-arm it
    const int f = *abort_flag;
    if (f) { out[4] = (double)f; *abort_flag = 0; }
  }
  double sa = 0, sb = 0, sc = 0, nz = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = a ? a[i] : 0.f, y = b ? b[i] : 0.f, z = c ? c[i] : 0.f;
    sa += (double)x * x;
    sb += (double)y * y;
    sc += (double)z * z;
    nz += y != 0.f ? 1.0 : 0.0;
  }
  __shared__ double sh[4][8];
  double v[4] = {sa, sb, sc, nz};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[threadIdx.x][w];
    atomicAdd(out + threadIdx.x, t);
  }
}

}  // namespace ac

extern "C" int ac_track_window_count(const ac_chunk_desc* h_chunks, int n_chunks, const ac_track_params* p) {
  if (!h_chunks || !p || n_chunks < 0 || !ac::track_geom_ok(*p)) return -1;
  std::vector<ac::WinDesc> w;
  ac::build_windows(h_chunks, n_chunks, *p, w);
  return (int)w.size();
}

extern "C" size_t ac_track_workspace_bytes(const ac_unet* net, const ac_chunk_desc* h_chunks, int n_chunks,
                                           const ac_track_params* p) {
  if (!net || !h_chunks || !p) return 0;
  const int nw = ac_track_window_count(h_chunks, n_chunks, p);
  if (nw < 0) return 0;
  int mb = p->max_batch > 0 ? p->max_batch : ac::default_batch(p->dtype);
  if (mb > nw) mb = nw > 0 ? nw : 1;
  const size_t es = p->dtype == AC_F32 ? 4 : 2;
  const size_t spec = (size_t)mb * p->mdx.dim_t * p->mdx.dim_f * 4 * es;
  return ac::align_up(sizeof(ac::WinDesc) * (size_t)(nw > 0 ? nw : 1), 256) + ac::align_up(spec, 256) +
         ac_unet_workspace_bytes(net, mb, p->dtype) + 1024;
}

extern "C" int ac_separate_track(ac_unet* net, const float* d_mix, long long n_samples, const ac_chunk_desc* h_chunks,
                                 int n_chunks, const ac_track_params* p, float* d_vocal, float* d_instr,
                                 float* d_weight, void* d_ws, size_t ws_bytes, void* stream) {
  return ac_separate_track_ex(net, d_mix, n_samples, h_chunks, n_chunks, p, d_vocal, d_instr, d_weight, nullptr, d_ws, ws_bytes,
                              stream);
}

namespace ac {
// events of the copy pipeline (one device per process; re-recording an event a stream already waited on is safe: a wait
// refers to the record that preceded it)
struct CopyEvents {
  std::vector<cudaEvent_t> ev;
  cudaEvent_t get(size_t i) {
    while (ev.size() <= i) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      ev.push_back(e);
    }
    return ev[i];
  }
};
static thread_local CopyEvents g_copy_events;

// h_mix / h_vocal / h_instr (optional, page-locked): the host<->device copies of the track are pipelined against the window
// batches on `cs` - the samples batch k reads are uploaded before it (later pieces while earlier batches compute), and every
// stretch of the stems that no later window can touch is finalised and downloaded while the next batch runs.
static int separate_track_impl(ac_unet* net, float* d_mix, const float* h_mix, long long n_samples, const ac_chunk_desc* h_chunks,
                               int n_chunks, const ac_track_params* p, float* d_vocal, float* d_instr, float* d_weight,
                               float* d_chunk_vocal, float* h_vocal, float* h_instr, void* d_ws, size_t ws_bytes, cudaStream_t st,
                               cudaStream_t cs, cudaEvent_t uploaded) {
  AC_REQUIRE(net && d_mix && h_chunks && p && d_vocal && d_instr && d_weight && d_ws, "null pointer");
  AC_REQUIRE(track_geom_ok(*p), "MDX geometry: hop*(dim_t-1) must exceed n_fft, all sizes positive");
  AC_REQUIRE(n_samples > 0 && n_chunks >= 0, "bad sizes");
  AC_REQUIRE(p->n_channels == 1 || p->n_channels == 2, "n_channels must be 1 or 2");
  AC_REQUIRE(p->dtype == AC_F32 || p->dtype == AC_BF16 || p->dtype == AC_F16, "dtype");
  AC_REQUIRE((h_vocal == nullptr) == (h_instr == nullptr), "host stems: both or none");
  for (int c = 0; c < n_chunks; ++c) {
    AC_REQUIRE(h_chunks[c].chunk_start >= 0 && h_chunks[c].chunk_start + h_chunks[c].chunk_len <= n_samples,
               "chunk outside the track");
    AC_REQUIRE(h_chunks[c].eff_start >= 0 && h_chunks[c].eff_end <= n_samples, "effective region outside the track");
  }
  if (ws_bytes < ac_track_workspace_bytes(net, h_chunks, n_chunks, p)) {
    set_error("track workspace too small");
    return AC_E_WORKSPACE;
  }
  const MdxPlan* plan = get_mdx_plan(p->mdx);
  if (!plan) return AC_E_INVALID;
  std::vector<WinDesc> wins;
  build_windows(h_chunks, n_chunks, *p, wins);
  const int nw = (int)wins.size();
  int mb = p->max_batch > 0 ? p->max_batch : default_batch(p->dtype);
  if (mb > nw) mb = nw > 0 ? nw : 1;
  const size_t es = p->dtype == AC_F32 ? 4 : 2;
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(d_ws) + 255) & ~uintptr_t(255));
  WinDesc* d_wins = reinterpret_cast<WinDesc*>(base);
  char* d_spec = base + align_up(sizeof(WinDesc) * (size_t)(nw > 0 ? nw : 1), 256);
  const size_t spec_bytes = align_up((size_t)mb * p->mdx.dim_t * p->mdx.dim_f * 4 * es, 256);
  char* d_uws = d_spec + spec_bytes;
  const size_t uws_bytes = ws_bytes - (size_t)(d_uws - reinterpret_cast<char*>(d_ws));
  const int n_batches = nw > 0 ? (nw + mb - 1) / mb : 0;

  // ---- upload plan: batch k reads track samples below up_end[k] (window position p of window w = sample base + p)
  size_t n_ev = 0;
  std::vector<cudaEvent_t> up_ev((size_t)n_batches, nullptr);
  if (h_mix) {
    long long done = 0;
    auto upload = [&](long long to) -> int {
      if (to > n_samples) to = n_samples;
      if (to <= done) return AC_OK;
      for (int c = 0; c < p->n_channels; ++c)
        AC_CHECK_CUDA(cudaMemcpyAsync(d_mix + (size_t)c * n_samples + done, h_mix + (size_t)c * n_samples + done,
                                      sizeof(float) * (size_t)(to - done), cudaMemcpyHostToDevice, cs));
      done = to;
      return AC_OK;
    };
    for (int k = 0; k < n_batches; ++k) {
      long long need = 0;
      for (int w = k * mb; w < nw && w < (k + 1) * mb; ++w) need = std::max(need, wins[w].base + (long long)wins[w].p_hi);
      int rc = upload(need);
      if (rc) return rc;
      up_ev[k] = g_copy_events.get(n_ev++);
      AC_REQUIRE(up_ev[k] != nullptr, "event creation failed");
      AC_CHECK_CUDA(cudaEventRecord(up_ev[k], cs));
    }
    int rc = upload(n_samples);  // whatever no window reads (the feature kernels still do)
    if (rc) return rc;
    if (uploaded) AC_CHECK_CUDA(cudaEventRecord(uploaded, cs));
  }
  // ---- download plan: after windows [0, w) no later window writes below fin_lo[w] = min eff_start of windows >= w
  std::vector<long long> fin_lo((size_t)nw + 1, n_samples);
  for (int w = nw - 1; w >= 0; --w) fin_lo[w] = std::min(fin_lo[w + 1], (long long)wins[w].eff_start);

  AC_CHECK_CUDA(cudaMemsetAsync(d_vocal, 0, sizeof(float) * n_samples, st));
  AC_CHECK_CUDA(cudaMemsetAsync(d_instr, 0, sizeof(float) * n_samples, st));
  AC_CHECK_CUDA(cudaMemsetAsync(d_weight, 0, sizeof(float) * n_samples, st));
  if (nw > 0) {
    // stage the descriptors in pinned memory: the copy is asynchronous and the host never waits here
    int slot = 0;
    WinDesc* h_pin = g_win_stage.acquire((size_t)nw, slot);
    if (!h_pin) {
      set_error("pinned staging for window descriptors failed");
      return AC_E_CUDA;
    }
    std::copy(wins.begin(), wins.end(), h_pin);
    AC_CHECK_CUDA(cudaMemcpyAsync(d_wins, h_pin, sizeof(WinDesc) * nw, cudaMemcpyHostToDevice, st));
    AC_CHECK_CUDA(cudaEventRecord(g_win_stage.ev[slot], st));
  }
  long long fin_done = 0;
  auto finish_to = [&](long long to) -> int {  // finalise [fin_done, to) and send it home
    if (to > n_samples) to = n_samples;
    if (to <= fin_done) return AC_OK;
    const long long n = to - fin_done;
    finalize_stems_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_vocal, d_instr, d_weight, fin_done, n);
    AC_LAUNCH_CHECK();
    if (h_vocal) {
      cudaEvent_t e = g_copy_events.get(n_ev++);
      AC_REQUIRE(e != nullptr, "event creation failed");
      AC_CHECK_CUDA(cudaEventRecord(e, st));
      AC_CHECK_CUDA(cudaStreamWaitEvent(cs, e, 0));
      AC_CHECK_CUDA(cudaMemcpyAsync(h_vocal + fin_done, d_vocal + fin_done, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, cs));
      AC_CHECK_CUDA(cudaMemcpyAsync(h_instr + fin_done, d_instr + fin_done, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, cs));
    }
    fin_done = to;
    return AC_OK;
  };
  int k = 0;
  for (int w0 = 0; w0 < nw; w0 += mb, ++k) {
    const int b = nw - w0 < mb ? nw - w0 : mb;
    if (h_mix) AC_CHECK_CUDA(cudaStreamWaitEvent(st, up_ev[k], 0));
    // STFT -> U-Net: where the network and the frame size allow it the first 1x1 conv runs in the STFT's epilogue (the [f][4]
    // spectrogram never goes to HBM, one launch and a 1.3 GB round trip less per 16 windows); else two launches.
    int rc;
    void* d_first = nullptr;
    const float *fw = nullptr, *fsc = nullptr, *fsh = nullptr;
    if (stft_first_conv_supported(plan, unet_base_channels(net), p->dtype) &&
        unet_first_conv_target(net, b, p->dtype, d_uws, uws_bytes, &d_first, &fw, &fsc, &fsh) == AC_OK) {
      rc = launch_stft_first_conv(plan, d_mix, n_samples, p->n_channels, d_wins + w0, b, d_first, unet_base_channels(net), fw,
                                  fsc, fsh, p->dtype, st);
      if (rc) return rc;
      rc = unet_forward_after_first(net, d_spec, b, p->dtype, d_uws, uws_bytes, st);
    } else {
      rc = launch_stft(plan, d_mix, n_samples, p->n_channels, d_wins + w0, b, d_spec, p->dtype, st);
      if (rc) return rc;
      rc = ac_unet_forward(net, d_spec, d_spec, b, p->dtype, d_uws, uws_bytes, st);
    }
    if (rc) return rc;
    rc = launch_istft(plan, d_spec, p->dtype, d_wins + w0, b, 1, nullptr, d_mix, n_samples, p->n_channels,
                      p->output_is_vocal, d_vocal, d_instr, d_weight, st, d_chunk_vocal);
    if (rc) return rc;
    if (h_vocal && (rc = finish_to(fin_lo[w0 + b]))) return rc;  // without a host target one finalise at the end is cheaper
  }
  if (h_mix && n_batches == 0 && uploaded) AC_CHECK_CUDA(cudaStreamWaitEvent(st, uploaded, 0));
  return finish_to(n_samples);
}
}  // namespace ac

extern "C" int ac_separate_track_ex(ac_unet* net, const float* d_mix, long long n_samples, const ac_chunk_desc* h_chunks,
                                    int n_chunks, const ac_track_params* p, float* d_vocal, float* d_instr,
                                    float* d_weight, float* d_chunk_vocal, void* d_ws, size_t ws_bytes, void* stream) {
  return ac::separate_track_impl(net, const_cast<float*>(d_mix), nullptr, n_samples, h_chunks, n_chunks, p, d_vocal, d_instr, d_weight,
                                 d_chunk_vocal, nullptr, nullptr, d_ws, ws_bytes, (cudaStream_t)stream, (cudaStream_t)stream, nullptr);
}

extern "C" int ac_separate_track_pipelined(ac_unet* net, float* d_mix, const float* h_mix, long long n_samples,
                                           const ac_chunk_desc* h_chunks, int n_chunks, const ac_track_params* p, float* d_vocal,
                                           float* d_instr, float* d_weight, float* d_chunk_vocal, float* h_vocal, float* h_instr,
                                           void* d_ws, size_t ws_bytes, void* stream, void* copy_stream, void* uploaded_event) {
  AC_REQUIRE(copy_stream != stream || (!h_mix && !h_vocal), "the copy stream must differ from the compute stream");
  return ac::separate_track_impl(net, d_mix, h_mix, n_samples, h_chunks, n_chunks, p, d_vocal, d_instr, d_weight, d_chunk_vocal,
                                 h_vocal, h_instr, d_ws, ws_bytes, (cudaStream_t)stream, (cudaStream_t)copy_stream,
                                 (cudaEvent_t)uploaded_event);
}

extern "C" int ac_downmix_mono(const float* d_mix, int n_channels, long long n, float* d_out, void* stream) {
  using namespace ac;
  AC_REQUIRE(d_mix && d_out && n >= 0, "null pointer");
  AC_REQUIRE(n_channels == 1 || n_channels == 2, "n_channels must be 1 or 2");
  if (n == 0) return AC_OK;
  const long long threads = (n + 3) / 4;
  downmix_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_mix, n_channels, n, d_out);
  AC_LAUNCH_CHECK();
  return AC_OK;
}

namespace ac { int* tc_abort_flag_if_any(); }
extern "C" int ac_track_stats(const float* d_a, const float* d_b, const float* d_c, long long n, double* d_out5, void* stream) {
  using namespace ac;
  AC_REQUIRE(d_out5 && n >= 0, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  AC_CHECK_CUDA(cudaMemsetAsync(d_out5, 0, 5 * sizeof(double), st));
  double* d_out4 = d_out5;
  if (n == 0) n = 1, d_a = d_b = d_c = nullptr;  // still report (and re-arm) the watchdog flag
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  track_stats_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_a, d_b, d_c, n, d_out4, tc_abort_flag_if_any());
  AC_LAUNCH_CHECK();
  return AC_OK;
}
