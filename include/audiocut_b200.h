/*
 * audiocut_b200.h - C ABI of libaudiocut_b200.so, the B200 (sm_100a) implementation of the
 * separation-and-feature hot path of BDMstudio/audio-cut.
 *
 * The reference is pure Python and has no FFI of its own for this path; the
 * entry points below are what a ctypes binding inside the reference's files would bind.  Each
 * one names the reference code it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer owned by the caller (torch tensors in the
 *     Python host); h_* is a HOST pointer; nothing is allocated inside a hot-path call, the
 *     caller passes a workspace sized by the matching *_workspace_bytes().
 *   - stream is a cudaStream_t passed as void* (0 = legacy default stream); calls are
 *     asynchronous with respect to the host unless stated otherwise.
 *   - return value: 0 = ok, negative = error (AC_E_*); ac_last_error() gives the text
 *     (thread-local).  There is NO CPU fallback: without a usable sm_100 device every call fails.
 *   - spectrogram layout ("TFC"): [window][t = 0..dim_t)[f = 0..dim_f)[4] with the 4 innermost
 *     values {L_re, L_im, R_re, R_im}.  It is a permutation of the ONNX tensor
 *     [B,4,dim_f,dim_t] (backends.py:355-358): onnx[b][c][f][t] == tfc[b][t][f][c].
 */
#ifndef AUDIOCUT_B200_H
#define AUDIOCUT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AC_API __attribute__((visibility("default")))
#else
#define AC_API
#endif

#define AC_OK 0
#define AC_E_INVALID (-1)   /* bad argument */
#define AC_E_CUDA (-2)      /* CUDA runtime/driver error */
#define AC_E_NODEVICE (-3)  /* no sm_100 device */
#define AC_E_WORKSPACE (-4) /* workspace too small */

#define AC_F32 0  /* fp32 storage, fp32 FFMA math (the >=60 dB parity path) */
#define AC_BF16 1 /* bfloat16 storage, tcgen05 tensor-core math with fp32 accumulation (fp32 range, 8-bit significand) */
#define AC_F16 2  /* IEEE half storage, the same tcgen05 kernels and rate (11-bit significand: the 16-bit path that
                     meets the >= 40 dB stem-SDR gate; values beyond +-65504 saturate) */

/* Library / device bring-up.  Replaces torch.cuda device selection in gpu_pipeline.py:87-130. */
AC_API int ac_init(int device);
AC_API const char* ac_last_error(void);
AC_API int ac_abi_version(void);
/* Number of kernels this library has launched since load (for bench.py "gpu_launches"). */
AC_API long long ac_launch_count(void);
/* Host staging helpers for the Python drop-in (enhanced_vocal_separator.py:378-389 stages every chunk through a
 * pinned pool and copies it straight back).  ac_host_is_pinned: 1 when [p, p+bytes) is page-locked memory CUDA
 * knows about (cudaHostAlloc / cudaHostRegister / torch pin_memory), 0 otherwise.  ac_copy_h2d_async: a
 * cudaMemcpyAsync host -> device on `stream`; truly asynchronous only for page-locked sources. */
AC_API int ac_host_is_pinned(const void* h_ptr, size_t bytes);
AC_API int ac_copy_h2d_async(void* d_dst, const void* h_src, size_t bytes, void* stream);

/* Per-kernel-class device timing with CUDA events on the launching stream (for bench.py's
 * roofline line): begin, run the workload, collect (synchronises the device).  `flops` / `bytes`
 * are the ALGORITHMIC counts of the launches, not hardware counters. */
typedef struct ac_kernel_stat {
  const char* name;
  long long launches;
  double total_ms;
  double flops;
  double bytes;
} ac_kernel_stat;
AC_API int ac_profile_begin(void);
AC_API int ac_profile_collect(ac_kernel_stat* out, int max_classes); /* returns the number of classes filled */

/* ---- framewise RMS ------------------------------------------------------------------------
 * librosa.feature.rms(y, frame_length, hop_length, center=True, pad_mode="constant") as called at
 * features_cache.py:182, pure_vocal_pause_detector.py:1111-1113 and :1397, seamless_splitter.py:1714
 * and :1848, vocal_separator.py:483.  d_out has ac_frame_count(n, frame, hop, center) floats. */
AC_API long long ac_frame_count(long long n, int frame, int hop, int center);
AC_API int ac_frame_rms(const float* d_x, long long n, int frame, int hop, int center, float* d_out, void* stream);
/* The same for MANY slices of one signal in one launch (every pipeline chunk of a track: ChunkFeatureBuilder.add_chunk
 * calls librosa.feature.rms once per chunk, features_cache.py:182).  Slice i = d_x[start, start + len) is framed and
 * padded as if it were the whole signal; its ac_frame_count(len, ...) values go to d_out + frame_off.
 * h_segs is a HOST array of ac_feat_segment (declared below). */
struct ac_feat_segment;
AC_API int ac_frame_rms_segments(const float* d_x, long long n, const struct ac_feat_segment* h_segs, int n_segs, int frame,
                          int hop, int center, float* d_out, void* stream);

/* ---- MDX23 STFT / iSTFT ---------------------------------------------------------------------
 * Conv_TDF_net_trim_model.stft / .istft of the external MVSEP-MDX23 inference.py, called at
 * backends.py:355 and :376 (torch.stft / torch.istft, hann periodic, center, reflect pad). */
typedef struct ac_mdx_geom {
  int n_fft; /* 6144 (backends.py:264) or 7680 (Kim_Vocal geometry); any 2^a 3^b 5^c multiple of 2*hop.. */
  int hop;   /* 1024 */
  int dim_f; /* 3072 bins kept (<= n_fft/2) */
  int dim_t; /* 256 frames; window length W = hop*(dim_t-1) */
} ac_mdx_geom;

/* d_wave [B][2][W] f32 -> d_spec [B][dim_t][dim_f][4] (dtype AC_F32, AC_F16 or AC_BF16). */
AC_API int ac_stft_mdx(const float* d_wave, void* d_spec, int B, const ac_mdx_geom* g, int dtype, void* stream);
/* d_spec [B][dim_t][dim_f][4] -> d_wave [B][2][W] f32 (full torch.istft output, no trim). */
AC_API int ac_istft_mdx(const void* d_spec, float* d_wave, int B, const ac_mdx_geom* g, int dtype, void* stream);

/* ---- TFC-TDF U-Net ---------------------------------------------------------------------------
 * Replaces self._session.run(None, {input: stft}) at backends.py:358 (onnxruntime).  */
typedef struct ac_unet_geom {
  int dim_f, dim_t; /* 3072, 256 */
  int dim_c;        /* 4 */
  int g;            /* growth, 48 */
  int n;            /* down/up levels, 5 (L = 11) */
  int l;            /* convs per TFC, 3 */
  int bn;           /* TDF bottleneck factor, 8 */
} ac_unet_geom;
typedef struct ac_unet ac_unet;

/* h_blob: all parameters as float32 with BatchNorm folded to (scale, shift), in execution order:
 *   first_conv:  W[g][dim_c], scale[g], shift[g]
 *   for i in 0..n-1:  block(c_i, f_i);  ds_i: W[c+g][c][2][2], scale[c+g], shift[c+g]
 *   block(c_n, f_n)                                                   (bottleneck)
 *   for i in 0..n-1:  us_i: W[c_in][c_out][2][2], scale[c_out], shift[c_out];  block(c_out, f)
 *   final_conv:  W[dim_c][g], bias[dim_c]
 * block(c, f) = 3 x { W[c][c][3][3], scale[c], shift[c] },
 *               tdf1 W[f/bn][f], scale[c], shift[c],  tdf2 W[f][f/bn], scale[c], shift[c]
 * (audio-cut_b200/unet_weights.py:pack_blob builds it).  Uploads and repacks for both dtypes. */
AC_API int ac_unet_create(const ac_unet_geom* g, const float* h_blob, size_t n_floats, ac_unet** out);
AC_API void ac_unet_destroy(ac_unet* net);
AC_API size_t ac_unet_param_floats(const ac_unet_geom* g);
AC_API size_t ac_unet_workspace_bytes(const ac_unet* net, int B, int dtype);
/* d_in, d_out: [B][dim_t][dim_f][4] in `dtype`.  d_out may alias d_in. */
AC_API int ac_unet_forward(ac_unet* net, const void* d_in, void* d_out, int B, int dtype, void* d_ws, size_t ws_bytes,
                    void* stream);
/* Test hook: 0 = let the library choose, 1 = force the CUDA-core kernels for every layer (bf16
 * storage kept), so the tcgen05 kernels can be checked layer by layer, 2 = tcgen05 kernels but
 * without the weight-stationary convolution (A/B timing), 3 = the same kernels as 0 but without the
 * cross-kernel fusions (level-0 conv chain in one launch, final 1x1 conv in the last TDF2, first 1x1 conv in
 * the STFT epilogue): every fusion is bit-identical to its unfused form, so 0 and 3 must give identical bits. */
AC_API int ac_unet_set_debug(ac_unet* net, int force_simt);
/* Test / profiling hook: one bf16 3x3 convolution layer y = relu(scale * conv(x, W) + shift).
 * impl: 0 = CUDA-core implicit GEMM on channels-last [B][T][F][C] tensors; 1 = streaming tcgen05
 * kernel, 2 = weight-stationary tcgen05 kernel (C = 48 / 96), 3 = the same with the CTA-pair
 * (cta_group::2) kernel where one exists (C = 96), 4 = CTA-pair streaming kernel (C >= 144), all on the
 * tensor-core path's
 * channel-group planar layout [B][T][C/8][F][8] for input and output.  h_w = W[C][C][3][3] float32 on the
 * host.  Runs `iters` launches; *h_ms (optional) = mean ms of launches 2..iters.  Synchronises. */
AC_API int ac_debug_conv3x3(const void* d_in, void* d_out, int B, int T, int F, int C, const float* h_w,
                     const float* d_scale, const float* d_shift, int impl, int iters, float* h_ms, void* stream);
/* Test / profiling hook: a chain of three 3x3 convolution layers (C = 48, the block shape of U-Net level 0) on
 * [B][T][C/8][F][8] tensors.  impl 0 = three launches of the weight-stationary kernel (d_out / d_tmp ping-pong),
 * 1 = the fused kernel that keeps both intermediates in shared memory; impl + 16 = IEEE-half operands.
 * h_w = 3 x W[C][C][3][3] float32 on the host, d_scale / d_shift = 3 x [C] on the device, d_tmp = scratch of the
 * tensor's size.  Runs `iters` launches; *h_ms (optional) = mean ms of launches 2..iters.  Synchronises. */
AC_API int ac_debug_conv3x3_chain(const void* d_in, void* d_out, void* d_tmp, int B, int T, int F, int C, const float* h_w,
                           const float* d_scale, const float* d_shift, int impl, int iters, float* h_ms, void* stream);
/* Test hook: synchronises and returns 1 when a tensor-core kernel gave up on an mbarrier wait
 * (a pipeline bug; the watchdog keeps such a bug from hanging the GPU), 0 otherwise. */
AC_API int ac_debug_tc_aborted(void);

/* ---- chunked separation of a whole track -------------------------------------------------------
 * Replaces the chunk loop of EnhancedVocalSeparator._separate_with_pipeline
 * (enhanced_vocal_separator.py:366-458) together with MDX23OnnxBackend.infer_chunk
 * (backends.py:299-406): window build, STFT, network, iSTFT, trim/concat, crop, stem
 * subtraction, mono mean, halo trim, accumulate, uniform average. */
typedef struct ac_chunk_desc {
  long long chunk_start; /* first track sample of the chunk: round(start_s*sr)                   */
  long long eff_start;   /* effective region [eff_start, eff_end) in track samples (halo removed) */
  long long eff_end;
  int chunk_len;         /* samples in the chunk before align_hop padding                         */
  int reserved;
} ac_chunk_desc;

typedef struct ac_track_params {
  ac_mdx_geom mdx;
  int align_hop;        /* 4096 (gpu_pipeline.align_hop / MDX23_ALIGN_HOP, backends.py:113)  */
  int n_channels;       /* 1: mono duplicated to both network channels (backends.py:269-270); 2: stereo */
  int output_is_vocal;  /* 1: network output is the vocal stem, other = mix - out (backends.py:395-401) */
  int dtype;            /* AC_F32 / AC_F16 / AC_BF16 */
  int max_batch;        /* windows per network launch (0 = library default) */
  int reserved;
} ac_track_params;

AC_API int ac_track_window_count(const ac_chunk_desc* h_chunks, int n_chunks, const ac_track_params* p);
AC_API size_t ac_track_workspace_bytes(const ac_unet* net, const ac_chunk_desc* h_chunks, int n_chunks,
                                const ac_track_params* p);
/* d_mix [n_channels][n_samples]; d_vocal/d_instr/d_weight [n_samples] (overwritten).  On return
 * (stream order) d_vocal and d_instr hold accum/max(weight,1) and d_weight the overlap counts. */
AC_API int ac_separate_track(ac_unet* net, const float* d_mix, long long n_samples, const ac_chunk_desc* h_chunks,
                      int n_chunks, const ac_track_params* p, float* d_vocal, float* d_instr, float* d_weight,
                      void* d_ws, size_t ws_bytes, void* stream);
/* The same, additionally writing every chunk's OWN vocal output - before halo trimming and overlap averaging,
 * chunks with chunk_len > 0 back to back in h_chunks order - to d_chunk_vocal [sum of chunk_len] (nullable).
 * This is what the reference hands its per-chunk VAD hook inside the chunk loop:
 * chunk_vad.process_chunk(plan, outputs.vocal, ...) at enhanced_vocal_separator.py:412-417. */
AC_API int ac_separate_track_ex(ac_unet* net, const float* d_mix, long long n_samples, const ac_chunk_desc* h_chunks,
                         int n_chunks, const ac_track_params* p, float* d_vocal, float* d_instr, float* d_weight,
                         float* d_chunk_vocal, void* d_ws, size_t ws_bytes, void* stream);
/* ac_separate_track_ex with the host<->device copies of the track pipelined against the window batches (what the
 * reference's chunk loop does with its per-chunk pinned copies, enhanced_vocal_separator.py:366-458, gpu_pipeline.py
 * PinnedBufferPool).  h_mix (optional, page-locked, [n_channels][n_samples] float32): uploaded into d_mix in pieces on
 * copy_stream, each piece before the window batch that reads it; uploaded_event (optional cudaEvent_t) is recorded on
 * copy_stream once the whole mix is resident.  h_vocal / h_instr (optional, page-locked, [n_samples] each): every stretch of
 * the stems that no later window can touch is finalised and downloaded on copy_stream while the next batch runs.  On
 * return everything is enqueued: wait for `stream` (compute) AND `copy_stream` (last download).  copy_stream must not be
 * `stream`. */
AC_API int ac_separate_track_pipelined(ac_unet* net, float* d_mix, const float* h_mix, long long n_samples,
                                const ac_chunk_desc* h_chunks, int n_chunks, const ac_track_params* p, float* d_vocal,
                                float* d_instr, float* d_weight, float* d_chunk_vocal, float* h_vocal, float* h_instr,
                                void* d_ws, size_t ws_bytes, void* stream, void* copy_stream, void* uploaded_event);

/* Stereo -> mono mean (np.mean(mix, axis=0) at features_cache.py:137-139 / enhanced_vocal_separator.py
 * mono handling); n_channels == 1 copies.  d_mix [n_channels][n] -> d_out [n]. */
AC_API int ac_downmix_mono(const float* d_mix, int n_channels, long long n, float* d_out, void* stream);
/* Energies for EnhancedVocalSeparator._estimate_separation_confidence (enhanced_vocal_separator.py:490-501)
 * and the all-zero test of the instrumental accumulator (:456-458), without a host pass over the stems:
 * d_out5 = { sum a^2, sum b^2, sum c^2, count(b != 0), watchdog } in fp64.  Any of d_a/d_b/d_c may be NULL.
 * watchdog != 0: a tcgen05 kernel launched before this call (stream order) gave up on an mbarrier wait (a wrong
 * descriptor / byte count would otherwise hang the GPU) and every tensor-core kernel after it drained out early -
 * the stems of this track are INVALID.  Reading the slot re-arms the flag; the Python drop-in raises on it. */
AC_API int ac_track_stats(const float* d_a, const float* d_b, const float* d_c, long long n, double* d_out5, void* stream);

/* ---- STFT-2048 framewise features ----------------------------------------------------------------
 * One pass over the signal per call; n_fft = 2048, periodic hann, center=True, zero padding
 * (librosa.stft defaults).  Segments reproduce the per-call (= per-chunk) scope of
 * power_to_db(top_db=80) in librosa.onset.onset_strength (features_cache.py:184): the clip
 * reference is the maximum over the segment's own frames.
 *   flatness   librosa.feature.spectral_flatness   features_cache.py:183, pure_vocal_pause_detector.py:1117
 *   onset      librosa.onset.onset_strength (mean / median over 128 slaney mels)
 *              features_cache.py:184, adaptive_vad_enhancer.py:61-67,143-148
 *   centroid   librosa.feature.spectral_centroid    pure_vocal_pause_detector.py:434
 *   low_ratio  _calculate_harmonic_ratio_direct     pure_vocal_pause_detector.py:937-959
 * Any output pointer may be NULL.  Segment s covers samples [seg_start[s], seg_start[s]+seg_len[s])
 * of d_x and writes n_s = 1 + seg_len/hop frames at frame offset seg_frame_off[s] of every output. */
typedef struct ac_feat_segment {
  long long start;     /* first sample                     */
  long long len;       /* samples                          */
  long long frame_off; /* first output frame of the segment */
} ac_feat_segment;
AC_API size_t ac_stft_features_workspace_bytes(const ac_feat_segment* h_segs, int n_segs, int hop);
AC_API int ac_stft_features(const float* d_x, const ac_feat_segment* h_segs, int n_segs, int hop, int sr, float* d_flatness,
                     float* d_onset_mean, float* d_onset_median, float* d_centroid, float* d_low_ratio, void* d_ws,
                     size_t ws_bytes, void* stream);

/* ---- F0 and formants of the legacy pause-detector branch --------
 * ac_pyin == librosa.pyin(y, fmin, fmax, sr, frame_length=2048, hop_length=hop) with librosa's defaults
 * (win_length 1024, 100 thresholds, beta(2,18), Boltzmann 2, 0.1-semitone bins, max_transition_rate
 * 35.92, switch_prob 0.01, no_trough_prob 0.01, center=True zero padding), as called at
 * pure_vocal_pause_detector.py:422-428.  n_frames = ac_pyin_frame_count = 1 + n/hop.
 *   d_f0           [n_frames] Hz, NaN where unvoiced (may be NULL: skips the Viterbi pass)
 *   d_voiced_flag  [n_frames] 0/1                    (may be NULL)
 *   d_voiced_prob  [n_frames]
 * The YIN/candidate stage is one CTA per frame; the HMM decode (1202 states, fp64) is sequential in
 * time and runs as one CTA. */
AC_API long long ac_pyin_frame_count(long long n, int hop);
AC_API size_t ac_pyin_workspace_bytes(long long n, int hop, int sr, float fmin, float fmax);
AC_API int ac_pyin(const float* d_x, long long n, int sr, int hop, float fmin, float fmax, float* d_f0,
                   unsigned char* d_voiced_flag, float* d_voiced_prob, void* d_ws, size_t ws_bytes, void* stream);
/* ac_lpc_formants == the per-frame body of PureVocalPauseDetector._extract_formants
 * (pure_vocal_pause_detector.py:961-1018): frames start at i = 0, hop, ... < n - frame (no centring),
 * pre-emphasis 0.95, librosa.lpc (Burg, `order`), |freqz(1, a, worN=512)|, scipy find_peaks above
 * 0.1 * max; d_mags[f][j] = magnitude of the j-th peak by frequency (0 beyond d_counts[f]).  The host
 * rebuilds the reference's ragged F1/F2/F3 lists: count k >= 1 appends to the first k tracks, k == 0
 * appends 0.0 to all three.  n_frames = ac_lpc_frame_count = len(range(0, n - frame, hop)). */
AC_API long long ac_lpc_frame_count(long long n, int frame, int hop);
AC_API int ac_lpc_formants(const float* d_x, long long n, int frame, int hop, int order, float* d_mags /*[n_frames][3]*/,
                           int* d_counts, void* stream);

/* ---- rhythm front end ------------------------------------------------
 * Autocorrelation tempogram of an onset envelope (librosa.feature.tempogram: hann window of `win`
 * frames, hop 1, centred with a linear ramp, autocorrelation max-normalised per frame) reduced on
 * the device to what librosa.feature.rhythm.tempo needs (features_cache.py:283-288,
 * adaptive_vad_enhancer.py:61-67, 151-156): d_best[t] = argmax_lag(log1p(1e6*tg[lag,t]) + logprior[lag])
 * (aggregate=None) and d_tg_sum[lag] = sum_t tg[lag,t] (aggregate=mean is d_tg_sum / n). */
AC_API int ac_tempogram_stats(const float* d_env, long long n, int win, const float* d_logprior, float* d_tg_sum,
                              int* d_best, void* stream);
/* HOST function: the sequential scan of librosa.beat.__beat_track_dp (Ellis DP).  h_cumscore must be
 * zero-initialised; h_backlink[i] = previous beat frame or -1. */
AC_API int ac_host_beat_dp(const float* h_localscore, int n, int period, float tightness, long long* h_backlink,
                           float* h_cumscore);

/* ---- cut-point refinement -------------------------------------------
 * The per-point loop of finalize_cut_points (src/audio_cut/cutting/refine.py:318-371) in one launch, one
 * CTA per NMS-pruned candidate: align_to_zero_cross (:72-110) on the vocal stem -> _apply_quiet_guard_fast
 * (:184-214; the 10 ms boxcar RMS-dB of _prepare_quiet_lookup :161-173 is evaluated only over the
 * [idx, idx+search) samples the point consults, never for the whole track) -> apply_quiet_guard (:113-158)
 * when the fast guard did not move the point -> the same three steps on the mix -> clip to [0, n/sr].
 * d_mix / d_vocal: mono float32 [n] device arrays (d_vocal may be NULL = ctx.vocal_wave is None);
 * d_times [n_points] fp64 seconds in, d_guard_times / d_final_times [n_points] fp64 out (CutAdjustment.
 * guard_time / final_time).  zero_cross_half = max(1, round(zero_cross_win_ms/1000*sr)), search = max(1,
 * round(search_right_ms/1000*sr)), win = max(1, round(guard_win_ms/1000*sr)) - the reference's own integers. */
AC_API int ac_refine_cut_points(const float* d_mix, const float* d_vocal, long long n, int sr, const double* d_times,
                                int n_points, int zero_cross_half, int search, int win, double guard_db,
                                double floor_db, int use_vocal_guard_first, int enable_vocal_guard,
                                int enable_mix_guard, double* d_guard_times, double* d_final_times, void* stream);
/* QuietGuardLookup.rms_db of _prepare_quiet_lookup (refine.py:161-173) for the whole track, fp64 [n]:
 * 20*log10(sqrt(convolve(wave^2, ones(win)/win, "same") + 1e-12) + 1e-12). */
AC_API int ac_quiet_lookup_db(const float* d_wave, long long n, int win, double* d_rms_db, void* stream);

/* librosa.feature.zero_crossing_rate(y, frame_length, hop_length, center=True) (edge padding,
 * |y| <= 1e-10 -> 0): pure_vocal_pause_detector.py:444. */
AC_API int ac_zero_crossing_rate(const float* d_x, long long n, int frame, int hop, float* d_out, void* stream);

/* ---- data formats either side of the path -----------------------------------
 * Decode side of AudioProcessor.load_audio (utils/audio_processor.py:32-60: librosa.load(mono=True) then audio / max|audio|):
 * d_bytes = interleaved little-endian PCM frames [n_frames][channels], bits 16 or 24; libsndfile's float normalisation
 * (/ 2^15, / 2^23).  mono != 0: d_out [n_frames] = channel mean (librosa.to_mono); else planar [channels][n_frames]. */
AC_API int ac_pcm_decode(const void* d_bytes, long long n_frames, int channels, int bits, int mono, float* d_out, void* stream);
/* x /= max|x| when the maximum is > 0 (audio_processor.py:54-56); d_scratch: one 32-bit word of device memory. */
AC_API int ac_peak_normalize(float* d_x, long long n, unsigned int* d_scratch, void* stream);
/* Encode side of export_audio (utils/audio_export.py:70-133): d_x planar [channels][n_frames] float32 -> interleaved frames.
 * format 0: PCM_24 exactly as sf.write(subtype="PCM_24") = libsndfile f2let_array, lrintf(x * 0x7FFFFF), low 3 bytes LE, no
 * clipping; 1: libsndfile's clipping variant (SFC_SET_CLIPPING: scale 2^31, saturate, top 3 bytes); 2: int16 of the MP3 path,
 * np.round(clip(x, -1, 1) * 32767).  d_out: n_frames * channels * (3 | 3 | 2) bytes. */
AC_API int ac_pcm_pack(const float* d_x, long long n_frames, int channels, int format, void* d_out, void* stream);
/* Polyphase resampler in front of the per-chunk VAD (core/vocal_pause_detector.py:175-296 resamples every chunk 44.1 -> 16 kHz
 * with librosa.resample and zero-pads it to a multiple of 4096): all segments (chunks) of d_x in one launch, semantics of
 * scipy.signal.resample_poly(x, up, down) = librosa res_type="polyphase".  Segment s = d_x[off_s, off_s + len_s) goes to row s
 * of d_out [n_segs][row_len], ac_resample_out_len(len_s) samples followed by zeros.  d_taps: the FIR (already multiplied
 * by `up`), n_pre_pad / n_pre_remove as resample_poly derives them (down - half_len % down, (half_len + n_pre_pad) / down). */
AC_API long long ac_resample_out_len(long long n_in, int up, int down);
AC_API int ac_resample_poly(const float* d_x, const long long* h_seg_off, const long long* h_seg_len, int n_segs, int up, int down,
                     const float* d_taps, int n_taps, int n_pre_pad, int n_pre_remove, long long row_len, float* d_out,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AUDIOCUT_B200_H */
