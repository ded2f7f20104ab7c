// Internal interface of the MDX STFT / fused iSTFT-OLA kernels (shared with track.cu).
#pragma once
#include "common.cuh"

namespace ac {

// One model window (261120 samples at Kim_Vocal geometry) cut out of a signal buffer.
struct WinDesc {
  long long base;       // index (per channel) of window position p = 0 in the source buffer
  long long out_base;   // stems mode: track sample that window output n = trim maps to
  long long eff_start;  // stems mode: only track samples in [eff_start, eff_end) are written
  long long eff_end;
  int p_lo, p_hi;       // window positions holding real samples; outside -> 0 (zero padding)
  int out_len;          // stems mode: outputs n = trim .. trim+out_len-1 are valid (<= gen)
  int pad_;
  long long side_base;  // stems mode: index in the per-chunk side buffer of this window's first output (see launch_istft)
};

struct MdxPlan {
  ac_mdx_geom g;
  int W;               // hop*(dim_t-1)
  const FftPlan* fft;  // length n_fft
  const float* d_env;  // sum of squared windows per padded position, (dim_t-1)*hop + n_fft values
};
const MdxPlan* get_mdx_plan(const ac_mdx_geom& g);

// src: [n_ch][ch_stride] (n_ch 1 or 2).  spec: [n_win][dim_t][dim_f][4] of T (float / bf16).
int launch_stft(const MdxPlan* plan, const float* d_src, long long ch_stride, int n_ch, const WinDesc* d_wins,
                int n_win, void* d_spec, int dtype, cudaStream_t st);

// STFT with the network's first 1x1 conv (4 -> 48 channels, folded BN + ReLU) applied in the epilogue: writes the CG8 tensor
// [n_win][dim_t][48/8][dim_f][8] the level-0 conv chain reads, bit-identical to launch_stft + first_conv_cg8_kernel (the four
// spectrogram values are rounded to the 16-bit format first, as the separate store would).  AC_E_INVALID when the shape has no
// fused kernel (n_fft other than 7680 / 6144, g != 48, fp32): the caller then runs the two launches.
int launch_stft_first_conv(const MdxPlan* plan, const float* d_src, long long ch_stride, int n_ch, const WinDesc* d_wins, int n_win,
                           void* d_cg8, int g, const float* d_w /*[g][4]*/, const float* d_scale, const float* d_shift, int dtype,
                           cudaStream_t st);
bool stft_first_conv_supported(const MdxPlan* plan, int g, int dtype);

// mode 0: raw torch.istft output into d_wave [n_win][2][W].
// mode 1: stems: trims n_fft/2 per side, maps to the track through WinDesc, subtracts from the
//         mix, averages the two channels and atomically accumulates vocal / instrumental / weight.
int launch_istft(const MdxPlan* plan, const void* d_spec, int dtype, const WinDesc* d_wins, int n_win, int mode,
                 float* d_wave, const float* d_mix, long long mix_stride, int n_ch, int output_is_vocal,
                 float* d_vocal, float* d_instr, float* d_weight, cudaStream_t st, float* d_chunk_vocal = nullptr);
// d_chunk_vocal (stems mode, nullable): every chunk's OWN vocal output before halo trimming and overlap averaging,
// chunks back to back (WinDesc::side_base) - what the reference hands its per-chunk VAD (enhanced_vocal_separator.py:412-417).

}  // namespace ac
