// Internal interface between the U-Net orchestration (unet.cu) and its kernels.
// CUDA-core kernels (fp32 path, bf16 cross-check path): activations are channels-last [B][T][F][C].
// tcgen05 kernels (16-bit production path, operands f16 or bf16 = `fmt`, common.cuh): activations are "CG8" = [B][T][C/8][F][8], see tc_common.cuh.
#pragma once
#include "common.cuh"

namespace ac {

enum AMode { A_CONV3 = 0, A_DOWN2 = 1, A_PLAIN = 2 };
enum EpiMode { EPI_AFFINE_RELU = 0, EPI_UP_SKIP = 1, EPI_RESIDUAL = 2 };

// C[batch][M][N] = A[M][K] * B[batch][K][N]  (+ epilogue), fp32 accumulate.
struct GemmArgs {
  int M, N, K;
  int batch;               // blockIdx.z
  // ---- A operand
  int a_mode;
  const void* A;           // A_PLAIN: [M][K] row-major (+ batch*a_batch_stride); else input activations
  long long a_batch_stride;
  int T, F, C;             // A_CONV3 / A_DOWN2: OUTPUT grid (T x F) and input channels; M = nB*T*F
  // ---- B operand: [K][N] row-major
  const void* Bm;
  long long b_batch_stride;
  // ---- epilogue: y = relu(acc*scale[n % cmod] + shift[n % cmod])
  int epi;
  const float* scale;
  const float* shift;
  int cmod;
  void* out;               // EPI_AFFINE_RELU / EPI_RESIDUAL: [batch][M][N]
  long long c_batch_stride;
  const void* extra;       // EPI_UP_SKIP: skip tensor (same layout as out); EPI_RESIDUAL: residual [batch][M][N]
  int up_T, up_F;          // EPI_UP_SKIP: INPUT grid; out is [nB][2*up_T][2*up_F][N/4]
  int kclass;              // profiler class (KC_*)
};

int launch_gemm_simt(const GemmArgs& a, int dtype, cudaStream_t st);

// first 1x1 conv: [P][4] -> [P][g] with affine+relu;  final 1x1 conv: [P][g] -> [P][4] + bias
int launch_first_conv(const void* in, void* out, long long P, int g, const float* w /*[g][4]*/, const float* scale,
                      const float* shift, int dtype, cudaStream_t st);
int launch_final_conv(const void* in, void* out, long long P, int g, const float* w /*[4][g]*/, const float* bias,
                      int dtype, cudaStream_t st);

// ---- hooks for a producer that applies the network's first 1x1 conv itself (track.cu + the fused STFT epilogue) -----------------
int unet_first_conv_target(ac_unet* net, int B, int dtype, void* d_ws, size_t ws_bytes, void** d_target, const float** w,
                           const float** scale, const float** shift);
int unet_base_channels(const ac_unet* net);  // g: channels after the first conv
int unet_forward_after_first(ac_unet* net, void* d_out, int B, int dtype, void* d_ws, size_t ws_bytes, cudaStream_t st);

// ---- tcgen05 path (f16 / bf16 operands: every *_pack takes the format, kFmtF16 / kFmtBF16) -------------------------------------------------------------------
struct TcConvWeights;  // opaque: packed smem images for one 3x3 conv layer
struct TcConvArgs {
  const h16* in;  // CG8 [nB][T][C/8][F][8]
  h16* out;       // CG8 [nB][T][C/8][F][8]
  int nB, T, F, C;
  const TcConvWeights* w;
  const float* scale;
  const float* shift;
};
// returns AC_OK, or AC_E_INVALID when the shape is not supported by the tensor-core kernel
int tc_conv3x3_supported(int T, int F, int C);
int tc_conv3x3_pack(const float* h_w /*[C][C][3][3]*/, int C, int fmt, TcConvWeights** out);
void tc_conv3x3_free(TcConvWeights* w);
int launch_tc_conv3x3(const TcConvArgs& a, cudaStream_t st);

// weight-stationary variant for C = 48 / 96 (unet_tc_conv_ws.cu); *out stays nullptr for other widths
struct TcConvWsWeights;
int tc_conv3x3_ws_supported(int T, int F, int C);
int tc_conv3x3_ws_pack(const float* h_w /*[C][C][3][3]*/, int C, int fmt, TcConvWsWeights** out);
void tc_conv3x3_ws_free(TcConvWsWeights* w);
int launch_tc_conv3x3_ws(const TcConvWsWeights* w, const TcConvArgs& a, cudaStream_t st);
void tc_conv3x3_ws_set_rs(int enabled);    // test hook: 0 = never use the row-stacked (N = 144) kernel for C = 48
void tc_conv3x3_ws_set_pair(int enabled);  // test hook: 0 = never use the CTA-pair (cta_group::2) kernel

// fused chain of three 3x3 convs for C = 48 (unet_tc_conv_f3.cu): intermediates stay in shared memory; *out stays nullptr
// for other widths.  Output bit-identical to three launch_tc_conv3x3_ws calls.
struct TcConvF3Weights;
int tc_conv3x3_f3_supported(int T, int F, int C, int n_convs);
int tc_conv3x3_f3_pack(const float* const h_w[3] /*[C][C][3][3] each*/, int C, int fmt, TcConvF3Weights** out);
void tc_conv3x3_f3_free(TcConvF3Weights* w);
int launch_tc_conv3x3_f3(const TcConvF3Weights* w, const h16* in, h16* out, int nB, int T, int F, const float* const scale[3],
                         const float* const shift[3], cudaStream_t st);

// CUDA-core pieces of the CG8 path (unet_cg8.cu)
int cg8_ends_supported(int g);
// spec [rows*F][4] bf16 -> CG8 [rows][g/8][F][8] (rows = nB*T);  and back with bias
int launch_first_conv_cg8(const void* in, void* out, long long rows, int F, int g, const float* w /*[g][4]*/,
                          const float* scale, const float* shift, int fmt, cudaStream_t st);
int launch_final_conv_cg8(const void* in, void* out, long long rows, int F, int g, const float* w /*[4][g]*/,
                          const float* bias, int fmt, cudaStream_t st);
// TDF layer too small for a UMMA tile: in CG8 [nB][T][C/8][K][8], w bf16 [M][K], residual/out CG8 [nB][T][C/8][M][8]
int launch_tdf_small_cg8(const h16* in, const h16* w, const h16* residual, h16* out,
                         int nB, int T, int C, int M, int K, const float* scale, const float* shift, int fmt, cudaStream_t st);

// CTA-pair streaming variant for C >= 144 (unet_tc_conv_pair.cu); *out stays nullptr for other widths
struct TcConvPairWeights;
int tc_conv3x3_pair_supported(int T, int F, int C);
int tc_conv3x3_pair_pack(const float* h_w /*[C][C][3][3]*/, int C, int fmt, TcConvPairWeights** out);
void tc_conv3x3_pair_free(TcConvPairWeights* w);
int launch_tc_conv3x3_pair(const TcConvPairWeights* w, const TcConvArgs& a, cudaStream_t st);

struct TcResampleWeights;  // opaque: packed smem images of a 2x2/s2 conv (down) or transposed conv (up)
int tc_resample_pack(int up, const float* h_w, int Cin, int Cout, int fmt, TcResampleWeights** out);
void tc_resample_free(TcResampleWeights* w);
// all tensors CG8.  DOWN: in (2T x 2F, Cin) -> out (T x F, Cout).  UP: in (T x F, Cin); skip, out (2T x 2F, Cout)
int launch_tc_resample(const TcResampleWeights* w, const h16* in, const h16* skip, h16* out,
                       int nB, int T, int F, const float* scale, const float* shift, cudaStream_t st);

struct TcTdfWeights;  // opaque: packed smem images of one TDF linear layer
// *out stays nullptr when the shape is left to the CUDA-core kernel (tiny deep-level layers)
int tc_tdf_pack(const float* h_w /*[M][K]*/, int M, int K, int C, int T, int fmt, TcTdfWeights** out);
void tc_tdf_free(TcTdfWeights* w);
// out[b][t][m][c] = relu(scale[c]*sum_k W[m][k]*in[b][t][k][c] + shift[c]) (+ residual[b][t][m][c]); tensors CG8
int launch_tc_tdf(const TcTdfWeights* w, const h16* in, const h16* residual, h16* out,
                  int nB, int T, const float* scale, const float* shift, cudaStream_t st);

// CTA-pair (cta_group::2), activation-stationary kernel for the second TDF layer (with residual);
// *out stays nullptr when the shape is not supported (M % 256, K % 32, N = NTt*C <= 256 ...)
struct TcTdf2PairWeights;
int tc_tdf2_pair_pack(const float* h_w /*[M][K]*/, int M, int K, int C, int T, int fmt, TcTdf2PairWeights** out);
void tc_tdf2_pair_free(TcTdf2PairWeights* w);
// final_w/final_b != nullptr: fused final 1x1 conv (C -> 4): `out` is then the [nB*T*M][4] network output
bool tc_tdf2_pair_can_fuse_final(const TcTdf2PairWeights* w);
int launch_tc_tdf2_pair(const TcTdf2PairWeights* w, const h16* in, const h16* residual, h16* out,
                        int nB, int T, const float* scale, const float* shift, cudaStream_t st, const float* final_w = nullptr,
                        const float* final_b = nullptr);

// CTA-pair kernel for the first TDF layer (no residual); *out stays nullptr for unsupported shapes
struct TcTdf1PairWeights;
int tc_tdf1_pair_pack(const float* h_w /*[M][K]*/, int M, int K, int C, int T, int fmt, TcTdf1PairWeights** out);
void tc_tdf1_pair_free(TcTdf1PairWeights* w);
int launch_tc_tdf1_pair(const TcTdf1PairWeights* w, const h16* in, h16* out, int nB, int T, const float* scale,
                        const float* shift, cudaStream_t st);

}  // namespace ac
