// tcgen05 / TMEM / TMA implementation of the 3x3 convolutions of the TFC blocks (75 % of the
// network's FLOPs), bf16 operands, fp32 accumulation in tensor memory.
//
// Implicit GEMM, output stationary:  D[128 positions][NT out-channels] += A_tap[128][16] * W_tap[16][NT]
//   * activations use the channel-group planar layout [B][T][C/8][F][8] (tc_common.cuh): an A
//     operand tile is ONE 3-D TMA box [KC/8 channel groups][130 positions][8 channels] whose rows
//     are 2080 contiguous bytes in global memory and which lands as the canonical no-swizzle
//     K-major UMMA layout (16-byte rows back to back, SBO = 128 B), so the three horizontal taps
//     are the SAME smem tile addressed with the descriptor start advanced by 16 B per position,
//     and conv zero padding is TMA out-of-bounds fill (f = -1, F and t = -1, T);
//   * weights are pre-packed on the host into the exact smem image ([K/8][NT][8], K-major) and
//     fetched with one 1-D bulk copy per pipeline stage;
//   * one CTA per SM, persistent over work units (n-tile, b, t, group of MT 128-position tiles);
//     MT accumulators live in TMEM so every weight tile fetched from L2 is reused MT times;
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocator), warps 2..13 =
//     epilogue (tcgen05.ld -> scale/shift/ReLU -> bf16 -> global); smem ring and (when the
//     accumulators fit twice) a double-buffered TMEM hand-off, all through mbarriers.
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "tc_common.cuh"
#include "unet_kernels.cuh"

namespace ac {

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
constexpr int kTcEpiGroups = 3;   // epilogue warps per TMEM lane quadrant: group g owns the 16-channel chunks g, g+3, ...
constexpr int kTcEpiWarps = 4 * kTcEpiGroups;
constexpr int kTcThreads = (2 + kTcEpiWarps) * 32;
constexpr int kTcHeader = 5120;   // barriers + this layer's scale/shift (<= 512 channels)
constexpr int kTileM = 128;
constexpr int kRowPos = kTileM + 2;  // positions per staged A tile (1-position halo each side)
constexpr int kRowStride = kRowPos;   // rows per channel group inside a staged tile (one dense 3-D box)

struct TcCfg {
  int C, NT, nsplit, MT, KC, nkc, stages, nbuf;
  int a_tile_bytes;   // one M tile of one stage: (KC/8) * 130 * 16, rounded up to 128
  int b_stage_bytes;  // 3 * KC * NT * 2
  int stage_bytes;    // MT * a_tile_bytes + b_stage_bytes, rounded up to 128
  int smem_bytes;
};

struct TcParams {
  TcCfg cfg;
  int nB, T, F;
  int n_fg;       // groups of MT tiles per row
  int n_units;    // nsplit * nB * T * n_fg
  const h16* wpack;
  const float* scale;
  const float* shift;
  h16* out;
  int* abort_flag;
};

template <int FMT>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_conv3x3_kernel(const __grid_constant__ CUtensorMap in_map, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  const TcCfg& c = p.cfg;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [stages]
  uint64_t* empty = full + 8;                           // [stages]
  uint64_t* tfull = full + 16;                          // [nbuf]
  uint64_t* tempty = full + 20;                         // [nbuf]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 24);
  float* s_scale = reinterpret_cast<float*>(smem + 1024);  // [C] (C <= 512)
  float* s_shift = s_scale + 512;
  uint8_t* stage0 = smem + kTcHeader;
  volatile int* abort_flag = p.abort_flag;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < c.C; i += blockDim.x) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < c.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < c.nbuf; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], kTcEpiWarps);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its output is visible from here on
  const int steps = 3 * c.nkc;  // pipeline stages consumed per work unit (dt x channel chunk)

  auto decode = [&](int u, int& nt, int& b, int& t, int& f0) {
    const int fg = u % p.n_fg;
    int q = u / p.n_fg;
    t = q % p.T;
    q /= p.T;
    b = q % p.nB;
    nt = q / p.nB;
    f0 = fg * c.MT * kTileM;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      bool alive = true;
      for (int u = blockIdx.x; u < p.n_units && alive; u += gridDim.x) {
        int nt, b, t, f0;
        decode(u, nt, b, t, f0);
        for (int dt = 0; dt < 3 && alive; ++dt) {
          for (int kc = 0; kc < c.nkc; ++kc) {
            if (!mbar_wait(&empty[s], ph ^ 1, abort_flag)) { alive = false; break; }
            uint8_t* st = stage0 + (size_t)s * c.stage_bytes;
            mbar_expect_tx(&full[s], (uint32_t)(c.MT * (c.KC / 8) * (kRowPos * 16) + c.b_stage_bytes));
            for (int mt = 0; mt < c.MT; ++mt)
              tma_load_5d(st + mt * c.a_tile_bytes, &in_map, &full[s], 0, f0 + mt * kTileM - 1, kc * (c.KC / 8), t + dt - 1, b);
            const h16* wsrc =
                p.wpack + ((size_t)((nt * 3 + dt) * c.nkc + kc)) * (size_t)(3 * c.KC * c.NT);
            bulk_load_1d(st + c.MT * c.a_tile_bytes, wsrc, (uint32_t)c.b_stage_bytes, &full[s]);
            if (++s == c.stages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // warp-uniform loop, one elected lane issues (keeps descriptors in uniform registers)
    {
      const uint32_t idesc = make_idesc<FMT>(c.NT);
      const uint32_t a_lbo = kRowStride * 16, b_lbo = (uint32_t)c.NT * 16;
      const uint64_t a_proto = make_desc(0, a_lbo, 128), b_proto = make_desc(0, b_lbo, 128);
      auto wait_all = [&](uint64_t* bar, uint32_t parity) {
        return __all_sync(0xffffffffu, mbar_wait(bar, parity, abort_flag)) != 0;
      };
      int s = 0;
      uint32_t ph = 0;
      int buf = 0;
      uint32_t tph = 0;
      bool alive = true;
      for (int u = blockIdx.x; u < p.n_units && alive; u += gridDim.x) {
        if (!wait_all(&tempty[buf], tph ^ 1)) break;
        tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(buf * c.MT * c.NT);
        for (int step = 0; step < steps; ++step) {
          if (!wait_all(&full[s], ph)) { alive = false; break; }
          tc_fence_after();
          const uint32_t sa = smem_u32(stage0 + (size_t)s * c.stage_bytes);
          const uint32_t sb = sa + (uint32_t)(c.MT * c.a_tile_bytes);
          if (elect_one()) {
            for (int df = 0; df < 3; ++df) {
              for (int mt = 0; mt < c.MT; ++mt) {
                const uint64_t ad0 = a_proto + ((sa + mt * c.a_tile_bytes + df * 16) >> 4);
                const uint64_t bd0 = b_proto + ((sb + df * (c.KC * c.NT * 2)) >> 4);
                for (int k = 0; k < c.KC / 16; ++k)
                  umma_f16(acc0 + (uint32_t)(mt * c.NT), ad0 + (uint64_t)((k * 2 * a_lbo) >> 4), bd0 + (uint64_t)((k * 2 * b_lbo) >> 4),
                           idesc, (step | df | k) != 0);
              }
            }
            umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
          }
          __syncwarp();
          if (++s == c.stages) { s = 0; ph ^= 1; }
        }
        if (!alive) break;
        if (elect_one()) umma_commit(&tfull[buf]);  // accumulators complete
        __syncwarp();
        if (++buf == c.nbuf) { buf = 0; tph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..13) =====================
    // One warp per (TMEM lane quadrant, chunk group): three warps per scheduler hide each other's
    // latencies (with one warp per scheduler the dependent ALU chains of the epilogue set the pace).
    const int quad = warp & 3;  // hardware rule: warp w may read TMEM lanes 32*(w%4) .. +31
    const int grp = (warp - 2) >> 2;
    const size_t plane = (size_t)p.F * 8;  // elements between 8-channel groups (CG8)
    int buf = 0;
    uint32_t tph = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int nt, b, t, f0;
      decode(u, nt, b, t, f0);
      const int n0 = nt * c.NT;
      const int f_lane = f0 + quad * 32 + lane;
      h16* dst0 = p.out + cg8_index(b, t, n0 >> 3, f_lane, p.T, c.C, p.F);
      if (!mbar_wait(&tfull[buf], tph, abort_flag)) break;
      tc_fence_after();
      for (int mt = 0; mt < c.MT; ++mt) {
        const bool valid = f_lane + mt * kTileM < p.F;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * c.MT * c.NT + mt * c.NT);
        h16* dst = dst0 + (size_t)mt * kTileM * 8;
        for (int j = grp * 16; j < c.NT; j += 16 * kTcEpiGroups) {
          uint32_t r[16];
          tmem_ld16(taddr + j, r);
          tmem_ld_wait();
          if (valid) {
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int ch = n0 + j + 2 * e;
              const float v0 = fmaxf(fmaf(__uint_as_float(r[2 * e]), s_scale[ch], s_shift[ch]), 0.f);
              const float v1 = fmaxf(fmaf(__uint_as_float(r[2 * e + 1]), s_scale[ch + 1], s_shift[ch + 1]), 0.f);
              pk[e] = pack2<FMT>(v0, v1);
            }
            *reinterpret_cast<uint4*>(dst + (size_t)(j >> 3) * plane) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(dst + (size_t)((j >> 3) + 1) * plane) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(&tempty[buf]);
      if (++buf == c.nbuf) { buf = 0; tph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TcConvWeights {
  int C;
  int fmt;
  TcCfg cfg;
  h16* d_pack;
};

static bool make_cfg(int C, int F, TcCfg& c) {
  if (C % 16 || C < 16 || C > 512) return false;
  c.C = C;
  c.nsplit = 1;
  if (C > 160) {
    c.nsplit = 0;
    for (int s = 2; s <= 8; ++s)
      if (C % (16 * s) == 0 && C / s <= 160) { c.nsplit = s; break; }
    if (!c.nsplit) return false;
  }
  c.NT = C / c.nsplit;
  const int mt2 = 256 / c.NT, mt1 = 512 / c.NT;
  if (mt2 >= 2) { c.MT = mt2 > 4 ? 4 : mt2; c.nbuf = 2; }
  else { c.MT = mt1 > 4 ? 4 : mt1; c.nbuf = 1; }
  const int tiles_per_row = (F + kTileM - 1) / kTileM;
  if (c.MT > tiles_per_row) c.MT = tiles_per_row;
  c.KC = 16;
  for (int k = 48; k >= 16; k -= 16)
    if (C % k == 0) { c.KC = k; break; }
  c.nkc = C / c.KC;
  c.a_tile_bytes = (int)align_up((size_t)(c.KC / 8) * kRowStride * 16, 128);
  c.b_stage_bytes = 3 * c.KC * c.NT * 2;
  c.stage_bytes = (int)align_up((size_t)c.MT * c.a_tile_bytes + c.b_stage_bytes, 128);
  const int budget = 224 * 1024 - kTcHeader;
  c.stages = budget / c.stage_bytes;
  if (c.stages > 8) c.stages = 8;
  if (c.stages < 2) return false;
  c.smem_bytes = kTcHeader + c.stages * c.stage_bytes;
  return true;
}

int tc_conv3x3_supported(int T, int F, int C) {
  TcCfg c;
  (void)T;
  return make_cfg(C, F, c) ? AC_OK : AC_E_INVALID;
}

int tc_conv3x3_pack(const float* h_w, int C, int fmt, TcConvWeights** out) {
  *out = nullptr;
  TcCfg c;
  if (!make_cfg(C, 1 << 20, c)) return AC_OK;  // unsupported shape: caller keeps the CUDA-core kernel
  // [nt][dt][kc][df][KC/8][NT][8]  <-  W[co][ci][kh=dt][kw=df]
  std::vector<h16> pack((size_t)9 * C * C);
  size_t o = 0;
  for (int nt = 0; nt < c.nsplit; ++nt)
    for (int dt = 0; dt < 3; ++dt)
      for (int kc = 0; kc < c.nkc; ++kc)
        for (int df = 0; df < 3; ++df)
          for (int kg = 0; kg < c.KC / 8; ++kg)
            for (int n = 0; n < c.NT; ++n)
              for (int e = 0; e < 8; ++e) {
                const int co = nt * c.NT + n, ci = kc * c.KC + kg * 8 + e;
                pack[o++] = h16_rn(h_w[(((size_t)co * C + ci) * 3 + dt) * 3 + df], fmt);
              }
  TcConvWeights* w = new TcConvWeights();
  w->C = C;
  w->fmt = fmt;
  w->d_pack = nullptr;
  if (cudaMalloc(&w->d_pack, pack.size() * 2) != cudaSuccess ||
      cudaMemcpy(w->d_pack, pack.data(), pack.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("tc weight upload failed");
    delete w;
    return AC_E_CUDA;
  }
  *out = w;
  return AC_OK;
}

void tc_conv3x3_free(TcConvWeights* w) {
  if (!w) return;
  if (w->d_pack) cudaFree(w->d_pack);
  delete w;
}

EncodeTiledFn get_tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

bool tc_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("AC_TC_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

static int* g_abort_flag = nullptr;  // device
int* tc_abort_flag() {
  if (!g_abort_flag) {
    if (cudaMalloc(&g_abort_flag, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(g_abort_flag, 0, sizeof(int));
  }
  return g_abort_flag;
}

int launch_tc_conv3x3(const TcConvArgs& a, cudaStream_t st) {
  AC_REQUIRE(a.w && a.w->C == a.C, "tc conv: weights do not match the layer");
  TcCfg c;
  AC_REQUIRE(make_cfg(a.C, a.F, c), "tc conv: unsupported shape");
  // the packing depends only on (C): NT, KC, nkc, nsplit are functions of C alone
  EncodeTiledFn enc = get_tensor_map_encoder();
  AC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available");
  AC_REQUIRE(tc_abort_flag() != nullptr, "abort flag allocation failed");
  CUtensorMap map;
  // CG8 tensor [nB][T][C/8][F][8] as a 5-D map (c%8, f, c/8, t, b); one box = [KC/8][130 positions][8]
  const cuuint64_t dims[5] = {8, (cuuint64_t)a.F, (cuuint64_t)(a.C / 8), (cuuint64_t)a.T, (cuuint64_t)a.nB};
  const cuuint64_t strides[4] = {16, (cuuint64_t)a.F * 16, (cuuint64_t)a.F * a.C * 2, (cuuint64_t)a.T * a.F * a.C * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)kRowPos, (cuuint32_t)(c.KC / 8), 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<h16*>(a.in), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return AC_E_CUDA;
  }
  TcParams p;
  p.cfg = c;
  p.nB = a.nB; p.T = a.T; p.F = a.F;
  p.n_fg = ((a.F + kTileM - 1) / kTileM + c.MT - 1) / c.MT;
  p.n_units = c.nsplit * a.nB * a.T * p.n_fg;
  p.wpack = a.w->d_pack;
  p.scale = a.scale; p.shift = a.shift;
  p.out = a.out;
  p.abort_flag = g_abort_flag;
  static int max_smem_set = 0;
  if (max_smem_set < c.smem_bytes) {
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3x3_kernel<kFmtF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3x3_kernel<kFmtBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    max_smem_set = 227 * 1024;
  }
  int grid = device_sm_count();
  if (grid > p.n_units) grid = p.n_units;
  ProfScope ps(KC_CONV_TC, 2.0 * 9.0 * a.nB * (double)a.T * a.F * a.C * a.C, 4.0 * a.nB * (double)a.T * a.F * a.C, st);
  AC_CHECK_CUDA(tc_launch(a.w->fmt == kFmtBF16 ? tc_conv3x3_kernel<kFmtBF16> : tc_conv3x3_kernel<kFmtF16>, grid, kTcThreads, c.smem_bytes, st, 1, map, p));
  AC_LAUNCH_CHECK();
  return AC_OK;
}

int* tc_abort_flag_if_any() { return g_abort_flag; }

// returns 1 if any tensor-core kernel hit its wait watchdog since the last call (synchronises)
int tc_check_abort() {
  if (!g_abort_flag) return 0;
  int v = 0;
  if (cudaMemcpy(&v, g_abort_flag, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (v) cudaMemset(g_abort_flag, 0, sizeof(int));
  return v;
}

}  // namespace ac
