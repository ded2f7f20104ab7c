// CTA-pair (cta_group::2) streaming 3x3 convolution for the deep U-Net levels (C = 144 ... 288), whose
// weight tensors (373 KB ... 1.5 MB) cannot stay in shared memory.
//
// The single-CTA streaming kernel (unet_tc.cu) is L2 -> shared-memory ingest bound there: per pipeline
// step a CTA fetches MT input tiles AND the complete [3 taps][KC][N] weight slab (79 KB per 1944 MMA
// cycles at C = 144; ncu: tensor pipe 52 % active).  As a CTA pair (M = 256 = one 128-position tile per
// CTA, full N) each CTA fetches only the weights of ITS N/2 output channels - the half of B the pair
// MMA reads from this CTA - so the weight stream per CTA halves while every MMA does twice the work.
// Everything else follows unet_tc.cu (output-stationary implicit GEMM, [KC/8][130][8] TMA tiles whose
// three horizontal taps are 16-byte descriptor shifts, MT accumulators per weight fetch) and the pair
// protocol of unet_tc_conv_ws.cu (leader-side full barriers fed by both CTAs' TMA, multicast commits,
// relaxed remote TMEM release).
#include <vector>

#include "tc_common.cuh"
#include "unet_kernels.cuh"

namespace ac {

constexpr int kCpEpiGroups = 3;
constexpr int kCpEpiWarps = 4 * kCpEpiGroups;
constexpr int kCpProducers = 4;
constexpr int kCpWeightWarp = 2 + kCpEpiWarps;  // warps: 0 producer, 1 MMA, 2..13 epilogue, 14..16 producers
constexpr int kCpThreads = (2 + kCpEpiWarps + kCpProducers - 1) * 32;
constexpr int kCpHeader = 5120;
constexpr int kCpTileM = 128;
constexpr int kCpRowPos = kCpTileM + 2;
constexpr int kCpMaxStages = 8;

struct CpCfg {
  int C, NT, nsplit, MT, KC, nkc, stages, nbuf;
  int a_tile_bytes;   // (KC/8) * 130 * 16 rounded up to 128
  int b_tap_bytes;    // KC * (NT/2) * 2: one horizontal tap of this CTA's half of the weights
  int stage_bytes;    // MT * a_tile_bytes + 3 * b_tap_bytes
  int smem_bytes;
};

struct CpParams {
  CpCfg cfg;
  int nB, T, F;
  int n_fg;      // groups of 2*MT tiles per row (the pair's strip)
  int n_units;   // nsplit * nB * T * n_fg
  const float* scale;
  const float* shift;
  h16* out;
  int* abort_flag;
};

__device__ __forceinline__ void cp_tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

constexpr int kCpKC = 48;                               // input channels per pipeline stage
constexpr int kCpALbo = kCpRowPos * 16;
constexpr int kCpATile = ((kCpKC / 8) * kCpALbo + 127) / 128 * 128;

__device__ __forceinline__ void cp_umma_2sm(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
  const uint64_t ad = ((uint64_t)a_hi << 32) | a_lo, bd = ((uint64_t)b_hi << 32) | b_lo;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(ad), "l"(bd), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the 9 MMAs of one tile in one pipeline stage: 3 horizontal taps x KC/16 K steps (a_lo = the tile's descriptor word)
__device__ __forceinline__ void cp_issue_tile(uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t tap16,
                                              uint32_t bk16, uint32_t acc, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int df = 0; df < 3; ++df) {
#pragma unroll
    for (int k = 0; k < kCpKC / 16; ++k) {
      cp_umma_2sm(acc, a_lo + (uint32_t)((df * 16 + k * 2 * kCpALbo) >> 4), a_hi, b_lo + (uint32_t)df * tap16 + (uint32_t)k * bk16,
                  b_hi, idesc, (df == 0 && k == 0) ? acc_first : 1u);
    }
  }
}

template <int FMT>
__global__ void __launch_bounds__(kCpThreads, 1)
tc_conv3x3_pair_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap w_map, const CpParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  const CpCfg& c = p.cfg;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [stages]  leader: both CTAs' stage landed
  uint64_t* empty = full + kCpMaxStages;                // [stages]
  uint64_t* tfull = empty + kCpMaxStages;               // [2 sets][4 tiles]  tile complete
  uint64_t* tempty = tfull + 8;                         // [2 sets][4 tiles]  leader: tile drained by both CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 8);
  float* s_scale = reinterpret_cast<float*>(smem + 1024);  // [C] (<= 512)
  float* s_shift = s_scale + 512;
  uint8_t* stage0 = smem + kCpHeader;
  volatile int* abort_flag = p.abort_flag;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  for (int i = threadIdx.x; i < c.C; i += blockDim.x) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < c.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 8; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 2 * kCpEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its output is visible from here on
  const int steps = 3 * c.nkc;  // (dt, channel chunk) pipeline steps per unit
  const int b_rows = c.b_tap_bytes / 128;  // rows of the weight tensor map per tap

  auto decode = [&](int u, int& nt, int& b, int& t, int& f0) {
    const int fg = u % p.n_fg;
    int q = u / p.n_fg;
    t = q % p.T;
    q /= p.T;
    b = q % p.nB;
    nt = q / p.nB;
    f0 = fg * (2 * c.MT * kCpTileM);
  };

  if (warp == 0 || warp >= kCpWeightWarp) {
    // ===================== TMA producers (both CTAs): four issuing threads share the requests of every stage ==========
    // A stage is MT activation tiles + 3 weight taps (this CTA's half of the slab) = 58 KB at C = 144, consumed by 1944
    // cycles of MMAs.  One issuing thread moves ~14 B/clk (scripts/microbench/tma_box.cu), two moved 29 B/clk - exactly the
    // consumption rate, and ncu had the MMA warp waiting on `full` a third of its time.  Request i of a stage goes to
    // producer i mod 4; producer 0 of the leader posts the expect_tx for the whole stage (a complete_tx that lands first just
    // runs the count negative).
    const int role = warp == 0 ? 0 : warp - kCpWeightWarp + 1;
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      bool alive = true;
      for (int u = pair; u < p.n_units && alive; u += n_pairs) {
        int nt, b, t, f0;
        decode(u, nt, b, t, f0);
        for (int dt = 0; dt < 3 && alive; ++dt) {
          for (int kc = 0; kc < c.nkc; ++kc) {
            if (!mbar_wait(&empty[s], ph ^ 1, abort_flag)) { alive = false; break; }
            if (leader && role == 0)
              mbar_expect_tx(&full[s], 2u * (uint32_t)(c.MT * (c.KC / 8) * (kCpRowPos * 16) + 3 * c.b_tap_bytes));
            const uint32_t bar = mapa_u32(smem_u32(&full[s]), 0);
            uint8_t* st = stage0 + (size_t)s * c.stage_bytes;
            // weights: slab (nt, rank, dt, kc) = 3 taps of b_rows rows each
            const int slab = ((nt * 2 + (int)rank) * 3 + dt) * c.nkc + kc;
            for (int i = role; i < c.MT + 3; i += kCpProducers) {
              if (i < c.MT)
                tma_load_5d_2sm(st + i * c.a_tile_bytes, &in_map, bar, 0, f0 + (2 * i + (int)rank) * kCpTileM - 1, kc * (c.KC / 8),
                                t + dt - 1, b);
              else
                cp_tma_load_2d_2sm(st + c.MT * c.a_tile_bytes + (i - c.MT) * c.b_tap_bytes, &w_map, bar, 0,
                                   (slab * 3 + (i - c.MT)) * b_rows);
            }
            if (++s == c.stages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader; warp-uniform loop, one elected lane issues) =====================
    if (leader) {
      // Straight-line issue: ncu (profiles/r02_conv_pair_issue.md) had this warp as the kernel's bottleneck - 26 SASS
      // instructions per MMA (runtime loop bounds from the parameter struct, 64-bit descriptor arithmetic), 139 cycles per
      // 72-cycle MMA, producers waiting on full stages, tensor pipe 54 %.  The tap loops are unrolled per MT with 32-bit
      // descriptor words: two adds and the instruction per MMA.
      const uint32_t idesc = make_idesc_2sm<FMT>(c.NT);
      const uint32_t b_lbo = (uint32_t)(c.NT / 2) * 16;
      const uint64_t a_proto = make_desc(0, kCpALbo, 128), b_proto = make_desc(0, b_lbo, 128);
      const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
      const uint32_t a_lo0 = (uint32_t)a_proto + (smem_u32(stage0) >> 4);
      const uint32_t b_lo0 = (uint32_t)b_proto + ((smem_u32(stage0) + (uint32_t)(c.MT * kCpATile)) >> 4);
      const uint32_t stage16 = (uint32_t)(c.stage_bytes >> 4), tap16 = (uint32_t)(c.b_tap_bytes >> 4), bk16 = (2 * b_lbo) >> 4;
      auto wait_all = [&](uint64_t* bar, uint32_t parity) {
        return __all_sync(0xffffffffu, mbar_wait(bar, parity, abort_flag)) != 0;
      };
      int s = 0;
      uint32_t ph = 0;
      uint32_t n_acc = 0;
      bool alive = true;
      // Tiles are issued one after the other inside a stage (tile-major: the accumulation order per output is unchanged), with
      // one complete / drained barrier pair PER TILE: with a single accumulator set (C >= 144: MT * NT > 256 columns) the drain
      // of a unit used to stop the MMAs for three tile drains; now tile 0 drains under the last stage's MMAs of tiles 1 and 2, and
      // the next unit's first stage starts on tile 0 while tile 2 is still draining.
      for (int u = pair; u < p.n_units && alive; u += n_pairs, ++n_acc) {
        const int buf = c.nbuf == 2 ? (int)(n_acc & 1) : 0;
        const uint32_t use = c.nbuf == 2 ? (n_acc >> 1) : n_acc;
        const uint32_t acc0 = tmem_base + (uint32_t)(buf * c.MT * c.NT);
        for (int step = 0; step < steps && alive; ++step) {
          if (!wait_all(&full[s], ph)) { alive = false; break; }
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)s * stage16, b_lo = b_lo0 + (uint32_t)s * stage16;
          for (int mt = 0; mt < c.MT; ++mt) {
            if (step == 0) {
              if (!wait_all(&tempty[buf * 4 + mt], (use & 1) ^ 1)) { alive = false; break; }
              tc_fence_after();
            }
            if (elect_one()) {
              cp_issue_tile(a_lo + (uint32_t)mt * (kCpATile >> 4), a_hi, b_lo, b_hi, tap16, bk16, acc0 + (uint32_t)(mt * c.NT), idesc,
                            step == 0 ? 0u : 1u);  // the unit's first MMA per tile overwrites the accumulator
              if (step == steps - 1) umma_commit_2sm(&tfull[buf * 4 + mt]);
              if (mt == c.MT - 1) umma_commit_2sm(&empty[s]);
            }
            __syncwarp();
          }
          if (++s == c.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (12 warps per CTA: own tiles, all NT channels) =====================
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;
    const size_t plane = (size_t)p.F * 8;
    const uint32_t tempty_leader = mapa_u32(smem_u32(&tempty[0]), 0);
    uint32_t n_acc = 0;
    for (int u = pair; u < p.n_units; u += n_pairs, ++n_acc) {
      const int buf = c.nbuf == 2 ? (int)(n_acc & 1) : 0;
      const uint32_t use = c.nbuf == 2 ? (n_acc >> 1) : n_acc;
      int nt, b, t, f0;
      decode(u, nt, b, t, f0);
      const int n0 = nt * c.NT;
      // This warp's 16-column chunks of a tile: columns (grp + 3 q) * 16.  Tile by tile: wait for the tile, pull its chunks
      // (the TMEM load of chunk q+1 in flight while chunk q is scaled, packed and stored), hand the tile back after the last load.
      const int per_tile = (c.NT / 16 - grp + kCpEpiGroups - 1) / kCpEpiGroups;
      bool alive = true;
      for (int mt = 0; mt < c.MT && alive; ++mt) {
        if (!mbar_wait(&tfull[buf * 4 + mt], use & 1, abort_flag)) { alive = false; break; }
        tc_fence_after();
        const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * c.MT * c.NT + mt * c.NT);
        const int f = f0 + (2 * mt + (int)rank) * kCpTileM + quad * 32 + lane;
        uint32_t ra[16], rb[16];
        auto chunk_addr = [&](int q) { return tbase + (uint32_t)((grp + kCpEpiGroups * q) * 16); };
        auto finish = [&](int q, const uint32_t* r) {
          const int j = (grp + kCpEpiGroups * q) * 16;
          if (f < p.F) {
            h16* dst = p.out + cg8_index(b, t, (n0 + j) >> 3, f, p.T, c.C, p.F);
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int ch = n0 + j + 2 * e;
              const float v0 = fmaxf(fmaf(__uint_as_float(r[2 * e]), s_scale[ch], s_shift[ch]), 0.f);
              const float v1 = fmaxf(fmaf(__uint_as_float(r[2 * e + 1]), s_scale[ch + 1], s_shift[ch + 1]), 0.f);
              pk[e] = pack2<FMT>(v0, v1);
            }
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(dst + plane) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        };
        auto release = [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(tempty_leader + (uint32_t)(buf * 4 + mt) * 8);
        };
        if (per_tile == 0) {
          release();
          continue;
        }
        tmem_ld16(chunk_addr(0), ra);
        tmem_ld_wait();
        for (int i = 0; i < per_tile; i += 2) {
          if (i + 1 < per_tile) tmem_ld16(chunk_addr(i + 1), rb);
          else release();
          finish(i, ra);
          if (i + 1 < per_tile) {
            tmem_ld_wait();
            if (i + 2 < per_tile) tmem_ld16(chunk_addr(i + 2), ra);
            else release();
            finish(i + 1, rb);
            if (i + 2 < per_tile) tmem_ld_wait();
          }
        }
      }
      if (!alive) break;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TcConvPairWeights {
  int C;
  int fmt;
  CpCfg cfg;
  h16* d_pack;
  size_t pack_elems;
};

static bool cp_make_cfg(int C, int F, CpCfg& c) {
  if (C % 48 || C < 144 || C > 512) return false;  // C = 48 / 96 have the weight-stationary kernels
  c.C = C;
  c.nsplit = 0;
  for (int s = 1; s <= 4; ++s)
    if (C % (32 * s) == 0 && C / s <= 256) { c.nsplit = s; break; }  // N/2 must be a multiple of 8 (N % 16 for M = 256 -> % 32)
  if (!c.nsplit) {
    for (int s = 1; s <= 4; ++s)
      if (C % (16 * s) == 0 && C / s <= 256) { c.nsplit = s; break; }
  }
  if (!c.nsplit) return false;
  c.NT = C / c.nsplit;
  if (c.NT % 16 || (c.NT / 2) % 8) return false;
  c.MT = 512 / c.NT;
  if (c.MT > 4) c.MT = 4;
  const int tiles_per_row = (F + kCpTileM - 1) / kCpTileM;
  if (tiles_per_row < 2) return false;  // a single 128-position tile per row would leave the peer CTA idle
  const int pair_tiles = (tiles_per_row + 1) / 2;
  if (c.MT > pair_tiles) c.MT = pair_tiles;
  if (c.MT < 1) return false;
  c.nbuf = (2 * c.MT * c.NT <= 512) ? 2 : 1;
  c.KC = kCpKC;  // fixed: the kernel's stage issue code is unrolled for it
  c.nkc = C / c.KC;
  c.a_tile_bytes = (int)align_up((size_t)(c.KC / 8) * kCpRowPos * 16, 128);
  c.b_tap_bytes = c.KC * (c.NT / 2) * 2;
  if (c.b_tap_bytes % 128 || c.b_tap_bytes / 128 > 256) return false;
  c.stage_bytes = c.MT * c.a_tile_bytes + 3 * c.b_tap_bytes;
  c.stages = (226 * 1024 - kCpHeader) / c.stage_bytes;
  if (c.stages > kCpMaxStages) c.stages = kCpMaxStages;
  if (c.stages < 2) return false;
  c.smem_bytes = kCpHeader + c.stages * c.stage_bytes;
  return true;
}

int tc_conv3x3_pair_supported(int T, int F, int C) {
  CpCfg c;
  (void)T;
  return cp_make_cfg(C, F, c) ? AC_OK : AC_E_INVALID;
}

int tc_conv3x3_pair_pack(const float* h_w, int C, int fmt, TcConvPairWeights** out) {
  *out = nullptr;
  CpCfg c;
  if (!cp_make_cfg(C, 1 << 20, c)) return AC_OK;
  // [nt][rank][dt][kc][df][KC/8][NT/2][8]  <-  W[co][ci][kh=dt][kw=df]
  const int NH = c.NT / 2;
  std::vector<h16> pack((size_t)9 * C * C);
  size_t o = 0;
  for (int nt = 0; nt < c.nsplit; ++nt)
    for (int r = 0; r < 2; ++r)
      for (int dt = 0; dt < 3; ++dt)
        for (int kc = 0; kc < c.nkc; ++kc)
          for (int df = 0; df < 3; ++df)
            for (int kg = 0; kg < c.KC / 8; ++kg)
              for (int n = 0; n < NH; ++n)
                for (int e = 0; e < 8; ++e) {
                  const int co = nt * c.NT + r * NH + n, ci = kc * c.KC + kg * 8 + e;
                  pack[o++] = h16_rn(h_w[(((size_t)co * C + ci) * 3 + dt) * 3 + df], fmt);
                }
  TcConvPairWeights* w = new TcConvPairWeights();
  w->C = C;
  w->fmt = fmt;
  w->d_pack = nullptr;
  w->pack_elems = pack.size();
  if (cudaMalloc(&w->d_pack, pack.size() * 2) != cudaSuccess ||
      cudaMemcpy(w->d_pack, pack.data(), pack.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("tc pair conv weight upload failed");
    delete w;
    return AC_E_CUDA;
  }
  *out = w;
  return AC_OK;
}

void tc_conv3x3_pair_free(TcConvPairWeights* w) {
  if (!w) return;
  if (w->d_pack) cudaFree(w->d_pack);
  delete w;
}

int launch_tc_conv3x3_pair(const TcConvPairWeights* w, const TcConvArgs& a, cudaStream_t st) {
  AC_REQUIRE(w && w->C == a.C, "tc pair conv: weights do not match the layer");
  CpCfg c;
  AC_REQUIRE(cp_make_cfg(a.C, a.F, c), "tc pair conv: unsupported shape");
  EncodeTiledFn enc = get_tensor_map_encoder();
  AC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available");
  AC_REQUIRE(tc_abort_flag() != nullptr, "abort flag allocation failed");
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap in_map, w_map;
  {
    const cuuint64_t dims[5] = {8, (cuuint64_t)a.F, (cuuint64_t)(a.C / 8), (cuuint64_t)a.T, (cuuint64_t)a.nB};
    const cuuint64_t strides[4] = {16, (cuuint64_t)a.F * 16, (cuuint64_t)a.F * a.C * 2, (cuuint64_t)a.T * a.F * a.C * 2};
    const cuuint32_t box[5] = {8, (cuuint32_t)kCpRowPos, (cuuint32_t)(c.KC / 8), 1, 1};
    CUresult r = enc(&in_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<h16*>(a.in), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (pair conv, input) failed with code " + std::to_string((int)r));
      return AC_E_CUDA;
    }
  }
  {
    // packed weights as rows of 64 bf16 (128 B); one tap of one slab = b_tap_bytes / 128 consecutive rows
    const cuuint64_t dims[2] = {64, (cuuint64_t)(w->pack_elems / 64)};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, (cuuint32_t)(c.b_tap_bytes / 128)};
    CUresult r = enc(&w_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w->d_pack, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (pair conv, weights) failed with code " + std::to_string((int)r));
      return AC_E_CUDA;
    }
  }
  CpParams p;
  p.cfg = c;
  p.nB = a.nB; p.T = a.T; p.F = a.F;
  const int tiles = (a.F + kCpTileM - 1) / kCpTileM;
  p.n_fg = (tiles + 2 * c.MT - 1) / (2 * c.MT);
  p.n_units = c.nsplit * a.nB * a.T * p.n_fg;
  p.scale = a.scale; p.shift = a.shift;
  p.out = a.out;
  p.abort_flag = tc_abort_flag();
  static bool attr_set = false;
  if (!attr_set) {
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3x3_pair_kernel<kFmtF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_conv3x3_pair_kernel<kFmtBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int pairs = device_sm_count() / 2;
  if (pairs > p.n_units) pairs = p.n_units;
  ProfScope ps(KC_CONV_TC, 2.0 * 9.0 * a.nB * (double)a.T * a.F * a.C * a.C, 4.0 * a.nB * (double)a.T * a.F * a.C, st);
  AC_CHECK_CUDA(tc_launch(w->fmt == kFmtBF16 ? tc_conv3x3_pair_kernel<kFmtBF16> : tc_conv3x3_pair_kernel<kFmtF16>, 2 * pairs, kCpThreads, c.smem_bytes, st, 2, in_map, w_map, p));
  AC_LAUNCH_CHECK();
  return AC_OK;
}

}  // namespace ac
