// CTA-pair (cta_group::2) tcgen05 kernel for the FIRST TDF layer of a block (F -> F/8 features):
//   H[b][t][m][c] = relu(scale[c] * sum_k W[m][k] * X[b][t][k][c] + shift[c]),   M = F/8, K = F
//
// Both operands stream along the long K axis; nothing can be stationary except the accumulators, so
// the only lever against the L2 -> shared-memory ingest bound of the single-CTA kernel (the whole
// weight matrix is re-fetched for every N = 96-column unit) is a bigger output tile per weight pass.
// A CTA pair doubles it twice over: the pair's TMEM holds all of M (<= 512 rows as M/256 row pairs) for
// N <= 256 columns, each CTA fetches only its 128-row halves of the weights and its N/2 columns of X.
// Per output column a CTA ingests ~2.7x fewer bytes than before and the M = 256 MMAs are tensor-pipe
// bound.  Rows >= M of the last row pair are zero blobs (M = 384 -> 512, 192 -> 256).
// Pipeline / barriers as in unet_tc_tdf2_pair.cu: leader-side full barriers collect both CTAs' TMA bytes
// (cp.async.bulk.tensor ... cta_group::2), commits are multicast, TMEM is released by relaxed remote arrives.
#include <vector>

#include "tc_common.cuh"
#include "unet_kernels.cuh"

namespace ac {

constexpr int kT1EpiGroups = 3;
constexpr int kT1EpiWarps = 4 * kT1EpiGroups;
constexpr int kT1FirstEpiWarp = 2;  // warps: 0 producer, 1 MMA, 2..13 epilogue
constexpr int kT1WeightWarp = kT1FirstEpiWarp + kT1EpiWarps;  // second TMA producer: the weight blobs of every stage
constexpr int kT1Threads = (kT1FirstEpiWarp + kT1EpiWarps + 1) * 32;
constexpr int kT1Header = 4096;
constexpr int kT1MaxStages = 8;

struct T1Cfg {
  int C, M, K;
  int NTt, N, split_t;
  int Kt, nk;
  int n_mp;           // row pairs of 256 (last one zero padded)
  int nbuf;           // TMEM accumulator sets (2 when n_mp * N * 2 <= 512)
  int stages;
  int a_blob_bytes;   // 128 * Kt * 2
  int b_bytes;        // (N/2) * Kt * 2
  int stage_bytes;
  int smem_bytes;
};

struct T1Params {
  T1Cfg cfg;
  int nB, T;
  int n_tg, n_units;
  const float* scale;
  const float* shift;
  h16* out;
  int* abort_flag;
};

__device__ __forceinline__ void t1_tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void t1_umma_2sm(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  const uint64_t ad = ((uint64_t)a_hi << 32) | a_lo, bd = ((uint64_t)b_hi << 32) | b_lo;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(ad), "l"(bd), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int FMT>
__global__ void __launch_bounds__(kT1Threads, 1)
tc_tdf1_pair_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap w_map, const T1Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  const T1Cfg& c = p.cfg;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);   // [stages]  leader: both CTAs' stage landed
  uint64_t* empty = full + kT1MaxStages;                 // [stages]
  uint64_t* tfull = empty + kT1MaxStages;                // [2]
  uint64_t* tempty = tfull + 2;                          // [2]       leader: both CTAs' epilogues drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_scale = reinterpret_cast<float*>(smem + 1024);  // [C] (<= 256)
  float* s_shift = s_scale + 256;
  uint8_t* ring = smem + kT1Header;
  volatile int* abort_flag = p.abort_flag;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int n_my = pair < p.n_units ? (p.n_units - pair + n_pairs - 1) / n_pairs : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < c.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 2 * kT1EpiWarps);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < c.C; i += blockDim.x) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
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
  const int acc_cols = c.n_mp * c.N;  // columns of one accumulator set

  if (warp == 0) {
    // ===================== producer: per K chunk n_mp weight blobs (this CTA's 128 rows) + this CTA's half of X
    if (lane == 0) {
      long long i = 0;
      bool alive = true;
      for (int lu = 0; lu < n_my && alive; ++lu) {
        const int u = pair + lu * n_pairs;
        const int b = u / p.n_tg, t0 = (u - b * p.n_tg) * c.NTt;
        for (int kc = 0; kc < c.nk; ++kc, ++i) {
          const int s = (int)(i % c.stages);
          if (!mbar_wait(&empty[s], (uint32_t)(((i / c.stages) & 1) ^ 1), abort_flag)) { alive = false; break; }
          if (leader) mbar_expect_tx(&full[s], 2u * (uint32_t)c.stage_bytes);
          const uint32_t bar = mapa_u32(smem_u32(&full[s]), 0);
          uint8_t* st = ring + (size_t)s * c.stage_bytes;
          uint8_t* xb = st + (size_t)c.n_mp * c.a_blob_bytes;
          if (c.split_t)
            tma_load_5d_2sm(xb, &x_map, bar, 0, kc * c.Kt, 0, t0 + (int)rank * (c.NTt / 2), b);
          else
            tma_load_5d_2sm(xb, &x_map, bar, 0, kc * c.Kt, (int)rank * (c.C / 16), t0, b);
        }
      }
    }
  } else if (warp == kT1WeightWarp) {
    // ===================== second producer: the n_mp weight blobs of every stage (their bytes are part of the
    // expect_tx the first producer posts; one issuing thread alone tops out near 14 B/clk) =====================
    if (lane == 0) {
      long long i = 0;
      bool alive = true;
      for (int lu = 0; lu < n_my && alive; ++lu) {
        for (int kc = 0; kc < c.nk; ++kc, ++i) {
          const int s = (int)(i % c.stages);
          if (!mbar_wait(&empty[s], (uint32_t)(((i / c.stages) & 1) ^ 1), abort_flag)) { alive = false; break; }
          const uint32_t bar = mapa_u32(smem_u32(&full[s]), 0);
          uint8_t* st = ring + (size_t)s * c.stage_bytes;
          for (int mp = 0; mp < c.n_mp; ++mp) {
            const int blob = (mp * 2 + (int)rank) * c.nk + kc;
            t1_tma_load_2d_2sm(st + (size_t)mp * c.a_blob_bytes, &w_map, bar, 0, blob * (c.a_blob_bytes / 512));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) ======
    if (leader) {
      auto wait_all = [&](uint64_t* bar, uint32_t parity) {
        return __all_sync(0xffffffffu, mbar_wait(bar, parity, abort_flag)) != 0;
      };
      // ring stage / phase carried incrementally and 32-bit descriptor words (the 64-bit division and modulo per stage and
      // the 64-bit descriptor sums cost this warp more instructions per MMA than a 96-cycle MMA leaves room for)
      const uint32_t idesc = make_idesc_2sm<FMT>(c.N) | (1u << 16);  // B is MN-major
      const uint64_t a_proto = make_desc(0, 128 * 16, 128);
      const uint64_t b_proto = make_desc_mn(0, 128, (uint32_t)c.Kt * 16);
      const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
      const uint32_t a_lo0 = (uint32_t)a_proto + (smem_u32(ring) >> 4);
      const uint32_t b_lo0 = (uint32_t)b_proto + ((smem_u32(ring) + (uint32_t)(c.n_mp * c.a_blob_bytes)) >> 4);
      const uint32_t stage16 = (uint32_t)c.stage_bytes >> 4, blob16 = (uint32_t)c.a_blob_bytes >> 4;
      const int ksteps = c.Kt / 16;
      int s = 0;
      uint32_t ph = 0;
      bool alive = true;
      for (int lu = 0; lu < n_my && alive; ++lu) {
        const int buf = c.nbuf == 2 ? (lu & 1) : 0;
        const uint32_t use = c.nbuf == 2 ? (uint32_t)(lu >> 1) : (uint32_t)lu;
        if (!wait_all(&tempty[buf], (use & 1) ^ 1)) break;
        const uint32_t acc0 = tmem_base + (uint32_t)(buf * acc_cols);
        for (int kc = 0; kc < c.nk; ++kc) {
          if (!wait_all(&full[s], ph)) { alive = false; break; }
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = a_lo0 + (uint32_t)s * stage16, sb = b_lo0 + (uint32_t)s * stage16;
            for (int mp = 0; mp < c.n_mp; ++mp) {
              const uint32_t a_lo = sa + (uint32_t)mp * blob16;
              const uint32_t acc = acc0 + (uint32_t)(mp * c.N);
#pragma unroll 4
              for (int k = 0; k < ksteps; ++k)
                t1_umma_2sm(acc, a_lo + (uint32_t)k * 256u, a_hi, sb + (uint32_t)k * 16u, b_hi, idesc, (kc | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(&empty[s]);
            if (kc == c.nk - 1) umma_commit_2sm(&tfull[buf]);
          }
          __syncwarp();
          if (++s == c.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (12 warps per CTA: own 128 rows of every row pair x all N columns) ====
    const int quad = warp & 3;
    const int grp = (warp - kT1FirstEpiWarp) >> 2;
    const int chunks_c = c.C >> 4;
    const int per_row = chunks_c > grp ? (chunks_c - grp + kT1EpiGroups - 1) / kT1EpiGroups : 0;
    const int n_mine = c.NTt * per_row;
    constexpr int kMaxMy = 4;
    const size_t plane = (size_t)c.M * 8;
    const size_t t_stride = (size_t)(c.C >> 3) * plane;
    int my_col[kMaxMy], my_ch[kMaxMy];
    size_t my_off[kMaxMy];
#pragma unroll
    for (int i = 0; i < kMaxMy; ++i) {
      const int tl = per_row ? i / per_row : 0, k = per_row ? i - tl * per_row : 0;
      const int cq = grp + kT1EpiGroups * k;
      my_ch[i] = cq * 16;
      my_col[i] = tl * c.C + cq * 16;
      my_off[i] = (size_t)tl * t_stride + (size_t)(cq * 2) * plane;
    }
    const uint32_t tempty_leader = mapa_u32(smem_u32(&tempty[0]), 0);
    for (int lu = 0; lu < n_my; ++lu) {
      const int buf = c.nbuf == 2 ? (lu & 1) : 0;
      const uint32_t use = c.nbuf == 2 ? (uint32_t)(lu >> 1) : (uint32_t)lu;
      const int u = pair + lu * n_pairs;
      const int b = u / p.n_tg, t0 = (u - b * p.n_tg) * c.NTt;
      if (!mbar_wait(&tfull[buf], use & 1, abort_flag)) break;
      tc_fence_after();
      for (int mp = 0; mp < c.n_mp; ++mp) {
        const int m = mp * 256 + (int)rank * 128 + quad * 32 + lane;
        const size_t base = cg8_index(b, t0, 0, m, p.T, c.C, c.M);
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * acc_cols + mp * c.N);
        uint32_t r[kMaxMy][16];
#pragma unroll
        for (int i = 0; i < kMaxMy; ++i)
          if (i < n_mine) tmem_ld16(taddr + my_col[i], r[i]);
        tmem_ld_wait();
        if (mp == c.n_mp - 1) {  // every accumulator of this set is in registers: hand the TMEM set back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(tempty_leader + (uint32_t)buf * 8);
        }
        if (m < c.M) {
#pragma unroll
          for (int i = 0; i < kMaxMy; ++i) {
            if (i < n_mine) {
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int ch = my_ch[i] + 2 * e;
                const float v0 = fmaxf(fmaf(__uint_as_float(r[i][2 * e]), s_scale[ch], s_shift[ch]), 0.f);
                const float v1 = fmaxf(fmaf(__uint_as_float(r[i][2 * e + 1]), s_scale[ch + 1], s_shift[ch + 1]), 0.f);
                pk[e] = pack2<FMT>(v0, v1);
              }
              const size_t idx = base + my_off[i];
              *reinterpret_cast<uint4*>(p.out + idx) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(p.out + idx + plane) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
      }
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
struct TcTdf1PairWeights {
  int fmt;
  T1Cfg cfg;
  h16* d_pack;
  size_t pack_elems;
};

static bool make_t1_cfg(int M, int K, int C, int T, T1Cfg& c) {
  if (C % 16 || C > 256 || M < 128 || M > 512 || K % 32 || K < 256) return false;
  c.C = C; c.M = M; c.K = K;
  c.n_mp = (M + 255) / 256;
  int ntt = 1;
  while ((ntt * 2) * C <= 256 && c.n_mp * (ntt * 2) * C <= 512 && T % (ntt * 2) == 0 && ntt * 2 <= 8) ntt *= 2;
  c.NTt = ntt;
  c.split_t = ntt >= 2;
  if (!c.split_t && C % 32) return false;
  c.N = ntt * C;
  if (c.N % 32 || c.N > 256 || c.n_mp * c.N > 512) return false;
  if (((C / 16 + kT1EpiGroups - 1) / kT1EpiGroups) * ntt > 4) return false;
  c.nbuf = (2 * c.n_mp * c.N <= 512) ? 2 : 1;
  // one stage = n_mp weight blobs + the X half; >= ~1000 cycles of MMA per mbarrier hand-off
  c.Kt = 0;
  for (int kt : {96, 64, 128, 32})
    if (K % kt == 0) { c.Kt = kt; break; }
  if (!c.Kt) return false;
  c.nk = K / c.Kt;
  c.a_blob_bytes = 128 * c.Kt * 2;
  c.b_bytes = (c.N / 2) * c.Kt * 2;
  c.stage_bytes = c.n_mp * c.a_blob_bytes + c.b_bytes;
  c.stages = (227 * 1024 - kT1Header) / c.stage_bytes;
  if (c.stages > kT1MaxStages) c.stages = kT1MaxStages;
  if (c.stages < 3) return false;
  c.smem_bytes = kT1Header + c.stages * c.stage_bytes;
  return true;
}

int tc_tdf1_pair_pack(const float* h_w /*[M][K]*/, int M, int K, int C, int T, int fmt, TcTdf1PairWeights** out) {
  *out = nullptr;
  T1Cfg c;
  if (!make_t1_cfg(M, K, C, T, c)) return AC_OK;
  // [mp][rank][kc][Kt/8][128][8]; rows >= M are zero
  const size_t total = (size_t)c.n_mp * 256 * K;
  std::vector<h16> pack(total, h16_rn(0.f, fmt));
  size_t o = 0;
  for (int mp = 0; mp < c.n_mp; ++mp)
    for (int r = 0; r < 2; ++r)
      for (int kc = 0; kc < c.nk; ++kc)
        for (int kg = 0; kg < c.Kt / 8; ++kg)
          for (int row = 0; row < 128; ++row)
            for (int e = 0; e < 8; ++e, ++o) {
              const int m = mp * 256 + r * 128 + row;
              if (m < M) pack[o] = h16_rn(h_w[(size_t)m * K + kc * c.Kt + kg * 8 + e], fmt);
            }
  TcTdf1PairWeights* w = new TcTdf1PairWeights();
  w->cfg = c;
  w->fmt = fmt;
  w->d_pack = nullptr;
  w->pack_elems = total;
  if (cudaMalloc(&w->d_pack, total * 2) != cudaSuccess ||
      cudaMemcpy(w->d_pack, pack.data(), total * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("tdf1 pair weight upload failed");
    delete w;
    return AC_E_CUDA;
  }
  *out = w;
  return AC_OK;
}

void tc_tdf1_pair_free(TcTdf1PairWeights* w) {
  if (!w) return;
  if (w->d_pack) cudaFree(w->d_pack);
  delete w;
}

int launch_tc_tdf1_pair(const TcTdf1PairWeights* w, const h16* in, h16* out, int nB, int T, const float* scale,
                        const float* shift, cudaStream_t st) {
  AC_REQUIRE(w && in && out, "tc tdf1 pair: null");
  const T1Cfg& c = w->cfg;
  AC_REQUIRE(T % c.NTt == 0, "tc tdf1 pair: T not divisible by the time tile");
  EncodeTiledFn enc = get_tensor_map_encoder();
  AC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available");
  AC_REQUIRE(tc_abort_flag() != nullptr, "abort flag allocation failed");
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap x_map, w_map;
  {
    // X: CG8 [nB][T][C/8][K][8] as (c%8, k, c/8, t, b); one box = this CTA's column half for one K chunk
    const cuuint64_t dims[5] = {8, (cuuint64_t)c.K, (cuuint64_t)(c.C / 8), (cuuint64_t)T, (cuuint64_t)nB};
    const cuuint64_t strides[4] = {16, (cuuint64_t)c.K * 16, (cuuint64_t)c.K * c.C * 2, (cuuint64_t)T * c.K * c.C * 2};
    const cuuint32_t box[5] = {8, (cuuint32_t)c.Kt, (cuuint32_t)(c.split_t ? c.C / 8 : c.C / 16),
                               (cuuint32_t)(c.split_t ? c.NTt / 2 : 1), 1};
    CUresult r = enc(&x_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<h16*>(in), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (tdf1 pair, X) failed with code " + std::to_string((int)r));
      return AC_E_CUDA;
    }
  }
  {
    const cuuint64_t dims[2] = {256, (cuuint64_t)(w->pack_elems / 256)};
    const cuuint64_t strides[1] = {512};
    const cuuint32_t box[2] = {256, (cuuint32_t)(c.a_blob_bytes / 512)};
    CUresult r = enc(&w_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w->d_pack, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (tdf1 pair, W) failed with code " + std::to_string((int)r));
      return AC_E_CUDA;
    }
  }
  T1Params p;
  p.cfg = c;
  p.nB = nB; p.T = T;
  p.n_tg = T / c.NTt;
  p.n_units = p.n_tg * nB;
  p.scale = scale; p.shift = shift;
  p.out = out;
  p.abort_flag = tc_abort_flag();
  static bool attr_set = false;
  if (!attr_set) {
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_tdf1_pair_kernel<kFmtF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_tdf1_pair_kernel<kFmtBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int pairs = device_sm_count() / 2;
  if (pairs > p.n_units) pairs = p.n_units;
  ProfScope ps(KC_TDF_TC, 2.0 * c.M * (double)c.K * c.C * T * nB, 2.0 * nB * (double)T * c.C * (c.K + c.M), st);
  AC_CHECK_CUDA(tc_launch(w->fmt == kFmtBF16 ? tc_tdf1_pair_kernel<kFmtBF16> : tc_tdf1_pair_kernel<kFmtF16>, 2 * pairs, kT1Threads, c.smem_bytes, st, 2, x_map, w_map, p));
  AC_LAUNCH_CHECK();
  return AC_OK;
}

}  // namespace ac
