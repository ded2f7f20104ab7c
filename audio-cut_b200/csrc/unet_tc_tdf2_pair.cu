// CTA-pair (cta_group::2) tcgen05 kernel for the second TDF layer of a block (F/8 -> F features, plus
// the residual):   Y[b][t][m][c] = relu(scale[c] * sum_k W[m][k] * H[b][t][k][c] + shift[c]) + X[b][t][m][c]
//
// The single-CTA kernel (unet_tc_tdf.cu) re-fetches the activation tile H for every 128-row slice
// of the wide output and was L2 -> shared-memory ingest bound (ncu: half of all stall samples are the
// epilogue waiting for accumulators).  Here the activations are stationary and the weights stream:
//   * a work unit is one (b, NTt time rows) tile = N = NTt*C <= 256 columns; each CTA of the pair keeps
//     ITS half of the columns of H resident ([N/2/8 groups][K][8], MN-major B operand, double buffered
//     across units) and the pair walks all M/256 output row pairs of the layer over it;
//   * per row pair every CTA streams only its own 128 rows of W (pre-packed K-major smem images,
//     [Kt/8][128][8] per stage) through a ring; M = 256, N <= 256 MMAs are issued by the leader CTA;
//   * so per 128 x N output tile a CTA ingests 128*K*2 B of weights + 1/(M/256) of a half H tile,
//     2.4x less than before, and the MMA shape is tensor-pipe bound instead of smem bound;
//   * three producer warps take ring stages round-robin (a single issuing thread tops out near
//     18 B/clk, scripts/microbench/tma_box.cu), a fourth loads H, warp 4 issues MMAs, 12 epilogue
//     warps (TMEM lane quadrant x channel-chunk group) prefetch the residual, apply BN/ReLU, add, store.
#include <vector>

#include "tc_common.cuh"
#include "unet_kernels.cuh"

namespace ac {

constexpr int t2_producers(bool) { return 2; }  // weight producer warps (round-robin over the ring stages)
constexpr int kT2EpiGroups = 3;
// warps: [0, P) weight producers, P = H producer, P + 1 = MMA issuer, P + 2 ... epilogue
// FINAL variant (the network's last TDF2, C = 48, 4 time rows per unit): 4 epilogue groups, one per time row, so
// that a thread sees all 48 channels of its (t, f) position and can apply the final 1x1 convolution itself.
constexpr int t2_groups(bool fin) { return fin ? 4 : kT2EpiGroups; }
constexpr int t2_epi_warps(bool fin) { return 4 * t2_groups(fin); }
constexpr int t2_threads(bool fin) { return (t2_producers(fin) + 2 + t2_epi_warps(fin)) * 32; }
constexpr int kT2Header = 4096;
constexpr int kT2MaxStages = 12;

struct T2Cfg {
  int C, M, K;
  int NTt, N;       // time rows per unit, N = NTt * C columns (pair); each CTA holds N/2
  int split_t;      // 1: the CTAs split the unit by time rows (NTt even); 0: by channel halves (NTt == 1)
  int Kt, nk;       // K chunk per ring stage
  int Kb, nkb;      // K extent of one H box (<= 256) and boxes per H tile
  int n_mp;         // M / 256
  int stages;
  int n_hbuf;         // H buffers (2 when they leave room for a deep enough weight ring)
  int a_stage_bytes;  // 128 * Kt * 2
  int h_bytes;        // (N/2) * K * 2
  int smem_bytes;
};

struct T2Params {
  T2Cfg cfg;
  int nB, T;
  int n_tg, n_units;
  const float* scale;
  const float* shift;
  const h16* residual;
  h16* out;
  int* abort_flag;
  // FINAL variant only: out4[(b*T + t)*M + m][0..3] = bias[o] + sum_c final_w[o][c] * bf16(Y[b][t][m][c]); Y is not stored
  const float* final_w;
  const float* final_b;
  h16* out4;
};

__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void t2_umma_2sm(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  const uint64_t ad = ((uint64_t)a_hi << 32) | a_lo, bd = ((uint64_t)b_hi << 32) | b_lo;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(ad), "l"(bd), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool FINAL, int FMT>
__global__ void __launch_bounds__(t2_threads(FINAL), 1)
tc_tdf2_pair_kernel(const __grid_constant__ CUtensorMap h_map, const __grid_constant__ CUtensorMap w_map, const T2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  const T2Cfg& c = p.cfg;
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem);  // [stages]  leader: both CTAs' A stage landed
  uint64_t* aempty = afull + kT2MaxStages;              // [stages]
  uint64_t* hfull = aempty + kT2MaxStages;              // [2]       leader: both CTAs' H half landed
  uint64_t* hempty = hfull + 2;                         // [2]
  uint64_t* tfull = hempty + 2;                         // [2]
  uint64_t* tempty = tfull + 2;                         // [2]       leader: both CTAs' epilogues drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  // per channel pair {scale[c], shift[c], scale[c+1], shift[c+1]}: one 128-bit shared load per two outputs
  float4* s_ss = reinterpret_cast<float4*>(smem + 1024);   // [C/2] (C <= 256: 2 KB)
  float4* s_fw = s_ss + 128;  // FINAL: [C] x {w0..w3 of that input channel} + the 4 biases (C = 48: 784 B, below kT2Header)
  constexpr int kEpiWarps = t2_epi_warps(FINAL);
  constexpr int kT2AProducers = t2_producers(FINAL);
  constexpr int kT2FirstEpiWarp = kT2AProducers + 2;
  uint8_t* h_smem = smem + kT2Header;                   // 2 x h_bytes
  uint8_t* ring = h_smem + c.n_hbuf * c.h_bytes;
  volatile int* abort_flag = p.abort_flag;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int n_my = pair < p.n_units ? (p.n_units - pair + n_pairs - 1) / n_pairs : 0;  // units of this pair

  if (threadIdx.x == 0) {
    for (int s = 0; s < c.stages; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&aempty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&hfull[b], 1);
      mbar_init(&hempty[b], 1);
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 2 * kEpiWarps);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < c.C; i += blockDim.x) {
    reinterpret_cast<float*>(s_ss)[(i >> 1) * 4 + (i & 1) * 2] = p.scale[i];
    reinterpret_cast<float*>(s_ss)[(i >> 1) * 4 + (i & 1) * 2 + 1] = p.shift[i];
  }
  if (FINAL)
    for (int i = threadIdx.x; i < 4 * c.C + 4; i += blockDim.x) {  // final_w is [4][C]
      const int o = i / c.C, ch = i - o * c.C;
      reinterpret_cast<float*>(s_fw)[i < 4 * c.C ? ch * 4 + o : i] = i < 4 * c.C ? p.final_w[i] : p.final_b[i - 4 * c.C];
    }
  if (warp == kT2AProducers + 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its output is visible from here on
  const int stages_per_unit = c.n_mp * c.nk;

  if (warp < kT2AProducers) {
    // ===================== weight producer (one 48-KB stage per mbarrier phase) =====================
    if (lane == 0) {
      const long long total = (long long)n_my * stages_per_unit;
      for (long long i = warp; i < total; i += kT2AProducers) {
        const int s = (int)(i % c.stages);
        const uint32_t ph = (uint32_t)((i / c.stages) & 1);
        if (!mbar_wait(&aempty[s], ph ^ 1, abort_flag)) break;
        const int within = (int)(i % stages_per_unit);  // (mp, kc) - the weight stream repeats for every unit
        const int mp = within / c.nk, kc = within - mp * c.nk;
        if (leader) mbar_expect_tx(&afull[s], 2u * (uint32_t)c.a_stage_bytes);
        // packed weights: blob index ((mp*2 + rank)*nk + kc), each blob = a_stage_bytes = rows of 512 B
        const int blob = (mp * 2 + (int)rank) * c.nk + kc;
        tma_load_2d_2sm(ring + (size_t)s * c.a_stage_bytes, &w_map, mapa_u32(smem_u32(&afull[s]), 0), 0,
                        blob * (c.a_stage_bytes / 512));
      }
    }
  } else if (warp == kT2AProducers) {
    // ===================== activation producer: this CTA's half of the unit's H tile =====================
    if (lane == 0) {
      for (int lu = 0; lu < n_my; ++lu) {
        const int hb = c.n_hbuf == 2 ? (lu & 1) : 0;
        const uint32_t use = c.n_hbuf == 2 ? (uint32_t)(lu >> 1) : (uint32_t)lu;  // how often this buffer was filled before
        if (!mbar_wait(&hempty[hb], (use & 1) ^ 1, abort_flag)) break;
        const int u = pair + lu * n_pairs;
        const int b = u / p.n_tg, t0 = (u - b * p.n_tg) * c.NTt;
        if (leader) mbar_expect_tx(&hfull[hb], 2u * (uint32_t)c.h_bytes);
        const uint32_t bar = mapa_u32(smem_u32(&hfull[hb]), 0);
        uint8_t* dst = h_smem + (size_t)hb * c.h_bytes;
        const int box_bytes = c.h_bytes / c.nkb;
        for (int kb = 0; kb < c.nkb; ++kb) {
          if (c.split_t)
            tma_load_5d_2sm(dst + (size_t)kb * box_bytes, &h_map, bar, 0, kb * c.Kb, 0, t0 + (int)rank * (c.NTt / 2), b);
          else
            tma_load_5d_2sm(dst + (size_t)kb * box_bytes, &h_map, bar, 0, kb * c.Kb, (int)rank * (c.C / 16), t0, b);
        }
      }
    }
  } else if (warp == kT2AProducers + 1) {
    // ===================== MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) ======
    if (leader) {
      auto wait_all = [&](uint64_t* bar, uint32_t parity) {
        return __all_sync(0xffffffffu, mbar_wait(bar, parity, abort_flag)) != 0;
      };
      // ncu (profiles/r02_tdf2_issue.md): this warp never waited - 34 SASS instructions per MMA (64-bit ring-stage division
      // and modulo, a division per K step for the H box, 64-bit descriptor sums) = 206 cycles per 96-cycle MMA, tensor pipe
      // 38 %, the epilogue idle a third of its time.  Ring stage, phase and the H-box position are now carried incrementally
      // and the descriptors are 32-bit words: two adds and the instruction per MMA.
      const uint32_t idesc = make_idesc_2sm<FMT>(c.N) | (1u << 16);  // B is MN-major
      const uint64_t a_proto = make_desc(0, 128 * 16, 128);
      const uint64_t b_proto = make_desc_mn(0, 128, (uint32_t)c.Kb * 16);
      const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
      const uint32_t a_lo0 = (uint32_t)a_proto + (smem_u32(ring) >> 4);
      const uint32_t stage16 = (uint32_t)c.a_stage_bytes >> 4;
      const uint32_t hbox16 = (uint32_t)(c.h_bytes / c.nkb) >> 4;
      const int ksteps = c.Kt / 16;
      int s = 0;          // ring stage
      uint32_t ph = 0;
      uint32_t acc_n = 0;  // accumulator tiles issued
      bool alive = true;
      for (int lu = 0; lu < n_my && alive; ++lu) {
        const int hb = c.n_hbuf == 2 ? (lu & 1) : 0;
        const uint32_t use = c.n_hbuf == 2 ? (uint32_t)(lu >> 1) : (uint32_t)lu;
        if (!wait_all(&hfull[hb], use & 1)) break;
        const uint32_t b_lo0 = (uint32_t)b_proto + (smem_u32(h_smem + (size_t)hb * c.h_bytes) >> 4);
        for (int mp = 0; mp < c.n_mp && alive; ++mp, ++acc_n) {
          const int buf = acc_n & 1;
          if (!wait_all(&tempty[buf], ((acc_n >> 1) & 1) ^ 1)) { alive = false; break; }
          const uint32_t acc = tmem_base + (uint32_t)(buf * c.N);
          // position inside the H tile: box kb (hbox16 apart), K offset kin inside it (16 bytes per K row in the MN-major box)
          uint32_t box_lo = b_lo0;
          int kin = 0;
          for (int kc = 0; kc < c.nk; ++kc) {
            if (!wait_all(&afull[s], ph)) { alive = false; break; }
            tc_fence_after();
            const bool el = elect_one();
            uint32_t a_lo = a_lo0 + (uint32_t)s * stage16;
#pragma unroll 4
            for (int k = 0; k < ksteps; ++k) {
              if (el) t2_umma_2sm(acc, a_lo, a_hi, box_lo + (uint32_t)kin, b_hi, idesc, (kc | k) != 0 ? 1u : 0u);
              a_lo += (2 * 128 * 16) >> 4;
              kin += 16;  // 16 K rows x 16 bytes >> 4
              if (kin == c.Kb) {
                kin = 0;
                box_lo += hbox16;
              }
            }
            if (el) {
              umma_commit_2sm(&aempty[s]);
              if (kc == c.nk - 1) {
                umma_commit_2sm(&tfull[buf]);
                if (mp == c.n_mp - 1) umma_commit_2sm(&hempty[hb]);  // the unit's H halves are free in both CTAs
              }
            }
            __syncwarp();
            if (++s == c.stages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else {
    // ===================== epilogue (12 warps per CTA: own 128 rows x all N columns) =====================
    const int quad = warp & 3;  // hardware rule: warp w may read TMEM lanes 32*(w%4) .. +31
    const int grp = (warp - kT2FirstEpiWarp) >> 2;
    // Tiles of this pair in order (unit lu, output row pair mp).  Index arithmetic is incremental (no divisions on the
    // per-tile path) and the residual registers of a chunk are re-requested for the NEXT tile right after they have
    // been consumed, one tile period ahead of their use: one register set instead of two, no spills at 128 registers
    // (the spilled version spent ~25 % of its stall samples on local-memory reloads in the epilogue's serial chain).
    const uint32_t tempty_leader = mapa_u32(smem_u32(&tempty[0]), 0);
    const size_t plane = (size_t)c.M * 8;
    const int m_in = (int)rank * 128 + quad * 32 + lane;
    auto unit_base = [&](int lu, int& b, int& t0) -> size_t {  // once per unit
      const int u = pair + lu * n_pairs;
      b = u / p.n_tg;
      t0 = (u - b * p.n_tg) * c.NTt;
      return cg8_index(b, t0, 0, m_in, p.T, c.C, c.M);
    };
    const long long n_tiles = (long long)n_my * c.n_mp;
    int lu = 0, mp = 0, ub = 0, ut0 = 0;
    size_t base = n_tiles ? unit_base(0, ub, ut0) : 0;
    if constexpr (FINAL) {
      // ---- last layer of the network: 16 warps, group = time row; residual add, then the 1x1 conv to 4 channels ----
      const int tl = grp;  // NTt == 4
      const size_t t_off = (size_t)tl * (size_t)(c.C >> 3) * plane;
      uint4 q[6];
      if (n_tiles) {
#pragma unroll
        for (int k = 0; k < 6; ++k) q[k] = ldg_stream_u4(p.residual + base + t_off + (size_t)k * plane);
      }
      for (long long tile = 0; tile < n_tiles; ++tile) {
        const int buf = (int)(tile & 1);
        const bool has_next = tile + 1 < n_tiles;
        const size_t pos = ((size_t)ub * p.T + ut0 + tl) * c.M + mp * 256 + m_in;
        size_t next_base = base + 256 * 8;
        if (++mp == c.n_mp) {
          mp = 0;
          ++lu;
          if (has_next) next_base = unit_base(lu, ub, ut0);
        }
        if (!mbar_wait(&tfull[buf], (uint32_t)((tile >> 1) & 1), abort_flag)) break;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * c.N + tl * 48);
        uint32_t r[2][16];
        tmem_ld16(taddr, r[0]);
        tmem_ld16(taddr + 16, r[1]);
        tmem_ld_wait();
        const float4 bias = s_fw[48];
        float o0 = bias.x, o1 = bias.y, o2 = bias.z, o3 = bias.w;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (k == 2) {  // third chunk was requested into r[0] while the second was processed
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(tempty_leader + (uint32_t)buf * 8);
          }
          const uint4 q0 = q[2 * k], q1 = q[2 * k + 1];
          if (has_next) {  // the residual of the next tile: a full tile period ahead of its use
            q[2 * k] = ldg_stream_u4(p.residual + next_base + t_off + (size_t)(2 * k) * plane);
            q[2 * k + 1] = ldg_stream_u4(p.residual + next_base + t_off + (size_t)(2 * k + 1) * plane);
          }
          const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
          float y[16];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 ss = s_ss[k * 8 + e];
            const float2 res = unpack2<FMT>(w[e]);
            const float v0 = fmaxf(fmaf(__uint_as_float(r[k & 1][2 * e]), ss.x, ss.y), 0.f) + res.x;
            const float v1 = fmaxf(fmaf(__uint_as_float(r[k & 1][2 * e + 1]), ss.z, ss.w), 0.f) + res.y;
            // rounded to bf16 exactly like the stored activation the separate 1x1 kernel would read
            const float2 yy = unpack2<FMT>(pack2<FMT>(v0, v1));
            y[2 * e] = yy.x;
            y[2 * e + 1] = yy.y;
          }
          if (k == 0) tmem_ld16(taddr + 32, r[0]);  // r[0] is consumed: fetch the third chunk behind the second's math
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float4 fw = s_fw[k * 16 + e];
            o0 = fmaf(fw.x, y[e], o0);
            o1 = fmaf(fw.y, y[e], o1);
            o2 = fmaf(fw.z, y[e], o2);
            o3 = fmaf(fw.w, y[e], o3);
          }
        }
        uint2 ov;
        ov.x = pack2<FMT>(o0, o1);
        ov.y = pack2<FMT>(o2, o3);
        *reinterpret_cast<uint2*>(p.out4 + pos * 4) = ov;
        base = next_base;
      }
    } else {
      const int chunks_c = c.C >> 4;
      const int per_row = chunks_c > grp ? (chunks_c - grp + kT2EpiGroups - 1) / kT2EpiGroups : 0;
      const int n_mine = c.NTt * per_row;  // 16-column chunks this warp owns per accumulator tile (<= 4)
      constexpr int kMaxMy = 4;
      const size_t t_stride = (size_t)(c.C >> 3) * plane;
      // chunk i -> (time row tl, k-th chunk of this group in the row); i is a compile-time constant after unrolling
      auto chunk_tl = [&](int i) { return (int)(i >= per_row) + (int)(i >= 2 * per_row) + (int)(i >= 3 * per_row); };
      auto chunk_cq = [&](int i) { return grp + kT2EpiGroups * (i - chunk_tl(i) * per_row); };
      auto chunk_off = [&](int i) { return (size_t)chunk_tl(i) * t_stride + (size_t)(chunk_cq(i) * 2) * plane; };
      // the residual of tile n + 1 is requested at the start of tile n (before the accumulator wait); accumulators are
      // read one 16-column chunk at a time so that {16 accumulators, 2 x residual sets} stay within 128 registers
      auto fetch = [&](size_t from, uint4* q) {
#pragma unroll
        for (int i = 0; i < kMaxMy; ++i)
          if (i < n_mine) {
            q[2 * i] = ldg_stream_u4(p.residual + from + chunk_off(i));
            q[2 * i + 1] = ldg_stream_u4(p.residual + from + chunk_off(i) + plane);
          }
      };
      uint4 q_cur[2 * kMaxMy], q_next[2 * kMaxMy];
      if (n_tiles) fetch(base, q_cur);
      for (long long tile = 0; tile < n_tiles; ++tile) {
        const int buf = (int)(tile & 1);
        const bool has_next = tile + 1 < n_tiles;
        size_t next_base = base + 256 * 8;
        if (++mp == c.n_mp) {
          mp = 0;
          ++lu;
          if (has_next) next_base = unit_base(lu, ub, ut0);
        }
        if (has_next) fetch(next_base, q_next);
        if (!mbar_wait(&tfull[buf], (uint32_t)((tile >> 1) & 1), abort_flag)) break;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * c.N);
#pragma unroll
        for (int i = 0; i < kMaxMy; ++i) {
          if (i < n_mine) {
            uint32_t r[16];
            tmem_ld16(taddr + chunk_tl(i) * c.C + chunk_cq(i) * 16, r);
            tmem_ld_wait();
            if (i == n_mine - 1) {  // every accumulator column of this warp is in registers: hand the TMEM buffer back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster_relaxed(tempty_leader + (uint32_t)buf * 8);
            }
            const size_t off = chunk_off(i);
            const uint4 q0 = q_cur[2 * i], q1 = q_cur[2 * i + 1];
            const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            const int ch0 = chunk_cq(i) * 16;
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 ss = s_ss[(ch0 >> 1) + e];
              const float2 res = unpack2<FMT>(w[e]);
              const float v0 = fmaxf(fmaf(__uint_as_float(r[2 * e]), ss.x, ss.y), 0.f) + res.x;
              const float v1 = fmaxf(fmaf(__uint_as_float(r[2 * e + 1]), ss.z, ss.w), 0.f) + res.y;
              pk[e] = pack2<FMT>(v0, v1);
            }
            const size_t idx = base + off;
            *reinterpret_cast<uint4*>(p.out + idx) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(p.out + idx + plane) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
#pragma unroll
        for (int i = 0; i < 2 * kMaxMy; ++i) q_cur[i] = q_next[i];
        base = next_base;
      }
    }  // !FINAL
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == kT2AProducers + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TcTdf2PairWeights {
  int fmt;
  T2Cfg cfg;
  h16* d_pack;
  size_t pack_elems;
};

static bool make_t2_cfg(int M, int K, int C, int T, T2Cfg& c) {
  if (C % 16 || C > 256 || M % 256 || K % 32 || K > 512) return false;
  c.C = C; c.M = M; c.K = K;
  // columns per unit: as many time rows as fit in N <= 256; the pair splits an even NTt by time rows,
  // NTt == 1 by channel halves (then C/2 must be a multiple of 8)
  int ntt = 1;
  while ((ntt * 2) * C <= 256 && T % (ntt * 2) == 0 && ntt * 2 <= 8) ntt *= 2;
  c.NTt = ntt;
  c.split_t = ntt >= 2;
  if (!c.split_t && C % 32) return false;
  c.N = ntt * C;
  if (c.N % 32 || c.N > 256) return false;
  // the 12 epilogue warps own at most 4 chunks of 16 columns each (registers: 4 x 16 accumulators + residual)
  if (((C / 16 + kT2EpiGroups - 1) / kT2EpiGroups) * ntt > 4) return false;
  c.Kb = K <= 256 ? K : K / 2;
  if (c.Kb > 256 || c.Kb % 16) return false;
  c.nkb = K / c.Kb;
  // One ring stage = one H box worth of K (<= 48 KB of weights): every mbarrier hand-off costs the
  // issuing thread several hundred cycles, so a stage has to carry >= ~1000 cycles of MMA work.
  c.Kt = c.Kb;
  while (128 * c.Kt * 2 > 48 * 1024 && c.Kt % 32 == 0) c.Kt /= 2;
  if (128 * c.Kt * 2 > 48 * 1024 || c.Kb % c.Kt) return false;
  c.nk = K / c.Kt;
  c.n_mp = M / 256;
  c.a_stage_bytes = 128 * c.Kt * 2;
  c.h_bytes = (c.N / 2) * K * 2;
  c.n_hbuf = 2;
  int budget = 227 * 1024 - kT2Header - 2 * c.h_bytes;
  if (budget / c.a_stage_bytes < 3) {  // a single H buffer (a short bubble per unit) buys a deeper weight ring
    c.n_hbuf = 1;
    budget = 227 * 1024 - kT2Header - c.h_bytes;
  }
  c.stages = budget / c.a_stage_bytes;
  if (c.stages > kT2MaxStages) c.stages = kT2MaxStages;
  if (c.stages < 3) return false;
  c.smem_bytes = kT2Header + c.n_hbuf * c.h_bytes + c.stages * c.a_stage_bytes;
  return true;
}

int tc_tdf2_pair_pack(const float* h_w /*[M][K]*/, int M, int K, int C, int T, int fmt, TcTdf2PairWeights** out) {
  *out = nullptr;
  T2Cfg c;
  if (!make_t2_cfg(M, K, C, T, c)) return AC_OK;
  // [mp][rank][kc][Kt/8][128][8]
  const size_t total = (size_t)M * K;
  std::vector<h16> pack(total);
  size_t o = 0;
  for (int mp = 0; mp < c.n_mp; ++mp)
    for (int r = 0; r < 2; ++r)
      for (int kc = 0; kc < c.nk; ++kc)
        for (int kg = 0; kg < c.Kt / 8; ++kg)
          for (int row = 0; row < 128; ++row)
            for (int e = 0; e < 8; ++e)
              pack[o++] = h16_rn(h_w[(size_t)(mp * 256 + r * 128 + row) * K + kc * c.Kt + kg * 8 + e], fmt);
  TcTdf2PairWeights* w = new TcTdf2PairWeights();
  w->cfg = c;
  w->fmt = fmt;
  w->d_pack = nullptr;
  w->pack_elems = total;
  if (cudaMalloc(&w->d_pack, total * 2) != cudaSuccess ||
      cudaMemcpy(w->d_pack, pack.data(), total * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("tdf2 pair weight upload failed");
    delete w;
    return AC_E_CUDA;
  }
  *out = w;
  return AC_OK;
}

void tc_tdf2_pair_free(TcTdf2PairWeights* w) {
  if (!w) return;
  if (w->d_pack) cudaFree(w->d_pack);
  delete w;
}

bool tc_tdf2_pair_can_fuse_final(const TcTdf2PairWeights* w) { return w && w->cfg.C == 48 && w->cfg.NTt == 4; }

// final_w != nullptr: the FINAL variant - `out` is then the network output [nB*T*M][4] bf16 and the layer's own
// activation is never written (unet.cu uses it for the last TDF2 of the network).
int launch_tc_tdf2_pair(const TcTdf2PairWeights* w, const h16* in, const h16* residual, h16* out,
                        int nB, int T, const float* scale, const float* shift, cudaStream_t st, const float* final_w,
                        const float* final_b) {
  AC_REQUIRE(w && in && residual && out, "tc tdf2 pair: null");
  const T2Cfg& c = w->cfg;
  const bool fin = final_w != nullptr;
  AC_REQUIRE(!fin || (final_b && tc_tdf2_pair_can_fuse_final(w)), "tc tdf2 pair: final 1x1 fusion needs C == 48, 4 time rows per unit");
  AC_REQUIRE(T % c.NTt == 0, "tc tdf2 pair: T not divisible by the time tile");
  EncodeTiledFn enc = get_tensor_map_encoder();
  AC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available");
  AC_REQUIRE(tc_abort_flag() != nullptr, "abort flag allocation failed");
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  // H: CG8 [nB][T][C/8][K][8] as (c%8, k, c/8, t, b); one box = this CTA's column half for one K box
  CUtensorMap h_map, w_map;
  {
    const cuuint64_t dims[5] = {8, (cuuint64_t)c.K, (cuuint64_t)(c.C / 8), (cuuint64_t)T, (cuuint64_t)nB};
    const cuuint64_t strides[4] = {16, (cuuint64_t)c.K * 16, (cuuint64_t)c.K * c.C * 2, (cuuint64_t)T * c.K * c.C * 2};
    const cuuint32_t box[5] = {8, (cuuint32_t)c.Kb, (cuuint32_t)(c.split_t ? c.C / 8 : c.C / 16),
                               (cuuint32_t)(c.split_t ? c.NTt / 2 : 1), 1};
    CUresult r = enc(&h_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<h16*>(in), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (tdf2 pair, H) failed with code " + std::to_string((int)r));
      return AC_E_CUDA;
    }
  }
  {
    // packed weights as rows of 256 bf16 (512 B); one ring stage = a_stage_bytes/512 consecutive rows
    const cuuint64_t dims[2] = {256, (cuuint64_t)(w->pack_elems / 256)};
    const cuuint64_t strides[1] = {512};
    const cuuint32_t box[2] = {256, (cuuint32_t)(c.a_stage_bytes / 512)};
    CUresult r = enc(&w_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w->d_pack, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (tdf2 pair, W) failed with code " + std::to_string((int)r));
      return AC_E_CUDA;
    }
  }
  T2Params p;
  p.cfg = c;
  p.nB = nB; p.T = T;
  p.n_tg = T / c.NTt;
  p.n_units = p.n_tg * nB;
  p.scale = scale; p.shift = shift;
  p.residual = residual;
  p.out = fin ? nullptr : out;
  p.out4 = fin ? out : nullptr;
  p.final_w = final_w;
  p.final_b = final_b;
  p.abort_flag = tc_abort_flag();
  static bool attr_set = false;
  if (!attr_set) {
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_tdf2_pair_kernel<false, kFmtF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_tdf2_pair_kernel<true, kFmtF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_tdf2_pair_kernel<false, kFmtBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_tdf2_pair_kernel<true, kFmtBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int pairs = device_sm_count() / 2;
  if (pairs > p.n_units) pairs = p.n_units;
  ProfScope ps(KC_TDF_TC, 2.0 * c.M * (double)c.K * c.C * T * nB + (fin ? 8.0 * c.M * c.C * T * nB : 0.0),
               2.0 * nB * (double)T * (c.C * (c.K + c.M * (fin ? 1 : 2)) + (fin ? 4 * c.M : 0)), st);
  if (fin)
    AC_CHECK_CUDA(tc_launch(w->fmt == kFmtBF16 ? tc_tdf2_pair_kernel<true, kFmtBF16> : tc_tdf2_pair_kernel<true, kFmtF16>, 2 * pairs, t2_threads(true), c.smem_bytes, st, 2, h_map, w_map, p));
  else
    AC_CHECK_CUDA(tc_launch(w->fmt == kFmtBF16 ? tc_tdf2_pair_kernel<false, kFmtBF16> : tc_tdf2_pair_kernel<false, kFmtF16>, 2 * pairs, t2_threads(false), c.smem_bytes, st, 2, h_map, w_map, p));
  AC_LAUNCH_CHECK();
  return AC_OK;
}

}  // namespace ac
