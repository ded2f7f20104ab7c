// Weight-stationary tcgen05 3x3 convolution for the two widest U-Net levels (C = 48 and C = 96,
// 68 % of the conv FLOPs and the levels where the convolution sits at the HBM/tensor knee).
//
// The streaming kernel in unet_tc.cu re-fetches every input row three times (once per vertical tap)
// and the whole weight tensor once per work unit; at C = 48 that makes it L2->smem bound (ncu: tensor
// pipe 18 % active).  Here instead
//   * the CTA's slice of the weights ([9 taps][C][NT] bf16, 41 KB at C = 48, 83 KB at C = 96/NT = 48)
//     is loaded ONCE and stays in shared memory;
//   * the CTA walks DOWN a strip of MT x 128 positions: a ring of R input rows lives in shared
//     memory, each row is fetched from global memory exactly once (one 3-D TMA box per 128-position
//     tile of the CG8 tensor, [C/8][130][8] = canonical no-swizzle K-major, 2080-byte contiguous
//     global rows) and feeds the 9 taps of three output rows;
//   * rows are handed out as one contiguous range per CTA of the linearised (batch, strip, t)
//     space, so the load is balanced to +-1 row and only 2 halo rows per CTA are read twice;
//   * accumulators (MT x NT columns) are double buffered in TMEM; 12 epilogue warps (one per TMEM lane
//     quadrant and 16-channel group) apply the folded BatchNorm + ReLU and store bf16 while the next
//     row's MMAs run - with a single warp per scheduler the epilogue's dependent ALU chains, not the
//     tensor pipe, set the pace (ncu: 75 % of the stall samples sat in the epilogue).
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer / TMEM owner, warps 2..13 = epilogue.
#include <stdlib.h>

#include <vector>

#include "tc_common.cuh"
#include "unet_kernels.cuh"

namespace ac {

constexpr int kWsEpiGroups = 3;                          // column groups of 16 channels, one epilogue warp per (lane quadrant, group)
constexpr int kWsEpiWarps = 4 * kWsEpiGroups;
constexpr int kWsThreads = (2 + kWsEpiWarps) * 32;       // producer + MMA issuer + 12 epilogue warps
constexpr int kWsTileM = 128;
constexpr int kWsRowPos = kWsTileM + 2;
constexpr int kWsMaxR = 8;

struct WsCfg {
  int C, NT, nsplit, MT, R;
  int a_tile_bytes;  // (C/8) * 130 * 16, rounded up to 128
  int a_lbo;         // 130 * 16
  int slot_bytes;    // MT * a_tile_bytes
  int w_bytes;       // 9 * C * NT * 2
  int smem_bytes;
};

struct WsParams {
  WsCfg cfg;
  int nB, T, F;
  int n_strips;          // strips of MT*128 positions per image row
  long long total_rows;  // nB * n_strips * T
  const h16* wpack;
  const float* scale;
  const float* shift;
  h16* out;
  int* abort_flag;
};

struct WsSeg {
  int b, f0, t0, t1;
};

// CTA -> contiguous range [lo, hi) of rows in the linearised (b, strip, t) space
__device__ __forceinline__ void ws_range(const WsParams& p, int group_size, int j, long long& lo, long long& hi) {
  lo = p.total_rows * j / group_size;
  hi = p.total_rows * (j + 1) / group_size;
}
__device__ __forceinline__ WsSeg ws_segment(const WsParams& p, long long L, long long hi) {
  WsSeg s;
  const long long strip = L / p.T;
  s.t0 = (int)(L - strip * p.T);
  const long long left = hi - L;
  s.t1 = (left < (long long)(p.T - s.t0)) ? s.t0 + (int)left : p.T;
  s.b = (int)(strip / p.n_strips);
  s.f0 = (int)(strip - (long long)s.b * p.n_strips) * p.cfg.MT * kWsTileM;
  return s;
}

// descriptor helpers for the issue loop: the 64-bit smem descriptor only changes in its 14-bit start
// field, so a tile's descriptor is (constant high word, low word + byte_offset/16)
__device__ __forceinline__ uint64_t desc_at(uint32_t lo_base, uint32_t hi, uint32_t byte_off) {
  return ((uint64_t)hi << 32) | (uint64_t)(lo_base + (byte_off >> 4));
}
template <bool kAcc>
__device__ __forceinline__ void umma_f16_c(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  if constexpr (kAcc)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, 1, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc)
        : "memory");
}

template <int C, int MT, int R, int FMT>
__global__ void __launch_bounds__(kWsThreads, 1)
tc_conv3x3_ws_kernel(const __grid_constant__ CUtensorMap in_map, const WsParams p) {
  constexpr int NT = 48;
  constexpr int kALbo = kWsRowPos * 16;
  constexpr int kATile = ((C / 8) * kALbo + 127) / 128 * 128;
  constexpr int kSlot = MT * kATile;
  constexpr int kTap = C * NT * 2;
  constexpr int kWBytes = 9 * kTap;
  constexpr int kBLbo = NT * 16;
  constexpr int K16 = C / 16;
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [R]
  uint64_t* empty = full + kWsMaxR;                     // [R]
  uint64_t* tfull = full + 2 * kWsMaxR;                 // [2]
  uint64_t* tempty = tfull + 2;                         // [2]
  uint64_t* wbar = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  float* s_scale = reinterpret_cast<float*>(smem + 256);  // [NT]
  float* s_shift = s_scale + 96;
  uint8_t* w_smem = smem + 1024;
  uint8_t* ring = w_smem + kWBytes;
  volatile int* abort_flag = p.abort_flag;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group_size = gridDim.x / (C / NT);
  const int nt = blockIdx.x / group_size;
  const int j_in_group = blockIdx.x - nt * group_size;
  long long lo, hi;
  ws_range(p, group_size, j_in_group, lo, hi);

  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], kWsEpiWarps);
    }
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < NT; i += blockDim.x) {
    s_scale[i] = p.scale[nt * NT + i];
    s_shift[i] = p.shift[nt * NT + i];
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(wbar, (uint32_t)kWBytes);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)nt * kWBytes;
      for (int tap = 0; tap < 9; ++tap) bulk_load_1d(w_smem + tap * kTap, wsrc + tap * kTap, kTap, wbar);
      int s = 0;
      uint32_t ph = 0;
      bool alive = true;
      for (long long L = lo; L < hi && alive;) {
        const WsSeg sg = ws_segment(p, L, hi);
        for (int r = sg.t0 - 1; r <= sg.t1; ++r) {
          if (!mbar_wait(&empty[s], ph ^ 1, abort_flag)) { alive = false; break; }
          uint8_t* dst = ring + (size_t)s * kSlot;
          mbar_expect_tx(&full[s], (uint32_t)(MT * (C / 8) * kWsRowPos * 16));
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            tma_load_5d(dst + mt * kATile, &in_map, &full[s], 0, sg.f0 + mt * kWsTileM - 1, 0, r, sg.b);
          if (++s == R) { s = 0; ph ^= 1; }
        }
        L += sg.t1 - sg.t0;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the (uniform) loop and one elected lane issues, so descriptors and TMEM
    // addresses stay in uniform registers instead of being broadcast (R2UR) before every MMA.
    {
      const uint32_t idesc = make_idesc<FMT>(NT);
      const uint64_t a_proto = make_desc(0, kALbo, 128), b_proto = make_desc(0, kBLbo, 128);
      const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
      const uint32_t a_lo0 = (uint32_t)a_proto + (smem_u32(ring) >> 4);
      const uint32_t b_lo0 = (uint32_t)b_proto + (smem_u32(w_smem) >> 4);
      auto wait_all = [&](uint64_t* bar, uint32_t parity) {
        return __all_sync(0xffffffffu, mbar_wait(bar, parity, abort_flag)) != 0;
      };
      bool alive = wait_all(wbar, 0);
      // ring bookkeeping without divisions: s0 = slot of the oldest row of the current output row
      int s0 = 0;
      uint32_t ph0 = 0;  // parity of slot s0's current fill
      uint32_t orow = 0;
      auto slot_after = [&](int s, int d, uint32_t ph, uint32_t& ph_out) {
        int q = s + d;
        ph_out = ph;
        if (q >= R) { q -= R; ph_out ^= 1; }
        return q;
      };
      for (long long L = lo; L < hi && alive;) {
        const WsSeg sg = ws_segment(p, L, hi);
        const int rows = sg.t1 - sg.t0;
        {
          uint32_t ph1;
          const int s1 = slot_after(s0, 1, ph0, ph1);
          if (!wait_all(&full[s0], ph0) || !wait_all(&full[s1], ph1)) { alive = false; break; }
        }
        for (int j = 0; j < rows; ++j, ++orow) {
          const int buf = orow & 1;
          uint32_t ph1, ph2;
          const int s1 = slot_after(s0, 1, ph0, ph1);
          const int s2 = slot_after(s0, 2, ph0, ph2);
          if (!wait_all(&tempty[buf], ((orow >> 1) & 1) ^ 1)) { alive = false; break; }
          if (!wait_all(&full[s2], ph2)) { alive = false; break; }
          tc_fence_after();
          const uint32_t acc0 = tmem_base + (uint32_t)(buf * MT * NT);
          const int slots[3] = {s0, s1, s2};
          if (elect_one()) {
#pragma unroll
            for (int dt = 0; dt < 3; ++dt) {
              const uint32_t a_lo = a_lo0 + (uint32_t)slots[dt] * (kSlot >> 4);
#pragma unroll
              for (int df = 0; df < 3; ++df) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                  for (int k = 0; k < K16; ++k) {
                    const uint64_t ad = desc_at(a_lo, a_hi, mt * kATile + df * 16 + k * 2 * kALbo);
                    const uint64_t bd = desc_at(b_lo0, b_hi, (dt * 3 + df) * kTap + k * 2 * kBLbo);
                    if (dt == 0 && df == 0 && k == 0)
                      umma_f16_c<false>(acc0 + (uint32_t)(mt * NT), ad, bd, idesc);
                    else
                      umma_f16_c<true>(acc0 + (uint32_t)(mt * NT), ad, bd, idesc);
                  }
                }
              }
              if (dt == 0) umma_commit(&empty[s0]);  // oldest row is done: it is refilled during dt = 1, 2
            }
            if (j == rows - 1) {
              umma_commit(&empty[s1]);
              umma_commit(&empty[s2]);
            }
            umma_commit(&tfull[buf]);
          }
          __syncwarp();
          s0 = s1;
          ph0 = ph1;
        }
        // skip the two trailing rows of this segment
        {
          uint32_t ph;
          s0 = slot_after(s0, 2, ph0, ph);
          ph0 = ph;
        }
        L += rows;
      }
    }
  } else {
    // ===================== epilogue (warps 2..13) =====================
    // One warp per (TMEM lane quadrant, 16-channel column group): 3 warps per scheduler hide each
    // other's latencies, and a warp's 16 scale/shift pairs live in registers for the whole kernel.
    static_assert(NT == 16 * kWsEpiGroups, "column groups");
    const int quad = warp & 3;          // hardware rule: warp w may read TMEM lanes 32*(w%4) .. +31
    const int grp = (warp - 2) >> 2;    // channels [16*grp, 16*grp + 16) of this CTA's NT
    float sc[16], sh[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      sc[e] = s_scale[grp * 16 + e];
      sh[e] = s_shift[grp * 16 + e];
    }
    const size_t plane = (size_t)p.F * 8;  // elements between consecutive 8-channel groups (CG8)
    uint32_t orow = 0;
    bool alive = true;
    for (long long L = lo; L < hi && alive;) {
      const WsSeg sg = ws_segment(p, L, hi);
      // CG8: plane (b, t, c/8) is [F][8], so the 32 lanes of a warp store 512 contiguous bytes
      h16* row0 = p.out + cg8_index(sg.b, sg.t0, nt * (NT / 8) + grp * 2, sg.f0 + quad * 32 + lane, p.T, C, p.F);
      const size_t row_stride = (size_t)(C / 8) * plane;
      for (int t = sg.t0; t < sg.t1; ++t, ++orow, row0 += row_stride) {
        const int buf = orow & 1;
        if (!mbar_wait(&tfull[buf], (orow >> 1) & 1, abort_flag)) { alive = false; break; }
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * MT * NT + grp * 16);
        uint32_t r[MT][16];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) tmem_ld16(taddr + mt * NT, r[mt]);
        tmem_ld_wait();
        // the accumulators are in registers: hand the TMEM buffer back before the stores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(&tempty[buf]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          if (sg.f0 + mt * kWsTileM + quad * 32 + lane < p.F) {
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float v0 = fmaxf(fmaf(__uint_as_float(r[mt][2 * e]), sc[2 * e], sh[2 * e]), 0.f);
              const float v1 = fmaxf(fmaf(__uint_as_float(r[mt][2 * e + 1]), sc[2 * e + 1], sh[2 * e + 1]), 0.f);
              pk[e] = pack2<FMT>(v0, v1);
            }
            h16* dst = row0 + (size_t)mt * kWsTileM * 8;
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(dst + plane) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      }
      L += sg.t1 - sg.t0;
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
// CTA-pair variant (cta_group::2) for C = 96: the two CTAs of a cluster form one M = 256 tile (128
// positions each) with the full N = 96.  Each CTA keeps its half of the weights resident (output
// channels 48*rank .. +48 = the N/2 rows of B the pair MMA reads from this CTA) and loads only its own
// 130-position input rows, so nothing is fetched twice; an N = 96 MMA is tensor-pipe bound (48 cycles)
// where the single-CTA N = 48 shape is shared-memory bound (44 cycles for half the work).
//   leader (rank 0): issues the MMAs; its full[] barriers collect the TMA bytes of BOTH CTAs;
//   commits are multicast to the empty[] / tfull[] barriers of both CTAs; the peer's epilogue warps
//   release the accumulators by arriving remotely on the leader's tempty[] barrier.
// ------------------------------------------------------------------------------------------------
template <int C, int R, int FMT>
__global__ void __launch_bounds__(kWsThreads, 1)
tc_conv3x3_ws2_kernel(const __grid_constant__ CUtensorMap in_map, const WsParams p) {
  constexpr int NH = C / 2;               // output channels whose weights live in this CTA
  constexpr int kGrpCh = C / kWsEpiGroups;  // channels per epilogue column group
  static_assert(C % 32 == 0 && kGrpCh % 16 == 0, "shape");
  constexpr int kALbo = kWsRowPos * 16;
  constexpr int kATile = ((C / 8) * kALbo + 127) / 128 * 128;
  constexpr int kSlot = kATile;
  constexpr int kTap = C * NH * 2;
  constexpr int kWBytes = 9 * kTap;
  constexpr int kBLbo = NH * 16;
  constexpr int K16 = C / 16;
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [R]   (used in the leader)
  uint64_t* empty = full + kWsMaxR;                     // [R]
  uint64_t* tfull = full + 2 * kWsMaxR;                 // [2]
  uint64_t* tempty = tfull + 2;                         // [2]   (used in the leader)
  uint64_t* wbar = tempty + 2;                          // own weights landed
  uint64_t* wbar_peer = wbar + 1;                       // leader: the peer's weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar_peer + 1);
  float* s_scale = reinterpret_cast<float*>(smem + 256);  // [C]
  float* s_shift = s_scale + 96;
  uint8_t* w_smem = smem + 1024;
  uint8_t* ring = w_smem + kWBytes;
  volatile int* abort_flag = p.abort_flag;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_pairs = gridDim.x >> 1;
  long long lo, hi;
  ws_range(p, n_pairs, blockIdx.x >> 1, lo, hi);

  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 2 * kWsEpiWarps);
    }
    mbar_init(wbar, 1);
    mbar_init(wbar_peer, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before anything can signal them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its output is visible from here on

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      mbar_expect_tx(wbar, (uint32_t)kWBytes);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)rank * kWBytes;
      for (int tap = 0; tap < 9; ++tap) bulk_load_1d(w_smem + tap * kTap, wsrc + tap * kTap, kTap, wbar);
      int s = 0;
      uint32_t ph = 0;
      bool alive = true;
      for (long long L = lo; L < hi && alive;) {
        const WsSeg sg = ws_segment(p, L, hi);
        for (int r = sg.t0 - 1; r <= sg.t1; ++r) {
          if (!mbar_wait(&empty[s], ph ^ 1, abort_flag)) { alive = false; break; }
          // the leader's barrier counts the bytes of both CTAs' tiles
          if (leader) mbar_expect_tx(&full[s], 2u * (uint32_t)((C / 8) * kWsRowPos * 16));
          tma_load_5d_2sm(ring + (size_t)s * kSlot, &in_map, mapa_u32(smem_u32(&full[s]), 0), 0,
                          sg.f0 + (int)rank * kWsTileM - 1, 0, r, sg.b);
          if (++s == R) { s = 0; ph ^= 1; }
        }
        L += sg.t1 - sg.t0;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) ======
    auto wait_all = [&](uint64_t* bar, uint32_t parity) {
      return __all_sync(0xffffffffu, mbar_wait(bar, parity, abort_flag)) != 0;
    };
    bool alive = wait_all(wbar, 0);
    if (!leader) {
      if (alive && lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(wbar_peer), 0));  // "my weights are in place"
    } else {
      alive = alive && wait_all(wbar_peer, 0);
      const uint32_t idesc = make_idesc_2sm<FMT>(C);
      const uint64_t a_proto = make_desc(0, kALbo, 128), b_proto = make_desc(0, kBLbo, 128);
      const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
      const uint32_t a_lo0 = (uint32_t)a_proto + (smem_u32(ring) >> 4);
      const uint32_t b_lo0 = (uint32_t)b_proto + (smem_u32(w_smem) >> 4);
      int s0 = 0;
      uint32_t ph0 = 0;
      uint32_t orow = 0;
      auto slot_after = [&](int s, int d, uint32_t ph, uint32_t& ph_out) {
        int q = s + d;
        ph_out = ph;
        if (q >= R) { q -= R; ph_out ^= 1; }
        return q;
      };
      for (long long L = lo; L < hi && alive;) {
        const WsSeg sg = ws_segment(p, L, hi);
        const int rows = sg.t1 - sg.t0;
        {
          uint32_t ph1;
          const int s1 = slot_after(s0, 1, ph0, ph1);
          if (!wait_all(&full[s0], ph0) || !wait_all(&full[s1], ph1)) { alive = false; break; }
        }
        for (int j = 0; j < rows; ++j, ++orow) {
          const int buf = orow & 1;
          uint32_t ph1, ph2;
          const int s1 = slot_after(s0, 1, ph0, ph1);
          const int s2 = slot_after(s0, 2, ph0, ph2);
          if (!wait_all(&tempty[buf], ((orow >> 1) & 1) ^ 1)) { alive = false; break; }
          if (!wait_all(&full[s2], ph2)) { alive = false; break; }
          tc_fence_after();
          const uint32_t acc = tmem_base + (uint32_t)(buf * C);
          const int slots[3] = {s0, s1, s2};
          if (elect_one()) {
#pragma unroll
            for (int dt = 0; dt < 3; ++dt) {
              const uint32_t a_lo = a_lo0 + (uint32_t)slots[dt] * (kSlot >> 4);
#pragma unroll
              for (int df = 0; df < 3; ++df) {
#pragma unroll
                for (int k = 0; k < K16; ++k) {
                  const uint64_t ad = desc_at(a_lo, a_hi, df * 16 + k * 2 * kALbo);
                  const uint64_t bd = desc_at(b_lo0, b_hi, (dt * 3 + df) * kTap + k * 2 * kBLbo);
                  if (dt == 0 && df == 0 && k == 0)
                    umma_f16_2sm<false>(acc, ad, bd, idesc);
                  else
                    umma_f16_2sm<true>(acc, ad, bd, idesc);
                }
              }
              if (dt == 0) umma_commit_2sm(&empty[s0]);  // oldest row is done in both CTAs: refill during dt = 1, 2
            }
            if (j == rows - 1) {
              umma_commit_2sm(&empty[s1]);
              umma_commit_2sm(&empty[s2]);
            }
            umma_commit_2sm(&tfull[buf]);
          }
          __syncwarp();
          s0 = s1;
          ph0 = ph1;
        }
        {
          uint32_t ph;
          s0 = slot_after(s0, 2, ph0, ph);
          ph0 = ph;
        }
        L += rows;
      }
    }
  } else {
    // ===================== epilogue (warps 2..13, both CTAs: own 128 positions x all C channels) =====
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;  // channels [kGrpCh*grp, +kGrpCh)
    const size_t plane = (size_t)p.F * 8;
    const uint32_t tempty_leader = mapa_u32(smem_u32(&tempty[0]), 0);
    uint32_t orow = 0;
    bool alive = true;
    for (long long L = lo; L < hi && alive;) {
      const WsSeg sg = ws_segment(p, L, hi);
      const int f = sg.f0 + (int)rank * kWsTileM + quad * 32 + lane;
      h16* row0 = p.out + cg8_index(sg.b, sg.t0, grp * (kGrpCh / 8), f, p.T, C, p.F);
      const size_t row_stride = (size_t)(C / 8) * plane;
      for (int t = sg.t0; t < sg.t1; ++t, ++orow, row0 += row_stride) {
        const int buf = orow & 1;
        if (!mbar_wait(&tfull[buf], (orow >> 1) & 1, abort_flag)) { alive = false; break; }
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * C + grp * kGrpCh);
        uint32_t r[kGrpCh];
#pragma unroll
        for (int j = 0; j < kGrpCh; j += 16) tmem_ld16(taddr + j, r + j);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(tempty_leader + (uint32_t)buf * 8);  // accumulators are in registers
        if (f < p.F) {
#pragma unroll
          for (int j = 0; j < kGrpCh; j += 8) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int ch = grp * kGrpCh + j + 2 * e;
              const float v0 = fmaxf(fmaf(__uint_as_float(r[j + 2 * e]), s_scale[ch], s_shift[ch]), 0.f);
              const float v1 = fmaxf(fmaf(__uint_as_float(r[j + 2 * e + 1]), s_scale[ch + 1], s_shift[ch + 1]), 0.f);
              pk[e] = pack2<FMT>(v0, v1);
            }
            *reinterpret_cast<uint4*>(row0 + (size_t)(j >> 3) * plane) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      L += sg.t1 - sg.t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// Row-stacked variant for C = 48 ("rs"): the three VERTICAL taps are stacked along N.
//
// With N = 48 an SS-MMA spends 44 cycles reading its operands for 24 cycles of math (ncu: L1/smem 80 %
// busy, tensor pipe 44 %).  Input row r contributes to output rows r+1, r, r-1 through the taps dt = 0, 1, 2
// at the SAME position, so one MMA of N = 144 with the weight columns ordered (dt, co) computes all three
// contributions of row r from one read of the A tile - provided the accumulators of the output rows
// r+1, r, r-1 sit side by side in TMEM.  They do: every M tile owns a ring of kRsBlocks 48-column blocks,
// output row x lives in block (-x mod kRsBlocks), so the three rows touched by step g are consecutive
// blocks (a window that straddles the end of the ring is issued as two MMAs).  Blocks are handed back
// ZEROED by the epilogue (tcgen05.st), so every MMA accumulates and no flag has to differ across N.
// An input row is now used by exactly one step: the rolling row ring becomes a plain stream.
//   per 128 positions: 9 MMAs x 72 cycles (tensor bound) instead of 27 x 44 (shared-memory bound).
// Virtual step index g runs over (rows + 4) steps per segment: rows + 2 real input rows (with the two halo
// rows) and two empty steps that flush the last two blocks; block x is complete after step x + 1.
// ------------------------------------------------------------------------------------------------
constexpr int kRsBlocks = 5;
constexpr int kRsMT = 2;
constexpr int kRsSlots = 6;

template <int FMT>
__global__ void __launch_bounds__(kWsThreads, 1)
tc_conv3x3_rs_kernel(const __grid_constant__ CUtensorMap in_map, const WsParams p) {
  constexpr int C = 48, NT = 48, MT = kRsMT, RB = kRsBlocks, RA = kRsSlots;
  constexpr int kALbo = kWsRowPos * 16;
  constexpr int kATile = ((C / 8) * kALbo + 127) / 128 * 128;
  constexpr int kSlot = MT * kATile;
  constexpr int kBLbo = 3 * NT * 16;              // rows of B = (dt, co): 144 rows of 16 B per channel group
  constexpr int kDfBytes = (C / 8) * kBLbo;        // one horizontal tap: [C/8][144][8]
  constexpr int kWBytes = 3 * kDfBytes;
  constexpr int K16 = C / 16;
  constexpr int kTileCols = RB * NT;               // TMEM columns of one M tile's block ring
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [RA]
  uint64_t* empty = full + kWsMaxR;                     // [RA]
  uint64_t* done = full + 2 * kWsMaxR;                  // [RB]  MMA -> epilogue: block complete
  uint64_t* bfree = done + 8;                           // [RB]  epilogue -> MMA: block drained and zeroed
  uint64_t* wbar = bfree + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  float* s_scale = reinterpret_cast<float*>(smem + 512);  // [NT]
  float* s_shift = s_scale + 64;
  uint8_t* w_smem = smem + 1024;
  uint8_t* ring = w_smem + kWBytes;
  volatile int* abort_flag = p.abort_flag;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long lo, hi;
  ws_range(p, gridDim.x, blockIdx.x, lo, hi);

  if (threadIdx.x == 0) {
    for (int s = 0; s < RA; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < RB; ++b) {
      mbar_init(&done[b], 1);
      mbar_init(&bfree[b], kWsEpiWarps);
    }
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < NT; i += blockDim.x) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
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

  if (warp == 0) {
    // ===================== TMA producer: weights once, then one slot per input row =====================
    if (lane == 0) {
      mbar_expect_tx(wbar, (uint32_t)kWBytes);
      for (int df = 0; df < 3; ++df)
        bulk_load_1d(w_smem + df * kDfBytes, reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)df * kDfBytes, kDfBytes, wbar);
      int s = 0;
      uint32_t ph = 0;
      bool alive = true;
      for (long long L = lo; L < hi && alive;) {
        const WsSeg sg = ws_segment(p, L, hi);
        for (int r = sg.t0 - 1; r <= sg.t1; ++r) {
          if (!mbar_wait(&empty[s], ph ^ 1, abort_flag)) { alive = false; break; }
          uint8_t* dst = ring + (size_t)s * kSlot;
          mbar_expect_tx(&full[s], (uint32_t)(MT * (C / 8) * kWsRowPos * 16));
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            tma_load_5d(dst + mt * kATile, &in_map, &full[s], 0, sg.f0 + mt * kWsTileM - 1, 0, r, sg.b);
          if (++s == RA) { s = 0; ph ^= 1; }
        }
        L += sg.t1 - sg.t0;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    const uint32_t idesc144 = make_idesc<FMT>(144), idesc96 = make_idesc<FMT>(96), idesc48 = make_idesc<FMT>(48);
    const uint64_t a_proto = make_desc(0, kALbo, 128), b_proto = make_desc(0, kBLbo, 128);
    const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
    const uint32_t a_lo0 = (uint32_t)a_proto + (smem_u32(ring) >> 4);
    const uint32_t b_lo0 = (uint32_t)b_proto + (smem_u32(w_smem) >> 4);
    auto wait_all = [&](uint64_t* bar, uint32_t parity) {
      return __all_sync(0xffffffffu, mbar_wait(bar, parity, abort_flag)) != 0;
    };
    bool alive = wait_all(wbar, 0);
    // the epilogue warps zero the block ring before the first MMA (named barrier 1: 12 epilogue warps + this warp)
    asm volatile("bar.sync 1, %0;" ::"r"((kWsEpiWarps + 1) * 32) : "memory");
    tc_fence_after();
    int s = 0;
    uint32_t ph = 0;
    int gm = 0;         // g mod RB
    uint32_t cyc = 0;   // g / RB
    for (long long L = lo; L < hi && alive;) {
      const WsSeg sg = ws_segment(p, L, hi);
      const int rows = sg.t1 - sg.t0;
      for (int v = 0; v < rows + 4 && alive; ++v) {
        // virtual step g touches output indices g+1, g, g-1 = blocks sb, sb+1, sb+2 (mod RB), sb = (-(g+1)) mod RB
        const int sb = RB - 1 - gm;
        // Index g+1 enters block sb at this step (also at the two empty steps, so that every index is entered in
        // order): its n earlier owners (indices g+1-RB, ...; the ring starts zeroed, owners start at index -1)
        // must have been drained and zeroed: n = floor((g + 2) / RB).
        {
          const uint32_t n = (gm + 2 >= RB) ? cyc + 1 : cyc;
          if (n > 0 && !wait_all(&bfree[sb], (n - 1) & 1)) { alive = false; break; }
        }
        if (v < rows + 2) {
          if (!wait_all(&full[s], ph)) { alive = false; break; }
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + (uint32_t)s * (kSlot >> 4);
            const int n1 = (sb + 3 <= RB) ? 3 : RB - sb;  // blocks before the ring wraps
#pragma unroll
            for (int df = 0; df < 3; ++df) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint32_t d0 = tmem_base + (uint32_t)(mt * kTileCols + sb * NT);
                const uint32_t d1 = tmem_base + (uint32_t)(mt * kTileCols);
#pragma unroll
                for (int k = 0; k < K16; ++k) {
                  const uint64_t ad = desc_at(a_lo, a_hi, mt * kATile + df * 16 + k * 2 * kALbo);
                  const uint32_t boff = df * kDfBytes + k * 2 * kBLbo;
                  if (n1 == 3) {
                    umma_f16_c<true>(d0, ad, desc_at(b_lo0, b_hi, boff), idesc144);
                  } else if (n1 == 2) {
                    umma_f16_c<true>(d0, ad, desc_at(b_lo0, b_hi, boff), idesc96);
                    umma_f16_c<true>(d1, ad, desc_at(b_lo0, b_hi, boff + 2 * NT * 16), idesc48);
                  } else {
                    umma_f16_c<true>(d0, ad, desc_at(b_lo0, b_hi, boff), idesc48);
                    umma_f16_c<true>(d1, ad, desc_at(b_lo0, b_hi, boff + NT * 16), idesc96);
                  }
                }
              }
            }
            umma_commit(&empty[s]);
          }
          __syncwarp();
          if (++s == RA) { s = 0; ph ^= 1; }
        }
        // output index g-1 (block sb+2 mod RB) is complete after this step
        if (elect_one()) umma_commit(&done[(sb + 2) % RB]);
        __syncwarp();
        if (++gm == RB) { gm = 0; ++cyc; }
      }
      L += rows;
    }
  } else {
    // ===================== epilogue (warps 2..13): drain + zero one block per virtual step =====================
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;  // channels [16*grp, +16)
    float sc[16], sh[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      sc[e] = s_scale[grp * 16 + e];
      sh[e] = s_shift[grp * 16 + e];
    }
    const size_t plane = (size_t)p.F * 8;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * 16);
    // the ring starts zeroed: every warp clears its 16 columns of every block of both tiles
    {
      const uint32_t z = 0;
#pragma unroll 1
      for (int c0 = 0; c0 < MT * kTileCols; c0 += NT)
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(lane_addr + c0),
            "r"(z)
            : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
    }
    // every warp has zeroed its columns before the first MMA: named barrier 1 = 12 epilogue warps + the MMA warp
    asm volatile("bar.sync 1, %0;" ::"r"((kWsEpiWarps + 1) * 32) : "memory");
    uint32_t uses[RB];   // completions of each block seen so far (parity of its `done` barrier)
#pragma unroll
    for (int b = 0; b < RB; ++b) uses[b] = 0;
    // virtual output index x runs in lockstep with the MMA warp: one drained block per virtual step, in order
    int blk = 1 % RB;    // first drained index is x = g - 1 with g = 0  ->  x = -1  ->  block (1 mod RB)
    bool alive = true;
    for (long long L = lo; L < hi && alive;) {
      const WsSeg sg = ws_segment(p, L, hi);
      const int rows = sg.t1 - sg.t0;
      for (int v = 0; v < rows + 4; ++v) {
        // step v completes output row t = t0 + v - 2 (valid for 2 <= v < rows + 2)
        const int t = sg.t0 + v - 2;
        const bool store = v >= 2 && v < rows + 2;
        uint32_t u = 0;
#pragma unroll
        for (int b = 0; b < RB; ++b)
          if (b == blk) u = uses[b];
        if (!mbar_wait(&done[blk], u & 1, abort_flag)) { alive = false; break; }
        tc_fence_after();
        uint32_t r[MT][16];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) tmem_ld16(lane_addr + (uint32_t)(mt * kTileCols + blk * NT), r[mt]);
        tmem_ld_wait();
        {
          const uint32_t z = 0;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(
                    lane_addr + (uint32_t)(mt * kTileCols + blk * NT)),
                "r"(z)
                : "memory");
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(&bfree[blk]);
        if (store) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int f = sg.f0 + mt * kWsTileM + quad * 32 + lane;
            if (f < p.F) {
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float v0 = fmaxf(fmaf(__uint_as_float(r[mt][2 * e]), sc[2 * e], sh[2 * e]), 0.f);
                const float v1 = fmaxf(fmaf(__uint_as_float(r[mt][2 * e + 1]), sc[2 * e + 1], sh[2 * e + 1]), 0.f);
                pk[e] = pack2<FMT>(v0, v1);
              }
              h16* dst = p.out + cg8_index(sg.b, t, grp * 2, f, p.T, C, p.F);
              *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(dst + plane) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
#pragma unroll
        for (int b = 0; b < RB; ++b)
          if (b == blk) uses[b] += 1;
        blk = blk == 0 ? RB - 1 : blk - 1;  // next index x+1 lives in block blk-1
      }
      L += rows;
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
struct TcConvWsWeights {
  int C;
  int fmt;
  WsCfg cfg;
  h16* d_pack;
  h16* d_pack_rs = nullptr;  // C = 48: row-stacked packing
};

static bool ws_make_cfg(int C, int F, WsCfg& c) {
  if (C != 48 && C != 96) return false;
  c.C = C;
  c.NT = 48;
  c.nsplit = C / c.NT;
  c.a_lbo = kWsRowPos * 16;
  c.a_tile_bytes = (int)align_up((size_t)(C / 8) * c.a_lbo, 128);
  c.w_bytes = 9 * C * c.NT * 2;
  const int tiles = (F + kWsTileM - 1) / kWsTileM;
  const int budget = 227 * 1024 - 1024 - c.w_bytes;
  c.MT = 1;
  static const int mt_cap = getenv("AC_WS_MT") ? atoi(getenv("AC_WS_MT")) : 3;  // dev hook: tiles per row step
  for (int mt = mt_cap < 1 ? 1 : (mt_cap > 3 ? 3 : mt_cap); mt >= 1; --mt) {
    if (mt > tiles) continue;
    if (2 * mt * c.NT > 512) continue;
    if (4 * mt * c.a_tile_bytes <= budget) { c.MT = mt; break; }
  }
  c.slot_bytes = c.MT * c.a_tile_bytes;
  c.R = C == 48 ? (c.MT == 3 ? 5 : 6) : 5;  // must match the instantiations in launch_tc_conv3x3_ws
  if (c.R * c.slot_bytes > budget) return false;
  c.smem_bytes = 1024 + c.w_bytes + c.R * c.slot_bytes;
  return true;
}

int tc_conv3x3_ws_supported(int T, int F, int C) {
  WsCfg c;
  (void)T;
  return ws_make_cfg(C, F, c) ? AC_OK : AC_E_INVALID;
}

static int ws_pack_rs(const float* h_w, int fmt, h16** d_out);

int tc_conv3x3_ws_pack(const float* h_w, int C, int fmt, TcConvWsWeights** out) {
  *out = nullptr;
  WsCfg c;
  if (!ws_make_cfg(C, 1 << 20, c)) return AC_OK;
  // [nt][dt][df][C/8][NT][8]  <-  W[co][ci][kh=dt][kw=df]
  std::vector<h16> pack((size_t)9 * C * C);
  size_t o = 0;
  for (int nt = 0; nt < c.nsplit; ++nt)
    for (int dt = 0; dt < 3; ++dt)
      for (int df = 0; df < 3; ++df)
        for (int kg = 0; kg < C / 8; ++kg)
          for (int n = 0; n < c.NT; ++n)
            for (int e = 0; e < 8; ++e) {
              const int co = nt * c.NT + n, ci = kg * 8 + e;
              pack[o++] = h16_rn(h_w[(((size_t)co * C + ci) * 3 + dt) * 3 + df], fmt);
            }
  TcConvWsWeights* w = new TcConvWsWeights();
  w->C = C;
  w->fmt = fmt;
  w->d_pack = nullptr;
  if (cudaMalloc(&w->d_pack, pack.size() * 2) != cudaSuccess ||
      cudaMemcpy(w->d_pack, pack.data(), pack.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("tc ws weight upload failed");
    delete w;
    return AC_E_CUDA;
  }
  if (C == 48 && ws_pack_rs(h_w, fmt, &w->d_pack_rs) != AC_OK) {
    cudaFree(w->d_pack);
    delete w;
    return AC_E_CUDA;
  }
  *out = w;
  return AC_OK;
}

void tc_conv3x3_ws_free(TcConvWsWeights* w) {
  if (!w) return;
  if (w->d_pack_rs) cudaFree(w->d_pack_rs);
  if (w->d_pack) cudaFree(w->d_pack);
  delete w;
}

// row-stacked packing for C = 48: [df][C/8][dt*48 + co][8]
static int ws_pack_rs(const float* h_w, int fmt, h16** d_out) {
  const int C = 48;
  std::vector<h16> pack((size_t)9 * C * C);
  size_t o = 0;
  for (int df = 0; df < 3; ++df)
    for (int kg = 0; kg < C / 8; ++kg)
      for (int dt = 0; dt < 3; ++dt)
        for (int co = 0; co < C; ++co)
          for (int e = 0; e < 8; ++e)
            pack[o++] = h16_rn(h_w[(((size_t)co * C + kg * 8 + e) * 3 + dt) * 3 + df], fmt);
  *d_out = nullptr;
  if (cudaMalloc(d_out, pack.size() * 2) != cudaSuccess ||
      cudaMemcpy(*d_out, pack.data(), pack.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("tc rs weight upload failed");
    return AC_E_CUDA;
  }
  return AC_OK;
}

static bool g_ws_pair_enabled = true;
// Off by default: at C = 48 the weight-stationary kernel already runs at the practical HBM rate (4.2 TB/s of
// algorithmic traffic, the streaming 1x1 kernels reach 4.7-4.9), so removing its shared-memory bottleneck
// changes nothing measurable (572 vs 576 us per 16-window layer, bit-identical output).  Kept as the
// tensor-bound formulation for narrower channel counts / faster memory; exercised by the tests.
static bool g_ws_rs_enabled = false;
void tc_conv3x3_ws_set_rs(int enabled) { g_ws_rs_enabled = enabled != 0; }
void tc_conv3x3_ws_set_pair(int enabled) { g_ws_pair_enabled = enabled != 0; }

int launch_tc_conv3x3_ws(const TcConvWsWeights* w, const TcConvArgs& a, cudaStream_t st) {
  AC_REQUIRE(w && w->C == a.C, "tc ws conv: weights do not match the layer");
  WsCfg c;
  AC_REQUIRE(ws_make_cfg(a.C, a.F, c), "tc ws conv: unsupported shape");
  EncodeTiledFn enc = get_tensor_map_encoder();
  AC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available");
  AC_REQUIRE(tc_abort_flag() != nullptr, "abort flag allocation failed");
  // 5-D view of the CG8 tensor [nB][T][C/8][F][8]: (c%8, f, c/8, t, b); one box = [C/8][130 positions][8 ch]
  CUtensorMap map;
  const cuuint64_t dims[5] = {8, (cuuint64_t)a.F, (cuuint64_t)(a.C / 8), (cuuint64_t)a.T, (cuuint64_t)a.nB};
  const cuuint64_t strides[4] = {16, (cuuint64_t)a.F * 16, (cuuint64_t)a.F * a.C * 2, (cuuint64_t)a.T * a.F * a.C * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)kWsRowPos, (cuuint32_t)(a.C / 8), 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<h16*>(a.in), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (ws conv) failed with code " + std::to_string((int)r));
    return AC_E_CUDA;
  }
  const bool pair = g_ws_pair_enabled && a.C == 96 && device_sm_count() >= 2;
  WsParams p;
  p.cfg = c;
  if (pair) p.cfg.MT = 2;  // a strip = the pair's two 128-position tiles
  p.nB = a.nB; p.T = a.T; p.F = a.F;
  p.n_strips = ((a.F + kWsTileM - 1) / kWsTileM + p.cfg.MT - 1) / p.cfg.MT;
  p.total_rows = (long long)a.nB * p.n_strips * a.T;
  p.wpack = w->d_pack;
  p.scale = a.scale; p.shift = a.shift;
  p.out = a.out;
  p.abort_flag = tc_abort_flag();
  ProfScope ps(KC_CONV_TC, 2.0 * 9.0 * a.nB * (double)a.T * a.F * a.C * a.C, 4.0 * a.nB * (double)a.T * a.F * a.C, st);
  if (pair) {
    auto kern = w->fmt == kFmtBF16 ? tc_conv3x3_ws2_kernel<96, 5, kFmtBF16> : tc_conv3x3_ws2_kernel<96, 5, kFmtF16>;
    const int smem = 1024 + 9 * 96 * 48 * 2 + 5 * c.a_tile_bytes;
    AC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int pairs = device_sm_count() / 2;
    if ((long long)pairs > p.total_rows) pairs = (int)p.total_rows;
    AC_CHECK_CUDA(tc_launch(kern, 2 * pairs, kWsThreads, smem, st, 2, map, p));
    AC_LAUNCH_CHECK();
    return AC_OK;
  }
  if (g_ws_rs_enabled && a.C == 48 && w->d_pack_rs && (a.F + kWsTileM - 1) / kWsTileM >= kRsMT) {
    p.cfg.MT = kRsMT;
    p.n_strips = ((a.F + kWsTileM - 1) / kWsTileM + kRsMT - 1) / kRsMT;
    p.total_rows = (long long)a.nB * p.n_strips * a.T;
    p.wpack = w->d_pack_rs;
    const int smem = 1024 + 9 * 48 * 48 * 2 + kRsSlots * kRsMT * c.a_tile_bytes;
    auto rs_kern = w->fmt == kFmtBF16 ? tc_conv3x3_rs_kernel<kFmtBF16> : tc_conv3x3_rs_kernel<kFmtF16>;
    AC_CHECK_CUDA(cudaFuncSetAttribute(rs_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int grid = device_sm_count();
    if ((long long)grid > p.total_rows) grid = (int)p.total_rows;
    AC_CHECK_CUDA(tc_launch(rs_kern, grid, kWsThreads, smem, st, 1, map, p));
    AC_LAUNCH_CHECK();
    return AC_OK;
  }
  void (*kern)(const CUtensorMap, const WsParams) = nullptr;
  const bool bf = w->fmt == kFmtBF16;
  if (c.C == 48 && c.MT == 3) kern = bf ? tc_conv3x3_ws_kernel<48, 3, 5, kFmtBF16> : tc_conv3x3_ws_kernel<48, 3, 5, kFmtF16>;
  else if (c.C == 48 && c.MT == 2) kern = bf ? tc_conv3x3_ws_kernel<48, 2, 6, kFmtBF16> : tc_conv3x3_ws_kernel<48, 2, 6, kFmtF16>;
  else if (c.C == 48 && c.MT == 1) kern = bf ? tc_conv3x3_ws_kernel<48, 1, 6, kFmtBF16> : tc_conv3x3_ws_kernel<48, 1, 6, kFmtF16>;
  else if (c.C == 96 && c.MT == 1) kern = bf ? tc_conv3x3_ws_kernel<96, 1, 5, kFmtBF16> : tc_conv3x3_ws_kernel<96, 1, 5, kFmtF16>;
  AC_REQUIRE(kern != nullptr, "tc ws conv: no instantiation for this shape");
  AC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int group = device_sm_count() / c.nsplit;
  if ((long long)group > p.total_rows) group = (int)p.total_rows;
  const int grid = group * c.nsplit;
  AC_CHECK_CUDA(tc_launch(kern, grid, kWsThreads, c.smem_bytes, st, 1, map, p));
  AC_LAUNCH_CHECK();
  return AC_OK;
}

}  // namespace ac
