// tcgen05 kernels for the 2x2 / stride-2 down-sampling convs and the 2x2 / stride-2 transposed
// convs (with the multiplicative skip fused into the epilogue) between U-Net levels.
//
// Both are implicit GEMMs over 128-position tiles of one output (down) / input (up) row:
//   DOWN  D[128][NT] += X[b][2t+dt][2f+df][ci] * W[co][ci][dt][df]     K = 4*C_in, taps walk K
//         the stride-2 gather is a 5-D tensor map over the CG8 tensor (c%8, df, f, c/8, row): a box
//         [KC/8][128][1][8] lands as the K-major no-swizzle tile [KC/8][128 positions][8 channels];
//   UP    D_tap[128][C_out] += X[b][t][f][ci] * W[ci][co][dt][df]      K = C_in, taps walk N
//         each tap has its own TMEM accumulator; the epilogue scatters tap (dt,df) of position
//         (t,f) to out[b][2t+dt][2f+df][:] and multiplies by the encoder skip tensor.
// These layers are HBM bound (<= 8 FLOP/B), so the design goal is streaming: one pass over the
// input, one over the skip, one write - the MMA work hides under the memory time.
#include <stdlib.h>

#include <vector>

#include "tc_common.cuh"
#include "unet_kernels.cuh"

namespace ac {

constexpr int kRsEpiGroups = 3;  // epilogue warps per TMEM lane quadrant: group g owns the 16-channel chunks g, g+3, ...
constexpr int kRsEpiWarps = 4 * kRsEpiGroups;
constexpr int kRsThreads = (2 + kRsEpiWarps) * 32;
constexpr int kRsHeader = 4096;   // barriers + scale/shift staged in shared memory
enum { RS_DOWN = 0, RS_UP = 1 };

struct RsCfg {
  int mode;
  int Cin, Cout;
  int NT, nsplit;   // DOWN: N tile of C_out.  UP: NT = C_out, nsplit = tap groups
  int ntap;         // UP: taps per unit; DOWN: 2 (df taps per stage)
  int MT, KC, nkc, stages, nbuf;
  int a_tile_bytes;   // one [128 positions][KC] tile = (KC/8) * 2048
  int b_stage_bytes;
  int stage_bytes, smem_bytes;
};

struct RsParams {
  RsCfg cfg;
  int nB, T, F;  // DOWN: OUTPUT grid.  UP: INPUT grid
  int n_fg, n_units;
  const h16* wpack;
  const float* scale;
  const float* shift;
  const h16* skip;  // UP only
  h16* out;
  int* abort_flag;
};

// 256-bit accesses (sm_100): one lane moves the two horizontally adjacent output positions of one CG8 plane
struct U8 { uint32_t v[8]; };
__device__ __forceinline__ U8 ldg_stream_u8(const void* p) {
  U8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_u8(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

template <int FMT>
__global__ void __launch_bounds__(kRsThreads, 1) tc_resample_kernel(const __grid_constant__ CUtensorMap in_map, const RsParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  const RsCfg& c = p.cfg;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 8;
  uint64_t* tfull = full + 16;
  uint64_t* tempty = full + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 24);
  float* s_scale = reinterpret_cast<float*>(smem + 1024);  // [Cout] (<= 384)
  float* s_shift = s_scale + 384;
  uint8_t* stage0 = smem + kRsHeader;
  volatile int* abort_flag = p.abort_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < p.cfg.Cout && i < 384; i += blockDim.x) {
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
      mbar_init(&tempty[b], kRsEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its output is visible from here on
  const bool down = c.mode == RS_DOWN;
  const int ndt = down ? 2 : 1;             // K-walk over vertical taps (DOWN only)
  const int steps = ndt * c.nkc;
  const int a_per_stage = (down ? 2 : 1) * c.MT;  // A tiles per stage
  const int acc_per_unit = down ? c.MT : c.MT * c.ntap;

  auto decode = [&](int u, int& ns, int& b, int& t, int& f0) {
    const int fg = u % p.n_fg;
    int q = u / p.n_fg;
    t = q % p.T;
    q /= p.T;
    b = q % p.nB;
    ns = q / p.nB;
    f0 = fg * c.MT * 128;
  };

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      bool alive = true;
      for (int u = blockIdx.x; u < p.n_units && alive; u += gridDim.x) {
        int ns, b, t, f0;
        decode(u, ns, b, t, f0);
        for (int dt = 0; dt < ndt && alive; ++dt) {
          for (int kc = 0; kc < c.nkc; ++kc) {
            if (!mbar_wait(&empty[s], ph ^ 1, abort_flag)) { alive = false; break; }
            uint8_t* st = stage0 + (size_t)s * c.stage_bytes;
            mbar_expect_tx(&full[s], (uint32_t)(a_per_stage * c.a_tile_bytes + c.b_stage_bytes));
            for (int mt = 0; mt < c.MT; ++mt) {
              if (down) {
                for (int df = 0; df < 2; ++df)
                  tma_load_5d(st + (mt * 2 + df) * c.a_tile_bytes, &in_map, &full[s], 0, df, f0 + mt * 128, kc * (c.KC / 8),
                              b * 2 * p.T + 2 * t + dt);
              } else {
                tma_load_5d(st + mt * c.a_tile_bytes, &in_map, &full[s], 0, f0 + mt * 128, kc * (c.KC / 8), t, b);
              }
            }
            const size_t blob = (size_t)c.ntap * c.KC * c.NT;
            const h16* wsrc = p.wpack + ((size_t)((ns * ndt + dt) * c.nkc + kc)) * blob;
            bulk_load_1d(st + a_per_stage * c.a_tile_bytes, wsrc, (uint32_t)c.b_stage_bytes, &full[s]);
            if (++s == c.stages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform loop, one elected lane issues (keeps descriptors in uniform registers)
    {
      const uint32_t idesc = make_idesc<FMT>(c.NT);
      const uint32_t b_lbo = (uint32_t)c.NT * 16;
      const uint64_t a_proto = make_desc(0, 2048, 128), b_proto = make_desc(0, b_lbo, 128);
      auto wait_all = [&](uint64_t* bar, uint32_t parity) {
        return __all_sync(0xffffffffu, mbar_wait(bar, parity, abort_flag)) != 0;
      };
      int s = 0, buf = 0;
      uint32_t ph = 0, tph = 0;
      bool alive = true;
      for (int u = blockIdx.x; u < p.n_units && alive; u += gridDim.x) {
        if (!wait_all(&tempty[buf], tph ^ 1)) break;
        tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(buf * acc_per_unit * c.NT);
        for (int step = 0; step < steps; ++step) {
          if (!wait_all(&full[s], ph)) { alive = false; break; }
          tc_fence_after();
          const uint32_t sa = smem_u32(stage0 + (size_t)s * c.stage_bytes);
          const uint32_t sb = sa + (uint32_t)(a_per_stage * c.a_tile_bytes);
          if (elect_one()) {
            for (int tp = 0; tp < c.ntap; ++tp) {
              for (int mt = 0; mt < c.MT; ++mt) {
                const uint64_t ad0 = a_proto + ((sa + (down ? (mt * 2 + tp) : mt) * c.a_tile_bytes) >> 4);
                const uint64_t bd0 = b_proto + ((sb + tp * (c.KC * c.NT * 2)) >> 4);
                const uint32_t acc = down ? acc0 + (uint32_t)(mt * c.NT) : acc0 + (uint32_t)((mt * c.ntap + tp) * c.NT);
                for (int k = 0; k < c.KC / 16; ++k)
                  umma_f16(acc, ad0 + (uint64_t)(k * 2 * 128), bd0 + (uint64_t)((k * 2 * b_lbo) >> 4), idesc,
                           down ? (step | tp | k) != 0 : (step | k) != 0);
              }
            }
            umma_commit(&empty[s]);
          }
          __syncwarp();
          if (++s == c.stages) { s = 0; ph ^= 1; }
        }
        if (!alive) break;
        if (elect_one()) umma_commit(&tfull[buf]);
        __syncwarp();
        if (++buf == c.nbuf) { buf = 0; tph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..13) =====================
    // One warp per (TMEM lane quadrant, chunk group): group g owns the 16-channel chunks g, g+3, ... of
    // every accumulator; three warps per scheduler hide each other's latencies.
    const int quad = warp & 3;  // hardware rule: warp w may read TMEM lanes 32*(w%4) .. +31
    const int grp = (warp - 2) >> 2;
    const int chunks_a = c.NT >> 4;
    const int per_acc = chunks_a > grp ? (chunks_a - grp + kRsEpiGroups - 1) / kRsEpiGroups : 0;
    const int n_my = acc_per_unit * per_acc;
    constexpr int kMaxMy = 6;
    int my_acc[kMaxMy], my_col[kMaxMy];
#pragma unroll
    for (int i = 0; i < kMaxMy; ++i) {
      const int a = per_acc ? i / per_acc : 0, k = per_acc ? i - a * per_acc : 0;
      my_acc[i] = a;
      my_col[i] = (grp + kRsEpiGroups * k) * 16;
    }
    const bool fast = n_my <= kMaxMy;
    const bool pre_ok = fast && !down;
    const size_t plane = down ? (size_t)p.F * 8 : (size_t)(2 * p.F) * 8;  // elements between 8-channel groups of `out`
    // ---- UP with both horizontal taps in the unit (ntap >= 2): a thread owns BOTH output positions 2f, 2f+1 of its input
    // position, so one plane is 32 contiguous bytes per lane = one 256-bit skip load and one 256-bit store (the per-tap form
    // writes 16 bytes at a 32-byte stride: every sector filled by two different instructions).  Item = (vertical tap, plane).
    bool paired_done = false;
    if (!down && c.ntap >= 2) {
      const int n_cg = c.Cout >> 3;
      const int n_items = (c.ntap >> 1) * n_cg;
      const int n_mine = n_items > grp ? (n_items - grp + kRsEpiGroups - 1) / kRsEpiGroups : 0;
      constexpr int kMaxPair = 6;
      if ((n_items + kRsEpiGroups - 1) / kRsEpiGroups <= kMaxPair) {  // the same decision in every warp (the two forms enumerate differently)
        int buf = 0;
        uint32_t tph = 0;
        for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
          int ns, b, t, f0;
          decode(u, ns, b, t, f0);
          const int f = f0 + quad * 32 + lane;
          // element index of (row 2t + dt, plane cg, position 2f)
          auto addr = [&](int it) {
            const int idx = grp + kRsEpiGroups * it;
            const int dtl = idx / n_cg, cg = idx - dtl * n_cg;
            const int tap0 = ns * c.ntap + 2 * dtl;
            return cg8_index(b, 2 * t + (tap0 >> 1), cg, 2 * f, 2 * p.T, c.Cout, 2 * p.F);
          };
          U8 pre[kMaxPair];
          if (f < p.F) {
#pragma unroll
            for (int it = 0; it < kMaxPair; ++it)
              if (it < n_mine) pre[it] = ldg_stream_u8(p.skip + addr(it));
          }
          if (!mbar_wait(&tfull[buf], tph, abort_flag)) break;
          tc_fence_after();
          const uint32_t tb = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * acc_per_unit * c.NT);
#pragma unroll
          for (int it = 0; it < kMaxPair; ++it) {
            if (it < n_mine) {
              const int idx = grp + kRsEpiGroups * it;
              const int dtl = idx / n_cg, cg = idx - dtl * n_cg;
              uint32_t r0[8], r1[8];
              tmem_ld8(tb + (uint32_t)((2 * dtl) * c.NT + cg * 8), r0);      // tap (dt, df = 0): position 2f
              tmem_ld8(tb + (uint32_t)((2 * dtl + 1) * c.NT + cg * 8), r1);  // tap (dt, df = 1): position 2f + 1
              tmem_ld_wait();
              if (f < p.F) {
                uint32_t pk[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int ch = cg * 8 + 2 * e;
                  const float sc0 = s_scale[ch], sh0 = s_shift[ch], sc1 = s_scale[ch + 1], sh1 = s_shift[ch + 1];
                  const float2 m0 = unpack2<FMT>(pre[it].v[e]), m1 = unpack2<FMT>(pre[it].v[4 + e]);
                  pk[e] = pack2<FMT>(fmaxf(fmaf(__uint_as_float(r0[2 * e]), sc0, sh0), 0.f) * m0.x,
                                     fmaxf(fmaf(__uint_as_float(r0[2 * e + 1]), sc1, sh1), 0.f) * m0.y);
                  pk[4 + e] = pack2<FMT>(fmaxf(fmaf(__uint_as_float(r1[2 * e]), sc0, sh0), 0.f) * m1.x,
                                         fmaxf(fmaf(__uint_as_float(r1[2 * e + 1]), sc1, sh1), 0.f) * m1.y);
                }
                stg_u8(p.out + addr(it), pk);
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_relaxed(&tempty[buf]);
          if (++buf == c.nbuf) { buf = 0; tph ^= 1; }
        }
        paired_done = true;
      }
    }
    int buf = 0;
    uint32_t tph = 0;
    for (int u = blockIdx.x; u < p.n_units && !paired_done; u += gridDim.x) {
      int ns, b, t, f0;
      decode(u, ns, b, t, f0);
      // CG8 index of channel 0 (of this unit's N slice) for accumulator a
      auto locate = [&](int a, int& f, int& n0) -> size_t {
        const int mt = down ? a : a / c.ntap;
        const int tp = down ? 0 : a % c.ntap;
        f = f0 + mt * 128 + quad * 32 + lane;
        if (down) {
          n0 = ns * c.NT;
          return cg8_index(b, t, n0 >> 3, f, p.T, c.Cout, p.F);
        }
        const int tap = ns * c.ntap + tp;
        n0 = 0;
        return cg8_index(b, 2 * t + (tap >> 1), 0, 2 * f + (tap & 1), 2 * p.T, c.Cout, 2 * p.F);
      };
      // UP: the skip values do not depend on the MMAs; request them before waiting for the accumulators
      uint4 pre[2 * kMaxMy];
      if (pre_ok) {
#pragma unroll
        for (int i = 0; i < kMaxMy; ++i) {
          if (i < n_my) {
            int f, n0;
            const size_t o0 = locate(my_acc[i], f, n0) + (size_t)(my_col[i] >> 3) * plane;
            if (f < p.F) {
              pre[2 * i] = ldg_stream_u4(p.skip + o0);
              pre[2 * i + 1] = ldg_stream_u4(p.skip + o0 + plane);
            }
          }
        }
      }
      if (!mbar_wait(&tfull[buf], tph, abort_flag)) break;
      tc_fence_after();
      auto chunk16 = [&](int a, int j, bool have, uint4 q0, uint4 q1) {
        int f, n0;
        const size_t o0 = locate(a, f, n0) + (size_t)(j >> 3) * plane, o1 = o0 + plane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((buf * acc_per_unit + a) * c.NT);
        uint32_t r[16];
        tmem_ld16(taddr + j, r);
        tmem_ld_wait();
        if (f >= p.F) return;
        if (!down && !have) {
          q0 = *reinterpret_cast<const uint4*>(p.skip + o0);
          q1 = *reinterpret_cast<const uint4*>(p.skip + o1);
        }
        const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int ch = n0 + j + 2 * e;
          const float2 mul = down ? make_float2(1.f, 1.f) : unpack2<FMT>(w[e]);
          const float v0 = fmaxf(fmaf(__uint_as_float(r[2 * e]), s_scale[ch], s_shift[ch]), 0.f) * mul.x;
          const float v1 = fmaxf(fmaf(__uint_as_float(r[2 * e + 1]), s_scale[ch + 1], s_shift[ch + 1]), 0.f) * mul.y;
          pk[e] = pack2<FMT>(v0, v1);
        }
        *reinterpret_cast<uint4*>(p.out + o0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(p.out + o1) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      };
      if (fast) {
#pragma unroll
        for (int i = 0; i < kMaxMy; ++i)
          if (i < n_my) chunk16(my_acc[i], my_col[i], pre_ok, pre[2 * i], pre[2 * i + 1]);
      } else {
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int a = 0; a < acc_per_unit; ++a)
          for (int q = grp; q < chunks_a; q += kRsEpiGroups) chunk16(a, q * 16, false, z, z);
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

struct TcResampleWeights {
  int fmt;
  RsCfg cfg;
  h16* d_pack;
};

static bool make_rs_cfg(int mode, int Cin, int Cout, RsCfg& c) {
  if (Cin % 16 || Cout % 16 || Cin < 16 || Cout < 16 || Cout > 384) return false;
  c.mode = mode;
  c.Cin = Cin;
  c.Cout = Cout;
  if (mode == RS_DOWN) {
    c.nsplit = 1;
    if (Cout > 256) {
      c.nsplit = 0;
      for (int s = 2; s <= 8; ++s)
        if (Cout % (16 * s) == 0 && Cout / s <= 256) { c.nsplit = s; break; }
      if (!c.nsplit) return false;
    }
    c.NT = Cout / c.nsplit;
    c.ntap = 2;
    c.MT = 256 / c.NT;
    if (c.MT < 1) c.MT = 1;
    if (c.MT > 2) c.MT = 2;
    c.nbuf = (2 * c.MT * c.NT <= 512) ? 2 : 1;
  } else {
    if (Cout > 256) return false;
    c.NT = Cout;
    c.ntap = 4;
    while (c.ntap > 1 && 2 * c.ntap * c.NT > 512) c.ntap /= 2;
    // Two accumulator sets force one tap per unit at C_out = 144 (2 x 2 x 144 columns > 512): the 49 KB input tile is then
    // fetched four times and the units are too short to stream.  Two taps on ONE set (MMA and drain of a unit serialised, both
    // small next to its memory time) took that layer from 255 to 161 us; at C_out = 96 (401 -> 465 us) and at C_out >= 192
    // (79 -> 82, 28 -> 30 us) the single set loses, so they keep the double-buffered configuration.
    if (c.ntap == 1 && 2 * c.NT <= 320) c.ntap = 2;
    c.nsplit = 4 / c.ntap;  // tap groups
    c.MT = 1;
    c.nbuf = (2 * c.ntap * c.NT <= 512) ? 2 : 1;
  }
  c.KC = 16;
  for (int k = 48; k >= 16; k -= 16)
    if (Cin % k == 0) { c.KC = k; break; }
  c.nkc = Cin / c.KC;
  c.a_tile_bytes = (c.KC / 8) * 2048;
  c.b_stage_bytes = c.ntap * c.KC * c.NT * 2;
  const int a_per_stage = (mode == RS_DOWN ? 2 : 1) * c.MT;
  c.stage_bytes = (int)align_up((size_t)a_per_stage * c.a_tile_bytes + c.b_stage_bytes, 128);
  c.stages = (224 * 1024 - kRsHeader) / c.stage_bytes;
  if (c.stages > 8) c.stages = 8;
  if (c.stages < 2) return false;
  c.smem_bytes = kRsHeader + c.stages * c.stage_bytes;
  return true;
}

// DOWN: h_w = Conv2d weight [Cout][Cin][2][2].  UP: h_w = ConvTranspose2d weight [Cin][Cout][2][2].
int tc_resample_pack(int up, const float* h_w, int Cin, int Cout, int fmt, TcResampleWeights** out) {
  *out = nullptr;
  RsCfg c;
  if (!make_rs_cfg(up ? RS_UP : RS_DOWN, Cin, Cout, c)) return AC_OK;
  std::vector<h16> pack((size_t)4 * Cin * Cout);
  size_t o = 0;
  if (!up) {
    // [ns][dt][kc][df][KC/8][NT][8]
    for (int ns = 0; ns < c.nsplit; ++ns)
      for (int dt = 0; dt < 2; ++dt)
        for (int kc = 0; kc < c.nkc; ++kc)
          for (int df = 0; df < 2; ++df)
            for (int kg = 0; kg < c.KC / 8; ++kg)
              for (int n = 0; n < c.NT; ++n)
                for (int e = 0; e < 8; ++e) {
                  const int co = ns * c.NT + n, ci = kc * c.KC + kg * 8 + e;
                  pack[o++] = h16_rn(h_w[(((size_t)co * Cin + ci) * 2 + dt) * 2 + df], fmt);
                }
  } else {
    // [tap group][kc][tp][KC/8][NT][8]
    for (int tg = 0; tg < c.nsplit; ++tg)
      for (int kc = 0; kc < c.nkc; ++kc)
        for (int tp = 0; tp < c.ntap; ++tp)
          for (int kg = 0; kg < c.KC / 8; ++kg)
            for (int n = 0; n < c.NT; ++n)
              for (int e = 0; e < 8; ++e) {
                const int tap = tg * c.ntap + tp, ci = kc * c.KC + kg * 8 + e;
                pack[o++] = h16_rn(h_w[((size_t)ci * Cout + n) * 4 + tap], fmt);
              }
  }
  TcResampleWeights* w = new TcResampleWeights();
  w->cfg = c;
  w->fmt = fmt;
  w->d_pack = nullptr;
  if (cudaMalloc(&w->d_pack, pack.size() * 2) != cudaSuccess ||
      cudaMemcpy(w->d_pack, pack.data(), pack.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("resample weight upload failed");
    delete w;
    return AC_E_CUDA;
  }
  *out = w;
  return AC_OK;
}

void tc_resample_free(TcResampleWeights* w) {
  if (!w) return;
  if (w->d_pack) cudaFree(w->d_pack);
  delete w;
}

// All tensors CG8.  DOWN: in (2T x 2F, Cin) -> out (T x F, Cout).   UP: in (T x F, Cin); skip, out (2T x 2F, Cout).
int launch_tc_resample(const TcResampleWeights* w, const h16* in, const h16* skip, h16* out,
                       int nB, int T, int F, const float* scale, const float* shift, cudaStream_t st) {
  AC_REQUIRE(w && in && out, "tc resample: null");
  RsCfg c = w->cfg;
  const bool down = c.mode == RS_DOWN;
  AC_REQUIRE(down || skip, "tc resample: up needs the skip tensor");
  const int tiles = (F + 127) / 128;
  if (c.MT > tiles) {  // fewer tiles per unit: the packing does not depend on MT
    c.MT = tiles;
    const int a_per_stage = (down ? 2 : 1) * c.MT;
    c.stage_bytes = (int)align_up((size_t)a_per_stage * c.a_tile_bytes + c.b_stage_bytes, 128);
    c.stages = (224 * 1024 - kRsHeader) / c.stage_bytes;
    if (c.stages > 8) c.stages = 8;
    c.smem_bytes = kRsHeader + c.stages * c.stage_bytes;
    c.nbuf = down ? ((2 * c.MT * c.NT <= 512) ? 2 : 1) : c.nbuf;
  }
  EncodeTiledFn enc = get_tensor_map_encoder();
  AC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available");
  AC_REQUIRE(tc_abort_flag() != nullptr, "abort flag allocation failed");
  CUtensorMap map;
  CUresult r;
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (down) {
    // CG8 input [nB][2T][Cin/8][2F][8] as (c%8, df, f, c/8, row = b*2T + t'); box = [KC/8][128][1][8]
    const cuuint64_t dims[5] = {8, 2, (cuuint64_t)F, (cuuint64_t)(c.Cin / 8), (cuuint64_t)nB * 2 * T};
    const cuuint64_t strides[4] = {16, 32, (cuuint64_t)(2 * F) * 16, (cuuint64_t)(2 * F) * c.Cin * 2};
    const cuuint32_t box[5] = {8, 1, 128, (cuuint32_t)(c.KC / 8), 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<h16*>(in), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    // CG8 input [nB][T][Cin/8][F][8] as (c%8, f, c/8, t, b); box = [KC/8][128][8]
    const cuuint64_t dims[5] = {8, (cuuint64_t)F, (cuuint64_t)(c.Cin / 8), (cuuint64_t)T, (cuuint64_t)nB};
    const cuuint64_t strides[4] = {16, (cuuint64_t)F * 16, (cuuint64_t)F * c.Cin * 2, (cuuint64_t)T * F * c.Cin * 2};
    const cuuint32_t box[5] = {8, 128, (cuuint32_t)(c.KC / 8), 1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<h16*>(in), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (resample) failed with code " + std::to_string((int)r));
    return AC_E_CUDA;
  }
  RsParams p;
  p.cfg = c;
  p.nB = nB; p.T = T; p.F = F;
  p.n_fg = (tiles + c.MT - 1) / c.MT;
  p.n_units = c.nsplit * nB * T * p.n_fg;
  p.wpack = w->d_pack;
  p.scale = scale; p.shift = shift;
  p.skip = skip;
  p.out = out;
  p.abort_flag = tc_abort_flag();
  static bool attr = false;
  if (!attr) {
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_resample_kernel<kFmtF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_resample_kernel<kFmtBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  int grid = device_sm_count();
  if (grid > p.n_units) grid = p.n_units;
  const double pos = (double)nB * T * F;
  const double bytes = down ? pos * 2.0 * (4.0 * c.Cin + c.Cout) : pos * 2.0 * (c.Cin + 8.0 * c.Cout);
  ProfScope ps(KC_RESAMPLE_TC, 2.0 * pos * 4.0 * c.Cin * c.Cout, bytes, st);
  AC_CHECK_CUDA(tc_launch(w->fmt == kFmtBF16 ? tc_resample_kernel<kFmtBF16> : tc_resample_kernel<kFmtF16>, grid, kRsThreads, c.smem_bytes, st, 1, map, p));
  AC_LAUNCH_CHECK();
  return AC_OK;
}

}  // namespace ac
