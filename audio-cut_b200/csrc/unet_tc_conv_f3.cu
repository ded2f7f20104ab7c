// Fused TFC conv chain for the widest U-Net level (C = 48): conv3x3+BN+ReLU three times in ONE kernel,
// the two intermediate activations never leave the SM.
//
// Why: at C = 48 a single 3x3 conv moves 2 x 1.2 GB (B = 16 windows) for 522 GFLOP - the HBM/tensor knee - and its
// N = 48 MMA shape is shared-memory bound (44 cycles of operand reads for 24 of math).  The unfused block therefore
// costs 3 x 583 us.  Here
//   * every conv runs in the ROW-STACKED formulation of unet_tc_conv_ws.cu (tc_conv3x3_rs_kernel): the three vertical
//     taps are stacked along N = 144, input row r adds its contributions to the accumulators of output rows r+1, r,
//     r-1, which sit side by side in a ring of TMEM blocks.  9 tensor-bound MMAs per 128 positions instead of 27
//     shared-memory-bound ones, AND an input row is used by exactly one step, so each stage only needs a short FIFO
//     of rows instead of a 3-row window;
//   * conv c's epilogue writes its finished row (folded BN + ReLU, rounded to the 16-bit operand format exactly like
//     the unfused kernel's global store) straight into the next conv's A-operand FIFO in shared memory, in the
//     canonical no-swizzle K-major order [C/8][130][8] - which is also the CG8 global order, so the store is 512
//     contiguous bytes per warp instruction;
//   * zero padding of conv c+1 = rows / positions outside the image are written as zeros;
//   * horizontally a strip is ONE M = 128 tile: conv1 computes 128 positions from 130, conv2 126 valid from those,
//     conv3 124 valid: strips advance by 124 positions (3 % recompute, F = 3072 -> 25 strips);
//   * vertically a segment of `rows` output rows costs rows+6 / rows+4 / rows+2 steps of the three convs.
// Shared memory: 3 x 41.5 KB weights + (4 + 2 + 2) row tiles of 12.25 KB = 222 KB.  TMEM: rings of 4 + 3 + 3 blocks
// of 48 columns = 480.  Every conv is its own pipeline stage (one MMA-issuing warp + four epilogue warps), coupled to its
// neighbours only through the row FIFOs.
// Output is bit-identical to three launches of the weight-stationary kernel (same accumulation order per element).
// Roles: warp 0 = TMA producer, warps 1..3 = MMA issuers of conv 1..3 (warp 1 owns TMEM), warps 4..15 = epilogue (4 per conv).
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "tc_common.cuh"
#include "unet_kernels.cuh"

namespace ac {

constexpr int kF3C = 48;
constexpr int kF3EpiWarps = 12;                       // (conv) x (TMEM lane quadrant)
constexpr int kF3Threads = (4 + kF3EpiWarps) * 32;   // producer + 3 MMA issuers + 12 epilogue warps
constexpr int kF3Valid = 124;                         // output positions per strip
constexpr int kF3RowPos = 130;                        // rows of an A tile
constexpr int kF3ALbo = kF3RowPos * 16;               // bytes between 8-channel planes of a tile
constexpr int kF3ATile = ((kF3C / 8) * kF3ALbo + 127) / 128 * 128;
constexpr int kF3InSlots = 4;
constexpr int kF3MidSlots = 2;
constexpr int kF3BLbo = 3 * kF3C * 16;                // B rows = (dt, co): 144 rows of 16 B per 8-channel K group
constexpr int kF3DfBytes = (kF3C / 8) * kF3BLbo;      // one horizontal tap: [C/8][144][8]
constexpr int kF3WBytes = 3 * kF3DfBytes;             // one conv
constexpr int kF3Header = 2048;
constexpr int kF3Smem = kF3Header + 3 * kF3WBytes + (kF3InSlots + 2 * kF3MidSlots) * kF3ATile;
static_assert(kF3Smem <= 227 * 1024, "shared memory budget");

__host__ __device__ constexpr int f3_rb(int c) { return c == 0 ? 4 : 3; }           // TMEM blocks of conv c's ring
__host__ __device__ constexpr int f3_col(int c) { return c == 0 ? 0 : (c == 1 ? 4 * 48 : 7 * 48); }
constexpr int kF3Cols = 10 * 48;

struct F3Params {
  int nB, T, F;
  int n_strips;          // strips of kF3Valid positions per image row
  long long total_rows;  // nB * n_strips * T
  const h16* wpack;      // [conv][df][C/8][dt*48 + co][8]
  const float* scale[3];
  const float* shift[3];
  h16* out;
  int* abort_flag;
};

struct F3Seg {
  int b, f0, t0, t1;
};
__device__ __forceinline__ F3Seg f3_segment(const F3Params& p, long long L, long long hi) {
  F3Seg s;
  const long long strip = L / p.T;
  s.t0 = (int)(L - strip * p.T);
  const long long left = hi - L;
  s.t1 = (left < (long long)(p.T - s.t0)) ? s.t0 + (int)left : p.T;
  s.b = (int)(strip / p.n_strips);
  s.f0 = (int)(strip - (long long)s.b * p.n_strips) * kF3Valid;
  return s;
}

__device__ __forceinline__ uint64_t f3_desc_at(uint32_t lo_base, uint32_t hi, uint32_t byte_off) {
  return ((uint64_t)hi << 32) | (uint64_t)(lo_base + (byte_off >> 4));
}
__device__ __forceinline__ void f3_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}

// The 9 (df, k) MMAs of one input row: N1 columns at d0 (+ N2 columns at d1 with the B rows after the first N1 when the
// 3-block window wraps around the ring).  Straight-line code: per MMA two descriptor adds and the instruction.
template <int N1, int N2, int FMT>
__device__ __forceinline__ void f3_issue(uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t d0, uint32_t d1) {
  const uint32_t idesc1 = make_idesc<FMT>(N1);
#pragma unroll
  for (int df = 0; df < 3; ++df) {
#pragma unroll
    for (int k = 0; k < kF3C / 16; ++k) {
      const uint64_t ad = f3_desc_at(a_lo, a_hi, df * 16 + k * 2 * kF3ALbo);
      const uint32_t boff = df * kF3DfBytes + k * 2 * kF3BLbo;
      f3_umma(d0, ad, f3_desc_at(b_lo, b_hi, boff), idesc1);
      if constexpr (N2 > 0) f3_umma(d1, ad, f3_desc_at(b_lo, b_hi, boff + N1 * 16), make_idesc<FMT>(N2));
    }
  }
}

template <int FMT>
__global__ void __launch_bounds__(kF3Threads, 1)
tc_conv3x3_f3_kernel(const __grid_constant__ CUtensorMap in_map, const F3Params p) {
  constexpr int C = kF3C, NT = 48, K16 = C / 16;
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  uint64_t* in_full = reinterpret_cast<uint64_t*>(smem);   // [4]
  uint64_t* in_empty = in_full + 4;                         // [4]
  uint64_t* mid_full = in_full + 8;                         // [2 stages][2 slots]
  uint64_t* mid_empty = in_full + 12;                       // [2 stages][2 slots]
  uint64_t* done = in_full + 16;                            // [3 convs][4 blocks]  MMA -> epilogue: block complete
  uint64_t* bfree = in_full + 28;                           // [3 convs][4 blocks]  epilogue -> MMA: drained and zeroed
  uint64_t* wbar = in_full + 40;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_full + 41);
  float* s_ss = reinterpret_cast<float*>(smem + 512);       // [3][48] {scale, shift}
  uint8_t* w_smem = smem + kF3Header;
  uint8_t* in_ring = w_smem + 3 * kF3WBytes;
  uint8_t* mid_ring = in_ring + kF3InSlots * kF3ATile;      // stage 0 (conv1 -> conv2): 2 slots, stage 1: 2 slots
  volatile int* abort_flag = p.abort_flag;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long lo, hi;
  lo = p.total_rows * blockIdx.x / gridDim.x;
  hi = p.total_rows * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kF3InSlots; ++s) {
      mbar_init(&in_full[s], 1);
      mbar_init(&in_empty[s], 1);
    }
    for (int s = 0; s < 2 * kF3MidSlots; ++s) {
      mbar_init(&mid_full[s], 4);
      mbar_init(&mid_empty[s], 1);
    }
    for (int b = 0; b < 12; ++b) {
      mbar_init(&done[b], 1);
      mbar_init(&bfree[b], 4);
    }
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 3 * NT; i += blockDim.x) {
    const int j = i / NT, ch = i - j * NT;
    const float* sc = j == 0 ? p.scale[0] : (j == 1 ? p.scale[1] : p.scale[2]);
    const float* sh = j == 0 ? p.shift[0] : (j == 1 ? p.shift[1] : p.shift[2]);
    s_ss[2 * i] = sc[ch];
    s_ss[2 * i + 1] = sh[ch];
  }
  // the intermediate tiles start as zeros (rows 128, 129 of a tile are read by the last MMA rows and never written)
  for (int i = threadIdx.x; i < 2 * kF3MidSlots * kF3ATile / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(mid_ring)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
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
    // ===================== TMA producer: the three weight sets once, then one slot per input row ==========
    if (lane == 0) {
      mbar_expect_tx(wbar, (uint32_t)(3 * kF3WBytes));
      for (int i = 0; i < 9; ++i)
        bulk_load_1d(w_smem + i * kF3DfBytes, reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)i * kF3DfBytes, kF3DfBytes, wbar);
      int s = 0;
      uint32_t ph = 0;
      bool alive = true;
      for (long long L = lo; L < hi && alive;) {
        const F3Seg sg = f3_segment(p, L, hi);
        for (int r = sg.t0 - 3; r < sg.t1 + 3; ++r) {
          if (!mbar_wait(&in_empty[s], ph ^ 1, abort_flag)) { alive = false; break; }
          mbar_expect_tx(&in_full[s], (uint32_t)((C / 8) * kF3RowPos * 16));
          tma_load_5d(in_ring + (size_t)s * kF3ATile, &in_map, &in_full[s], 0, sg.f0 - 3, 0, r, sg.b);
          if (++s == kF3InSlots) { s = 0; ph ^= 1; }
        }
        L += sg.t1 - sg.t0;
      }
    }
  } else if (warp <= 3) {
    // ===================== MMA issuers: warp 1 + c issues conv c (warp-uniform loop, one elected lane issues) ==========
    // One issuing warp for all three convs was the bottleneck of the first build (ncu: 711 SASS instructions per macro
    // step of ~45 MMAs at ~8 cycles each, every other role waiting): the three convs are independent pipeline stages
    // coupled only by the row FIFOs, so each gets its own issuer, and the ring-wrap case is chosen OUTSIDE the unrolled
    // tap loops so that an MMA is two descriptor adds and the instruction.
    auto wait_all = [&](uint64_t* bar, uint32_t parity) {
      return __all_sync(0xffffffffu, mbar_wait(bar, parity, abort_flag)) != 0;
    };
    bool alive = wait_all(wbar, 0);
    // the epilogue warps zero the block rings before the first MMA (named barrier 1: 12 epilogue warps + 3 issuers)
    asm volatile("bar.sync 1, %0;" ::"r"((kF3EpiWarps + 3) * 32) : "memory");
    tc_fence_after();
    auto run = [&](auto ci) {
      constexpr int CI = decltype(ci)::value;
      constexpr int RB = f3_rb(CI);
      constexpr int kSlots = CI == 0 ? kF3InSlots : kF3MidSlots;
      const uint64_t a_proto = make_desc(0, kF3ALbo, 128), b_proto = make_desc(0, kF3BLbo, 128);
      const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
      const uint32_t a_lo0 = (uint32_t)a_proto + (smem_u32(CI == 0 ? in_ring : mid_ring + (CI - 1) * kF3MidSlots * kF3ATile) >> 4);
      const uint32_t b_lo = (uint32_t)b_proto + (smem_u32(w_smem + CI * kF3WBytes) >> 4);
      uint64_t* fullb = CI == 0 ? in_full : mid_full + (CI - 1) * kF3MidSlots;
      uint64_t* emptyb = CI == 0 ? in_empty : mid_empty + (CI - 1) * kF3MidSlots;
      const uint32_t ring0 = tmem_base + (uint32_t)f3_col(CI);
      int s = 0;          // FIFO slot read next
      uint32_t ph = 0;
      int gm = 0;         // virtual step g mod RB
      uint32_t cyc = 0;   // g / RB
      for (long long L = lo; L < hi && alive;) {
        const F3Seg sg = f3_segment(p, L, hi);
        const int rows = sg.t1 - sg.t0;
        const int real = rows + 6 - 2 * CI;  // input rows of this conv in the segment; + 2 flush steps
        for (int v = 0; v < real + 2 && alive; ++v) {
          // virtual step g: input row -> blocks of output rows g+1, g, g-1 = ring blocks sb, sb+1, sb+2 (mod RB)
          const int sb = RB - 1 - gm;
          {
            // row g+1 enters block sb: its previous owners (n of them) must have been drained and zeroed
            const uint32_t n = (gm + 2 >= RB) ? cyc + 1 : cyc;
            if (n > 0 && !wait_all(&bfree[CI * 4 + sb], (n - 1) & 1)) { alive = false; break; }
          }
          if (v < real) {
            if (!wait_all(&fullb[s], ph)) { alive = false; break; }
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_lo = a_lo0 + (uint32_t)s * (kF3ATile >> 4);
              const uint32_t d0 = ring0 + (uint32_t)(sb * NT);
              if (sb + 3 <= RB) {
                f3_issue<144, 0, FMT>(a_lo, a_hi, b_lo, b_hi, d0, ring0);
              } else if (sb + 2 == RB) {
                f3_issue<96, 48, FMT>(a_lo, a_hi, b_lo, b_hi, d0, ring0);
              } else {
                f3_issue<48, 96, FMT>(a_lo, a_hi, b_lo, b_hi, d0, ring0);
              }
              umma_commit(&emptyb[s]);
              umma_commit(&done[CI * 4 + (sb + 2) % RB]);
            }
            __syncwarp();
            if (++s == kSlots) { s = 0; ph ^= 1; }
          } else {
            if (elect_one()) umma_commit(&done[CI * 4 + (sb + 2) % RB]);
            __syncwarp();
          }
          if (++gm == RB) { gm = 0; ++cyc; }
        }
        L += rows;
      }
    };
    if (warp == 1) run(std::integral_constant<int, 0>{});
    else if (warp == 2) run(std::integral_constant<int, 1>{});
    else run(std::integral_constant<int, 2>{});
  } else {
    // ===================== epilogue (warps 4..15): four warps (one per TMEM lane quadrant) per conv ==========
    // A drain is a chain of latencies (barrier wake-up, tcgen05.ld, zeroing tcgen05.st, shared-memory stores, proxy
    // fence, arrive: ~1500 cycles); with every warp draining all three convs in turn that chain, three times per macro
    // step, was the critical path (first build: 2023 us against 1736 us unfused).  Specialised by conv the three drains
    // of a macro step run side by side and each thread carries all 48 channels of its position.
    const int quad = warp & 3;           // hardware rule: warp w may read TMEM lanes 32*(w%4) .. +31
    const int conv = (warp - 4) >> 2;    // which conv this warp drains
    const int mrow = quad * 32 + lane;   // MMA row = position within the tile
    const size_t plane = (size_t)p.F * 8;
    {
      // zero this conv's ring (4 warps x 32 lanes x all its columns)
      const uint32_t z = 0;
      const int c0 = conv == 0 ? f3_col(0) : (conv == 1 ? f3_col(1) : f3_col(2));
      const int c1 = conv == 0 ? f3_col(1) : (conv == 1 ? f3_col(2) : kF3Cols);
#pragma unroll 1
      for (int c = c0; c < c1; c += 16)
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(
                tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c),
            "r"(z)
            : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
    }
    asm volatile("bar.sync 1, %0;" ::"r"((kF3EpiWarps + 3) * 32) : "memory");

    auto run = [&](auto ci) {
      constexpr int CI = decltype(ci)::value;
      constexpr int RB = f3_rb(CI);
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)f3_col(CI);
      const float2* ss = reinterpret_cast<const float2*>(s_ss) + CI * NT;  // {scale, shift} per channel
      int Gm = 0;          // drained virtual steps: G mod RB, G / RB
      uint32_t Gc = 0;
      int ws_ = 0;         // FIFO slot written next (CI < 2)
      uint32_t wph = 0;
      bool alive = true;
      for (long long L = lo; L < hi && alive;) {
        const F3Seg sg = f3_segment(p, L, hi);
        const int rows = sg.t1 - sg.t0;
        // conv CI: rows + 6 - 2*CI real steps + 2 flush steps; step v completes output row t0 - 4 + CI + v; the rows
        // v in [2, real) are kept (conv1: t0-2 .. t1+1 feed conv2, conv2: t0-1 .. t1 feed conv3, conv3: t0 .. t1-1)
        const int real = rows + 6 - 2 * CI;
        const int pos = sg.f0 - 2 + CI + mrow;  // conv1's tile starts at f0 - 2, conv2's at f0 - 1, conv3's at f0
        for (int v = 0; v < real + 2 && alive; ++v) {
          const int t_out = sg.t0 - 4 + CI + v;
          const bool valid = v >= 2 && v < real;
          int blk = 1 - Gm;
          if (blk < 0) blk += RB;
          if (!mbar_wait(&done[CI * 4 + blk], Gc & 1, abort_flag)) { alive = false; break; }
          tc_fence_after();
          uint32_t r[NT];
          const uint32_t taddr = lane_addr + (uint32_t)(blk * NT);
          tmem_ld16(taddr, r);
          tmem_ld16(taddr + 16, r + 16);
          tmem_ld16(taddr + 32, r + 32);
          tmem_ld_wait();
          {
            const uint32_t z = 0;
#pragma unroll
            for (int j = 0; j < NT; j += 16)
              asm volatile(
                  "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + j),
                  "r"(z)
                  : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_relaxed(&bfree[CI * 4 + blk]);
          if (++Gm == RB) { Gm = 0; ++Gc; }
          if (!valid) continue;
          if constexpr (CI < 2) {
            // next conv's zero padding: nothing outside the image
            const bool inside = t_out >= 0 && t_out < p.T && pos >= 0 && pos < p.F;
            if (!mbar_wait(&mid_empty[CI * kF3MidSlots + ws_], wph ^ 1, abort_flag)) { alive = false; break; }
            uint8_t* dst = mid_ring + (size_t)(CI * kF3MidSlots + ws_) * kF3ATile + (size_t)mrow * 16;
#pragma unroll
            for (int cg = 0; cg < NT / 8; ++cg) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 q = *reinterpret_cast<const float4*>(ss + cg * 8 + 2 * e);  // {sc0, sh0, sc1, sh1}
                const float v0 = fmaxf(fmaf(__uint_as_float(r[cg * 8 + 2 * e]), q.x, q.y), 0.f);
                const float v1 = fmaxf(fmaf(__uint_as_float(r[cg * 8 + 2 * e + 1]), q.z, q.w), 0.f);
                pk[e] = inside ? pack2<FMT>(v0, v1) : 0u;
              }
              *reinterpret_cast<uint4*>(dst + cg * kF3ALbo) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&mid_full[CI * kF3MidSlots + ws_]);
            if (++ws_ == kF3MidSlots) { ws_ = 0; wph ^= 1; }
          } else {
            if (mrow < kF3Valid && pos < p.F) {
              h16* dst = p.out + cg8_index(sg.b, t_out, 0, pos, p.T, C, p.F);
#pragma unroll
              for (int cg = 0; cg < NT / 8; ++cg) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float4 q = *reinterpret_cast<const float4*>(ss + cg * 8 + 2 * e);
                  const float v0 = fmaxf(fmaf(__uint_as_float(r[cg * 8 + 2 * e]), q.x, q.y), 0.f);
                  const float v1 = fmaxf(fmaf(__uint_as_float(r[cg * 8 + 2 * e + 1]), q.z, q.w), 0.f);
                  pk[e] = pack2<FMT>(v0, v1);
                }
                *reinterpret_cast<uint4*>(dst + (size_t)cg * plane) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
        }
        L += rows;
      }
    };
    if (conv == 0) run(std::integral_constant<int, 0>{});
    else if (conv == 1) run(std::integral_constant<int, 1>{});
    else run(std::integral_constant<int, 2>{});
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two (batch, strip) columns side by side as ONE M = 256 tile.
//
// ncu on the single-CTA kernel above (profiles/r02_f3_ncu.md): tensor pipe 66 % active, L1/shared-memory 74 % - an N = 144
// SS-MMA reads 4 KB of A and 4.6 KB of B per 72 cycles of math (121 B/clk of the 128 B/clk the SM has), and a step whose
// 3-block accumulator window wraps around the TMEM ring is issued as N = 96 + N = 48 and reads A twice.  The pair removes
// both: each CTA reads only its half of B (72 of the 144 rows: 88 B/clk), and with rings of exactly three blocks - output
// row x ALWAYS in block x mod 3 - every step is one N = 144 MMA over blocks 0..2 whose weight rows come in one of three
// rotations of (dt0, dt1, dt2).  All three rotations are contiguous 144-row windows of the 240-row sequence
// E = (dt1, dt0, dt2, dt1, dt0), and the pair splits a window into rows [w, w+72) from CTA 0 and [w+72, w+144) from CTA 1
// AT THE SAME LOCAL ADDRESS, so CTA 0 stores E[0, 168), CTA 1 stores E[72, 240): 48 KB per conv and CTA.
//   leader (rank 0): three MMA-issuing warps (one per conv); its in_full / mid_full / bfree barriers collect both CTAs
//   commits are multicast to the in_empty / mid_empty / done barriers of both CTAs
//   each CTA: TMA producer for its own column, 12 epilogue warps draining its own 128 TMEM lanes into its own FIFOs
// Shared memory per CTA: 3 x 47.25 KB weights + (2 + 2 + 2) row tiles = 217 KB.  TMEM: 3 x 144 columns.
// ------------------------------------------------------------------------------------------------
constexpr int kF3pInSlots = 2;
constexpr int kF3pRows = 168;                          // rows of E a CTA keeps per (df, 8-channel K group)
constexpr int kF3pBLbo = kF3pRows * 16;
constexpr int kF3pDfBytes = (kF3C / 8) * kF3pBLbo;
constexpr int kF3pWBytes = 3 * kF3pDfBytes;            // one conv, one CTA
constexpr int kF3pSmem = kF3Header + 3 * kF3pWBytes + (kF3pInSlots + 2 * kF3MidSlots) * kF3ATile;
static_assert(kF3pSmem <= 227 * 1024, "shared memory budget (pair)");

struct F3pSeg {
  int b, f0, t0, t1;
  bool valid;  // false: the odd column of the last pair (no such column)
};
// rows are linearised over (column pair, t); column = b * n_strips + strip; this CTA takes column 2*pair + rank
__device__ __forceinline__ F3pSeg f3p_segment(const F3Params& p, long long L, long long hi, int rank) {
  F3pSeg s;
  const long long cp = L / p.T;
  s.t0 = (int)(L - cp * p.T);
  const long long left = hi - L;
  s.t1 = (left < (long long)(p.T - s.t0)) ? s.t0 + (int)left : p.T;
  const long long col = 2 * cp + rank;
  s.valid = col < (long long)p.nB * p.n_strips;
  s.b = (int)(col / p.n_strips);
  s.f0 = (int)(col - (long long)s.b * p.n_strips) * kF3Valid;
  return s;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait on a barrier that threads of the peer CTA arrive on (their writes must be visible: cluster-scope acquire)
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    if (mbar_try_wait_cluster(bar, parity)) return true;
    if ((it & 1023u) == 1023u && *abort_flag) return false;
  }
  *abort_flag = 1;
  return false;
}

template <int FMT>
__global__ void __launch_bounds__(kF3Threads, 1)
tc_conv3x3_f3p_kernel(const __grid_constant__ CUtensorMap in_map, const F3Params p) {
  constexpr int C = kF3C, NT = 48, K16 = C / 16, RB = 3;
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  uint64_t* in_full = reinterpret_cast<uint64_t*>(smem);   // [2]      leader: bytes of both CTAs' tiles
  uint64_t* in_empty = in_full + 4;                         // [2]      both (multicast commit)
  uint64_t* mid_full = in_full + 8;                         // [2][2]   own 4 epilogue warps (+ leader: the peer's forwarder)
  uint64_t* mid_empty = in_full + 12;                       // [2][2]   both (multicast commit)
  uint64_t* done = in_full + 16;                            // [3][4]   both (multicast commit): block complete
  uint64_t* bfree = in_full + 28;                           // [3][4]   leader: 4 epilogue warps x 2 CTAs: drained and zeroed
  uint64_t* wbar = in_full + 40;                            // own weights landed
  uint64_t* wbar_peer = in_full + 41;                       // leader: the peer's weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_full + 42);
  float* s_ss = reinterpret_cast<float*>(smem + 512);       // [3][48] {scale, shift}
  uint8_t* w_smem = smem + kF3Header;
  uint8_t* in_ring = w_smem + 3 * kF3pWBytes;
  uint8_t* mid_ring = in_ring + kF3pInSlots * kF3ATile;     // stage 0 (conv1 -> conv2): 2 slots, stage 1: 2 slots
  volatile int* abort_flag = p.abort_flag;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_pairs = gridDim.x >> 1;
  const long long lo = p.total_rows * (blockIdx.x >> 1) / n_pairs;
  const long long hi = p.total_rows * ((blockIdx.x >> 1) + 1) / n_pairs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kF3pInSlots; ++s) {
      mbar_init(&in_full[s], 1);
      mbar_init(&in_empty[s], 1);
    }
    for (int s = 0; s < 2 * kF3MidSlots; ++s) {
      mbar_init(&mid_full[s], leader ? 5 : 4);  // own 4 epilogue warps (+ leader: the peer's forwarder)
      mbar_init(&mid_empty[s], 1);
    }
    for (int b = 0; b < 12; ++b) {
      mbar_init(&done[b], 1);
      mbar_init(&bfree[b], 8);
    }
    mbar_init(wbar, 1);
    mbar_init(wbar_peer, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 3 * NT; i += blockDim.x) {
    const int j = i / NT, ch = i - j * NT;
    const float* sc = j == 0 ? p.scale[0] : (j == 1 ? p.scale[1] : p.scale[2]);
    const float* sh = j == 0 ? p.shift[0] : (j == 1 ? p.shift[1] : p.shift[2]);
    s_ss[2 * i] = sc[ch];
    s_ss[2 * i + 1] = sh[ch];
  }
  // the intermediate tiles start as zeros (rows 128, 129 of a tile are read by the last MMA rows and never written)
  for (int i = threadIdx.x; i < 2 * kF3MidSlots * kF3ATile / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(mid_ring)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= 4) {
    // every accumulator block starts as zeros: four warps (one per TMEM lane quadrant) per conv clear that conv's ring
    const uint32_t z = 0;
    const int conv = (warp - 4) >> 2;
#pragma unroll 1
    for (int c = conv * RB * NT; c < (conv + 1) * RB * NT; c += 16)
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(
              tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c),
          "r"(z)
          : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised and both accumulator rings zeroed before anything is signalled
  tc_fence_after();
  pdl_wait();  // everything above overlapped the previous kernel's tail; its output is visible from here on

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): own weight rows once, then one slot per input row of the own column ====
    if (lane == 0) {
      mbar_expect_tx(wbar, (uint32_t)(3 * kF3pWBytes));
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)rank * 3 * kF3pWBytes;
      for (int i = 0; i < 9; ++i) bulk_load_1d(w_smem + i * kF3pDfBytes, wsrc + (size_t)i * kF3pDfBytes, kF3pDfBytes, wbar);
      const uint32_t full0 = mapa_u32(smem_u32(&in_full[0]), 0);
      int s = 0;
      uint32_t ph = 0;
      bool alive = true;
      for (long long L = lo; L < hi && alive;) {
        const F3pSeg sg = f3p_segment(p, L, hi, (int)rank);
        for (int r = sg.t0 - 3; r < sg.t1 + 3; ++r) {
          if (!mbar_wait(&in_empty[s], ph ^ 1, abort_flag)) { alive = false; break; }
          if (leader) mbar_expect_tx(&in_full[s], 2u * (uint32_t)((C / 8) * kF3RowPos * 16));
          // a column that does not exist (odd column count) has b = nB: every coordinate out of range, the tile is zero-filled
          tma_load_5d_2sm(in_ring + (size_t)s * kF3ATile, &in_map, full0 + (uint32_t)s * 8, 0, sg.f0 - 3, 0, r, sg.b);
          if (++s == kF3pInSlots) { s = 0; ph ^= 1; }
        }
        L += sg.t1 - sg.t0;
      }
    }
  } else if (warp <= 3) {
    // ===================== MMA issuers (leader CTA): warp 1 + c issues conv c ==========
    // (cta-scope acquire: the leader never reads the peer's shared memory itself - the peer's tensor core does - and the
    // peer's forwarder released at cluster scope; a cluster-scope acquire in this polling loop costs a CCTL.IVALL per poll)
    auto wait_all = [&](uint64_t* bar, uint32_t parity) {
      return __all_sync(0xffffffffu, mbar_wait(bar, parity, abort_flag)) != 0;
    };
    bool alive = wait_all(wbar, 0);
    if (!leader) {
      if (warp == 1 && alive && lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(wbar_peer), 0));  // "my weights are in place"
      // Forwarders (warps 1, 2 of the peer, otherwise idle): the peer's epilogue warps signal a finished FIFO row on the
      // peer's LOCAL barrier (a cluster-scope release per epilogue warp and row showed up as MEMBAR / ERRBAR stalls, 20 % of
      // all samples); one thread here relays it to the leader's barrier, off everybody's critical path.
      if (warp <= 2 && lane == 0) {
        const int stage = warp - 1;
        const uint32_t remote0 = mapa_u32(smem_u32(&mid_full[stage * kF3MidSlots]), 0);
        int s = 0;
        uint32_t ph = 0;
        for (long long L = lo; L < hi && alive;) {
          const F3pSeg sg = f3p_segment(p, L, hi, 0);
          const int rows = sg.t1 - sg.t0;
          for (int e = 0; e < rows + 4 - 2 * stage && alive; ++e) {
            if (!mbar_wait(&mid_full[stage * kF3MidSlots + s], ph, abort_flag)) { alive = false; break; }
            mbar_arrive_cluster(remote0 + (uint32_t)s * 8);
            if (++s == kF3MidSlots) { s = 0; ph ^= 1; }
          }
          L += rows;
        }
      }
    } else {
      alive = alive && wait_all(wbar_peer, 0);
      auto run = [&](auto ci) {
        constexpr int CI = decltype(ci)::value;
        constexpr int kSlots = CI == 0 ? kF3pInSlots : kF3MidSlots;
        const uint32_t idesc = make_idesc_2sm<FMT>(3 * NT);
        const uint64_t a_proto = make_desc(0, kF3ALbo, 128), b_proto = make_desc(0, kF3pBLbo, 128);
        const uint32_t a_hi = (uint32_t)(a_proto >> 32), b_hi = (uint32_t)(b_proto >> 32);
        const uint32_t a_lo0 = (uint32_t)a_proto + (smem_u32(CI == 0 ? in_ring : mid_ring + (CI - 1) * kF3MidSlots * kF3ATile) >> 4);
        const uint32_t b_lo0 = (uint32_t)b_proto + (smem_u32(w_smem + CI * kF3pWBytes) >> 4);
        uint64_t* fullb = CI == 0 ? in_full : mid_full + (CI - 1) * kF3MidSlots;
        uint64_t* emptyb = CI == 0 ? in_empty : mid_empty + (CI - 1) * kF3MidSlots;
        const uint32_t ring0 = tmem_base + (uint32_t)(CI * RB * NT);
        int s = 0;          // FIFO slot read next
        uint32_t ph = 0;
        int gm = 0;         // virtual step g mod 3
        uint32_t cyc = 0;   // g / 3
        for (long long L = lo; L < hi && alive;) {
          const F3pSeg sg = f3p_segment(p, L, hi, 0);
          const int rows = sg.t1 - sg.t0;
          const int real = rows + 6 - 2 * CI;  // input rows of this conv in the segment; + 2 flush steps
          for (int v = 0; v < real + 2 && alive; ++v) {
            // virtual step g: the input row adds to output rows g+1 (dt 0), g (dt 1), g-1 (dt 2); row x lives in block x mod 3
            const int enter = gm == 2 ? 0 : gm + 1;   // block of row g+1
            const int compl_ = gm == 0 ? 2 : gm - 1;  // block of row g-1: complete after this step
            {
              // row g+1 enters: the block's previous owners (n of them) must have been drained and zeroed in BOTH CTAs
              const uint32_t n = (gm + 2 >= RB) ? cyc + 1 : cyc;
              if (n > 0 && !wait_all(&bfree[CI * 4 + enter], (n - 1) & 1)) { alive = false; break; }
            }
            if (v < real) {
              if (!wait_all(&fullb[s], ph)) { alive = false; break; }
              tc_fence_after();
              if (elect_one()) {
                const uint32_t a_lo = a_lo0 + (uint32_t)s * (kF3ATile >> 4);
                // weight rows in block order (dt of block 0, 1, 2): g mod 3 = 0 -> (1,0,2), 1 -> (2,1,0), 2 -> (0,2,1) =
                // the window of E = (dt1, dt0, dt2, dt1, dt0) starting at block 0, 2, 1
                const uint32_t b_lo = b_lo0 + (uint32_t)((gm == 0 ? 0 : (gm == 1 ? 2 : 1)) * NT);  // 48 rows x 16 B >> 4
#pragma unroll
                for (int df = 0; df < 3; ++df) {
#pragma unroll
                  for (int k = 0; k < K16; ++k) {
                    const uint64_t ad = f3_desc_at(a_lo, a_hi, df * 16 + k * 2 * kF3ALbo);
                    const uint64_t bd = f3_desc_at(b_lo, b_hi, df * kF3pDfBytes + k * 2 * kF3pBLbo);
                    umma_f16_2sm<true>(ring0, ad, bd, idesc);
                  }
                }
                umma_commit_2sm(&emptyb[s]);
                umma_commit_2sm(&done[CI * 4 + compl_]);
              }
              __syncwarp();
              if (++s == kSlots) { s = 0; ph ^= 1; }
            } else {
              if (elect_one()) umma_commit_2sm(&done[CI * 4 + compl_]);
              __syncwarp();
            }
            if (++gm == RB) { gm = 0; ++cyc; }
          }
          L += rows;
        }
      };
      if (warp == 1) run(std::integral_constant<int, 0>{});
      else if (warp == 2) run(std::integral_constant<int, 1>{});
      else run(std::integral_constant<int, 2>{});
    }
  } else {
    // ===================== epilogue (warps 4..15, both CTAs): four warps (one per TMEM lane quadrant) per conv ==========
    const int quad = warp & 3;
    const int conv = (warp - 4) >> 2;
    const int mrow = quad * 32 + lane;
    const size_t plane = (size_t)p.F * 8;
    const uint32_t bfree0 = mapa_u32(smem_u32(&bfree[0]), 0);
    auto run = [&](auto ci) {
      constexpr int CI = decltype(ci)::value;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(CI * RB * NT);
      const float2* ss = reinterpret_cast<const float2*>(s_ss) + CI * NT;
      int Gm = 0;          // drained virtual steps: G mod 3, G / 3
      uint32_t Gc = 0;
      int ws_ = 0;         // FIFO slot written next (CI < 2)
      uint32_t wph = 0;
      bool alive = true;
      for (long long L = lo; L < hi && alive;) {
        const F3pSeg sg = f3p_segment(p, L, hi, (int)rank);
        const int rows = sg.t1 - sg.t0;
        const int real = rows + 6 - 2 * CI;
        const int pos = sg.f0 - 2 + CI + mrow;
        for (int v = 0; v < real + 2 && alive; ++v) {
          const int t_out = sg.t0 - 4 + CI + v;
          const bool valid = v >= 2 && v < real;
          const int blk = Gm == 0 ? 2 : Gm - 1;  // row g-1 -> block (g-1) mod 3
          if (!mbar_wait(&done[CI * 4 + blk], Gc & 1, abort_flag)) { alive = false; break; }
          tc_fence_after();
          uint32_t r[NT];
          const uint32_t taddr = lane_addr + (uint32_t)(blk * NT);
          tmem_ld16(taddr, r);
          tmem_ld16(taddr + 16, r + 16);
          tmem_ld16(taddr + 32, r + 32);
          tmem_ld_wait();
          {
            const uint32_t z = 0;
#pragma unroll
            for (int j = 0; j < NT; j += 16)
              asm volatile(
                  "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + j),
                  "r"(z)
                  : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(bfree0 + (uint32_t)(CI * 4 + blk) * 8);
          if (++Gm == RB) { Gm = 0; ++Gc; }
          if (!valid) continue;
          if constexpr (CI < 2) {
            const bool inside = sg.valid && t_out >= 0 && t_out < p.T && pos >= 0 && pos < p.F;
            if (!mbar_wait(&mid_empty[CI * kF3MidSlots + ws_], wph ^ 1, abort_flag)) { alive = false; break; }
            uint8_t* dst = mid_ring + (size_t)(CI * kF3MidSlots + ws_) * kF3ATile + (size_t)mrow * 16;
#pragma unroll
            for (int cg = 0; cg < NT / 8; ++cg) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 q = *reinterpret_cast<const float4*>(ss + cg * 8 + 2 * e);  // {sc0, sh0, sc1, sh1}
                const float v0 = fmaxf(fmaf(__uint_as_float(r[cg * 8 + 2 * e]), q.x, q.y), 0.f);
                const float v1 = fmaxf(fmaf(__uint_as_float(r[cg * 8 + 2 * e + 1]), q.z, q.w), 0.f);
                pk[e] = inside ? pack2<FMT>(v0, v1) : 0u;
              }
              *reinterpret_cast<uint4*>(dst + cg * kF3ALbo) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&mid_full[CI * kF3MidSlots + ws_]);  // local; the peer's forwarder relays it
            if (++ws_ == kF3MidSlots) { ws_ = 0; wph ^= 1; }
          } else {
            if (sg.valid && mrow < kF3Valid && pos < p.F) {
              h16* dst = p.out + cg8_index(sg.b, t_out, 0, pos, p.T, C, p.F);
#pragma unroll
              for (int cg = 0; cg < NT / 8; ++cg) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float4 q = *reinterpret_cast<const float4*>(ss + cg * 8 + 2 * e);
                  const float v0 = fmaxf(fmaf(__uint_as_float(r[cg * 8 + 2 * e]), q.x, q.y), 0.f);
                  const float v1 = fmaxf(fmaf(__uint_as_float(r[cg * 8 + 2 * e + 1]), q.z, q.w), 0.f);
                  pk[e] = pack2<FMT>(v0, v1);
                }
                *reinterpret_cast<uint4*>(dst + (size_t)cg * plane) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
        }
        L += rows;
      }
    };
    if (conv == 0) run(std::integral_constant<int, 0>{});
    else if (conv == 1) run(std::integral_constant<int, 1>{});
    else run(std::integral_constant<int, 2>{});
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory and write its TMEM: nobody leaves early
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TcConvF3Weights {
  int fmt;
  h16* d_pack;
  h16* d_pack_pair = nullptr;  // [rank][conv][df][C/8][168 rows of E][8]
};

int tc_conv3x3_f3_supported(int T, int F, int C, int n_convs) {
  static const bool off = getenv("AC_NO_F3") && atoi(getenv("AC_NO_F3")) != 0;  // dev hook: A/B against the unfused kernels
  return (!off && C == kF3C && n_convs == 3 && T >= 1 && F >= kF3RowPos) ? AC_OK : AC_E_INVALID;
}

// h_w[j] = W_j[C][C][3][3] (conv j of the chain); packing per conv: [df][C/8][dt*48 + co][8]
int tc_conv3x3_f3_pack(const float* const h_w[3], int C, int fmt, TcConvF3Weights** out) {
  *out = nullptr;
  if (C != kF3C) return AC_OK;
  std::vector<h16> pack((size_t)3 * 9 * C * C);
  size_t o = 0;
  for (int j = 0; j < 3; ++j)
    for (int df = 0; df < 3; ++df)
      for (int kg = 0; kg < C / 8; ++kg)
        for (int dt = 0; dt < 3; ++dt)
          for (int co = 0; co < C; ++co)
            for (int e = 0; e < 8; ++e) pack[o++] = h16_rn(h_w[j][(((size_t)co * C + kg * 8 + e) * 3 + dt) * 3 + df], fmt);
  TcConvF3Weights* w = new TcConvF3Weights();
  w->fmt = fmt;
  w->d_pack = nullptr;
  if (cudaMalloc(&w->d_pack, pack.size() * 2) != cudaSuccess ||
      cudaMemcpy(w->d_pack, pack.data(), pack.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("tc f3 weight upload failed");
    if (w->d_pack) cudaFree(w->d_pack);
    delete w;
    return AC_E_CUDA;
  }
  // pair packing: E = (dt1, dt0, dt2, dt1, dt0) x 48 output channels; CTA `rank` keeps E[72 rank, 72 rank + 168)
  {
    static const int e_dt[5] = {1, 0, 2, 1, 0};
    std::vector<h16> pp((size_t)2 * 3 * 3 * (C / 8) * kF3pRows * 8);
    size_t q = 0;
    for (int rank = 0; rank < 2; ++rank)
      for (int j = 0; j < 3; ++j)
        for (int df = 0; df < 3; ++df)
          for (int kg = 0; kg < C / 8; ++kg)
            for (int row = 0; row < kF3pRows; ++row) {
              const int e_row = 72 * rank + row, dt = e_dt[e_row / C], co = e_row % C;
              for (int e = 0; e < 8; ++e) pp[q++] = h16_rn(h_w[j][(((size_t)co * C + kg * 8 + e) * 3 + dt) * 3 + df], fmt);
            }
    if (cudaMalloc(&w->d_pack_pair, pp.size() * 2) != cudaSuccess ||
        cudaMemcpy(w->d_pack_pair, pp.data(), pp.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("tc f3 pair weight upload failed");
      if (w->d_pack_pair) cudaFree(w->d_pack_pair);
      cudaFree(w->d_pack);
      delete w;
      return AC_E_CUDA;
    }
  }
  *out = w;
  return AC_OK;
}

void tc_conv3x3_f3_free(TcConvF3Weights* w) {
  if (!w) return;
  if (w->d_pack_pair) cudaFree(w->d_pack_pair);
  if (w->d_pack) cudaFree(w->d_pack);
  delete w;
}

int launch_tc_conv3x3_f3(const TcConvF3Weights* w, const h16* in, h16* out, int nB, int T, int F, const float* const scale[3],
                         const float* const shift[3], cudaStream_t st) {
  AC_REQUIRE(w && in && out, "tc f3 conv: null pointer");
  AC_REQUIRE(tc_conv3x3_f3_supported(T, F, kF3C, 3) == AC_OK, "tc f3 conv: unsupported shape");
  EncodeTiledFn enc = get_tensor_map_encoder();
  AC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available");
  AC_REQUIRE(tc_abort_flag() != nullptr, "abort flag allocation failed");
  const int C = kF3C;
  // 5-D view of the CG8 tensor [nB][T][C/8][F][8]: (c%8, f, c/8, t, b); one box = [C/8][130 positions][8 ch]
  CUtensorMap map;
  const cuuint64_t dims[5] = {8, (cuuint64_t)F, (cuuint64_t)(C / 8), (cuuint64_t)T, (cuuint64_t)nB};
  const cuuint64_t strides[4] = {16, (cuuint64_t)F * 16, (cuuint64_t)F * C * 2, (cuuint64_t)T * F * C * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)kF3RowPos, (cuuint32_t)(C / 8), 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<h16*>(in), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (f3 conv) failed with code " + std::to_string((int)r));
    return AC_E_CUDA;
  }
  F3Params p;
  p.nB = nB; p.T = T; p.F = F;
  p.n_strips = (F + kF3Valid - 1) / kF3Valid;
  p.total_rows = (long long)nB * p.n_strips * T;
  p.wpack = w->d_pack;
  for (int j = 0; j < 3; ++j) {
    p.scale[j] = scale[j];
    p.shift[j] = shift[j];
  }
  p.out = out;
  p.abort_flag = tc_abort_flag();
  // algorithmic work of the fused op: three convs' FLOPs, one read + one write of the activation
  ProfScope ps(KC_CONV_TC, 3 * 2.0 * 9.0 * nB * (double)T * F * C * C, 4.0 * nB * (double)T * F * C, st);
  static const bool pair_off = getenv("AC_F3_PAIR") && atoi(getenv("AC_F3_PAIR")) == 0;  // dev hook: single-CTA kernel
  if (!pair_off && w->d_pack_pair && device_sm_count() >= 2) {
    const long long cols = (long long)nB * p.n_strips;
    p.total_rows = ((cols + 1) / 2) * T;  // rows of column PAIRS
    p.wpack = w->d_pack_pair;
    auto pk = w->fmt == kFmtBF16 ? tc_conv3x3_f3p_kernel<kFmtBF16> : tc_conv3x3_f3p_kernel<kFmtF16>;
    AC_CHECK_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int pairs = device_sm_count() / 2;
    if ((long long)pairs > p.total_rows) pairs = (int)p.total_rows;
    AC_CHECK_CUDA(tc_launch(pk, 2 * pairs, kF3Threads, kF3pSmem, st, 2, map, p));
    AC_LAUNCH_CHECK();
    return AC_OK;
  }
  auto kern = w->fmt == kFmtBF16 ? tc_conv3x3_f3_kernel<kFmtBF16> : tc_conv3x3_f3_kernel<kFmtF16>;
  AC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int grid = device_sm_count();
  if ((long long)grid > p.total_rows) grid = (int)p.total_rows;
  AC_CHECK_CUDA(tc_launch(kern, grid, kF3Threads, kF3Smem, st, 1, map, p));
  AC_LAUNCH_CHECK();
  return AC_OK;
}

}  // namespace ac
