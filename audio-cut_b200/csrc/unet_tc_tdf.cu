// tcgen05 kernel for the TDF (time-distributed fully connected) layers: for every (b, t) row,
//   Y[b][t][m][c] = relu(scale[c] * sum_k W[m][k] * X[b][t][k][c] + shift[c]) (+ R[b][t][m][c])
// i.e. a GEMM over the frequency axis with channels riding along.  As UMMA:
//   D[128 out-features][N = NTt*C] += A[128][16] (weights, K-major) * B[16][N] (activations)
// Channels ride along N, so B is MN-major.  In the CG8 activation layout [B][T][C/8][F][8] a
// (b, t, channel group) plane is [F][8] = Kt*16 contiguous bytes per K chunk, which IS the canonical
// no-swizzle MN-major layout (16-byte rows, 8-row core matrices 128 B apart, SBO = Kt*16 between
// channel groups): ONE 4-D TMA box [NTt][C/8][Kt][8] stages a whole pipeline step.  Weights are
// pre-packed smem images fetched by 1-D bulk copies.  Pipeline / roles / TMEM hand-off as in
// unet_tc.cu; outputs (and the residual) are CG8 too, so a warp stores 512 contiguous bytes.
#include <vector>

#include "tc_common.cuh"
#include "unet_kernels.cuh"

namespace ac {

constexpr int kTdfEpiGroups = 3;  // epilogue warps per TMEM lane quadrant: group g owns the 16-channel chunks g, g+3, ...
constexpr int kTdfEpiWarps = 4 * kTdfEpiGroups;
constexpr int kTdfThreads = (2 + kTdfEpiWarps) * 32;
constexpr int kTdfHeader = 4096;   // barriers + scale/shift staged in shared memory

struct TdfCfg {
  int C, M, K;
  int NTt, N;            // time rows per unit, N = NTt*C
  int n_mtiles, mt, n_mg;  // 128-row tiles of M, tiles per unit, groups
  int Kt, nk;
  int stages, nbuf;
  int a_tile_bytes;   // 128*Kt*2
  int b_stage_bytes;  // N*Kt*2
  int stage_bytes;
  int smem_bytes;
};

struct TdfParams {
  TdfCfg cfg;
  int nB, T;
  int n_tg, n_units;
  const h16* wpack;
  const float* scale;
  const float* shift;
  const h16* residual;  // nullable
  h16* out;             // CG8 [nB][T][C/8][M][8]
  int* abort_flag;
};

template <int FMT>
__global__ void __launch_bounds__(kTdfThreads, 1) tc_tdf_kernel(const __grid_constant__ CUtensorMap in_map, const TdfParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  const TdfCfg& c = p.cfg;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 8;
  uint64_t* tfull = full + 16;
  uint64_t* tempty = full + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 24);
  float* s_scale = reinterpret_cast<float*>(smem + 1024);  // [C] (C <= 256)
  float* s_shift = s_scale + 256;
  uint8_t* stage0 = smem + kTdfHeader;
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
      mbar_init(&tempty[b], kTdfEpiWarps);
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


  auto decode = [&](int u, int& mg, int& b, int& t0) {
    mg = u % c.n_mg;
    int q = u / c.n_mg;
    t0 = (q % p.n_tg) * c.NTt;
    b = q / p.n_tg;
  };

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      bool alive = true;
      for (int u = blockIdx.x; u < p.n_units && alive; u += gridDim.x) {
        int mg, b, t0;
        decode(u, mg, b, t0);
        for (int kc = 0; kc < c.nk; ++kc) {
          if (!mbar_wait(&empty[s], ph ^ 1, abort_flag)) { alive = false; break; }
          uint8_t* st = stage0 + (size_t)s * c.stage_bytes;
          mbar_expect_tx(&full[s], (uint32_t)(c.mt * c.a_tile_bytes + c.b_stage_bytes));
          const h16* wsrc = p.wpack + ((size_t)mg * c.nk + kc) * (size_t)(c.mt * 128 * c.Kt);
          bulk_load_1d(st, wsrc, (uint32_t)(c.mt * c.a_tile_bytes), &full[s]);
          uint8_t* sb = st + c.mt * c.a_tile_bytes;
          tma_load_5d(sb, &in_map, &full[s], 0, kc * c.Kt, 0, t0, b);  // [NTt][C/8][Kt][8]
          if (++s == c.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform loop, one elected lane issues (keeps descriptors in uniform registers)
    {
      const uint32_t idesc = make_idesc_bmn<FMT>(c.N);
      const uint64_t a_proto = make_desc(0, 128 * 16, 128), b_proto = make_desc_mn(0, 128, (uint32_t)c.Kt * 16);
      auto wait_all = [&](uint64_t* bar, uint32_t parity) {
        return __all_sync(0xffffffffu, mbar_wait(bar, parity, abort_flag)) != 0;
      };
      int s = 0, buf = 0;
      uint32_t ph = 0, tph = 0;
      bool alive = true;
      for (int u = blockIdx.x; u < p.n_units && alive; u += gridDim.x) {
        if (!wait_all(&tempty[buf], tph ^ 1)) break;
        tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(buf * c.mt * c.N);
        for (int kc = 0; kc < c.nk; ++kc) {
          if (!wait_all(&full[s], ph)) { alive = false; break; }
          tc_fence_after();
          const uint32_t sa = smem_u32(stage0 + (size_t)s * c.stage_bytes);
          const uint32_t sb = sa + (uint32_t)(c.mt * c.a_tile_bytes);
          if (elect_one()) {
            for (int mi = 0; mi < c.mt; ++mi) {
              const uint64_t ad0 = a_proto + ((sa + mi * c.a_tile_bytes) >> 4);
              const uint64_t bd0 = b_proto + (sb >> 4);
              for (int k = 0; k < c.Kt / 16; ++k)
                umma_f16(acc0 + (uint32_t)(mi * c.N), ad0 + (uint64_t)(k * 2 * 128), bd0 + (uint64_t)(k * 16), idesc, (kc | k) != 0);
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
    // every time row, so all of a warp's column offsets are fixed for the whole kernel (no division or
    // 64-bit multiply in the unit loop), and three warps per scheduler hide each other's latencies.
    const int quad = warp & 3;        // hardware rule: warp w may read TMEM lanes 32*(w%4) .. +31
    const int grp = (warp - 2) >> 2;
    const int chunks_c = c.C >> 4;                                        // 16-channel chunks per time row
    const int per_row = chunks_c > grp ? (chunks_c - grp + kTdfEpiGroups - 1) / kTdfEpiGroups : 0;
    const int n_my = c.NTt * per_row;                                     // chunks this warp owns per accumulator tile
    constexpr int kMaxMy = 6;
    const size_t plane = (size_t)c.M * 8;                                 // elements between 8-channel groups (CG8)
    const size_t t_stride = (size_t)(c.C >> 3) * plane;                   // elements between time rows
    int my_col[kMaxMy], my_ch[kMaxMy];
    size_t my_off[kMaxMy];
#pragma unroll
    for (int i = 0; i < kMaxMy; ++i) {
      const int tl = per_row ? i / per_row : 0, k = per_row ? i - tl * per_row : 0;
      const int cq = grp + kTdfEpiGroups * k;
      my_ch[i] = cq * 16;
      my_col[i] = tl * c.C + cq * 16;
      my_off[i] = (size_t)tl * t_stride + (size_t)(cq * 2) * plane;
    }
    const bool fast = n_my <= kMaxMy;
    const bool pre_ok = fast && p.residual != nullptr && c.mt == 1;
    int buf = 0;
    uint32_t tph = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      int mg, b, t0;
      decode(u, mg, b, t0);
      const int m0 = mg * c.mt * 128 + quad * 32 + lane;
      const size_t base0 = cg8_index(b, t0, 0, m0, p.T, c.C, c.M);
      // The residual does not depend on the MMAs: request this warp's share of it before waiting for the
      // accumulators so the latency hides behind the MMA time.
      uint4 pre[2 * kMaxMy];
      if (pre_ok && m0 < c.M) {
#pragma unroll
        for (int i = 0; i < kMaxMy; ++i) {
          if (i < n_my) {
            pre[2 * i] = ldg_stream_u4(p.residual + base0 + my_off[i]);
            pre[2 * i + 1] = ldg_stream_u4(p.residual + base0 + my_off[i] + plane);
          }
        }
      }
      if (!mbar_wait(&tfull[buf], tph, abort_flag)) break;
      tc_fence_after();
      // one 16-column chunk: TMEM -> affine/ReLU (+ residual q0|q1) -> bf16 -> global (512 contiguous bytes per warp)
      auto chunk16 = [&](uint32_t taddr, int col, int ch0, bool valid, size_t idx, bool have_res, uint4 q0, uint4 q1) {
        uint32_t r[16];
        tmem_ld16(taddr + col, r);
        tmem_ld_wait();
        if (!valid) return;
        if (p.residual && !have_res) {
          q0 = *reinterpret_cast<const uint4*>(p.residual + idx);
          q1 = *reinterpret_cast<const uint4*>(p.residual + idx + plane);
        }
        const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int ch = ch0 + 2 * e;
          const float2 res = p.residual ? unpack2<FMT>(w[e]) : make_float2(0.f, 0.f);
          const float v0 = fmaxf(fmaf(__uint_as_float(r[2 * e]), s_scale[ch], s_shift[ch]), 0.f) + res.x;
          const float v1 = fmaxf(fmaf(__uint_as_float(r[2 * e + 1]), s_scale[ch + 1], s_shift[ch + 1]), 0.f) + res.y;
          pk[e] = pack2<FMT>(v0, v1);
        }
        *reinterpret_cast<uint4*>(p.out + idx) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(p.out + idx + plane) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      };
      for (int mi = 0; mi < c.mt; ++mi) {
        const bool valid = m0 + mi * 128 < c.M;
        const size_t base = base0 + (size_t)mi * 128 * 8;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * c.mt * c.N + mi * c.N);
        if (fast) {
#pragma unroll
          for (int i = 0; i < kMaxMy; ++i)
            if (i < n_my) chunk16(taddr, my_col[i], my_ch[i], valid, base + my_off[i], pre_ok, pre[2 * i], pre[2 * i + 1]);
        } else {
          const uint4 z = make_uint4(0, 0, 0, 0);
          for (int tl = 0; tl < c.NTt; ++tl)
            for (int cq = grp; cq < chunks_c; cq += kTdfEpiGroups)
              chunk16(taddr, tl * c.C + cq * 16, cq * 16, valid, base + (size_t)tl * t_stride + (size_t)(cq * 2) * plane, false, z, z);
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

struct TcTdfWeights {
  TdfCfg cfg;
  int fmt;
  h16* d_pack;
};

static bool make_tdf_cfg(int M, int K, int C, int T, TdfCfg& c) {
  if (C % 16 || C > 256 || K % 16 || M < 32) return false;  // the tiniest layers stay on the CUDA-core kernel
  c.C = C; c.M = M; c.K = K;
  c.n_mtiles = (M + 127) / 128;
  c.Kt = K % 64 == 0 ? 64 : (K % 32 == 0 ? 32 : 16);
  c.nk = K / c.Kt;
  auto pick_ntt = [&](int col_budget) {
    int ntt = 1;
    while (ntt * 2 <= 8 && (ntt * 2) * C <= 256 && (ntt * 2) * C <= col_budget && T % (ntt * 2) == 0) ntt *= 2;
    return ntt;
  };
  if (c.n_mtiles <= 4 && c.n_mtiles * C <= 512) {
    c.mt = c.n_mtiles;  // all of M in one unit: the activations are streamed exactly once
    c.NTt = pick_ntt(512 / c.mt);
    c.nbuf = (2 * c.mt * c.NTt * C <= 512) ? 2 : 1;
  } else {
    c.mt = 1;
    c.NTt = pick_ntt(256);
    c.nbuf = 2;
  }
  c.N = c.NTt * C;
  if (c.N % 16 || c.N > 256 || c.mt * c.N * c.nbuf > 512) return false;
  c.n_mg = (c.n_mtiles + c.mt - 1) / c.mt;
  c.a_tile_bytes = 128 * c.Kt * 2;
  c.b_stage_bytes = c.N * c.Kt * 2;
  c.stage_bytes = (int)align_up((size_t)c.mt * c.a_tile_bytes + c.b_stage_bytes, 128);
  c.stages = (224 * 1024 - kTdfHeader) / c.stage_bytes;
  if (c.stages > 8) c.stages = 8;
  if (c.stages < 2) return false;
  c.smem_bytes = kTdfHeader + c.stages * c.stage_bytes;
  return true;
}

int tc_tdf_pack(const float* h_w /*[M][K]*/, int M, int K, int C, int T, int fmt, TcTdfWeights** out) {
  *out = nullptr;
  TdfCfg c;
  if (!make_tdf_cfg(M, K, C, T, c)) return AC_OK;
  // [m_group][k_chunk][mt][Kt/8][128][8]; rows >= M are zero
  const size_t total = (size_t)c.n_mg * c.nk * c.mt * 128 * c.Kt;
  std::vector<h16> pack(total, h16_rn(0.f, fmt));
  size_t o = 0;
  for (int mg = 0; mg < c.n_mg; ++mg)
    for (int kc = 0; kc < c.nk; ++kc)
      for (int mi = 0; mi < c.mt; ++mi)
        for (int kg = 0; kg < c.Kt / 8; ++kg)
          for (int r = 0; r < 128; ++r)
            for (int e = 0; e < 8; ++e, ++o) {
              const int m = (mg * c.mt + mi) * 128 + r, k = kc * c.Kt + kg * 8 + e;
              if (m < M) pack[o] = h16_rn(h_w[(size_t)m * K + k], fmt);
            }
  TcTdfWeights* w = new TcTdfWeights();
  w->cfg = c;
  w->fmt = fmt;
  w->d_pack = nullptr;
  if (cudaMalloc(&w->d_pack, total * 2) != cudaSuccess ||
      cudaMemcpy(w->d_pack, pack.data(), total * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("tdf weight upload failed");
    delete w;
    return AC_E_CUDA;
  }
  *out = w;
  return AC_OK;
}

void tc_tdf_free(TcTdfWeights* w) {
  if (!w) return;
  if (w->d_pack) cudaFree(w->d_pack);
  delete w;
}

int launch_tc_tdf(const TcTdfWeights* w, const h16* in, const h16* residual, h16* out,
                  int nB, int T, const float* scale, const float* shift, cudaStream_t st) {
  AC_REQUIRE(w && in && out, "tc tdf: null");
  const TdfCfg& c = w->cfg;
  AC_REQUIRE(T % c.NTt == 0, "tc tdf: T not divisible by the time tile");
  EncodeTiledFn enc = get_tensor_map_encoder();
  AC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available");
  AC_REQUIRE(tc_abort_flag() != nullptr, "abort flag allocation failed");
  CUtensorMap map;
  // CG8 input [nB][T][C/8][K][8] as (c%8, k, c/8, t, b); one box = [NTt][C/8][Kt][8]
  const cuuint64_t dims[5] = {8, (cuuint64_t)c.K, (cuuint64_t)(c.C / 8), (cuuint64_t)T, (cuuint64_t)nB};
  const cuuint64_t strides[4] = {16, (cuuint64_t)c.K * 16, (cuuint64_t)c.K * c.C * 2, (cuuint64_t)T * c.K * c.C * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)c.Kt, (cuuint32_t)(c.C / 8), (cuuint32_t)c.NTt, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<h16*>(in), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (tdf) failed with code " + std::to_string((int)r));
    return AC_E_CUDA;
  }
  TdfParams p;
  p.cfg = c;
  p.nB = nB; p.T = T;
  p.n_tg = T / c.NTt;
  p.n_units = c.n_mg * p.n_tg * nB;
  p.wpack = w->d_pack;
  p.scale = scale; p.shift = shift;
  p.residual = residual;
  p.out = out;
  p.abort_flag = tc_abort_flag();
  static bool attr = false;
  if (!attr) {
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_tdf_kernel<kFmtF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AC_CHECK_CUDA(cudaFuncSetAttribute(tc_tdf_kernel<kFmtBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  int grid = device_sm_count();
  if (grid > p.n_units) grid = p.n_units;
  ProfScope ps(KC_TDF_TC, 2.0 * c.M * (double)c.K * c.C * T * nB, 2.0 * nB * (double)T * c.C * (c.K + c.M * (residual ? 2 : 1)), st);
  AC_CHECK_CUDA(tc_launch(w->fmt == kFmtBF16 ? tc_tdf_kernel<kFmtBF16> : tc_tdf_kernel<kFmtF16>, grid, kTdfThreads, c.smem_bytes, st, 1, map, p));
  AC_LAUNCH_CHECK();
  return AC_OK;
}

}  // namespace ac
