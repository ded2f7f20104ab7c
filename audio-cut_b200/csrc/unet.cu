// TFC-TDF U-Net: parameter upload / repacking and the forward schedule.
// Replaces self._session.run(...) at backends.py:358 (onnxruntime executing Kim_Vocal_1.onnx).
#include <algorithm>
#include <vector>

#include "unet_kernels.cuh"

namespace ac {

struct Affine {
  const float* scale;
  const float* shift;
};
// Everything 16-bit exists once per operand format (index kFmtF16 / kFmtBF16) and is built lazily by
// ensure_h16() on the first forward in that format: the arena copy and the per-layer tcgen05 packings.
struct ConvLayer {  // implicit-GEMM B operand [K][N]
  const float* w32;
  size_t w_off;      // arena offset of w32 (the 16-bit arena copies use the same offsets)
  size_t raw_off;    // offset of the layer's weights in the caller's blob order (host copy kept in ac_unet)
  int cin, cout;
  Affine af;
  int K, N;
  TcConvWeights* tc[2] = {nullptr, nullptr};      // tcgen05 packing (3x3 convs only)
  TcConvWsWeights* ws[2] = {nullptr, nullptr};    // weight-stationary tcgen05 packing (3x3 convs, C = 48 / 96)
  TcConvPairWeights* cp[2] = {nullptr, nullptr};  // CTA-pair streaming packing (3x3 convs, C >= 144)
  TcResampleWeights* rs[2] = {nullptr, nullptr};  // tcgen05 packing (down / up convs only)
};
struct TdfLayer {  // GEMM A operand [M][K] (PyTorch Linear weight as is)
  const float* w32;
  size_t w_off, raw_off;
  Affine af;
  int M, K;
  TcTdfWeights* tc[2] = {nullptr, nullptr};
  TcTdf2PairWeights* pair[2] = {nullptr, nullptr};   // CTA-pair kernel (second TDF layer of a block only)
  TcTdf1PairWeights* pair1[2] = {nullptr, nullptr};  // CTA-pair kernel (first TDF layer of a block only)
};
struct Block {
  int c, T, F;
  ConvLayer conv[8];
  TdfLayer tdf1, tdf2;
  TcConvF3Weights* f3[2] = {nullptr, nullptr};  // fused chain of the block's three 3x3 convs (C = 48 only)
};

}  // namespace ac

struct ac_unet {
  ac_unet_geom g;
  int n_blocks;
  std::vector<ac::Block> blocks;  // enc0..enc(n-1), bottleneck, dec0..dec(n-1)
  std::vector<ac::ConvLayer> ds, us;
  const float* first_w;
  ac::Affine first_af;
  const float* final_w;
  const float* final_b;
  float* d_f32 = nullptr;           // arena: all fp32 params (repacked)
  ac::h16* d_h16[2] = {nullptr, nullptr};  // arena: 16-bit copies (f16 / bf16), same offsets
  bool h16_ready[2] = {false, false};
  bool tc_ok[2] = {false, false};  // every layer has a CG8 / tcgen05 implementation: the 16-bit path runs on tensor cores
  std::vector<float> h_blob;        // the caller's parameter blob (the packers read the original weight order)
  size_t arena_floats = 0;
  int force_simt = 0;
  int device = 0;
};

namespace ac {

static size_t block_floats(const ac_unet_geom& g, int c, int f) {
  size_t s = 0;
  for (int j = 0; j < g.l; ++j) s += (size_t)c * c * 9 + 2 * c;
  s += (size_t)(f / g.bn) * f + 2 * c;
  s += (size_t)f * (f / g.bn) + 2 * c;
  return s;
}

static size_t param_floats(const ac_unet_geom& g) {
  size_t s = (size_t)g.g * g.dim_c + 2 * g.g;
  for (int i = 0; i < g.n; ++i) {
    int c = g.g * (i + 1), f = g.dim_f >> i;
    s += block_floats(g, c, f);
    s += (size_t)(c + g.g) * c * 4 + 2 * (c + g.g);
  }
  s += block_floats(g, g.g * (g.n + 1), g.dim_f >> g.n);
  for (int i = 0; i < g.n; ++i) {
    int lvl = g.n - 1 - i;
    int c = g.g * (lvl + 1), f = g.dim_f >> lvl;
    s += (size_t)(c + g.g) * c * 4 + 2 * c;
    s += block_floats(g, c, f);
  }
  s += (size_t)g.dim_c * g.g + g.dim_c;
  return s;
}

static bool geom_ok(const ac_unet_geom& g) {
  return g.dim_c == 4 && g.g > 0 && g.g % 16 == 0 && g.n >= 1 && g.n <= 6 && g.l >= 1 && g.l <= 8 && g.bn >= 1 &&
         g.dim_f % (g.bn << g.n) == 0 && g.dim_t % (1 << g.n) == 0 && g.dim_f > 0 && g.dim_t > 0;
}

}  // namespace ac

extern "C" size_t ac_unet_param_floats(const ac_unet_geom* g) { return (g && ac::geom_ok(*g)) ? ac::param_floats(*g) : 0; }

extern "C" int ac_unet_create(const ac_unet_geom* gp, const float* h_blob, size_t n_floats, ac_unet** out) {
  using namespace ac;
  AC_REQUIRE(gp && h_blob && out, "null pointer");
  AC_REQUIRE(geom_ok(*gp), "unsupported U-Net geometry");
  const ac_unet_geom g = *gp;
  AC_REQUIRE(n_floats == param_floats(g), "parameter blob has the wrong size");

  // Repack on the host into the device arena layout (same total size): conv weights become the
  // implicit-GEMM B operand [K][N]; everything else is copied through.
  std::vector<float> arena(n_floats);
  ac_unet* net = new ac_unet();
  net->g = g;
  net->arena_floats = n_floats;
  net->h_blob.assign(h_blob, h_blob + n_floats);
  cudaGetDevice(&net->device);
  size_t rd = 0, wr = 0;
  struct Fix {  // pointer fix-ups recorded as arena offsets
    size_t off;
  };
  auto take = [&](size_t n) {
    size_t o = wr;
    std::copy(h_blob + rd, h_blob + rd + n, arena.begin() + wr);
    rd += n;
    wr += n;
    return o;
  };
  // returns arena offset of the repacked [K][N] matrix
  auto conv3 = [&](int cin, int cout, int kh, int kw) {  // src W[cout][cin][kh][kw] -> [(tap*cin+ci)][cout]
    size_t o = wr;
    const float* src = h_blob + rd;
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < kh * kw; ++t)
          arena[o + ((size_t)t * cin + ci) * cout + co] = src[((size_t)co * cin + ci) * kh * kw + t];
    rd += (size_t)cout * cin * kh * kw;
    wr += (size_t)cout * cin * kh * kw;
    return o;
  };
  auto convT = [&](int cin, int cout) {  // src W[cin][cout][2][2] -> [ci][(tap*cout+co)]
    size_t o = wr;
    const float* src = h_blob + rd;
    for (int ci = 0; ci < cin; ++ci)
      for (int co = 0; co < cout; ++co)
        for (int t = 0; t < 4; ++t) arena[o + (size_t)ci * 4 * cout + (size_t)t * cout + co] = src[((size_t)ci * cout + co) * 4 + t];
    rd += (size_t)cin * cout * 4;
    wr += (size_t)cin * cout * 4;
    return o;
  };
  struct ConvOff { size_t w, sc, sh; int K, N; const float* raw; int cin, cout; };
  struct TdfOff { size_t w, sc, sh; int M, K; const float* raw; };
  struct BlockOff { int c, T, F; std::vector<ConvOff> conv; TdfOff t1, t2; };
  std::vector<BlockOff> boffs;
  std::vector<ConvOff> dsoffs, usoffs;

  auto read_block = [&](int c, int T, int F) {
    BlockOff b;
    b.c = c; b.T = T; b.F = F;
    for (int j = 0; j < g.l; ++j) {
      ConvOff co;
      co.raw = h_blob + rd;
      co.cin = co.cout = c;
      co.w = conv3(c, c, 3, 3);
      co.sc = take(c);
      co.sh = take(c);
      co.K = 9 * c; co.N = c;
      b.conv.push_back(co);
    }
    b.t1.raw = h_blob + rd;
    b.t1.w = take((size_t)(F / g.bn) * F); b.t1.sc = take(c); b.t1.sh = take(c); b.t1.M = F / g.bn; b.t1.K = F;
    b.t2.raw = h_blob + rd;
    b.t2.w = take((size_t)F * (F / g.bn)); b.t2.sc = take(c); b.t2.sh = take(c); b.t2.M = F; b.t2.K = F / g.bn;
    boffs.push_back(b);
  };

  size_t first_w = take((size_t)g.g * g.dim_c), first_sc = take(g.g), first_sh = take(g.g);
  for (int i = 0; i < g.n; ++i) {
    int c = g.g * (i + 1), T = g.dim_t >> i, F = g.dim_f >> i;
    read_block(c, T, F);
    ConvOff d;
    d.raw = h_blob + rd; d.cin = c; d.cout = c + g.g;
    d.w = conv3(c, c + g.g, 2, 2); d.sc = take(c + g.g); d.sh = take(c + g.g); d.K = 4 * c; d.N = c + g.g;
    dsoffs.push_back(d);
  }
  read_block(g.g * (g.n + 1), g.dim_t >> g.n, g.dim_f >> g.n);
  for (int i = 0; i < g.n; ++i) {
    int lvl = g.n - 1 - i;
    int c = g.g * (lvl + 1), T = g.dim_t >> lvl, F = g.dim_f >> lvl;
    ConvOff u;
    u.raw = h_blob + rd; u.cin = c + g.g; u.cout = c;
    u.w = convT(c + g.g, c); u.sc = take(c); u.sh = take(c); u.K = c + g.g; u.N = 4 * c;
    usoffs.push_back(u);
    read_block(c, T, F);
  }
  size_t final_w = take((size_t)g.dim_c * g.g), final_b = take(g.dim_c);
  if (rd != n_floats || wr != n_floats) {
    delete net;
    set_error("internal: blob walk mismatch");
    return AC_E_INVALID;
  }

  if (cudaMalloc(&net->d_f32, n_floats * 4) != cudaSuccess ||
      cudaMemcpy(net->d_f32, arena.data(), n_floats * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error(std::string("unet parameter upload: ") + cudaGetErrorString(cudaGetLastError()));
    ac_unet_destroy(net);
    return AC_E_CUDA;
  }
  auto mk_conv = [&](const ConvOff& o) {
    ConvLayer L;
    L.w32 = net->d_f32 + o.w; L.w_off = o.w; L.raw_off = (size_t)(o.raw - h_blob); L.cin = o.cin; L.cout = o.cout;
    L.af = Affine{net->d_f32 + o.sc, net->d_f32 + o.sh};
    L.K = o.K; L.N = o.N;
    return L;
  };
  auto mk_tdf = [&](const TdfOff& o) {
    TdfLayer L;
    L.w32 = net->d_f32 + o.w; L.w_off = o.w; L.raw_off = (size_t)(o.raw - h_blob);
    L.af = Affine{net->d_f32 + o.sc, net->d_f32 + o.sh};
    L.M = o.M; L.K = o.K;
    return L;
  };
  for (auto& bo : boffs) {
    Block b;
    b.c = bo.c; b.T = bo.T; b.F = bo.F;
    for (int j = 0; j < g.l; ++j) b.conv[j] = mk_conv(bo.conv[j]);
    b.tdf1 = mk_tdf(bo.t1);
    b.tdf2 = mk_tdf(bo.t2);
    net->blocks.push_back(b);
  }
  for (auto& o : dsoffs) net->ds.push_back(mk_conv(o));
  for (auto& o : usoffs) net->us.push_back(mk_conv(o));
  net->first_w = net->d_f32 + first_w;
  net->first_af = Affine{net->d_f32 + first_sc, net->d_f32 + first_sh};
  net->final_w = net->d_f32 + final_w;
  net->final_b = net->d_f32 + final_b;
  net->n_blocks = (int)net->blocks.size();
  *out = net;
  return AC_OK;
}

extern "C" void ac_unet_destroy(ac_unet* net) {
  if (!net) return;
  for (int f = 0; f < 2; ++f) {
    for (auto& b : net->blocks) {
      for (int j = 0; j < net->g.l; ++j) {
        if (b.conv[j].tc[f]) ac::tc_conv3x3_free(b.conv[j].tc[f]);
        if (b.conv[j].ws[f]) ac::tc_conv3x3_ws_free(b.conv[j].ws[f]);
        if (b.conv[j].cp[f]) ac::tc_conv3x3_pair_free(b.conv[j].cp[f]);
      }
      ac::tc_conv3x3_f3_free(b.f3[f]);
      ac::tc_tdf_free(b.tdf1.tc[f]);
      ac::tc_tdf_free(b.tdf2.tc[f]);
      ac::tc_tdf2_pair_free(b.tdf2.pair[f]);
      ac::tc_tdf1_pair_free(b.tdf1.pair1[f]);
    }
    for (auto& L : net->ds) ac::tc_resample_free(L.rs[f]);
    for (auto& L : net->us) ac::tc_resample_free(L.rs[f]);
    if (net->d_h16[f]) cudaFree(net->d_h16[f]);
  }
  if (net->d_f32) cudaFree(net->d_f32);
  delete net;
}

namespace ac {
template <int FMT>
__global__ void arena_to_h16_kernel(const float* __restrict__ src, h16* __restrict__ dst, size_t n) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    *reinterpret_cast<uint32_t*>(dst + i) = pack2<FMT>(src[i], src[i + 1]);
  } else if (i < n) {
    const uint32_t v = pack2<FMT>(src[i], 0.f);
    dst[i].bits = (uint16_t)(v & 0xffffu);
  }
}

// Builds, once per operand format, the 16-bit arena copy and every per-layer tcgen05 packing.  A pack that fails
// half-way is owned by the net already (the layer pointer is set by the packer's out-parameter), so
// ac_unet_destroy frees whatever exists.
static int ensure_h16(ac_unet* net, int fmt) {
  if (net->h16_ready[fmt]) return AC_OK;
  const ac_unet_geom& g = net->g;
  const size_t n = net->arena_floats;
  if (!net->d_h16[fmt]) {
    AC_CHECK_CUDA(cudaMalloc(&net->d_h16[fmt], (n + 2) * 2));
    const unsigned grid = (unsigned)((n / 2 + 1 + 255) / 256);
    if (fmt == kFmtBF16) arena_to_h16_kernel<kFmtBF16><<<grid, 256>>>(net->d_f32, net->d_h16[fmt], n);
    else arena_to_h16_kernel<kFmtF16><<<grid, 256>>>(net->d_f32, net->d_h16[fmt], n);
    AC_CHECK_CUDA(cudaDeviceSynchronize());
  }
  const float* blob = net->h_blob.data();
  int rc;
  for (auto& b : net->blocks) {
    for (int j = 0; j < g.l; ++j) {
      ConvLayer& L = b.conv[j];
      const float* raw = blob + L.raw_off;
      if (!L.tc[fmt] && tc_conv3x3_supported(b.T, b.F, b.c) == AC_OK && (rc = tc_conv3x3_pack(raw, b.c, fmt, &L.tc[fmt]))) return rc;
      if (!L.cp[fmt] && tc_conv3x3_pair_supported(b.T, b.F, b.c) == AC_OK && (rc = tc_conv3x3_pair_pack(raw, b.c, fmt, &L.cp[fmt]))) return rc;
      if (!L.ws[fmt] && tc_conv3x3_ws_supported(b.T, b.F, b.c) == AC_OK && (rc = tc_conv3x3_ws_pack(raw, b.c, fmt, &L.ws[fmt]))) return rc;
    }
    if (!b.f3[fmt] && g.l == 3 && tc_conv3x3_f3_supported(b.T, b.F, b.c, g.l) == AC_OK) {
      const float* raws[3] = {blob + b.conv[0].raw_off, blob + b.conv[1].raw_off, blob + b.conv[2].raw_off};
      if ((rc = tc_conv3x3_f3_pack(raws, b.c, fmt, &b.f3[fmt]))) return rc;
    }
    TdfLayer& t1 = b.tdf1;
    TdfLayer& t2 = b.tdf2;
    if (!t1.tc[fmt] && (rc = tc_tdf_pack(blob + t1.raw_off, t1.M, t1.K, b.c, b.T, fmt, &t1.tc[fmt]))) return rc;
    if (!t2.tc[fmt] && (rc = tc_tdf_pack(blob + t2.raw_off, t2.M, t2.K, b.c, b.T, fmt, &t2.tc[fmt]))) return rc;
    if (!t1.tc[fmt] && !t2.tc[fmt] && (t1.M < 32 || t1.M % 16)) {
      // Bottleneck width below a UMMA tile's granularity (Kim_Vocal level 4: F/8 = 24): run the pair on the tensor cores with the
      // hidden tensor padded to Mp rows - TDF1 gets Mp - M zero weight rows (its extra outputs are relu(shift), finite), TDF2 gets
      // Mp - M zero weight columns (so they contribute nothing).  Both layers or neither: they share the hidden layout.
      const int Mp = t1.M < 32 ? 32 : (t1.M + 15) / 16 * 16;
      std::vector<float> w1p((size_t)Mp * t1.K, 0.f), w2p((size_t)t2.M * Mp, 0.f);
      for (int m = 0; m < t1.M; ++m) std::copy(blob + t1.raw_off + (size_t)m * t1.K, blob + t1.raw_off + (size_t)(m + 1) * t1.K, w1p.begin() + (size_t)m * t1.K);
      for (int m = 0; m < t2.M; ++m) std::copy(blob + t2.raw_off + (size_t)m * t2.K, blob + t2.raw_off + (size_t)(m + 1) * t2.K, w2p.begin() + (size_t)m * Mp);
      if ((rc = tc_tdf_pack(w1p.data(), Mp, t1.K, b.c, b.T, fmt, &t1.tc[fmt]))) return rc;
      if ((rc = tc_tdf_pack(w2p.data(), t2.M, Mp, b.c, b.T, fmt, &t2.tc[fmt]))) return rc;
      if (!t1.tc[fmt] || !t2.tc[fmt]) {
        tc_tdf_free(t1.tc[fmt]);
        tc_tdf_free(t2.tc[fmt]);
        t1.tc[fmt] = t2.tc[fmt] = nullptr;
      }
    }
    if (!t2.pair[fmt] && (rc = tc_tdf2_pair_pack(blob + t2.raw_off, t2.M, t2.K, b.c, b.T, fmt, &t2.pair[fmt]))) return rc;
    if (!t1.pair1[fmt] && (rc = tc_tdf1_pair_pack(blob + t1.raw_off, t1.M, t1.K, b.c, b.T, fmt, &t1.pair1[fmt]))) return rc;
  }
  for (auto& L : net->ds)
    if (!L.rs[fmt] && (rc = tc_resample_pack(0, blob + L.raw_off, L.cin, L.cout, fmt, &L.rs[fmt]))) return rc;
  for (auto& L : net->us)
    if (!L.rs[fmt] && (rc = tc_resample_pack(1, blob + L.raw_off, L.cin, L.cout, fmt, &L.rs[fmt]))) return rc;
  bool ok = cg8_ends_supported(g.g) == AC_OK;
  for (auto& b : net->blocks)
    for (int j = 0; j < g.l; ++j) ok = ok && b.conv[j].tc[fmt] != nullptr;
  for (auto& L : net->ds) ok = ok && L.rs[fmt] != nullptr;
  for (auto& L : net->us) ok = ok && L.rs[fmt] != nullptr;
  net->tc_ok[fmt] = ok;
  net->h16_ready[fmt] = true;
  return AC_OK;
}
}  // namespace ac

namespace ac { int tc_check_abort(); }
extern "C" int ac_debug_tc_aborted(void) { return ac::tc_check_abort(); }

extern "C" int ac_unet_set_debug(ac_unet* net, int force_simt) {
  AC_REQUIRE(net, "null");
  net->force_simt = force_simt;
  return AC_OK;
}

// Test / profiling hook: one 3x3 conv layer (bf16, folded BN + ReLU) through a chosen implementation.
extern "C" int ac_debug_conv3x3(const void* d_in, void* d_out, int B, int T, int F, int C, const float* h_w,
                                const float* d_scale, const float* d_shift, int impl, int iters, float* h_ms,
                                void* stream) {
  using namespace ac;
  AC_REQUIRE(d_in && d_out && h_w && d_scale && d_shift && iters >= 1, "null pointer / iters");
  cudaStream_t st = (cudaStream_t)stream;
  const int fmt = (impl & 16) ? kFmtF16 : kFmtBF16;  // impl + 16: IEEE half operands
  impl &= 15;
  TcConvArgs ta{(const h16*)d_in, (h16*)d_out, B, T, F, C, nullptr, d_scale, d_shift};
  TcConvWeights* tc = nullptr;
  TcConvWsWeights* ws = nullptr;
  TcConvPairWeights* cp = nullptr;
  h16* d_w16 = nullptr;
  int rc = AC_OK;
  if (impl == 1) {
    AC_REQUIRE(tc_conv3x3_supported(T, F, C) == AC_OK, "streaming tc conv does not support this shape");
    if ((rc = tc_conv3x3_pack(h_w, C, fmt, &tc))) return rc;
    ta.w = tc;
  } else if (impl == 4) {
    AC_REQUIRE(tc_conv3x3_pair_supported(T, F, C) == AC_OK, "pair streaming tc conv does not support this shape");
    if ((rc = tc_conv3x3_pair_pack(h_w, C, fmt, &cp))) return rc;
  } else if (impl == 2 || impl == 3) {
    tc_conv3x3_ws_set_pair(impl == 3);
    tc_conv3x3_ws_set_rs(impl == 3);
    AC_REQUIRE(tc_conv3x3_ws_supported(T, F, C) == AC_OK, "ws tc conv does not support this shape");
    if ((rc = tc_conv3x3_ws_pack(h_w, C, fmt, &ws))) return rc;
  } else {
    std::vector<h16> w16((size_t)9 * C * C);  // [(tap*C+ci)][co]
    for (int co = 0; co < C; ++co)
      for (int ci = 0; ci < C; ++ci)
        for (int t = 0; t < 9; ++t) w16[((size_t)t * C + ci) * C + co] = h16_rn(h_w[((size_t)co * C + ci) * 9 + t], fmt);
    AC_CHECK_CUDA(cudaMalloc(&d_w16, w16.size() * 2));
    AC_CHECK_CUDA(cudaMemcpy(d_w16, w16.data(), w16.size() * 2, cudaMemcpyHostToDevice));
  }
  auto once = [&]() -> int {
    if (impl == 1) return launch_tc_conv3x3(ta, st);
    if (impl == 2 || impl == 3) return launch_tc_conv3x3_ws(ws, ta, st);
    if (impl == 4) return launch_tc_conv3x3_pair(cp, ta, st);
    GemmArgs a{};
    a.M = B * T * F; a.N = C; a.K = 9 * C; a.batch = 1;
    a.a_mode = A_CONV3; a.A = d_in; a.T = T; a.F = F; a.C = C;
    a.Bm = d_w16;
    a.epi = EPI_AFFINE_RELU; a.scale = d_scale; a.shift = d_shift; a.cmod = C; a.out = d_out;
    a.kclass = KC_CONV_SIMT;
    return launch_gemm_simt(a, fmt == kFmtF16 ? AC_F16 : AC_BF16, st);
  };
  rc = once();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (rc == AC_OK && iters > 1) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int i = 1; i < iters && rc == AC_OK; ++i) rc = once();
    cudaEventRecord(e1, st);
  }
  cudaError_t ce = cudaStreamSynchronize(st);
  if (e0) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (h_ms) *h_ms = ms / (iters - 1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  tc_conv3x3_ws_set_pair(1);
  tc_conv3x3_ws_set_rs(0);
  if (tc) tc_conv3x3_free(tc);
  if (ws) tc_conv3x3_ws_free(ws);
  if (cp) tc_conv3x3_pair_free(cp);
  if (d_w16) cudaFree(d_w16);
  if (rc == AC_OK && ce != cudaSuccess) {
    set_error(std::string("ac_debug_conv3x3: ") + cudaGetErrorString(ce));
    return AC_E_CUDA;
  }
  return rc;
}

// Test / profiling hook: a chain of three 3x3 conv layers (C = 48) on CG8 tensors; impl 0 = three launches of the
// weight-stationary kernel, 1 = the fused kernel (unet_tc_conv_f3.cu).  impl + 16: IEEE-half operands.  d_tmp = scratch of the
// tensor's size (impl 0 ping-pongs d_out / d_tmp).  h_w = 3 x W[C][C][3][3], d_scale / d_shift = 3 x [C].
extern "C" int ac_debug_conv3x3_chain(const void* d_in, void* d_out, void* d_tmp, int B, int T, int F, int C, const float* h_w,
                                      const float* d_scale, const float* d_shift, int impl, int iters, float* h_ms,
                                      void* stream) {
  using namespace ac;
  AC_REQUIRE(d_in && d_out && d_tmp && h_w && d_scale && d_shift && iters >= 1, "null pointer / iters");
  cudaStream_t st = (cudaStream_t)stream;
  const int fmt = (impl & 16) ? kFmtF16 : kFmtBF16;
  impl &= 15;
  const size_t wn = (size_t)C * C * 9;
  TcConvWsWeights* ws[3] = {nullptr, nullptr, nullptr};
  TcConvF3Weights* f3 = nullptr;
  int rc = AC_OK;
  if (impl == 1) {
    AC_REQUIRE(tc_conv3x3_f3_supported(T, F, C, 3) == AC_OK, "fused conv chain does not support this shape");
    const float* raws[3] = {h_w, h_w + wn, h_w + 2 * wn};
    if ((rc = tc_conv3x3_f3_pack(raws, C, fmt, &f3))) return rc;
  } else {
    AC_REQUIRE(tc_conv3x3_ws_supported(T, F, C) == AC_OK, "ws tc conv does not support this shape");
    for (int j = 0; j < 3 && rc == AC_OK; ++j) rc = tc_conv3x3_ws_pack(h_w + j * wn, C, fmt, &ws[j]);
  }
  auto once = [&]() -> int {
    if (impl == 1) {
      const float* sc[3] = {d_scale, d_scale + C, d_scale + 2 * C};
      const float* sh[3] = {d_shift, d_shift + C, d_shift + 2 * C};
      return launch_tc_conv3x3_f3(f3, (const h16*)d_in, (h16*)d_out, B, T, F, sc, sh, st);
    }
    const void* src = d_in;
    void* bufs[3] = {d_out, d_tmp, d_out};
    for (int j = 0; j < 3; ++j) {
      TcConvArgs ta{(const h16*)src, (h16*)bufs[j], B, T, F, C, nullptr, d_scale + j * C, d_shift + j * C};
      int r = launch_tc_conv3x3_ws(ws[j], ta, st);
      if (r) return r;
      src = bufs[j];
    }
    return AC_OK;
  };
  if (rc == AC_OK) rc = once();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (rc == AC_OK && iters > 1) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int i = 1; i < iters && rc == AC_OK; ++i) rc = once();
    cudaEventRecord(e1, st);
  }
  cudaError_t ce = cudaStreamSynchronize(st);
  if (e0) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (h_ms) *h_ms = ms / (iters - 1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  for (int j = 0; j < 3; ++j)
    if (ws[j]) tc_conv3x3_ws_free(ws[j]);
  if (f3) tc_conv3x3_f3_free(f3);
  if (rc == AC_OK && ce != cudaSuccess) {
    set_error(std::string("ac_debug_conv3x3_chain: ") + cudaGetErrorString(ce));
    return AC_E_CUDA;
  }
  return rc;
}

namespace ac {
// Workspace of one forward over B windows (elements of the activation type):
//   IO[i]   i = 1..n   level-i tensors for the whole batch: the encoder's input of level i, later reused for the decoder's
//                      output of level i (the encoder input is dead by then)
//   skip[i] i = 0..n-1 encoder outputs (multiplicative skips)
//   A, Bf, H           ping-pong scratch of the TFC convs and the TDF bottleneck, sized for the largest SUB-BATCH
// Sub-batching: level i runs sb[i] windows at a time through all the kernels of its block (first conv / up, 3 convs, TDF1,
// TDF2, down) before the next sb[i] windows start.  The idea - keep the layer-to-layer hand-over (75 MB per window at level 0
// in 16 bit) inside the 126 MB L2 instead of a 1.2 GB HBM round trip per layer - does NOT pay on B200: AC_UNET_SB=1 costs
// 17.3 ms per 16 windows against 16.3 ms for the whole batch (persistent kernels lose more on short grids than L2 gives
// back; two dies, 63 MB each).  Kept as a dev hook (AC_UNET_SB) with the negative result recorded; default sb = B.
struct WsPlan {
  size_t e[8];        // elements of one window's level-i tensor
  int sb[8];          // sub-batch of level i
  size_t io_off[8];   // IO[i], i >= 1
  size_t skip_off[8];
  size_t A, Bf, H, total;
};
static void default_sub_batches(const ac_unet_geom& g, int B, int dtype, int* sb) {
  // dev hook: AC_UNET_SB="1,2,4" = windows per pass at levels 0,1,2 (deeper levels: the whole batch)
  static const char* env = getenv("AC_UNET_SB");
  int cfg[8] = {B, B, B, B, B, B, B, B};
  (void)dtype;  // measured on B200 (profiles/r02_subbatch_sweep.log): the whole batch per pass is fastest - per-window passes
                // at level 0 (75 MB tensors, nominally L2-sized) are 6 % SLOWER, so the default stays sb = B everywhere
  if (env && *env) {
    int k = 0;
    const char* p = env;
    while (*p && k < 8) {
      int v = atoi(p);
      if (v > 0) cfg[k] = v;
      ++k;
      while (*p && *p != ',') ++p;
      if (*p == ',') ++p;
    }
  }
  for (int i = 0; i <= g.n; ++i) sb[i] = cfg[i] < 1 ? 1 : (cfg[i] > B ? B : cfg[i]);  // levels are independent passes
}
static WsPlan plan_ws(const ac_unet_geom& g, int B, int dtype) {
  WsPlan w{};
  default_sub_batches(g, B, dtype, w.sb);
  size_t off = 0;
  auto bump = [&](size_t n) {
    size_t o = off;
    off += (n + 127) / 128 * 128;
    return o;
  };
  size_t scratch = 0;
  for (int i = 0; i <= g.n; ++i) {
    w.e[i] = (size_t)(g.dim_t >> i) * (g.dim_f >> i) * g.g * (i + 1);
    if ((size_t)w.sb[i] * w.e[i] > scratch) scratch = (size_t)w.sb[i] * w.e[i];
  }
  w.A = bump(scratch);
  w.Bf = bump(scratch);
  w.H = bump(2 * (scratch / g.bn) + 128);  // x2: levels whose bottleneck is padded to a UMMA-friendly row count (ensure_h16)
  for (int i = 0; i < g.n; ++i) w.skip_off[i] = bump((size_t)B * w.e[i]);
  for (int i = 1; i <= g.n; ++i) w.io_off[i] = bump((size_t)B * w.e[i]);
  w.total = off;
  return w;
}
}  // namespace ac

extern "C" size_t ac_unet_workspace_bytes(const ac_unet* net, int B, int dtype) {
  if (!net || B <= 0) return 0;
  return ac::plan_ws(net->g, B, dtype).total * (dtype == AC_F32 ? 4 : 2) + 256;
}

namespace ac {
static int unet_forward_impl(ac_unet* net, const void* d_in, void* d_out, int B, int dtype, void* d_ws, size_t ws_bytes,
                             void* stream, bool first_done);
static bool unet_first_fusable(const ac_unet* net, int B, int dtype) {
  if (!net || dtype == AC_F32 || net->force_simt != 0) return false;
  const int fmt = fmt_of_dtype(dtype);
  if (!net->h16_ready[fmt] || !net->tc_ok[fmt]) return false;
  return plan_ws(net->g, B, dtype).sb[0] >= B;  // the whole batch goes through level 0 in one pass
}
// Where the network expects the output of its first 1x1 conv for a batch of B windows (CG8 [B][T][g/8][F][8]) and that
// conv's parameters - for a producer that applies it itself (the fused STFT epilogue, stft_mdx.cu).  AC_E_INVALID when
// this network / dtype / batch cannot take it (fp32, CUDA-core debug modes, sub-batched level 0).
int unet_first_conv_target(ac_unet* net, int B, int dtype, void* d_ws, size_t ws_bytes, void** d_target, const float** w,
                           const float** scale, const float** shift) {
  AC_REQUIRE(net && d_ws && d_target && w && scale && shift, "null pointer");
  if (dtype != AC_F32) {
    int rc0 = ensure_h16(net, fmt_of_dtype(dtype));
    if (rc0) return rc0;
  }
  if (!unet_first_fusable(net, B, dtype)) return AC_E_INVALID;
  if (ws_bytes < ac_unet_workspace_bytes(net, B, dtype)) {
    set_error("unet workspace too small");
    return AC_E_WORKSPACE;
  }
  const WsPlan wp = plan_ws(net->g, B, dtype);
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(d_ws) + 255) & ~uintptr_t(255));
  *d_target = base + wp.Bf * 2;
  *w = net->first_w;
  *scale = net->first_af.scale;
  *shift = net->first_af.shift;
  return AC_OK;
}
int unet_base_channels(const ac_unet* net) { return net ? net->g.g : 0; }
int unet_forward_after_first(ac_unet* net, void* d_out, int B, int dtype, void* d_ws, size_t ws_bytes, cudaStream_t st) {
  AC_REQUIRE(unet_first_fusable(net, B, dtype), "the first conv of this network / dtype cannot be applied by the caller");
  return unet_forward_impl(net, d_out, d_out, B, dtype, d_ws, ws_bytes, (void*)st, true);
}
}  // namespace ac

extern "C" int ac_unet_forward(ac_unet* net, const void* d_in, void* d_out, int B, int dtype, void* d_ws,
                               size_t ws_bytes, void* stream) {
  return ac::unet_forward_impl(net, d_in, d_out, B, dtype, d_ws, ws_bytes, stream, false);
}

static int ac::unet_forward_impl(ac_unet* net, const void* d_in, void* d_out, int B, int dtype, void* d_ws, size_t ws_bytes,
                                 void* stream, bool first_done) {
  using namespace ac;
  AC_REQUIRE(net && d_in && d_out && d_ws, "null pointer");
  AC_REQUIRE(B > 0, "batch must be positive");
  AC_REQUIRE(dtype == AC_F32 || dtype == AC_BF16 || dtype == AC_F16, "dtype");
  const int fmt = fmt_of_dtype(dtype);
  if (dtype != AC_F32) {
    int rc0 = ensure_h16(net, fmt);
    if (rc0) return rc0;
  }
  if (ws_bytes < ac_unet_workspace_bytes(net, B, dtype)) {
    set_error("unet workspace too small");
    return AC_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const ac_unet_geom& g = net->g;
  const size_t es = dtype == AC_F32 ? 4 : 2;
  const WsPlan wp = plan_ws(g, B, dtype);
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(d_ws) + 255) & ~uintptr_t(255));
  auto ptr = [&](size_t off) { return (void*)(base + off * es); };
  auto at = [&](void* p, size_t elems) { return (void*)((char*)p + elems * es); };
  auto cat = [&](const void* p, size_t elems) { return (const void*)((const char*)p + elems * es); };
  void* A = ptr(wp.A);
  void* Bf = ptr(wp.Bf);
  void* H = ptr(wp.H);
  auto wsel = [&](const float* w32, size_t w_off) { return dtype == AC_F32 ? (const void*)w32 : (const void*)(net->d_h16[fmt] + w_off); };
  int rc;

  // The 16-bit formats run on the tensor-core path (all activations in the CG8 layout) unless the geometry has a layer
  // without a tcgen05 kernel or the test hook forces the CUDA-core kernels (channels-last layout).
  const bool use_tc = dtype != AC_F32 && net->tc_ok[fmt] && net->force_simt != 1;

  auto conv3x3 = [&](const ConvLayer& L, const void* x, void* y, int T, int F, int C, int nb) -> int {
    if (use_tc) {
      TcConvArgs ta{(const h16*)x, (h16*)y, nb, T, F, C, L.tc[fmt], L.af.scale, L.af.shift};
      if (L.ws[fmt] && net->force_simt != 2 && tc_conv3x3_ws_supported(T, F, C) == AC_OK) return launch_tc_conv3x3_ws(L.ws[fmt], ta, st);
      if (L.cp[fmt] && net->force_simt != 2 && tc_conv3x3_pair_supported(T, F, C) == AC_OK) return launch_tc_conv3x3_pair(L.cp[fmt], ta, st);
      return launch_tc_conv3x3(ta, st);
    }
    GemmArgs a{};
    a.M = nb * T * F; a.N = L.N; a.K = L.K; a.batch = 1;
    a.a_mode = A_CONV3; a.A = x; a.T = T; a.F = F; a.C = C;
    a.Bm = wsel(L.w32, L.w_off);
    a.epi = EPI_AFFINE_RELU; a.scale = L.af.scale; a.shift = L.af.shift; a.cmod = L.N; a.out = y;
    a.kclass = KC_CONV_SIMT;
    return launch_gemm_simt(a, dtype, st);
  };
  // out = relu(affine(W * in)) (+ residual)
  auto tdf = [&](const TdfLayer& L, const Block& b, const void* in, const void* residual, void* out, int nb) -> int {
    if (use_tc) {
      if (L.pair[fmt] && residual && net->force_simt != 2)
        return launch_tc_tdf2_pair(L.pair[fmt], (const h16*)in, (const h16*)residual, (h16*)out, nb, b.T,
                                   L.af.scale, L.af.shift, st);
      if (L.pair1[fmt] && !residual && net->force_simt != 2)
        return launch_tc_tdf1_pair(L.pair1[fmt], (const h16*)in, (h16*)out, nb, b.T, L.af.scale, L.af.shift, st);
      if (L.tc[fmt])
        return launch_tc_tdf(L.tc[fmt], (const h16*)in, (const h16*)residual, (h16*)out, nb, b.T,
                             L.af.scale, L.af.shift, st);
      return launch_tdf_small_cg8((const h16*)in, net->d_h16[fmt] + L.w_off, (const h16*)residual, (h16*)out, nb, b.T,
                                  b.c, L.M, L.K, L.af.scale, L.af.shift, fmt, st);
    }
    GemmArgs a{};
    a.M = L.M; a.N = b.c; a.K = L.K; a.batch = nb * b.T;
    a.a_mode = A_PLAIN; a.A = wsel(L.w32, L.w_off); a.a_batch_stride = 0;
    a.Bm = in; a.b_batch_stride = (long long)L.K * b.c;
    a.epi = residual ? EPI_RESIDUAL : EPI_AFFINE_RELU; a.scale = L.af.scale; a.shift = L.af.shift; a.cmod = b.c;
    a.out = out; a.c_batch_stride = (long long)L.M * b.c; a.extra = residual;
    a.kclass = KC_TDF_SIMT;
    return launch_gemm_simt(a, dtype, st);
  };
  // One TFC-TDF block on nb windows: the l convs go in -> A -> Bf -> A ...; `in` may alias Bf (it is dead after the first
  // conv) but not A; the result lands in Z (anything but the conv chain's last buffer).  out4 != nullptr: the network's
  // last block - the final 1x1 conv is applied in the TDF2 epilogue when the kernel can, else by a separate launch.
  auto run_block = [&](const Block& b, const void* in, void* Z, int nb, void* out4) -> int {
    const void* src = in;
    void* dst = A;
    if (use_tc && net->force_simt == 0 && g.l == 3 && b.f3[fmt] && tc_conv3x3_f3_supported(b.T, b.F, b.c, g.l) == AC_OK) {
      // level 0: the three convs in one kernel, intermediates in shared memory (bit-identical to the chain below)
      const float* sc[3] = {b.conv[0].af.scale, b.conv[1].af.scale, b.conv[2].af.scale};
      const float* sh[3] = {b.conv[0].af.shift, b.conv[1].af.shift, b.conv[2].af.shift};
      if ((rc = launch_tc_conv3x3_f3(b.f3[fmt], (const h16*)src, (h16*)dst, nb, b.T, b.F, sc, sh, st))) return rc;
      src = dst;
    } else {
      for (int j = 0; j < g.l; ++j) {
        if ((rc = conv3x3(b.conv[j], src, dst, b.T, b.F, b.c, nb))) return rc;
        src = dst;
        dst = dst == A ? Bf : A;
      }
    }
    const void* tfc = src;
    if (Z == tfc) return (set_error("internal: block output aliases TFC output"), AC_E_INVALID);
    if ((rc = tdf(b.tdf1, b, tfc, nullptr, H, nb))) return rc;
    if (out4) {
      if (use_tc && net->force_simt == 0 && b.tdf2.pair[fmt] && tc_tdf2_pair_can_fuse_final(b.tdf2.pair[fmt]) && b.c == g.g &&
          b.T == g.dim_t && b.tdf2.M == g.dim_f)
        return launch_tc_tdf2_pair(b.tdf2.pair[fmt], (const h16*)H, (const h16*)tfc, (h16*)out4, nb, b.T, b.tdf2.af.scale,
                                   b.tdf2.af.shift, st, net->final_w, net->final_b);
      if ((rc = tdf(b.tdf2, b, H, tfc, Z, nb))) return rc;
      if (use_tc) return launch_final_conv_cg8(Z, out4, (long long)nb * g.dim_t, g.dim_f, g.g, net->final_w, net->final_b, fmt, st);
      return launch_final_conv(Z, out4, (long long)nb * g.dim_t * g.dim_f, g.g, net->final_w, net->final_b, dtype, st);
    }
    return tdf(b.tdf2, b, H, tfc, Z, nb);
  };
  auto down = [&](const ConvLayer& d, const Block& b, const void* skip, void* out, int nb) -> int {
    if (use_tc) return launch_tc_resample(d.rs[fmt], (const h16*)skip, nullptr, (h16*)out, nb, b.T / 2, b.F / 2, d.af.scale, d.af.shift, st);
    GemmArgs a{};
    a.M = nb * (b.T / 2) * (b.F / 2); a.N = d.N; a.K = d.K; a.batch = 1;
    a.a_mode = A_DOWN2; a.A = skip; a.T = b.T / 2; a.F = b.F / 2; a.C = b.c;
    a.Bm = wsel(d.w32, d.w_off);
    a.epi = EPI_AFFINE_RELU; a.scale = d.af.scale; a.shift = d.af.shift; a.cmod = d.N; a.out = out;
    a.kclass = KC_RESAMPLE_SIMT;
    return launch_gemm_simt(a, dtype, st);
  };
  // b = the block at the OUTPUT resolution; in = the level below
  auto up = [&](const ConvLayer& u, const Block& b, const void* in, const void* skip, void* out, int nb) -> int {
    if (use_tc) return launch_tc_resample(u.rs[fmt], (const h16*)in, (const h16*)skip, (h16*)out, nb, b.T / 2, b.F / 2, u.af.scale, u.af.shift, st);
    GemmArgs a{};
    a.M = nb * (b.T / 2) * (b.F / 2); a.N = u.N; a.K = u.K; a.batch = 1;
    a.a_mode = A_PLAIN; a.A = in; a.a_batch_stride = 0;
    a.Bm = wsel(u.w32, u.w_off);
    a.epi = EPI_UP_SKIP; a.scale = u.af.scale; a.shift = u.af.shift; a.cmod = b.c; a.out = out; a.extra = skip;
    a.up_T = b.T / 2; a.up_F = b.F / 2;
    a.kclass = KC_RESAMPLE_SIMT;
    return launch_gemm_simt(a, dtype, st);
  };

  const size_t e_spec = (size_t)g.dim_t * g.dim_f * g.dim_c;  // one window of the input / output spectrogram
  // ---- encoder: level i turns IO[i] (level 0: the spectrogram through the first 1x1 conv) into skip[i] and IO[i+1]
  for (int i = 0; i < g.n; ++i) {
    const Block& b = net->blocks[i];
    void* skip = ptr(wp.skip_off[i]);
    void* next = ptr(wp.io_off[i + 1]);
    for (int b0 = 0; b0 < B; b0 += wp.sb[i]) {
      const int nb = B - b0 < wp.sb[i] ? B - b0 : wp.sb[i];
      const void* in;
      if (i == 0) {
        const void* spec = cat(d_in, (size_t)b0 * e_spec);
        if (first_done)
          rc = AC_OK;  // the caller's producer (fused STFT epilogue) has already written Bf
        else if (use_tc)
          rc = launch_first_conv_cg8(spec, Bf, (long long)nb * g.dim_t, g.dim_f, g.g, net->first_w, net->first_af.scale,
                                     net->first_af.shift, fmt, st);
        else
          rc = launch_first_conv(spec, Bf, (long long)nb * g.dim_t * g.dim_f, g.g, net->first_w, net->first_af.scale,
                                 net->first_af.shift, dtype, st);
        if (rc) return rc;
        in = Bf;
      } else {
        in = at(ptr(wp.io_off[i]), (size_t)b0 * wp.e[i]);
      }
      void* sk = at(skip, (size_t)b0 * wp.e[i]);
      if ((rc = run_block(b, in, sk, nb, nullptr))) return rc;
      if ((rc = down(net->ds[i], b, sk, at(next, (size_t)b0 * wp.e[i + 1]), nb))) return rc;
    }
  }
  // ---- bottleneck: IO[n] -> IO[n] (the input is dead after the first conv)
  {
    const Block& b = net->blocks[g.n];
    void* io = ptr(wp.io_off[g.n]);
    for (int b0 = 0; b0 < B; b0 += wp.sb[g.n]) {
      const int nb = B - b0 < wp.sb[g.n] ? B - b0 : wp.sb[g.n];
      void* s = at(io, (size_t)b0 * wp.e[g.n]);
      if ((rc = run_block(b, s, s, nb, nullptr))) return rc;
    }
  }
  // ---- decoder: level lvl = n-1 .. 0: up(IO[lvl+1]) * skip[lvl] -> block -> IO[lvl] (level 0: the output spectrogram)
  for (int i = 0; i < g.n; ++i) {
    const int lvl = g.n - 1 - i;
    const Block& b = net->blocks[g.n + 1 + i];
    const ConvLayer& u = net->us[i];
    void* skip = ptr(wp.skip_off[lvl]);
    void* below = ptr(wp.io_off[lvl + 1]);
    for (int b0 = 0; b0 < B; b0 += wp.sb[lvl]) {
      const int nb = B - b0 < wp.sb[lvl] ? B - b0 : wp.sb[lvl];
      if ((rc = up(u, b, at(below, (size_t)b0 * wp.e[lvl + 1]), at(skip, (size_t)b0 * wp.e[lvl]), Bf, nb))) return rc;
      if (lvl == 0) {
        // Z (only used when the final conv cannot be fused): the skip slice is dead after the up-conv consumed it
        if ((rc = run_block(b, Bf, at(skip, (size_t)b0 * wp.e[0]), nb, at(d_out, (size_t)b0 * e_spec)))) return rc;
      } else {
        if ((rc = run_block(b, Bf, at(ptr(wp.io_off[lvl]), (size_t)b0 * wp.e[lvl]), nb, nullptr))) return rc;
      }
    }
  }
  return AC_OK;
}
