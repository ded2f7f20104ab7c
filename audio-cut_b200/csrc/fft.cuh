// Shared-memory mixed-radix (2,3,4,5,8) Stockham FFT, complex fp32, one transform per CTA.
// Ping-pong between two padded smem buffers; twiddles come from a device table
// exp(-2*pi*i*k/n) computed in double on the host (ac::get_fft_plan).
#pragma once
#include "common.cuh"

namespace ac {

struct FftDev {
  int n;
  int n_radix;
  int radix[12];
  unsigned magic[12];  // ceil(2^32 / Ns) of every pass: j / Ns == __umulhi(j, magic) for j < 2^16, Ns <= 2^15
  const float2* tw;
  const float2* tw_lo;  // exp(-2*pi*i*m/n), m < 64
  const float2* tw_hi;  // exp(-2*pi*i*64*h/n), h < ceil(n/64)
};

inline FftDev make_fft_dev(const FftPlan* p) {
  FftDev d;
  d.n = p->n;
  d.n_radix = p->n_radix;
  unsigned long long ns = 1;
  for (int i = 0; i < 12; ++i) {
    d.radix[i] = i < p->n_radix ? p->radix[i] : 1;
    d.magic[i] = ns == 1 ? 0u : (unsigned)((0x100000000ull + ns - 1) / ns);
    ns *= (unsigned long long)d.radix[i];
  }
  d.tw = p->d_twiddle;
  d.tw_lo = p->d_tw_lo;
  d.tw_hi = p->d_tw_hi;
  return d;
}

// padded smem index: one extra float2 every 32 keeps strided butterfly writes off one bank
__device__ __host__ __forceinline__ int fpad(int i) { return i + (i >> 5); }
inline size_t fft_smem_floats2(int n) { return (size_t)fpad(n) + 1; }

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <int R, bool INV>
__device__ __forceinline__ void dft_small(float2* v) {
  if constexpr (R == 2) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  } else if constexpr (R == 3) {
    const float s60 = 0.86602540378443864676f;
    float2 t1 = cadd(v[1], v[2]);
    float2 m = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
    float2 d = csub(v[1], v[2]);
    float2 s = make_float2(s60 * d.x, s60 * d.y);
    float2 ms = mul_mi<INV>(s);  // -i*s forward
    v[0] = cadd(v[0], t1);
    v[1] = cadd(m, ms);
    v[2] = csub(m, ms);
  } else if constexpr (R == 4) {
    float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    float2 t2 = cadd(v[1], v[3]), t3 = mul_mi<INV>(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[2] = csub(t0, t2);
    v[1] = cadd(t1, t3);
    v[3] = csub(t1, t3);
  } else if constexpr (R == 5) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    float2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    float2 m1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    float2 m2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    float2 n1 = mul_mi<INV>(make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y));
    float2 n2 = mul_mi<INV>(make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y));
    v[0] = cadd(v[0], cadd(a1, a2));
    v[1] = cadd(m1, n1);
    v[4] = csub(m1, n1);
    v[2] = cadd(m2, n2);
    v[3] = csub(m2, n2);
  } else if constexpr (R == 8) {
    const float h = 0.70710678118654752440f;
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    dft_small<4, INV>(e);
    dft_small<4, INV>(o);
    // o[k] *= w8^k ; w8 = exp(-+ 2*pi*i/8)
    float2 w1 = INV ? make_float2(h, h) : make_float2(h, -h);
    float2 w3 = INV ? make_float2(-h, h) : make_float2(-h, -h);
    o[1] = cmul(o[1], w1);
    o[2] = mul_mi<INV>(o[2]);
    o[3] = cmul(o[3], w3);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = cadd(e[k], o[k]);
      v[k + 4] = csub(e[k], o[k]);
    }
  }
}

// The R-1 twiddles W^(r*ts), r = 1..R-1, of one butterfly.
//   n <= 4096 (the 2048-point feature FFTs): gathered from the exactly rounded n-entry table, which is L1-sized there; the
//     log-domain features of near-silent bins need that accuracy (flatness at 1e-4 relative).
//   larger n (the 6144 / 7680-point MDX frames): the 61 KB table does not stay in L1 next to 3 x 61 KB of shared memory and
//     the gathers cost one L2 sector each (4.4 GB of L2 -> L1 traffic per 16-window STFT launch); W^ts comes from two small
//     L1-resident tables (one complex multiply) and its powers by products of depth <= log2 R (<= 5 ulp; STFT and iSTFT stay
//     >= 110 / 100 dB against torch).  STFT 0.51 -> 0.38 ms, iSTFT 0.76 -> 0.56 ms per 16 windows.
template <int R, bool INV>
__device__ __forceinline__ void butterfly_twiddles(const FftDev& p, int ts, float2* w /*[R], w[0] unused*/) {
  if (p.n <= 4096) {
#pragma unroll
    for (int r = 1; r < R; ++r) {
      float2 v = __ldg(&p.tw[r * ts]);
      if (INV) v.y = -v.y;
      w[r] = v;
    }
    return;
  }
  float2 w1 = cmul(__ldg(p.tw_lo + (ts & 63)), __ldg(p.tw_hi + (ts >> 6)));
  if (INV) w1.y = -w1.y;
  w[1] = w1;
#pragma unroll
  for (int r = 2; r < R; ++r) w[r] = cmul(w[r >> 1], w[r - (r >> 1)]);
}

template <int R, bool INV>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ in, float2* __restrict__ out, int n, int Ns,
                                         unsigned magic, const FftDev& p) {
  const int nb = n / R;
  const int tmul = n / (Ns * R);
  for (int j = threadIdx.x; j < nb; j += blockDim.x) {
    const int q = Ns == 1 ? j : (int)__umulhi((unsigned)j, magic);  // j / Ns without a division
    const int k = j - q * Ns;
    const int ts = k * tmul;
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = in[fpad(j + r * nb)];
    if (k != 0) {
      float2 w[R];
      butterfly_twiddles<R, INV>(p, ts, w);
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmul(v[r], w[r]);
    }
    dft_small<R, INV>(v);
    const int j0 = q * Ns * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) out[fpad(j0 + r * Ns)] = v[r];
  }
}

// Transforms the n points in buf0 (index through fpad); returns the buffer holding the result
// in natural order.  All threads of the CTA must call it; ends with a __syncthreads().
template <bool INV>
__device__ __forceinline__ float2* fft_smem(float2* buf0, float2* buf1, const FftDev& p) {
  float2* in = buf0;
  float2* out = buf1;
  int Ns = 1;
  for (int s = 0; s < p.n_radix; ++s) {
    const int R = p.radix[s];
    switch (R) {
      case 2: fft_pass<2, INV>(in, out, p.n, Ns, p.magic[s], p); break;
      case 3: fft_pass<3, INV>(in, out, p.n, Ns, p.magic[s], p); break;
      case 4: fft_pass<4, INV>(in, out, p.n, Ns, p.magic[s], p); break;
      case 5: fft_pass<5, INV>(in, out, p.n, Ns, p.magic[s], p); break;
      default: fft_pass<8, INV>(in, out, p.n, Ns, p.magic[s], p); break;
    }
    __syncthreads();
    float2* t = in;
    in = out;
    out = t;
    Ns *= R;
  }
  return in;
}

// ---- in-place variant: ONE shared-memory buffer --------------------------------------------------
// A Stockham pass reads and writes different index patterns, so the two-buffer version above ping-pongs.
// Here every thread first pulls ALL the butterflies it owns in this pass into registers (N / blockDim
// complex values, <= 16), the CTA synchronises, and the results go back into the same buffer.  Halving
// the footprint (61 KB instead of 123 KB at n_fft = 7680) is what lets three STFT CTAs share an SM and
// hide each other's shared-memory / twiddle latency.
template <int R, bool INV, int MAXB>
__device__ __forceinline__ void fft_pass_inplace(float2* __restrict__ buf, int n, int Ns, unsigned magic,
                                                 const FftDev& p) {
  const int nb = n / R;
  const int tmul = n / (Ns * R);
  float2 v[MAXB][R];
#pragma unroll
  for (int i = 0; i < MAXB; ++i) {
    const int j = threadIdx.x + i * blockDim.x;
    if (j < nb) {
      const int k = Ns == 1 ? 0 : j - (int)__umulhi((unsigned)j, magic) * Ns;
      const int ts = k * tmul;
#pragma unroll
      for (int r = 0; r < R; ++r) v[i][r] = buf[fpad(j + r * nb)];
      if (k != 0) {
        float2 w[R];
        butterfly_twiddles<R, INV>(p, ts, w);
#pragma unroll
        for (int r = 1; r < R; ++r) v[i][r] = cmul(v[i][r], w[r]);
      }
      dft_small<R, INV>(v[i]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MAXB; ++i) {
    const int j = threadIdx.x + i * blockDim.x;
    if (j < nb) {
      const int q = Ns == 1 ? j : (int)__umulhi((unsigned)j, magic);
      const int k = j - q * Ns;
      const int j0 = q * Ns * R + k;
#pragma unroll
      for (int r = 0; r < R; ++r) buf[fpad(j0 + r * Ns)] = v[i][r];
    }
  }
  __syncthreads();
}

// true when fft_smem_inplace can run this plan with `threads` threads per CTA (register staging bound)
inline bool fft_inplace_ok(const FftPlan* p, int threads) {
  for (int s = 0; s < p->n_radix; ++s) {
    const int R = p->radix[s];
    const int maxb = R == 2 ? 8 : (R == 3 ? 5 : (R == 4 ? 4 : (R == 5 ? 3 : 2)));
    if ((p->n / R + threads - 1) / threads > maxb) return false;
  }
  return true;
}

// In-place transform of the n points in buf (index through fpad); result in natural order in buf.
template <bool INV>
__device__ __forceinline__ void fft_smem_inplace(float2* buf, const FftDev& p) {
  int Ns = 1;
  for (int s = 0; s < p.n_radix; ++s) {
    const int R = p.radix[s];
    switch (R) {
      case 2: fft_pass_inplace<2, INV, 8>(buf, p.n, Ns, p.magic[s], p); break;
      case 3: fft_pass_inplace<3, INV, 5>(buf, p.n, Ns, p.magic[s], p); break;
      case 4: fft_pass_inplace<4, INV, 4>(buf, p.n, Ns, p.magic[s], p); break;
      case 5: fft_pass_inplace<5, INV, 3>(buf, p.n, Ns, p.magic[s], p); break;
      default: fft_pass_inplace<8, INV, 2>(buf, p.n, Ns, p.magic[s], p); break;
    }
    Ns *= R;
  }
}

}  // namespace ac
