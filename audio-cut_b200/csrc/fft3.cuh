// Three-pass shared-memory FFT for the MDX frame sizes (7680 = 16 x 20 x 24, 6144 = 16 x 16 x 24), complex fp32, one
// transform per CTA, ONE padded shared-memory buffer.
//
// The generic mixed-radix code in fft.cuh runs 5 passes of radix <= 8 over these sizes: 206 SASS instructions per point,
// issue- and shared-memory-bound (ncu: profiles/r02_stft_ncu_full.md).  Here every thread owns ONE butterfly per pass and does
// the whole radix-16 / 20 / 24 DFT in registers (two-level Cooley-Tukey out of the 3/4/5/8-point kernels, inner twiddles are
// compile-time constants, trivial ones cost nothing), so a frame crosses shared memory three times instead of five and
//   * pass 1 (Ns = 1, no twiddles) takes its inputs from a caller-supplied loader - the STFT reads the track through the
//     window descriptor and applies the Hann window there, the iSTFT builds the Hermitian extension from the kept bins - so
//     the frame is never staged in shared memory first;
//   * pass 2's twiddles W_{R1 R2}^{k r} come from a [R1][R2] table (<= 2.5 KB), pass 3's W_N^{k r} from a [R1 R2][R3] table
//     (61 KB, read as R3 consecutive values per thread): exactly rounded, no products of twiddles;
//   * pass 3 reads and writes the same addresses (k + r Ns), so it needs no barrier between its loads and stores and its
//     results can go to a caller-supplied sink (the iSTFT's windowed overlap-add) straight from registers;
//   * all shared-memory offsets along r are compile-time constants (the strides are multiples of 32, so the padding term of
//     fpad() separates).
// Stockham autosort, decimation in time: pass with radix R and Ns = product of the earlier radices maps
//   in[j + r N/R], r < R  (x W_{Ns R}^{k r}, k = j mod Ns)  --R-point DFT-->  out[(j / Ns) Ns R + k + r Ns].
#pragma once
#include <utility>

#include "fft.cuh"

namespace ac {

template <int N>
struct Fft3Geom;
template <>
struct Fft3Geom<7680> {
  static constexpr int R1 = 16, P1 = 4, Q1 = 4, R2 = 20, P2 = 4, Q2 = 5, R3 = 24, P3 = 8, Q3 = 3;
};
template <>
struct Fft3Geom<6144> {
  static constexpr int R1 = 16, P1 = 4, Q1 = 4, R2 = 16, P2 = 4, Q2 = 4, R3 = 24, P3 = 8, Q3 = 3;
};
constexpr bool fft3_supported(int n) { return n == 7680 || n == 6144; }
constexpr int kFft3Threads = 512;  // >= N / R for every pass of both sizes (480 / 384 / 320 and 384 / 384 / 256)

// ---- compile-time sine / cosine of 2*pi*m/r (Taylor series in double on an angle reduced to [-pi/4, pi/4]) -------------
constexpr double f3c_pi = 3.14159265358979323846264338327950288;
constexpr double f3c_sin_t(double x) {
  double term = x, sum = x;
  for (int i = 1; i < 14; ++i) {
    term *= -x * x / ((2.0 * i) * (2.0 * i + 1.0));
    sum += term;
  }
  return sum;
}
constexpr double f3c_cos_t(double x) {
  double term = 1.0, sum = 1.0;
  for (int i = 1; i < 14; ++i) {
    term *= -x * x / ((2.0 * i - 1.0) * (2.0 * i));
    sum += term;
  }
  return sum;
}
// cos / sin of 2*pi*m/r for 0 <= m < r, exact symmetries first (octant reduction on the integer fraction)
constexpr double f3c_cos(int m, int r) {
  m %= r;
  if (2 * m > r) m = r - m;                 // cos(2pi - a) = cos a     -> a in [0, pi]
  if (4 * m > r) return -f3c_cos(r - 2 * m, 2 * r);  // cos(pi - b) = -cos b, b = 2pi (r/2 - m)/r = 2pi (r - 2m)/(2r)
  if (8 * m > r) return f3c_sin_t(2.0 * f3c_pi * (r - 4 * m) / (4.0 * r));  // cos a = sin(pi/2 - a)
  return f3c_cos_t(2.0 * f3c_pi * m / r);
}
constexpr double f3c_sin(int m, int r) {
  m %= r;
  if (2 * m > r) return -f3c_sin(r - m, r);
  if (4 * m > r) return f3c_sin(r - 2 * m, 2 * r);  // sin(pi - b) = sin b
  if (8 * m > r) return f3c_cos_t(2.0 * f3c_pi * (r - 4 * m) / (4.0 * r));
  return f3c_sin_t(2.0 * f3c_pi * m / r);
}

// a * W_R^M (forward: W = exp(-2 pi i / R); INV: the conjugate), M a compile-time constant
template <int R, int M, bool INV>
__device__ __forceinline__ float2 f3_twc(float2 a) {
  constexpr int m = M % R;
  if constexpr (m == 0) {
    return a;
  } else if constexpr (4 * m == R) {
    return mul_mi<INV>(a);
  } else if constexpr (2 * m == R) {
    return make_float2(-a.x, -a.y);
  } else if constexpr (4 * m == 3 * R) {
    return mul_mi<!INV>(a);
  } else {
    constexpr float c = (float)f3c_cos(m, R);
    constexpr float s = (float)(INV ? f3c_sin(m, R) : -f3c_sin(m, R));  // W = c + i s
    if constexpr (8 * m == R || 8 * m == 3 * R || 8 * m == 5 * R || 8 * m == 7 * R) {
      // |c| = |s| = sqrt(1/2): two adds and two multiplies
      constexpr float h = 0.70710678118654752440f;
      const float xc = c > 0 ? a.x : -a.x, ys = s > 0 ? a.y : -a.y;  // signs fold into the adds
      const float xs = s > 0 ? a.x : -a.x, yc = c > 0 ? a.y : -a.y;
      return make_float2(h * (xc - ys), h * (xs + yc));
    } else {
      return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
    }
  }
}

// ---- R = P*Q point DFT in registers: v[n] -> v[k], natural order in and out ---------------------------------------------
//   X[k1 + P k2] = sum_{n2 < Q} W_Q^{n2 k2} ( W_R^{n2 k1} sum_{n1 < P} W_P^{n1 k1} x[Q n1 + n2] )
template <int R, int P, int Q, bool INV, int N2, int... K1>
__device__ __forceinline__ void f3_dft_col(const float2* v, float2 (*y)[P], std::integer_sequence<int, K1...>) {
  float2 t[P] = {v[Q * K1 + N2]...};
  dft_small<P, INV>(t);
  ((y[N2][K1] = f3_twc<R, N2 * K1, INV>(t[K1])), ...);
}
template <int R, int P, int Q, bool INV, int... N2>
__device__ __forceinline__ void f3_dft_cols(const float2* v, float2 (*y)[P], std::integer_sequence<int, N2...>) {
  (f3_dft_col<R, P, Q, INV, N2>(v, y, std::make_integer_sequence<int, P>{}), ...);
}
template <int P, int Q, bool INV>
__device__ __forceinline__ void f3_dft(float2* v) {
  constexpr int R = P * Q;
  float2 y[Q][P];
  f3_dft_cols<R, P, Q, INV>(v, y, std::make_integer_sequence<int, Q>{});
#pragma unroll
  for (int k1 = 0; k1 < P; ++k1) {
    float2 t[Q];
#pragma unroll
    for (int n2 = 0; n2 < Q; ++n2) t[n2] = y[n2][k1];
    dft_small<Q, INV>(t);
#pragma unroll
    for (int k2 = 0; k2 < Q; ++k2) v[k1 + P * k2] = t[k2];
  }
}

// v[r] *= tw[r] (INV: conj), r = 1..R-1; tw = R consecutive float2 (16-byte aligned: R is even)
template <int R, bool INV>
__device__ __forceinline__ void f3_apply_tw(float2* v, const float2* __restrict__ tw) {
  const float4* t4 = reinterpret_cast<const float4*>(tw);
#pragma unroll
  for (int h = 0; h < R / 2; ++h) {
    const float4 w = __ldg(t4 + h);
    if (h > 0) {
      const float2 a = v[2 * h];
      const float wy = INV ? -w.y : w.y;
      v[2 * h] = make_float2(a.x * w.x - a.y * wy, a.x * wy + a.y * w.x);
    }
    const float2 b = v[2 * h + 1];
    const float wz = INV ? -w.w : w.w;
    v[2 * h + 1] = make_float2(b.x * w.z - b.y * wz, b.x * wz + b.y * w.z);
  }
}

struct Fft3Tw {
  const float2* p2;  // [R1][R2]     W_{R1 R2}^{k r}
  const float2* p3;  // [R1 R2][R3]  W_N^{k r}
};

// Forward (INV = false) or unnormalised inverse transform.  `load(m)` returns input element m (pass 1 calls it for
// m = j + r N/R1); `sink(j, r, value)` receives output element m = j + r N/R3 (natural order; r is a compile-time constant
// at every call site).  `buf` holds fpad(N) + 1 float2; all kFft3Threads threads of the CTA must call.  The function ends
// WITHOUT a barrier after the sink.
template <int N, bool INV, typename Load, typename Sink>
__device__ __forceinline__ void fft3_run(float2* buf, const Fft3Tw tw, Load load, Sink sink) {
  using G = Fft3Geom<N>;
  constexpr int nb1 = N / G::R1, nb2 = N / G::R2, nb3 = N / G::R3;
  static_assert(nb1 % 32 == 0 && nb2 % 32 == 0 && nb3 % 32 == 0, "row strides must be multiples of the padding period");
  static_assert(G::R1 == 16 && nb3 == G::R1 * G::R2 && nb1 <= kFft3Threads && nb2 <= kFft3Threads, "geometry");
  const int j = threadIdx.x;
  // ---- pass 1: radix R1 = 16, Ns = 1: in[j + r nb1] -> out[16 j + r]
  if (j < nb1) {
    float2 v[G::R1];
#pragma unroll
    for (int r = 0; r < G::R1; ++r) v[r] = load(j + r * nb1);
    f3_dft<G::P1, G::Q1, INV>(v);
    float2* o = buf + 16 * j + (j >> 1);  // fpad(16 j + r) = 16 j + r + (j >> 1) for r < 16
#pragma unroll
    for (int r = 0; r < G::R1; ++r) o[r] = v[r];
  }
  __syncthreads();
  // ---- pass 2: radix R2, Ns = 16: in[j + r nb2] x W_{16 R2}^{k r} -> out[q 16 R2 + k + 16 r]
  {
    float2 v[G::R2];
    const int q = j >> 4, k = j & 15;
    if (j < nb2) {
      const float2* in = buf + j + (j >> 5);
#pragma unroll
      for (int r = 0; r < G::R2; ++r) v[r] = in[r * (nb2 + nb2 / 32)];
      f3_apply_tw<G::R2, INV>(v, tw.p2 + k * G::R2);
      f3_dft<G::P2, G::Q2, INV>(v);
    }
    __syncthreads();
    if (j < nb2) {
      // i = q 16 R2 + k + 16 r; 16 R2 is a multiple of 32 and k < 16, so i >> 5 = q R2 / 2 + (r >> 1)
      float2* o = buf + q * (16 * G::R2 + G::R2 / 2) + k;
#pragma unroll
      for (int r = 0; r < G::R2; ++r) o[16 * r + (r >> 1)] = v[r];
    }
  }
  __syncthreads();
  // ---- pass 3: radix R3, Ns = 16 R2 = nb3: in[k + r nb3] x W_N^{k r} -> out[k + r nb3]  (same addresses)
  if (j < nb3) {
    float2 v[G::R3];
    const float2* in = buf + j + (j >> 5);
#pragma unroll
    for (int r = 0; r < G::R3; ++r) v[r] = in[r * (nb3 + nb3 / 32)];
    f3_apply_tw<G::R3, INV>(v, tw.p3 + j * G::R3);
    f3_dft<G::P3, G::Q3, INV>(v);
#pragma unroll
    for (int r = 0; r < G::R3; ++r) sink(j, r, v[r]);
  }
}

}  // namespace ac
