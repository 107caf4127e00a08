// fft_core.cuh -- in-register radix codelets and in-place shared-memory FFT passes (sm_100a).
//
// Design (DESIGN.md section "FFT core"):
//   * A length-N transform is a list of radix passes over a shared-memory array holding
//     T interleaved sequences (element n of sequence t lives at buf[n*T + t], T = 2^logT).
//   * Forward direction = decimation in frequency, in place: natural order in,
//     mixed-radix digit-reversed ("scrambled") order out.  The matching decimation-in-time
//     passes (reverse radix order) take scrambled order in and give natural order out.
//     Point-wise work between the two (multiply by the transfer function) happens in
//     scrambled order through a permutation table, so no reordering pass ever runs.
//   * Both directions use the forward kernel e^{-2 pi i/N}; an inverse transform is obtained
//     by swapping re/im at load and at store (ifft(x) = swap(fft(swap(x)))).
//   * Each butterfly is a radix-R DFT held in registers; composite radices are built from
//     2/3/4/5 codelets with compile-time twiddles.
#pragma once
#include <cuda_runtime.h>

namespace asmb {

constexpr int kMaxPass = 12;

struct Fft1d {
  int n;                // transform length
  int npass;
  int radix[kMaxPass];  // DIF order (of the length-m convolution transform when blue != 0)
  const float2* tw;     // tw[k] = exp(-2 pi i k / len), len = n (or m when blue != 0)
  const int* perm;      // perm[pos]  = natural index held at scrambled position pos
  const int* iperm;     // iperm[k]   = scrambled position of natural index k
  // Bluestein (chirp-z) fallback for lengths with a prime factor > 5: the length-n DFT is a
  // circular convolution of length m = 2^k >= 2n-1 carried out with the passes above.  Output is
  // in NATURAL order (perm/iperm are the identity) and buf must hold buf_len = m elements.
  int blue;
  int m;
  int buf_len;          // elements of shared memory one sequence needs (n, or m for Bluestein)
  const float2* chirp;  // c[k] = exp(-i pi k^2 / n), k < n
  const float2* hf;     // FFT_m of the wrapped conjugate chirp, scaled by 1/m, in scrambled order
};

// ---- complex arithmetic on the packed FP32x2 pipe of sm_100 ---------------------------------------
// add/sub/mul/fma.f32x2 issue ONE instruction for both halves of a complex number (FADD2 / FMUL2 / FFMA2;
// ptxas folds the pack/unpack moves, the re<->im swap, per-half negation and scalar broadcast into operand
// modifiers).  The butterflies are bound by instruction issue, not by FP32 lanes (measured: tools/ubench.cu),
// so this halves their cost.  A complex multiply is FMUL2 + FFMA2, a multiply by -i is free.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ u64 pk2(float2 a) { return pk2(a.x, a.y); }
__device__ __forceinline__ float2 up2(u64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  // (a.x b.x - a.y b.y, a.y b.x + a.x b.y)
  return up2(fma2(pk2(-a.y, a.x), pk2(b.y, b.y), mul2(pk2(a), pk2(b.x, b.x))));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return up2(add2(pk2(a), pk2(b))); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return up2(sub2(pk2(a), pk2(b))); }
__device__ __forceinline__ float2 cswap(float2 a) { return make_float2(a.y, a.x); }
// multiply by -i
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }
// s * a, a + s * b, a - i * s * b   (s real)
__device__ __forceinline__ float2 cscale(float s, float2 a) { return up2(mul2(pk2(a), pk2(s, s))); }
__device__ __forceinline__ float2 caxpy(float s, float2 b, float2 a) { return up2(fma2(pk2(b), pk2(s, s), pk2(a))); }
__device__ __forceinline__ float2 caxpy_mi(float s, float2 b, float2 a) {
  return up2(fma2(pk2(b.y, -b.x), pk2(s, s), pk2(a)));
}

// ---- compile-time trigonometry for codelet twiddles -------------------------------------
constexpr double kPi = 3.14159265358979323846264338327950288;
__host__ __device__ constexpr double c_sin(double x) {
  double term = x, sum = x;
  for (int i = 1; i < 40; ++i) {
    term *= -x * x / ((2.0 * i) * (2.0 * i + 1.0));
    sum += term;
  }
  return sum;
}
__host__ __device__ constexpr double c_cos(double x) {
  double term = 1.0, sum = 1.0;
  for (int i = 1; i < 40; ++i) {
    term *= -x * x / ((2.0 * i - 1.0) * (2.0 * i));
    sum += term;
  }
  return sum;
}
// v * exp(-2 pi i K / N), K and N compile-time
template <int N, int K>
__device__ __forceinline__ float2 twiddle_const(float2 v) {
  constexpr int k = ((K % N) + N) % N;
  if constexpr (k == 0) {
    return v;
  } else if constexpr (4 * k == N) {
    return make_float2(v.y, -v.x);  // * -i
  } else if constexpr (2 * k == N) {
    return make_float2(-v.x, -v.y);
  } else if constexpr (4 * k == 3 * N) {
    return make_float2(-v.y, v.x);  // * +i
  } else {
    // reduce the angle to (-pi, pi] before the series; v * (c + i s) with s = -sin: FMUL2 + FFMA2
    constexpr int kk = (2 * k > N) ? k - N : k;
    constexpr float c = (float)c_cos(2.0 * kPi * kk / N);
    constexpr float s = (float)(-c_sin(2.0 * kPi * kk / N));
    return cmul(v, make_float2(c, s));
  }
}

// ---- base codelets: forward DFT, natural order in and out -----------------------------------
template <int R>
struct Dft;

template <>
struct Dft<2> {
  __device__ __forceinline__ static void run(float2 (&v)[2]) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};

template <>
struct Dft<3> {
  __device__ __forceinline__ static void run(float2 (&v)[3]) {
    constexpr float s = 0.86602540378443864676f;  // sin(2 pi/3)
    const float2 t1 = cadd(v[1], v[2]);
    const float2 d = csub(v[1], v[2]);
    const float2 m1 = caxpy(-0.5f, t1, v[0]);
    v[0] = cadd(v[0], t1);
    v[1] = caxpy_mi(s, d, m1);   // m1 - i s d
    v[2] = caxpy_mi(-s, d, m1);  // m1 + i s d
  }
};

template <>
struct Dft<4> {
  __device__ __forceinline__ static void run(float2 (&v)[4]) {
    float2 t0 = cadd(v[0], v[2]);
    float2 t1 = csub(v[0], v[2]);
    float2 t2 = cadd(v[1], v[3]);
    float2 t3 = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
  }
};

template <>
struct Dft<5> {
  __device__ __forceinline__ static void run(float2 (&v)[5]) {
    constexpr float c1 = 0.30901699437494742410f;   // cos(2 pi/5)
    constexpr float c2 = -0.80901699437494742410f;  // cos(4 pi/5)
    constexpr float s1 = 0.95105651629515357212f;   // sin(2 pi/5)
    constexpr float s2 = 0.58778525229247312917f;   // sin(4 pi/5)
    const float2 t1 = cadd(v[1], v[4]);
    const float2 t2 = cadd(v[2], v[3]);
    const float2 t3 = csub(v[1], v[4]);
    const float2 t4 = csub(v[2], v[3]);
    const float2 a1 = caxpy(c2, t2, caxpy(c1, t1, v[0]));
    const float2 a2 = caxpy(c1, t2, caxpy(c2, t1, v[0]));
    const float2 b1 = caxpy(s2, t4, cscale(s1, t3));
    const float2 b2 = caxpy(-s1, t4, cscale(s2, t3));
    v[0] = cadd(cadd(v[0], t1), t2);
    v[1] = cadd(a1, mul_mi(b1));
    v[4] = csub(a1, mul_mi(b1));
    v[2] = cadd(a2, mul_mi(b2));
    v[3] = csub(a2, mul_mi(b2));
  }
};

// composite radix N = R1*R2 (Cooley-Tukey inside registers, compile-time twiddles)
//   input index n = R2*n1 + n2, output index k = k1 + R1*k2
template <int R1, int R2>
struct DftComposite {
  static constexpr int N = R1 * R2;
  template <int n2, int k1>
  __device__ __forceinline__ static void tw_col(float2 (&a)[R2][R1]) {
    if constexpr (k1 < R1) {
      a[n2][k1] = twiddle_const<N, n2 * k1>(a[n2][k1]);
      tw_col<n2, k1 + 1>(a);
    }
  }
  template <int n2>
  __device__ __forceinline__ static void tw_all(float2 (&a)[R2][R1]) {
    if constexpr (n2 < R2) {
      tw_col<n2, 1>(a);
      tw_all<n2 + 1>(a);
    }
  }
  __device__ __forceinline__ static void run(float2 (&v)[N]) {
    float2 a[R2][R1];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) {
#pragma unroll
      for (int n1 = 0; n1 < R1; ++n1) a[n2][n1] = v[R2 * n1 + n2];
      Dft<R1>::run(a[n2]);
    }
    tw_all<1>(a);
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
      float2 b[R2];
#pragma unroll
      for (int n2 = 0; n2 < R2; ++n2) b[n2] = a[n2][k1];
      Dft<R2>::run(b);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = b[k2];
    }
  }
};

template <> struct Dft<6> { __device__ __forceinline__ static void run(float2 (&v)[6]) { DftComposite<2, 3>::run(v); } };
template <> struct Dft<8> { __device__ __forceinline__ static void run(float2 (&v)[8]) { DftComposite<2, 4>::run(v); } };
template <> struct Dft<9> { __device__ __forceinline__ static void run(float2 (&v)[9]) { DftComposite<3, 3>::run(v); } };
template <> struct Dft<10> { __device__ __forceinline__ static void run(float2 (&v)[10]) { DftComposite<2, 5>::run(v); } };
template <> struct Dft<12> { __device__ __forceinline__ static void run(float2 (&v)[12]) { DftComposite<3, 4>::run(v); } };
template <> struct Dft<15> { __device__ __forceinline__ static void run(float2 (&v)[15]) { DftComposite<3, 5>::run(v); } };
template <> struct Dft<16> { __device__ __forceinline__ static void run(float2 (&v)[16]) { DftComposite<4, 4>::run(v); } };

// ---- in-place passes over shared memory -----------------------------------------------------
// One pass of radix R over blocks of length n_cur (= R*m).  kDit=false: butterfly then
// twiddle (decimation in frequency).  kDit=true: twiddle then butterfly (decimation in time).
// Every butterfly reads and writes the same R slots, so a pass needs one barrier after it.
template <int R, bool kDit>
__device__ __forceinline__ void radix_pass(float2* __restrict__ buf, int n, int n_cur, int logT,
                                           const float2* __restrict__ tw, int tid, int nthr) {
  const int m = n_cur / R;
  const int total = (n / R) << logT;
  const int tmask = (1 << logT) - 1;
  const int tstride = n / n_cur;
  const float inv_m = 1.0f / (float)m;
  for (int b = tid; b < total; b += nthr) {
    const int t = b & tmask;
    const int jj = b >> logT;
    int blk = (int)((float)jj * inv_m);
    int j = jj - blk * m;
    if (j < 0) { j += m; --blk; }
    if (j >= m) { j -= m; ++blk; }
    float2* p = buf + (((size_t)(blk * n_cur + j)) << logT) + t;
    const int estride = m << logT;
    float2 v[R];
#pragma unroll
    for (int k = 0; k < R; ++k) v[k] = p[k * estride];
    if (kDit) {
      if (m > 1) {
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = cmul(v[q], __ldg(tw + (size_t)j * q * tstride));
      }
      Dft<R>::run(v);
    } else {
      Dft<R>::run(v);
      if (m > 1) {
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = cmul(v[q], __ldg(tw + (size_t)j * q * tstride));
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) p[k * estride] = v[k];
  }
}

template <bool kDit>
__device__ __forceinline__ void radix_pass_dyn(int r, float2* buf, int n, int n_cur, int logT,
                                               const float2* tw, int tid, int nthr) {
  switch (r) {
    case 2: radix_pass<2, kDit>(buf, n, n_cur, logT, tw, tid, nthr); break;
    case 3: radix_pass<3, kDit>(buf, n, n_cur, logT, tw, tid, nthr); break;
    case 4: radix_pass<4, kDit>(buf, n, n_cur, logT, tw, tid, nthr); break;
    case 5: radix_pass<5, kDit>(buf, n, n_cur, logT, tw, tid, nthr); break;
    case 8: radix_pass<8, kDit>(buf, n, n_cur, logT, tw, tid, nthr); break;
    default: break;
  }
}

__device__ __forceinline__ void dif_passes(float2* buf, const Fft1d& f, int len, int logT, int tid, int nthr) {
  int n_cur = len;
  for (int p = 0; p < f.npass; ++p) {
    const int r = f.radix[p];
    radix_pass_dyn<false>(r, buf, len, n_cur, logT, f.tw, tid, nthr);
    n_cur /= r;
    __syncthreads();
  }
}

__device__ __forceinline__ void dit_passes(float2* buf, const Fft1d& f, int len, int logT, int tid, int nthr) {
  int n_cur = 1;
  for (int p = f.npass - 1; p >= 0; --p) {
    const int r = f.radix[p];
    n_cur *= r;
    radix_pass_dyn<true>(r, buf, len, n_cur, logT, f.tw, tid, nthr);
    __syncthreads();
  }
}

// Bluestein: X[k] = c[k] * sum_n (x[n] c[n]) conj(c)[k-n], natural order in and out, in place in buf[0..m)
__device__ __forceinline__ void fft_bluestein(float2* buf, const Fft1d& f, int logT, int tid, int nthr) {
  const int T = 1 << logT, tmask = T - 1;
  for (int e = tid; e < (f.m << logT); e += nthr) {
    const int i = e >> logT;
    buf[e] = i < f.n ? cmul(buf[e], __ldg(f.chirp + i)) : make_float2(0.0f, 0.0f);
  }
  __syncthreads();
  dif_passes(buf, f, f.m, logT, tid, nthr);
  for (int e = tid; e < (f.m << logT); e += nthr) buf[e] = cswap(cmul(buf[e], __ldg(f.hf + (e >> logT))));
  __syncthreads();
  dit_passes(buf, f, f.m, logT, tid, nthr);
  for (int e = tid; e < (f.n << logT); e += nthr) buf[e] = cmul(cswap(buf[e]), __ldg(f.chirp + (e >> logT)));
  __syncthreads();
  (void)tmask;
}

// natural order -> scrambled order (forward DFT).  Caller has synchronised before; returns synchronised.
__device__ __forceinline__ void fft_dif(float2* buf, const Fft1d& f, int logT, int tid, int nthr) {
  if (f.blue) fft_bluestein(buf, f, logT, tid, nthr);
  else dif_passes(buf, f, f.n, logT, tid, nthr);
}

// scrambled order -> natural order (forward DFT of the sequence whose scrambled layout is in buf)
__device__ __forceinline__ void fft_dit(float2* buf, const Fft1d& f, int logT, int tid, int nthr) {
  if (f.blue) fft_bluestein(buf, f, logT, tid, nthr);
  else dit_passes(buf, f, f.n, logT, tid, nthr);
}

}  // namespace asmb
