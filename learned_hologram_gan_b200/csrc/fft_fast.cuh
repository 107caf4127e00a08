// fft_fast.cuh -- compile-time planned in-place FFT passes (sm_100a).
//
// Same algorithm as fft_core.cuh (DIF forward: natural -> scrambled; DIT: scrambled -> natural,
// inverse via re/im swap) with everything the generic path resolves at run time fixed at compile
// time: transform length, radix list (up to 4 passes of radix <= 32), interleave T and CTA size.
// A pass takes a LOAD and a STORE functor, so the first pass can read global memory (or registers)
// and the last one can write global memory (or registers) without a round trip through shared
// memory; the spectral multiply of the column kernel is fused into such a functor.
#pragma once
#include "fft_core.cuh"

namespace asmb {

template <> struct Dft<18> { __device__ __forceinline__ static void run(float2 (&v)[18]) { DftComposite<2, 9>::run(v); } };
template <> struct Dft<20> { __device__ __forceinline__ static void run(float2 (&v)[20]) { DftComposite<4, 5>::run(v); } };
template <> struct Dft<24> { __device__ __forceinline__ static void run(float2 (&v)[24]) { DftComposite<3, 8>::run(v); } };
template <> struct Dft<25> { __device__ __forceinline__ static void run(float2 (&v)[25]) { DftComposite<5, 5>::run(v); } };
template <> struct Dft<27> { __device__ __forceinline__ static void run(float2 (&v)[27]) { DftComposite<3, 9>::run(v); } };
template <> struct Dft<30> { __device__ __forceinline__ static void run(float2 (&v)[30]) { DftComposite<5, 6>::run(v); } };
template <> struct Dft<32> { __device__ __forceinline__ static void run(float2 (&v)[32]) { DftComposite<4, 8>::run(v); } };

// w[q] = w1^q by a balanced product tree (depth <= log2 R, keeps the rounding error ~ 1e-7)
template <int R, int Q>
__device__ __forceinline__ void tw_chain_step(float2 (&w)[R]) {
  if constexpr (Q < R) {
    w[Q] = cmul(w[Q / 2], w[Q - Q / 2]);
    tw_chain_step<R, Q + 1>(w);
  }
}

// compile-time description of a transform: N = R0*R1*R2*R3 (unused trailing radices = 1)
template <int N_, int R0_, int R1_, int R2_ = 1, int R3_ = 1>
struct FastPlan {
  static constexpr int N = N_, R0 = R0_, R1 = R1_, R2 = R2_, R3 = R3_;
  static_assert(R0_ * R1_ * R2_ * R3_ == N_, "radices must multiply to N");
  static constexpr int NPASS = R3_ > 1 ? 4 : (R2_ > 1 ? 3 : 2);
  static constexpr int RL = R3_ > 1 ? R3_ : (R2_ > 1 ? R2_ : R1_);  // last DIF radix (block length of the M=1 pass)
  // natural index held at scrambled position pos (same recursion as build_perm on the host)
  __host__ __device__ static constexpr int perm(int pos) {
    int idx = 0, stride = 1, n = N_;
    const int r[4] = {R0_, R1_, R2_, R3_};
    for (int p = 0; p < 4; ++p) {
      if (r[p] == 1) break;
      const int m = n / r[p];
      const int q = pos / m;
      pos -= q * m;
      idx += q * stride;
      stride *= r[p];
      n = m;
    }
    return idx;
  }
};

// One radix-R pass over blocks of NCUR = R*M elements of T = 2^LOGT interleaved sequences.
//   ld(it, row, t, k) -> float2      st(it, row, t, k, value)
// "row" is the element index inside the sequence, it/k are compile-time after unrolling (so
// functors may index register arrays with them).
template <int N, int NCUR, int R, int LOGT, int NT, bool DIT, class Ld, class St>
__device__ __forceinline__ void fpass(const float2* __restrict__ tw, int tid, Ld ld, St st) {
  constexpr int T = 1 << LOGT;
  constexpr int M = NCUR / R;
  constexpr int NB = (N / R) * T;
  constexpr int ITERS = (NB + NT - 1) / NT;
  constexpr int TS = N / NCUR;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int b = tid + it * NT;
    if (ITERS * NT == NB || b < NB) {
      const int t = b & (T - 1);
      const int jj = b >> LOGT;
      const int blk = jj / M;
      const int j = jj - blk * M;
      const int base = blk * NCUR + j;
      float2 v[R];
#pragma unroll
      for (int k = 0; k < R; ++k) v[k] = ld(it, base + k * M, t, k);
      if constexpr (M > 1) {
        // twiddles w^q, q = 1..R-1: balanced product tree from one table value (depth <= log2 R)
        float2 w[R];
        w[0] = make_float2(1.0f, 0.0f);
        w[1] = __ldg(tw + j * TS);
        tw_chain_step<R, 2>(w);
        if constexpr (!DIT) Dft<R>::run(v);
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
        if constexpr (DIT) Dft<R>::run(v);
      } else {
        Dft<R>::run(v);
      }
#pragma unroll
      for (int k = 0; k < R; ++k) st(it, base + k * M, t, k, v[k]);
    }
  }
}

// exp(i*theta) for |theta| up to ~1e5 rad: Cody-Waite reduction to [-pi, pi] with a 3-term 2*pi,
// then the SFU sine/cosine (abs error ~4e-7 on a unit phasor; the parity gate on H is 1e-6 rel-L2).
__device__ __forceinline__ float2 fast_cis(float theta) {
  const float k = rintf(theta * 0.15915494309189535f);
  float r = fmaf(-k, 6.28125f, theta);                   // 2*pi head: 9 significant bits, k*head exact
  r = fmaf(-k, 1.9350051879882812e-3f, r);               // next 12 bits
  r = fmaf(-k, 3.0199159819567e-7f, r);                  // tail
  return make_float2(__cosf(r), __sinf(r));
}

}  // namespace asmb
