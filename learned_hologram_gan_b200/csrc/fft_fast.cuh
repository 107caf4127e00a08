// fft_fast.cuh -- compile-time planned in-place FFT passes (sm_100a).
//
// Same algorithm as fft_core.cuh (DIF forward: natural -> scrambled; DIT: scrambled -> natural,
// inverse via re/im swap) with everything the generic path resolves at run time fixed at compile
// time: transform length, radix list (up to 4 passes of radix <= 32), sequences per tile and CTA size.
//   * A pass takes a LOAD and a STORE functor, so the first pass can read global memory and the last
//     one can write global memory without a round trip through shared memory; the spectral multiply of
//     the column kernel is fused into such a functor.
//   * Twiddles come from per-pass tables in SHARED memory, tab[(q-1)*M + j] = exp(-2 pi i j q / NCUR),
//     filled once per CTA from the length-N master table (the measured FP32/INT issue rate, not shared
//     memory, bounds these kernels: a table lookup replaces a 4-instruction complex multiply).  Pass 0
//     may instead build its powers by a product tree when its table would not fit.
//   * Zero-pad pruning: when the non-pad samples are exactly the inputs k in [KLO, KHI) of every
//     first-pass butterfly (pad = KLO*M0, rows = (KHI-KLO)*M0), the zero inputs are neither loaded nor
//     added (DftIn), and the matching last inverse pass computes only the outputs that survive the crop.
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

// ---- butterflies with inputs known to be zero --------------------------------------------------
// MASK bit n set = input n may be non-zero.  IEEE arithmetic does not let the compiler drop "x + 0",
// so the zero structure is spelled out with compile-time flags.
template <bool ZA, bool ZB>
__device__ __forceinline__ float2 addz(float2 a, float2 b) {
  if constexpr (ZA && ZB) return make_float2(0.0f, 0.0f);
  else if constexpr (ZA) return b;
  else if constexpr (ZB) return a;
  else return cadd(a, b);
}
template <bool ZA, bool ZB>
__device__ __forceinline__ float2 subz(float2 a, float2 b) {
  if constexpr (ZA && ZB) return make_float2(0.0f, 0.0f);
  else if constexpr (ZA) return make_float2(-b.x, -b.y);
  else if constexpr (ZB) return a;
  else return csub(a, b);
}

template <int R, unsigned MASK>
struct DftIn {
  __device__ __forceinline__ static void run(float2 (&v)[R]) { Dft<R>::run(v); }
};
template <unsigned MASK>
struct DftIn<2, MASK> {
  __device__ __forceinline__ static void run(float2 (&v)[2]) {
    constexpr bool z0 = !(MASK & 1u), z1 = !(MASK & 2u);
    const float2 a = v[0], b = v[1];
    v[0] = addz<z0, z1>(a, b);
    v[1] = subz<z0, z1>(a, b);
  }
};
template <unsigned MASK>
struct DftIn<4, MASK> {
  __device__ __forceinline__ static void run(float2 (&v)[4]) {
    constexpr bool z0 = !(MASK & 1u), z1 = !(MASK & 2u), z2 = !(MASK & 4u), z3 = !(MASK & 8u);
    constexpr bool ze = z0 && z2, zo = z1 && z3;
    const float2 t0 = addz<z0, z2>(v[0], v[2]);
    const float2 t1 = subz<z0, z2>(v[0], v[2]);
    const float2 t2 = addz<z1, z3>(v[1], v[3]);
    const float2 t3 = mul_mi(subz<z1, z3>(v[1], v[3]));
    v[0] = addz<ze, zo>(t0, t2);
    v[1] = addz<ze, zo>(t1, t3);
    v[2] = subz<ze, zo>(t0, t2);
    v[3] = subz<ze, zo>(t1, t3);
  }
};

// input-index mask of the first-stage sub-transform n2 of a composite radix R1*R2 (n = R2*n1 + n2)
template <int R1, int R2, int KLO, int KHI>
__host__ __device__ constexpr unsigned sub_mask(int n2) {
  unsigned m = 0;
  for (int n1 = 0; n1 < R1; ++n1) {
    const int n = R2 * n1 + n2;
    if (n >= KLO && n < KHI) m |= 1u << n1;
  }
  return m;
}

// composite radix with inputs outside [KLO, KHI) known to be zero (first stage pruned)
template <int R1, int R2, int KLO, int KHI>
struct DftCompositeIn {
  static constexpr int N = R1 * R2;
  template <int n2>
  __device__ __forceinline__ static void stage1(float2 (&a)[R2][R1], const float2 (&v)[N]) {
    if constexpr (n2 < R2) {
#pragma unroll
      for (int n1 = 0; n1 < R1; ++n1) a[n2][n1] = v[R2 * n1 + n2];
      DftIn<R1, sub_mask<R1, R2, KLO, KHI>(n2)>::run(a[n2]);
      stage1<n2 + 1>(a, v);
    }
  }
  __device__ __forceinline__ static void run(float2 (&v)[N]) {
    float2 a[R2][R1];
    stage1<0>(a, v);
    DftComposite<R1, R2>::template tw_all<1>(a);
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

template <int R, int KLO, int KHI>
struct DftPruned {
  __device__ __forceinline__ static void run(float2 (&v)[R]) { Dft<R>::run(v); }
};
template <int KLO, int KHI>
struct DftPruned<8, KLO, KHI> {
  __device__ __forceinline__ static void run(float2 (&v)[8]) {
    if constexpr (KLO == 0 && KHI == 8) Dft<8>::run(v);
    else DftCompositeIn<2, 4, KLO, KHI>::run(v);
  }
};
template <int KLO, int KHI>
struct DftPruned<16, KLO, KHI> {
  __device__ __forceinline__ static void run(float2 (&v)[16]) {
    if constexpr (KLO == 0 && KHI == 16) Dft<16>::run(v);
    else DftCompositeIn<4, 4, KLO, KHI>::run(v);
  }
};

template <int KLO, int KHI>
struct DftPruned<32, KLO, KHI> {
  __device__ __forceinline__ static void run(float2 (&v)[32]) {
    if constexpr (KLO == 0 && KHI == 32) Dft<32>::run(v);
    else DftCompositeIn<4, 8, KLO, KHI>::run(v);
  }
};

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
  __host__ __device__ static constexpr int radix(int p) { return p == 0 ? R0_ : p == 1 ? R1_ : p == 2 ? R2_ : R3_; }
  // block length of pass p (DIF order): N, N/R0, N/(R0 R1), ...
  __host__ __device__ static constexpr int ncur(int p) {
    int n = N_;
    for (int i = 0; i < p; ++i) n /= radix(i);
    return n;
  }
  __host__ __device__ static constexpr int mlen(int p) { return ncur(p) / radix(p); }
  // twiddle tables: passes 1 .. NPASS-2 always have the full table; pass 0 according to tab0:
  //   0 = none (product tree from the global master table), 1 = full table,
  //   2 = first powers only (product tree from shared memory),
  //   3 = first powers only for EVERY pass (row kernels: their passes are bound by shared-memory
  //       bandwidth, and a table lookup per twiddle costs as much of it as the data itself)
  __host__ __device__ static constexpr int tab_len(int p, int tab0) {
    if (tab0 == 3) return mlen(p);
    if (p == 0) return tab0 == 1 ? (radix(0) - 1) * mlen(0) : (tab0 == 2 ? mlen(0) : 0);
    return (radix(p) - 1) * mlen(p);
  }
  __host__ __device__ static constexpr int tab_off(int p, int tab0) {
    int off = 0;
    for (int i = 0; i < p; ++i) off += tab_len(i, tab0);
    return off;
  }
  __host__ __device__ static constexpr int tab_total(int tab0) { return tab_off(NPASS - 1, tab0); }
  // natural index held at scrambled position pos (same recursion as build_perm on the host)
  __host__ __device__ static constexpr int perm(int pos) {
    int idx = 0, stride = 1, n = N_;
    for (int p = 0; p < NPASS; ++p) {
      const int m = n / radix(p);
      const int q = pos / m;
      pos -= q * m;
      idx += q * stride;
      stride *= radix(p);
      n = m;
    }
    return idx;
  }
  // scrambled position that holds natural index idx (inverse of perm)
  __host__ __device__ static constexpr int iperm(int idx) {
    int pos = 0, n = N_;
    for (int p = 0; p < NPASS; ++p) {
      const int m = n / radix(p);
      const int q = idx % radix(p);
      idx /= radix(p);
      pos += q * m;
      n = m;
    }
    return pos;
  }
};

// fill the shared-memory twiddle tables of plan P from the master table tw[k] = exp(-2 pi i k / N)
template <class P, int TAB0>
__device__ __forceinline__ void fill_tables(float2* tabs, const float2* __restrict__ tw, int tid, int nthr) {
#pragma unroll
  for (int p = 0; p < P::NPASS - 1; ++p) {
    const int m = P::mlen(p), ts = P::N / P::ncur(p);
    const int len = P::tab_len(p, TAB0);
    float2* t = tabs + P::tab_off(p, TAB0);
    for (int e = tid; e < len; e += nthr) {
      const int q = e / m + 1, j = e - (q - 1) * m;
      t[e] = __ldg(tw + (size_t)(j * q) * ts);  // j*q*ts < N: j < m, q < r, m*r*ts = N
    }
  }
}

// twiddle powers w[1..R-1] of butterfly j of a pass.  TWMODE: 0 = product tree from the global master
// table, 1 = full shared-memory table, 2 = product tree from a shared-memory table of first powers
template <int R, int M, int TS, int TWMODE>
__device__ __forceinline__ void load_twiddles(float2 (&w)[R], const float2* __restrict__ tw,
                                              const float2* __restrict__ tab, int j) {
  if constexpr (TWMODE == 1) {
#pragma unroll
    for (int q = 1; q < R; ++q) w[q] = tab[(q - 1) * M + j];
  } else {
    w[0] = make_float2(1.0f, 0.0f);
    if constexpr (TWMODE == 2) w[1] = tab[j];
    else w[1] = __ldg(tw + j * TS);
    tw_chain_step<R, 2>(w);
  }
}

#ifndef LHG_TW_HOIST
#define LHG_TW_HOIST 1
#endif
// One radix-R pass (PASS of plan P) over T = 2^LOGT sequences.
//   ld(row, t, k, it) -> float2      st(row, t, k, it, value)      k, it are compile-time after unrolling
// PLANAR selects the thread -> (sequence, butterfly) map: interleaved layouts ([row][t], the column
// tiles) want the sequence index fastest, planar layouts ([t][row], the row tiles) the butterfly index.
// DIF pass 0: inputs outside [KLO, KHI) are zero and not loaded.  DIT pass 0: only the outputs inside
// [KLO, KHI) are stored (the rest of the butterfly is dead code).
// WL (warp-local, passes >= 1 of a planar one-sequence tile): the first pass leaves R0 independent blocks of N/R0
// positions; warp w runs the butterflies of blocks [w*R0/NW, (w+1)*R0/NW) and of no other, so consecutive WL passes
// need __syncwarp() only.
template <class P, int PASS, int LOGT, int NT, bool DIT, bool PLANAR, int TWMODE, int KLO, int KHI, bool WL = false, class Ld,
          class St>
__device__ __forceinline__ void fpass(const float2* __restrict__ tw, const float2* __restrict__ tabs, int tid, Ld ld,
                                      St st) {
  constexpr int N = P::N, R = P::radix(PASS), NCUR = P::ncur(PASS);
  constexpr int T = 1 << LOGT;
  constexpr int M = NCUR / R;
  constexpr int NBS = N / R;
  constexpr int NB = NBS * T;
  constexpr bool PRUNE = (KLO > 0 || KHI < R);
  // A pass whose butterflies of one block are M = 15 apart: 15 of every 16 lanes take one block each, so a
  // half-warp (one 64-bit shared-memory wavefront) never straddles two blocks (2-way bank conflicts else).
  constexpr bool MAP15 = PLANAR && T == 1 && M == 15 && (NT % 16) == 0;
  constexpr int NW = NT / 32;
  constexpr int NBW = NBS / NW;  // WL: butterflies of this pass inside one warp's blocks
  static_assert(!WL || (PLANAR && T == 1 && PASS >= 1 && (NT % 32) == 0 && (P::radix(0) % NW) == 0 && NBW * NW == NBS),
                "warp-local passes: whole first-pass blocks per warp");
  static_assert(!WL || !MAP15 || (NBW % 15) == 0, "whole 15-butterfly blocks per warp");
  constexpr int PER_IT = WL ? (MAP15 ? 30 : 32) : (MAP15 ? (NT / 16) * 15 : NT);
  constexpr int ITERS = WL ? (NBW + PER_IT - 1) / PER_IT : (NB + PER_IT - 1) / PER_IT;
  // MAP15: a thread's butterflies all have j = tid & 15, so their twiddle powers are the same in every iteration:
  // built once (the compiler does not see it through the index arithmetic)
  float2 w15[R];
  if constexpr (MAP15 && LHG_TW_HOIST) load_twiddles<R, M, N / NCUR, TWMODE>(w15, tw, tabs, (tid & 15) < 15 ? (tid & 15) : 0);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    int b;
    bool on;
    if constexpr (WL && MAP15) {
      const int l16 = tid & 15, hw = ((tid >> 4) & 1) + 2 * it;
      b = ((tid >> 5) * (NBW / 15) + hw) * 15 + l16;
      on = l16 < 15 && (ITERS * 2 == NBW / 15 || hw < NBW / 15);
    } else if constexpr (WL) {
      const int l = (tid & 31) + 32 * it;
      b = (tid >> 5) * NBW + l;
      on = ITERS * 32 == NBW || l < NBW;
    } else if constexpr (MAP15) {
      const int l16 = tid & 15;
      b = ((tid >> 4) + it * (NT / 16)) * 15 + l16;
      on = l16 < 15 && (ITERS * PER_IT == NB || b < NB);
    } else {
      b = tid + it * NT;
      on = ITERS * NT == NB || b < NB;
    }
    if (on) {
      int t, jj;
      if constexpr (PLANAR) {
        t = b / NBS;
        jj = b - t * NBS;
      } else {
        t = b & (T - 1);
        jj = b >> LOGT;
      }
      const int blk = jj / M;
      const int j = jj - blk * M;
      const int base = blk * NCUR + j;
      float2 v[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (!DIT && PRUNE && (k < KLO || k >= KHI)) v[k] = make_float2(0.0f, 0.0f);
        else v[k] = ld(base + k * M, t, k, it);
      }
      if constexpr (M > 1) {
        float2 w[R];
        if constexpr (MAP15 && LHG_TW_HOIST) {
#pragma unroll
          for (int q = 1; q < R; ++q) w[q] = w15[q];
        } else {
          load_twiddles<R, M, N / NCUR, TWMODE>(w, tw, tabs, j);
        }
        if constexpr (!DIT) {
          if constexpr (PRUNE) DftPruned<R, KLO, KHI>::run(v);
          else Dft<R>::run(v);
        }
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
        if constexpr (DIT) Dft<R>::run(v);
      } else {
        Dft<R>::run(v);
      }
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (DIT && PRUNE && (k < KLO || k >= KHI)) continue;
        st(base + k * M, t, k, it, v[k]);
      }
    }
  }
}

// The last inverse pass (DIT pass 0, outputs pruned to [KLO, KHI)) and the first forward pass (DIF pass 0, inputs
// pruned to the same range) of plan P on the same butterfly, without leaving registers: both touch positions
// base + k*M of their sequence with the same twiddles.  Planar layout.
//   pre(base, t, it) -> Aux           anything the middle stage wants fetched before the butterfly is computed
//   ld(row, t, k, it) -> float2       (re/im-swapped representation of the inverse transform)
//   mid(base, t, it, v, aux)          turns v[KLO..KHI) (still swapped, un-normalised) into the forward inputs
//   st(row, t, k, it, value)
template <class P, int LOGT, int NT, int TWMODE, int KLO, int KHI, class Pre, class Ld, class Mid, class St>
__device__ __forceinline__ void fturn(const float2* __restrict__ tw, const float2* __restrict__ tabs, int tid, Pre pre,
                                      Ld ld, Mid mid, St st) {
  constexpr int N = P::N, R = P::radix(0);
  constexpr int T = 1 << LOGT;
  constexpr int M = N / R;
  constexpr int NB = M * T;
  constexpr int ITERS = (NB + NT - 1) / NT;
  static_assert(M > 1, "fturn needs a twiddled pass 0");
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int b = tid + it * NT;
    if (ITERS * NT == NB || b < NB) {
      const int t = b / M;
      const int j = b - t * M;
      auto aux = pre(j, t, it);
      float2 v[R], w[R];
#pragma unroll
      for (int k = 0; k < R; ++k) v[k] = ld(j + k * M, t, k, it);
      load_twiddles<R, M, 1, TWMODE>(w, tw, tabs, j);
#pragma unroll
      for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
      Dft<R>::run(v);  // outputs outside [KLO, KHI) are dead code
      mid(j, t, it, v, aux);
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (k < KLO || k >= KHI) v[k] = make_float2(0.0f, 0.0f);
      DftPruned<R, KLO, KHI>::run(v);
#pragma unroll
      for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
#pragma unroll
      for (int k = 0; k < R; ++k) st(j + k * M, t, k, it, v[k]);
    }
  }
}

// The same pass over PAIRS of adjacent sequences of an interleaved tile ([row][T] layout): one thread
// runs butterfly jj on two columns, so the twiddles are fetched (or built) once for both and every
// shared-memory access is 16 bytes wide.
//   ld(row, tp, k, it) -> float4 (x0, y0, x1, y1)      st(row, tp, k, it, float4)       tp = column pair
template <class P, int PASS, int LOGT, int NT, bool DIT, int TWMODE, int KLO, int KHI, class Ld, class St>
__device__ __forceinline__ void fpass2(const float2* __restrict__ tw, const float2* __restrict__ tabs, int tid, Ld ld,
                                       St st) {
  constexpr int N = P::N, R = P::radix(PASS), NCUR = P::ncur(PASS);
  constexpr int TP = 1 << (LOGT - 1);  // column pairs per tile
  constexpr int M = NCUR / R;
  constexpr int NBS = N / R;
  constexpr int NB = NBS * TP;
  constexpr int ITERS = (NB + NT - 1) / NT;
  constexpr bool PRUNE = (KLO > 0 || KHI < R);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int b = tid + it * NT;
    if (ITERS * NT == NB || b < NB) {
      const int tp = b & (TP - 1);
      const int jj = b >> (LOGT - 1);
      const int blk = jj / M;
      const int j = jj - blk * M;
      const int base = blk * NCUR + j;
      float2 v0[R], v1[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (!DIT && PRUNE && (k < KLO || k >= KHI)) {
          v0[k] = v1[k] = make_float2(0.0f, 0.0f);
        } else {
          const float4 x = ld(base + k * M, tp, k, it);
          v0[k] = make_float2(x.x, x.y);
          v1[k] = make_float2(x.z, x.w);
        }
      }
      if constexpr (M > 1) {
        float2 w[R];
        load_twiddles<R, M, N / NCUR, TWMODE>(w, tw, tabs, j);
        if constexpr (!DIT) {
          if constexpr (PRUNE) {
            DftPruned<R, KLO, KHI>::run(v0);
            DftPruned<R, KLO, KHI>::run(v1);
          } else {
            Dft<R>::run(v0);
            Dft<R>::run(v1);
          }
        }
#pragma unroll
        for (int q = 1; q < R; ++q) {
          v0[q] = cmul(v0[q], w[q]);
          v1[q] = cmul(v1[q], w[q]);
        }
        if constexpr (DIT) {
          Dft<R>::run(v0);
          Dft<R>::run(v1);
        }
      } else {
        Dft<R>::run(v0);
        Dft<R>::run(v1);
      }
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (DIT && PRUNE && (k < KLO || k >= KHI)) continue;
        st(base + k * M, tp, k, it, make_float4(v0[k].x, v0[k].y, v1[k].x, v1[k].y));
      }
    }
  }
}

// ---- asynchronous copies (LDGSTS) ----------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

// ---- TMA (cp.async.bulk.tensor) + mbarrier --------------------------------------------------------
// The blocked W layouts hand a row kernel 16/32-byte pieces 128/256 bytes apart: as LDGSTS / STG each piece
// is its own L1 tag lookup (one per clock per SM).  A 4-D tensor map over (complex-in-piece, row-in-block,
// piece, row-block) lets the TMA unit gather / scatter a whole row with 15 bulk copies that bypass L1.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- thread-block clusters: rank, cluster-wide barrier, arrival on the same mbarrier of another CTA of the cluster
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_peer(unsigned long long* bar, unsigned peer_rank) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(peer_rank)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE_%=;\n"
      "bra MBAR_WAIT_%=;\n"
      "MBAR_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// the same wait with a bound: a mis-programmed copy ends the kernel with a trap (a CUDA error the caller sees)
// instead of hanging the device
__device__ __forceinline__ void mbar_wait_bounded(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  for (int spin = 0; spin < (1 << 24) && !done; ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
  if (!done) asm volatile("trap;");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
// the same box, only as far as L2 (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2),
               "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, int c0, int c1, int c2, int c3, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(tmap),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// exp(i*theta) for |theta| up to ~1e5 rad: Cody-Waite reduction to [-pi, pi] with a 3-term 2*pi,
// then the SFU sine/cosine (abs error ~4e-7 on a unit phasor; the parity gate on H is 1e-6 rel-L2).
// The rounding to the nearest turn uses the 1.5*2^23 trick (FP32 pipe) instead of FRND (SFU-rate pipe).
__device__ __forceinline__ float2 fast_cis(float theta) {
  const float k = __fadd_rn(__fmaf_rn(theta, 0.15915494309189535f, 12582912.0f), -12582912.0f);
  float r = fmaf(-k, 6.28125f, theta);                   // 2*pi head: 9 significant bits, k*head exact
  r = fmaf(-k, 1.9350051879882812e-3f, r);               // next 12 bits
  r = fmaf(-k, 3.0199159819567e-7f, r);                  // tail
  return make_float2(__cosf(r), __sinf(r));
}
// the same for theta = fl(beta * w) with the number of turns taken from w * (beta / 2 pi) directly, so the
// rounding of theta and the turn count do not wait for each other (a turn more or less is harmless)
__device__ __forceinline__ float2 fast_cis_bw(float beta, float beta_turns, float w) {
  const float theta = __fmul_rn(beta, w);
  const float k = __fadd_rn(__fmaf_rn(w, beta_turns, 12582912.0f), -12582912.0f);
  float r = fmaf(-k, 6.28125f, theta);                   // 2*pi head: 9 significant bits, k*head exact
  r = fmaf(-k, 1.9350051879882812e-3f, r);               // next 12 bits
  r = fmaf(-k, 3.0199159819567e-7f, r);                  // tail
  return make_float2(__cosf(r), __sinf(r));
}

#ifndef LHG_CIS_TWO_TERM
#define LHG_CIS_TWO_TERM 1
#endif
// G transfer-function values at once, written stage by stage so that the G dependent chains
// (reduction -> SFU -> product) are in flight together instead of one after the other
template <int G>
__device__ __forceinline__ void fast_cis_group(float beta, float beta_turns, const float* w, float2* h) {
  float k[G], r[G];
#pragma unroll
  for (int i = 0; i < G; ++i) {
    k[i] = __fadd_rn(__fmaf_rn(w[i], beta_turns, 12582912.0f), -12582912.0f);
    r[i] = __fmul_rn(beta, w[i]);
  }
#if LHG_CIS_TWO_TERM
  // Two-term reduction: theta - k*fl(2 pi) is computed by ONE fused multiply-add, i.e. rounded once from the exact
  // value -- and that value is representable (a multiple of 2^-21 below 8 whenever ulp(theta) >= 2^-21), so the
  // first step is exact without a short head constant; the second term carries 2 pi - fl(2 pi).  Residual
  // k * 3e-15 + half an ulp of the reduced angle (2.4e-7), inside the SFU's own error.
#pragma unroll
  for (int i = 0; i < G; ++i) r[i] = fmaf(-k[i], 6.2831854820251465f, r[i]);
#pragma unroll
  for (int i = 0; i < G; ++i) r[i] = fmaf(-k[i], -1.7484555e-7f, r[i]);
#else
#pragma unroll
  for (int i = 0; i < G; ++i) r[i] = fmaf(-k[i], 6.28125f, r[i]);
#pragma unroll
  for (int i = 0; i < G; ++i) r[i] = fmaf(-k[i], 1.9350051879882812e-3f, r[i]);
#pragma unroll
  for (int i = 0; i < G; ++i) r[i] = fmaf(-k[i], 3.0199159819567e-7f, r[i]);
#endif
#pragma unroll
  for (int i = 0; i < G; ++i) h[i] = make_float2(__cosf(r[i]), __sinf(r[i]));
}

}  // namespace asmb
