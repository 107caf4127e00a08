// col_warp.cuh -- column kernel with warp-local inner passes (N = 18 * R1 * R2, T = 32 / R1 columns per tile:
// 4320 = 18 * 16 * 15 with 2 columns, 2160 = 18 * 8 * 15 with 4 columns).
//
// The transform is split 18 x (R1 x R2): the radix-18 pass runs across the CTA, after it every block of
// L = R1*R2 = N/18 consecutive positions is an independent L-point transform.  Warp q owns block q of all
// columns of the tile and runs the radix-R1 and radix-R2 passes, the transfer-function multiply and the
// matching inverse passes on it with nothing but __syncwarp() in between: one synchronisation of the whole CTA
// per transform is left (between the warp-local passes and the radix-18 pass).  For an even number of depths it
// is split-phase: "exchange buffer written / read by every warp" are mbarriers with one arrival per warp, and
// the depths are taken in pairs (W W R R in the forward launch, R R W W in the adjoint), so that a warp always
// has a phase of its own work between announcing and waiting and the shared-memory, SFU and FP32 phases of
// different warps overlap.  Odd depth counts use one CTA barrier per depth.  Two exchange buffers alternate, the
// strips arrive through the TMA unit (one thread, mbarrier completion) into the buffer a pair leaves idle.
//
// Zero-pad pruning for the 2x padded grid: pad = N/4 = 4.5 * M0 (M0 = N/18), so butterfly j of the radix-18
// pass sees its 9 non-pad samples at k in [5,14) when j < M0/2 and k in [4,13) otherwise.  In the (2,9)
// split of the radix-18 butterfly every first-stage 2-point transform then has exactly one non-zero input
// (free), and in the (9,2) split used for the last inverse pass every last-stage 2-point transform has
// exactly one output inside the crop.
#pragma once
#include "common.cuh"
#include "fft_fast.cuh"

// the adjoint launch's last inverse transform of a tile runs in place in the accumulator buffer, so both exchange
// buffers are free for the next tile's first two strips a whole transform earlier and no CTA barrier separates two
// tiles (A/B knob)
#ifndef LHG_COL_XINV
#define LHG_COL_XINV 1
#endif
// radix-18 twiddle powers kept in shared memory (1 = first powers only, product tree for the rest; 9 = half of them)
#ifndef LHG_COL_TABQ
#define LHG_COL_TABQ 9
#endif
#ifndef LHG_COL_XFWD
#define LHG_COL_XFWD 1
#endif
#ifndef LHG_COL_PREFETCH
#define LHG_COL_PREFETCH 1
#endif
// split-phase depth loop (A/B knob: -DLHG_COL_SPLIT=0 restores one CTA barrier per depth)
#ifndef LHG_COL_SPLIT
#define LHG_COL_SPLIT 1
#endif

namespace asmb {

// x[n2] = the non-zero input of first-stage pair n2 (from v[n2+9] when HI, else v[n2]); pair 4 flips
// between the two halves of the column with `hi4` (run-time, uniform for all but one warp).
// Timing-only ablation knobs (tools/exp.sh builds them into separate libraries with build_variant; results are
// invalid by construction): LHG_EXP_NOSTG drops the global stores of the last inverse pass.
__device__ __forceinline__ void dft18_in9(const float2 (&x)[9], bool hi4, float2 (&v)[18]) {
  float2 a0[9], a1[9];
#pragma unroll
  for (int n2 = 0; n2 < 9; ++n2) {
    a0[n2] = x[n2];
    const bool hi = n2 < 4 ? true : (n2 > 4 ? false : hi4);
    a1[n2] = hi ? make_float2(-x[n2].x, -x[n2].y) : x[n2];
  }
  // twiddle W_18^(n2*k1) for k1 = 1
  a1[1] = twiddle_const<18, 1>(a1[1]);
  a1[2] = twiddle_const<18, 2>(a1[2]);
  a1[3] = twiddle_const<18, 3>(a1[3]);
  a1[4] = twiddle_const<18, 4>(a1[4]);
  a1[5] = twiddle_const<18, 5>(a1[5]);
  a1[6] = twiddle_const<18, 6>(a1[6]);
  a1[7] = twiddle_const<18, 7>(a1[7]);
  a1[8] = twiddle_const<18, 8>(a1[8]);
  Dft<9>::run(a0);
  Dft<9>::run(a1);
#pragma unroll
  for (int k2 = 0; k2 < 9; ++k2) {
    v[2 * k2] = a0[k2];
    v[2 * k2 + 1] = a1[k2];
  }
}

// 18-point DFT of v (natural order), of which only outputs 4..13 are produced: out[i] = X[4 + i].
__device__ __forceinline__ void dft18_out4_13(const float2 (&v)[18], float2 (&out)[10]) {
  float2 b0[9], b1[9];
#pragma unroll
  for (int n1 = 0; n1 < 9; ++n1) {
    b0[n1] = v[2 * n1];
    b1[n1] = v[2 * n1 + 1];
  }
  Dft<9>::run(b0);
  Dft<9>::run(b1);
  b1[1] = twiddle_const<18, 1>(b1[1]);
  b1[2] = twiddle_const<18, 2>(b1[2]);
  b1[3] = twiddle_const<18, 3>(b1[3]);
  b1[4] = twiddle_const<18, 4>(b1[4]);
  b1[5] = twiddle_const<18, 5>(b1[5]);
  b1[6] = twiddle_const<18, 6>(b1[6]);
  b1[7] = twiddle_const<18, 7>(b1[7]);
  b1[8] = twiddle_const<18, 8>(b1[8]);
  // X[k1] = b0 + b1 (k1 = 4..8), X[k1 + 9] = b0 - b1 (k1 = 0..4)
#pragma unroll
  for (int k1 = 4; k1 < 9; ++k1) out[k1 - 4] = cadd(b0[k1], b1[k1]);
#pragma unroll
  for (int k1 = 0; k1 < 5; ++k1) out[5 + k1] = csub(b0[k1], b1[k1]);
}

// tmap / use_tma: the adjoint launch stages every strip after the first of a tile with the TMA unit: ONE thread issues
// one or two cp.async.bulk.tensor.4d boxes (T columns x 8 rows x up to 256 row blocks of the blocked W layout) that
// land densely in the idle exchange buffer -- exactly the [position][column] order the radix-18 pass reads -- and
// complete on an mbarrier, instead of 9 cp.async per thread pair through the LSU queue.
// REDUCE: the adjoint launch (D strips in, their filtered sum out) -- a template parameter so that each direction is
// its own kernel with its own register allocation.
template <int N, int R1, int R2, int LOGT, int NT, bool use_tma, bool REDUCE>
__global__ void __launch_bounds__(NT, 1) col_warp_kernel(ColParams a, const __grid_constant__ CUtensorMap tmap) {
  constexpr int R0 = 18, M0 = N / R0, L = M0, T = 1 << LOGT, NEL = N << LOGT;
  static_assert(R1 * R2 == L, "block length");
  static_assert(NT == 32 * R0, "one warp per block");
  static_assert(T * R1 == 32, "the radix-R2 pass of one block of all columns fills a warp");
  static_assert(T * M0 <= NT, "one radix-18 butterfly per thread");
  constexpr int IT1 = (R2 * T + 31) / 32;  // warp iterations of the radix-R1 pass (R2 butterflies per column)
  static_assert(N % 4 == 0 && (M0 % 2) == 0, "2x padded geometry");
  constexpr int PAD = N / 4, HALF = M0 / 2;
  constexpr int M1 = R2;                 // stride of the radix-R1 pass inside a block
  constexpr int TAB1 = (R1 - 1) * M1;    // W_L^(j q), q = 1..R1-1, j < M1
  extern __shared__ __align__(128) float2 smem[];
  float2* const bufA = smem;
  float2* const bufB = bufA + NEL;
  float2* const bufX = bufB + NEL;
  float2* const tab1 = bufX + NEL;
  // W_N^(j q), q = 1..TABQ, j < M0: the first half of the radix-18 twiddles of butterfly j; the other half are one
  // product each (q = q/2 + (q - q/2), both <= TABQ).  With only the first powers in the table the 16 products per
  // butterfly were ~4 % of the kernel's FP32 instructions.
  constexpr int TABQ = LHG_COL_TABQ;
  float2* const tab0 = tab1 + TAB1;
  float* const sbeta = reinterpret_cast<float*>(tab0 + TABQ * M0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float2* __restrict__ tw = a.f.tw;
  const int tiles_per_plane = a.Cp >> LOGT;
  const long long n_tiles = (long long)a.S * a.n_colour * tiles_per_plane;
  const int R = a.R, Cp = a.Cp;
  const int use_h = a.use_h;
  const bool masked = (a.flags & kFilterMask) != 0;
  const float bsign = (a.flags & kFilterConj) ? -1.0f : 1.0f;
  const size_t strip = (size_t)R * Cp;

  for (int e = tid; e < TAB1; e += NT) {
    const int q = e / M1 + 1, j = e - (q - 1) * M1;
    tab1[e] = __ldg(tw + (size_t)(j * q) * R0);
  }
  for (int e = tid; e < TABQ * M0; e += NT) {
    const int q = e / M0 + 1, j = e - (q - 1) * M0;
    tab0[e] = __ldg(tw + j * q);  // j q < N
  }
  __shared__ __align__(8) unsigned long long tma_bar[2];  // one per exchange buffer (bufA, bufB)
  unsigned tma_phase = 0;                                 // bit w = parity of buffer w's next completion (every thread)
  // split-phase depth loop (LHG_COL_SPLIT): "exchange buffer written by every warp" (full) and "read by every warp"
  // (empty) are mbarriers with one arrival per warp, so a warp announces its phase and goes on with independent work
  // instead of standing at a CTA barrier: index 0/1 = full bufB/bufA, 2/3 = empty bufB/bufA
  __shared__ __align__(8) unsigned long long ph_bar[4];
  unsigned ph_phase = 0;
  constexpr int TMA_TID = LHG_COL_SPLIT ? NT - 32 : 0;    // the thread that issues the bulk copies (a warp without radix-18 work)
  if (use_tma && tid == 0) {
    mbar_init(&tma_bar[0], 1);
    mbar_init(&tma_bar[1], 1);
  }
  if (LHG_COL_SPLIT && tid == 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(&ph_bar[i], R0);
  }
  __syncthreads();

  // ---- radix-18 pass across the CTA: item b -> column t = b & (T-1), butterfly j = b >> LOGT -----------
  const int j0 = tid >> LOGT, t0 = tid & (T - 1);
  // Row j0 + k*M0 - PAD of the strip: M0 is a multiple of 8, so in the blocked layouts k moves the address by
  // a constant; everything else is fixed per tile.  off5_* = offset of k = 5 (valid for every j0).
  auto kstride = [&](int b) { return b ? (((long long)(M0 / 8) * (Cp >> b)) << (3 + b)) : (long long)M0 * Cp; };
  const long long kstr_in = kstride(a.blocked_in), kstr_out = kstride(a.blocked_out);
  long long off5_in = 0, off5_out = 0;
  const bool p0_active = tid < T * M0;
  const bool hi4 = j0 < HALF;
  auto twiddles18 = [&](float2 (&w)[18]) {
    w[0] = make_float2(1.0f, 0.0f);
#pragma unroll
    for (int q = 1; q <= TABQ; ++q) w[q] = tab0[(q - 1) * M0 + j0];
    tw_chain_step<18, TABQ + 1>(w);
  };
  // DIF: the 9 non-pad samples of butterfly j0 from global memory -> buf (all 18 outputs)
  // the non-pad samples of butterfly j0 of BOTH columns of a pair (t0, t0 ^ 1: adjacent in every W layout, 16 bytes),
  // asynchronously from global memory to their slots of buf.  The even lane of the pair copies the even n2, the odd
  // lane the odd ones: 9 16-byte copies per pair instead of 18 8-byte ones (the adjoint launch's top stall was the
  // MIO queue).  A thread reads back what its neighbour lane copied: cp.async.wait_all + __syncwarp().
  auto stage_inputs = [&](const float2* __restrict__ src, float2* buf) {
    if (!p0_active) return;
#ifdef LHG_COL_STAGE8
#pragma unroll
    for (int n2 = 0; n2 < 9; ++n2) {
      const int k = n2 < 4 ? n2 + 9 : (n2 > 4 ? n2 : (hi4 ? 13 : 4));
      cp_async8(buf + ((j0 + k * M0) << LOGT) + t0, src + (off5_in + (k - 5) * kstr_in));
    }
#else
    const int odd = t0 & 1;
    const float2* s16 = src + (off5_in - odd);
    float2* d16 = buf + (t0 - odd);
#pragma unroll
    for (int n2 = 0; n2 < 9; ++n2) {
      const int k = n2 < 4 ? n2 + 9 : (n2 > 4 ? n2 : (hi4 ? 13 : 4));
      if ((n2 & 1) == odd) cp_async16(d16 + ((j0 + k * M0) << LOGT), s16 + (k - 5) * kstr_in);
    }
#endif
    cp_async_commit();
  };
  // the strip of global plane `plane` (R rows of this tile's T columns) into the non-pad positions of buf
  constexpr int NBOX = (N / 2 / 8 > 256) ? 2 : 1;  // R = N/2 rows = N/16 blocks of 8; a box dimension holds 256
  static_assert((N / 16) % NBOX == 0, "whole row blocks per box");
  // Measured and not kept (profiles/r02_summary.md section 5): the strip as 18 small boxes, each issued by the warp
  // that owns the block it lands in (5.54 vs 5.12 ms per C4 step: the boxes queue in the SM's one TMA unit), and
  // buffer B's strips by cp.async from the owning warps (equal).
  auto stage_tma = [&](size_t plane, float2* buf, int which, int col0) {
    if (tid == TMA_TID) {
      fence_proxy_async();  // the buffer's earlier generic-proxy accesses (ordered by the barrier) before the async writes
      mbar_expect_tx(&tma_bar[which], (unsigned)((N / 2) * T * sizeof(float2)));
      const int b = a.blocked_in;
      const int piece = col0 >> b, inner = (col0 & ((1 << b) - 1)) * 2;
      const int blk0 = (int)((plane * (size_t)(N / 2)) >> 3);
#pragma unroll
      for (int h = 0; h < NBOX; ++h)
        tma_load_4d(buf + ((PAD + h * (N / 2 / NBOX)) << LOGT), &tmap, inner, 0, piece, blk0 + h * (N / 16 / NBOX),
                    &tma_bar[which]);
    }
  };
  // the same boxes as far as L2 only: issued a pair of depths ahead of stage_tma, so that the copy into shared memory
  // (which can only start once the buffer is free) does not wait for DRAM
  auto prefetch_strip = [&](size_t plane, int col0) {
    if (tid == TMA_TID) {
      const int b = a.blocked_in;
      const int piece = col0 >> b, inner = (col0 & ((1 << b) - 1)) * 2;
      const int blk0 = (int)((plane * (size_t)(N / 2)) >> 3);
#pragma unroll
      for (int h = 0; h < NBOX; ++h) tma_prefetch_4d(&tmap, inner, 0, piece, blk0 + h * (N / 16 / NBOX));
    }
  };
  // rbuf: where a staged strip was put (the adjoint transforms it in place, the forward launch reads the tile's one
  // strip from the idle buffer and writes the butterflies into bufA)
  auto pass0_forward = [&](const float2* __restrict__ src, float2* buf, bool staged, const float2* rbuf) {
    if (use_tma && staged) {  // every thread keeps the parities; only the radix-18 threads read the strip
      const int which = rbuf == bufA ? 0 : 1;
      if (p0_active) mbar_wait_bounded(&tma_bar[which], (tma_phase >> which) & 1u);
      tma_phase ^= 1u << which;
    }
    if (!p0_active) return;
    float2 x[9];
    if (staged && !use_tma) {
      cp_async_wait_all();
#ifndef LHG_COL_STAGE8
      __syncwarp();
#endif
    }
#pragma unroll
    for (int n2 = 0; n2 < 9; ++n2) {
      const int k = n2 < 4 ? n2 + 9 : (n2 > 4 ? n2 : (hi4 ? 13 : 4));
      if (staged) x[n2] = rbuf[((j0 + k * M0) << LOGT) + t0];
      else x[n2] = __ldg(src + (off5_in + (k - 5) * kstr_in));
    }
    float2 v[18], w[18];
    dft18_in9(x, hi4, v);
    twiddles18(w);
#pragma unroll
    for (int q = 0; q < 18; ++q) {
      if (q > 0) v[q] = cmul(v[q], w[q]);
      buf[((j0 + q * M0) << LOGT) + t0] = v[q];
    }
  };
  // DIT: buf -> the 9 outputs of butterfly j0 inside the crop -> global memory (re/im swapped back)
  auto pass0_inverse = [&](const float2* buf, float2* __restrict__ dst) {
    if (!p0_active) return;
    float2 v[18], w[18];
    twiddles18(w);
#pragma unroll
    for (int q = 0; q < 18; ++q) {
      v[q] = buf[((j0 + q * M0) << LOGT) + t0];
      if (q > 0) v[q] = cmul(v[q], w[q]);
    }
    float2 o[10];
    dft18_out4_13(v, o);
    dst += off5_out - kstr_out;  // k = 4
#pragma unroll
    for (int i = 0; i < 10; ++i, dst += kstr_out) {
      const int k = 4 + i;
      if ((k == 4 && hi4) || (k == 13 && !hi4)) continue;
#ifdef LHG_EXP_NOSTG
      if (o[i].x == 12345.678f)
#endif
      *dst = cswap(o[i]);
    }
  };

  // ---- warp-local passes on block `warp` of both columns ----------------------------------------------
  const int lt = lane & (T - 1), lj = lane >> LOGT;
  const int bbase = warp * L;
  constexpr bool p2_active = true;  // radix-R2 pass: R1 butterflies per column, T * R1 = 32 lanes
  auto pass1 = [&](float2* buf, auto dit_tag) {  // radix-R1 pass: R2 butterflies per column
    constexpr bool DIT = decltype(dit_tag)::value;
#pragma unroll
    for (int it = 0; it < IT1; ++it) {
      const int j1 = lj + it * (32 >> LOGT);
      if (j1 < R2) {
        float2 v[R1];
        float2* p = buf + ((bbase + j1) << LOGT) + lt;
#pragma unroll
        for (int k = 0; k < R1; ++k) v[k] = p[(k * M1) << LOGT];
        if (!DIT) Dft<R1>::run(v);
#pragma unroll
        for (int q = 1; q < R1; ++q) v[q] = cmul(v[q], tab1[(q - 1) * M1 + j1]);
        if (DIT) Dft<R1>::run(v);
#pragma unroll
        for (int k = 0; k < R1; ++k) p[(k * M1) << LOGT] = v[k];
      }
    }
    __syncwarp();
  };

  // Tiles are taken in PAIRS (2m, 2m+1): in the blocked W layout the two tiles of a pair share every
  // 32-byte sector, and L2 only merges their half-sector writes if they arrive within its residency window
  // (measured: with the pair split over two CTAs the column kernel wrote every sector twice and read it
  // back in between: 11.6 GB of DRAM traffic for 4.0 GB of algorithmic bytes).
  // TMA builds: the FIRST strip of a tile is staged too -- by the previous tile of this CTA while its last depth is
  // still being transformed (into the exchange buffer that depth leaves idle; D even), or at the tile's own start
  // when there was no live predecessor.  staged_tile is uniform over the CTA.
  const bool first_tma = use_tma && (a.D & 1) == 0;
  int staged_tile = -1;  // low 32 bits of the tile index
  bool e3_pending = false;  // forward launch: the last "bufA read by every warp" phase has not been waited for yet
  auto tile_live = [&](long long tl) {
    return !(masked && a.tile_active && !a.tile_active[(int)(tl % tiles_per_plane)]);
  };
  auto stage_first = [&](long long tl) {
    if (tl < n_tiles && tile_live(tl)) {
      const long long g2 = tl / tiles_per_plane;
      const size_t plane = REDUCE ? (size_t)(g2 / a.n_colour) * a.D * a.n_colour + (size_t)(g2 % a.n_colour) : (size_t)g2;
      stage_tma(plane, REDUCE ? bufA : bufB, REDUCE ? 0 : 1, (int)(tl % tiles_per_plane) << LOGT);
      staged_tile = (int)tl;
    }
  };
  auto next_tile_of = [&](long long tl) { return (tl & 1) ? tl + 2 * (long long)gridDim.x - 1 : tl + 1; };
  const long long n_pairs = (n_tiles + 1) >> 1;
  for (long long it2 = 2 * (long long)blockIdx.x; it2 < 2 * n_pairs; it2 += ((it2 & 1) ? 2 * (long long)gridDim.x - 1 : 1)) {
    const long long tile = it2;
    if (tile >= n_tiles) continue;
    const int ct = (int)(tile % tiles_per_plane);
    const long long g = tile / tiles_per_plane;  // sample * n_colour + colour
    const int colour = (int)(g % a.n_colour);
    const long long s = g / a.n_colour;
    const int col0 = ct << LOGT;
    off5_in = (long long)woff(a.blocked_in, Cp, j0 + 5 * M0 - PAD, col0 + t0);
    off5_out = (long long)woff(a.blocked_out, Cp, j0 + 5 * M0 - PAD, col0 + t0);

    if (masked && a.tile_active && !a.tile_active[ct]) {
      // every bin of these columns is outside the circular mask: the result is zero (and the row kernel
      // knows, when both are the compile-time planned ones)
      const int n_out = a.rows_skip_dead ? 0 : (REDUCE ? 1 : a.D);
      for (int d = 0; d < n_out; ++d) {
        const size_t plane = REDUCE ? (size_t)g : ((size_t)s * a.D + d) * a.n_colour + colour;
        float2* dst = a.out + plane * strip;
        for (int e = tid; e < R * (T / 2); e += NT)
          *reinterpret_cast<float4*>(dst + woff(a.blocked_out, Cp, e / (T / 2), col0 + 2 * (e % (T / 2)))) =
              make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
      continue;
    }

    // per-depth phase slopes (visible after the first barrier below)
    for (int d = tid; d < a.D; d += NT) {
      const int zi = a.depth_index ? a.depth_index[s * a.D + d] : d;
      sbeta[d] = use_h ? bsign * beta_of(a.z[zi]) : 0.0f;
    }
    const bool pre_staged = staged_tile == (int)tile;  // the previous tile of this CTA staged this one's first strip(s)
    if (first_tma && !pre_staged) stage_first(tile);   // no live predecessor: all buffers are free here
    // the 9 samples this thread starts the next tile of this CTA from: into L2 while this tile is transformed
    if (!first_tma) {
      const long long nt = next_tile_of(tile);
      if (nt < n_tiles && p0_active) {
        const int nct = (int)(nt % tiles_per_plane);
        const long long ng = nt / tiles_per_plane;
        const size_t nplane = REDUCE ? (size_t)(ng / a.n_colour) * a.D * a.n_colour + (size_t)(ng % a.n_colour) : (size_t)ng;
        const float2* nsrc = a.in + nplane * strip + woff(a.blocked_in, Cp, j0 + 5 * M0 - PAD, (nct << LOGT) + t0);
#pragma unroll
        for (int k = 4; k < 14; ++k) prefetch_l2(nsrc + (k - 5) * kstr_in);
      }
    }
    // w (sign bit = outside the mask) of the R2 bins this lane owns in the radix-R2 pass
    float wreg[R2];
    if (a.wmt && p2_active) {
      const float* wsrc = a.wmt + ((size_t)colour * tiles_per_plane + ct) * NEL;
#pragma unroll
      for (int k = 0; k < R2; ++k) wreg[k] = __ldg(wsrc + ((bbase + lj * R2 + k) << LOGT) + lt);
    } else {
#pragma unroll
      for (int k = 0; k < R2; ++k) wreg[k] = 0.0f;
    }
    float2* const xp = bufX + ((bbase + lj * R2) << LOGT) + lt;  // this lane's R2 spectrum bins (stride T)
#ifdef LHG_COL_DEAD_VOTES
    // (kept only as an A/B variant) bit k of `dead` = bin k is outside the mask in all lanes of the warp.  The mask
    // is never consulted (skipping bins inside the transfer-function loop serialised its chains), but its 15 votes
    // made every warp wait for its w loads here, BEFORE the strip loads of the radix-18 pass were issued.
    unsigned dead = 0;
    if (masked) {
#pragma unroll
      for (int k = 0; k < R2; ++k)
        if (__all_sync(0xffffffffu, signbit(wreg[k]))) dead |= 1u << k;
    }
    if (dead == 0xdeadbeefu) sbeta[0] = 0.0f;
#endif

    if constexpr (!REDUCE) {
      // The tile's forward transform runs in bufX (LHG_COL_XFWD): the previous tile's spectrum there is dead once any
      // warp has passed "every warp has written its last depth", whereas bufA is still being read by the radix-18
      // pass of that depth -- so a warp that is through with the previous tile starts this one without waiting for
      // the others (the wait for "bufA read by every warp" is taken after the barrier below, where it is over).
      float2* const fbuf = LHG_COL_XFWD ? bufX : bufA;
      pass0_forward(a.in + (size_t)g * strip, fbuf, first_tma, bufB);
      __syncthreads();
      if (e3_pending) {
        mbar_wait_bounded(&ph_bar[3], (ph_phase >> 3) & 1u);
        ph_phase ^= 1u << 3;
        e3_pending = false;
      }
      pass1(fbuf, std::false_type{});
      if (p2_active) {  // radix-R2 DIF, masked spectrum into bufX
        float2 v[R2];
        const float2* p = fbuf + ((bbase + lj * R2) << LOGT) + lt;
#pragma unroll
        for (int k = 0; k < R2; ++k) v[k] = p[k << LOGT];
        Dft<R2>::run(v);
#pragma unroll
        for (int k = 0; k < R2; ++k) {
          if (masked && signbit(wreg[k])) v[k] = make_float2(0.0f, 0.0f);
          xp[k << LOGT] = v[k];
        }
      }
      // warp-local half of the inverse transform of depth d: spectrum x transfer function, radix-R2 DIT, radix-R1 DIT
      auto depth_local = [&](int d, float2* buf) {
        if (p2_active) {
          const float beta = sbeta[d], beta_t = beta * 0.15915494309189535f;
          float2 v[R2];
          if (use_h) {
            constexpr int G = 5;  // chains in flight together (R2 = 3 groups of 5)
            static_assert(R2 % G == 0, "group size");
#pragma unroll
            for (int k0 = 0; k0 < R2; k0 += G) {
              float wa[G];
              float2 h[G];
#pragma unroll
              for (int i = 0; i < G; ++i) wa[i] = fabsf(wreg[k0 + i]);
              fast_cis_group<G>(beta, beta_t, wa, h);
#pragma unroll
              for (int i = 0; i < G; ++i) v[k0 + i] = cswap(cmul(xp[(k0 + i) << LOGT], h[i]));
            }
          } else {
#pragma unroll
            for (int k = 0; k < R2; ++k) v[k] = cswap(xp[k << LOGT]);
          }
          Dft<R2>::run(v);
          float2* p = buf + ((bbase + lj * R2) << LOGT) + lt;
#pragma unroll
          for (int k = 0; k < R2; ++k) p[k << LOGT] = v[k];
        }
        __syncwarp();
        pass1(buf, std::true_type{});
      };
      if (LHG_COL_SPLIT && (a.D & 1) == 0) {
        // Depths in pairs, W W R R: between announcing a phase and waiting for everybody else's there is always a
        // whole phase of this warp's own work, so the warps drift instead of meeting at a CTA barrier per depth.
        auto warp_arrive = [&](int i) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&ph_bar[i]);
        };
        auto ph_wait = [&](int i) {
          mbar_wait_bounded(&ph_bar[i], (ph_phase >> i) & 1u);
          ph_phase ^= 1u << i;
        };
        for (int d = 0; d < a.D; d += 2) {
          if (LHG_COL_PREFETCH && first_tma && d + 2 >= a.D) {  // last pair: the next tile's strip, as far as L2
            const long long tl = next_tile_of(tile);
            if (tl < n_tiles && tile_live(tl)) prefetch_strip((size_t)(tl / tiles_per_plane), (int)(tl % tiles_per_plane) << LOGT);
          }
          // (the two halves of a pair are loops, not copies: one set of live registers)
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {  // depth d in bufB, d + 1 in bufA
            if (d > 0) ph_wait(2 + h);
            depth_local(d + h, h ? bufA : bufB);
            warp_arrive(h);
          }
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            ph_wait(h);
            pass0_inverse(h ? bufA : bufB, a.out + (((size_t)s * a.D + d + h) * a.n_colour + colour) * strip);
            warp_arrive(2 + h);
          }
        }
        // every radix-18 pass over bufB is through: the next tile's strip travels into it (issued by a warp that has
        // no radix-18 work, i.e. while the others are still in the last depth)
        ph_wait(2);
        if (first_tma) stage_first(next_tile_of(tile));
        if (LHG_COL_XFWD) e3_pending = true;  // bufA is not touched before the next tile's first barrier
        else ph_wait(3);                      // ... and over bufA, which the next tile's radix-18 pass rewrites
        continue;
      }
      for (int d = 0; d < a.D; ++d) {
        const size_t out_plane = ((size_t)s * a.D + d) * a.n_colour + colour;
        float2* buf = (d & 1) ? bufA : bufB;
        depth_local(d, buf);
        __syncthreads();
        // last depth (it used bufA, D even): bufB is idle from here on, the next tile's strip travels into it
        if (first_tma && d == a.D - 1) stage_first(next_tile_of(tile));
        pass0_inverse(buf, a.out + out_plane * strip);
      }
    } else {
      // warp-local half of the forward transform of depth d: radix-R1 DIF, radix-R2 DIF, x conj-able transfer function,
      // accumulated over depth in bufX
      auto depth_local = [&](int d, float2* buf) {
        pass1(buf, std::false_type{});
        if (p2_active) {
          const float beta = sbeta[d], beta_t = beta * 0.15915494309189535f;
          float2 v[R2];
          const float2* p = buf + ((bbase + lj * R2) << LOGT) + lt;
#pragma unroll
          for (int k = 0; k < R2; ++k) v[k] = p[k << LOGT];
          Dft<R2>::run(v);
          constexpr int G = 5;
#pragma unroll
          for (int k0 = 0; k0 < R2; k0 += G) {
            // (skipping the groups of 5 bins that lie outside the mask in a whole warp -- 29 % of the bins of the live
            // columns are outside -- measured WORSE, 4.99 vs 4.92 ms per C4 step: profiles/r02_summary.md section 5)
            if (use_h) {
              float wa[G];
              float2 h[G];
#pragma unroll
              for (int i = 0; i < G; ++i) wa[i] = fabsf(wreg[k0 + i]);
              fast_cis_group<G>(beta, beta_t, wa, h);
#pragma unroll
              for (int i = 0; i < G; ++i) v[k0 + i] = cmul(v[k0 + i], h[i]);
            }
#pragma unroll
            for (int i = 0; i < G; ++i) {
              float2 x = v[k0 + i];
              if (d > 0) x = cadd(x, xp[(k0 + i) << LOGT]);
              xp[(k0 + i) << LOGT] = x;
            }
          }
        }
      };
      if (LHG_COL_SPLIT && use_tma && (a.D & 1) == 0) {
        // Depths in pairs, R R W W (see the forward launch).  Only the warp that issues the bulk copies waits for
        // "every warp has read buffer x" (ph_bar 2/3); the strip of depth d+2 travels while depth d+1 is transformed.
        auto warp_arrive = [&](int i) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&ph_bar[i]);
        };
        auto ph_wait = [&](int i) {
          mbar_wait_bounded(&ph_bar[i], (ph_phase >> i) & 1u);
          ph_phase ^= 1u << i;
        };
        const bool producer = warp == (TMA_TID >> 5);
        const size_t plane0 = (size_t)s * a.D * a.n_colour + colour;
        // the second strip: staged by the previous tile too (LHG_COL_XINV), else now (bufB is free: every warp has
        // passed the barrier inside the previous tile's last inverse transform, after its last read of bufB)
        if (!(LHG_COL_XINV && pre_staged)) stage_tma(plane0 + a.n_colour, bufB, 1, col0);
        const long long ntile = next_tile_of(tile);
        const bool nlive = ntile < n_tiles && tile_live(ntile);
        const size_t nplane0 = (size_t)((ntile / tiles_per_plane) / a.n_colour) * a.D * a.n_colour +
                               (size_t)((ntile / tiles_per_plane) % a.n_colour);
        const int ncol0 = (int)(ntile % tiles_per_plane) << LOGT;
        if (LHG_COL_PREFETCH && a.D > 2) {
          prefetch_strip(plane0 + 2 * (size_t)a.n_colour, col0);
          prefetch_strip(plane0 + 3 * (size_t)a.n_colour, col0);
        }
        for (int d = 0; d < a.D; d += 2) {
          if (LHG_COL_PREFETCH) {
            if (d + 4 < a.D) {
              prefetch_strip(plane0 + (size_t)(d + 4) * a.n_colour, col0);
              prefetch_strip(plane0 + (size_t)(d + 5) * a.n_colour, col0);
            } else if (d + 2 >= a.D && nlive) {  // last pair: the second strip of the next tile
              prefetch_strip(nplane0 + a.n_colour, ncol0);
            }
          }
          // (the two halves of a pair are loops, not copies: one set of live registers)
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {  // depth d in bufA, d + 1 in bufB
            float2* buf = h ? bufB : bufA;
            pass0_forward(nullptr, buf, true, buf);
            warp_arrive(1 - h);
          }
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            ph_wait(1 - h);
            depth_local(d + h, h ? bufB : bufA);
            // one thread stages whole strips: it waits until every warp has read the buffer
            warp_arrive(3 - h);
            if (producer) ph_wait(3 - h);
            // the buffer is idle until depth d + h + 2 (or, after the last even depth, the next tile)
            if (d + h + 2 < a.D) stage_tma(plane0 + (size_t)(d + h + 2) * a.n_colour, h ? bufB : bufA, h, col0);
            else if (h == 0) stage_first(ntile);
            else if (LHG_COL_XINV && nlive) stage_tma(nplane0 + a.n_colour, bufB, 1, ncol0);
          }
        }
      } else
      for (int d = 0; d < a.D; ++d) {
        const size_t in_plane = ((size_t)s * a.D + d) * a.n_colour + colour;
        const float2* src = a.in + in_plane * strip;
        float2* buf = (d & 1) ? bufB : bufA;
        pass0_forward(src, buf, d > 0 || first_tma, buf);
        __syncthreads();
        // the other exchange buffer is idle until the next depth's radix-18 pass (its last readers passed
        // the barrier above): the next strip travels into it while this one is transformed
        if (d + 1 < a.D) {
          if constexpr (use_tma) stage_tma(in_plane + a.n_colour, (d & 1) ? bufA : bufB, (d & 1) ? 0 : 1, col0);
          else stage_inputs(src + (size_t)a.n_colour * strip, (d & 1) ? bufA : bufB);
        } else if (first_tma) {
          stage_first(next_tile_of(tile));  // last depth (in bufB, D even): bufA is idle until the next tile
        }
        depth_local(d, buf);
      }
      // the exchange buffer of the last depth: this warp's block was last read by this warp.  Paired loop: the
      // accumulator buffer itself (a lane's radix-R2 butterfly returns to the bins it came from), which leaves both
      // exchange buffers to the next tile's strips and needs no barrier before the next tile: its first write to
      // bufX comes after a phase barrier that every warp reaches only after this tile's radix-18 pass.
      const bool xinv = LHG_COL_XINV && LHG_COL_SPLIT && use_tma && (a.D & 1) == 0;
      float2* buf = xinv ? bufX : (((a.D - 1) & 1) ? bufB : bufA);
      if (p2_active) {
        float2 v[R2];
#pragma unroll
        for (int k = 0; k < R2; ++k) {
          float2 x = xp[k << LOGT];
          if (masked && signbit(wreg[k])) x = make_float2(0.0f, 0.0f);
          v[k] = cswap(x);
        }
        Dft<R2>::run(v);
        float2* p = buf + ((bbase + lj * R2) << LOGT) + lt;
#pragma unroll
        for (int k = 0; k < R2; ++k) p[k << LOGT] = v[k];
      }
      __syncwarp();
      pass1(buf, std::true_type{});
      __syncthreads();
      pass0_inverse(buf, a.out + (size_t)g * strip);
      if (xinv) continue;
    }
    __syncthreads();  // the next tile's radix-18 pass rewrites bufA
  }
}

}  // namespace asmb
