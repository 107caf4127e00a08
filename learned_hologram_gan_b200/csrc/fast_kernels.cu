// fast_kernels.cu -- compile-time planned kernels for the transform lengths of the BASELINE configs.
//
// Same three-kernel pipeline as asm_b200.cu, specialised per (length, pad):
//   * radix 8..18 butterflies in registers on the packed FP32x2 pipe, 3 (or 4) passes per transform instead
//     of 5-6, twiddles from shared-memory tables or product trees;
//   * the zero-pad rows/columns are never loaded or added (pruned first butterfly) and the cropped-away
//     outputs of the last inverse butterfly are never computed;
//   * row kernels (here): one row per CTA, global memory touched only by cooperative 16-byte accesses
//     (kind-specialised prologue / epilogue over groups of 4 samples), passes in place in shared memory;
//   * column kernels: col_warp.cuh / col_warp16.cuh (warp-local inner passes; the 4320-, padded 2160- and
//     1024-point columns) and col_fast_kernel below (CTA-synchronous, pairs of columns; everything else): the
//     (masked) forward spectrum of a tile stays in shared memory across the depth loop, the transfer function
//     is generated per depth from the w values with the SFU sin/cos and multiplied in as the load stage of the
//     first inverse pass; the adjoint accumulates the depth sum in the same buffer.  Tiles that lie completely
//     outside the circular mask are not transformed.
//   * W1/W2 keep the scrambled column order of the row transform (no reordering pass) and are stored in 8-row
//     blocks when the column tiles are narrow (common.cuh woff); the w/mask grid is pre-permuted into the tile
//     order of the column kernel once per geometry (asm_build_wm_tiled).
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <set>
#include <utility>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "fft_fast.cuh"
#include "col_warp.cuh"
#include "col_warp16.cuh"

namespace asmb {

// ------------------------------------------------------------------------------------------------
// column kernel
// ------------------------------------------------------------------------------------------------
// One CTA owns a tile of T adjacent columns; a thread runs each butterfly on a PAIR of columns (16-byte
// shared and global accesses, one set of twiddles).  Shared memory: two exchange buffers used alternately
// (so the global stores of one depth overlap the transfer-function pass of the next and only two barriers
// per transform remain), bufX (forward mode: masked spectrum of the tile, kept across the depth loop;
// reduce mode: the depth-sum accumulator) and the twiddle tables.  The w values of the bins a thread owns
// in the last forward / first inverse pass live in its registers.
template <class P, int LOGT, int NT, int KLO, int KHI>
__global__ void __launch_bounds__(NT, 1) col_fast_kernel(ColParams a) {
  extern __shared__ float2 smem[];
  constexpr int N = P::N, T = 1 << LOGT, TP = T / 2, NEL = N << LOGT, LAST = P::NPASS - 1, R0 = P::R0, M0 = N / R0;
  constexpr int RL = P::radix(LAST);
  constexpr int NBL = (N / RL) * TP;                  // work items of the last pass
  constexpr int ITL = (NBL + NT - 1) / NT;            // ... per thread
  constexpr int TW0 = 2;                              // pass 0: product tree from a shared-memory table
  float2* const bufA = smem;
  float2* const bufB = bufA + NEL;
  float2* const bufX = bufB + NEL;
  float2* const tabs = bufX + NEL;
  float* const sbeta = reinterpret_cast<float*>(tabs + P::tab_total(TW0));  // [D] fl(-2 pi z_d) of this sample
  const int tid = threadIdx.x;
  const float2* __restrict__ tw = a.f.tw;
  const int tiles_per_plane = a.Cp >> LOGT;
  const long long n_tiles = (long long)a.S * a.n_colour * tiles_per_plane;
  const int R = a.R, Cp = a.Cp;
  const int use_h = a.use_h;
  const bool masked = (a.flags & kFilterMask) != 0;
  const float bsign = (a.flags & kFilterConj) ? -1.0f : 1.0f;
  const size_t strip = (size_t)R * Cp;
  // spectrum-in / spectrum-out calls: that side is a natural-order padded spectrum [plane][N][Cp] in global
  // memory (rows looked up through the plan's digit reversal), W1 / W2 then hold natural-order columns
  const bool in_full = a.in_full != 0, out_full = a.out_full != 0;
  const size_t full = (size_t)N * Cp;
  const bool late_mask = masked && in_full;

  fill_tables<P, TW0>(tabs, tw, tid, NT);
  __syncthreads();
  int col0g = 0;  // first column of the current tile

  auto s4 = [](float2* buf, int row, int tp) { return reinterpret_cast<float4*>(buf + ((row << LOGT) + 2 * tp)); };
  // middle passes (1 .. NPASS-2), each followed by a barrier
  auto middle = [&](float2* buf, auto dit_tag) {
    constexpr bool DIT = decltype(dit_tag)::value;
    auto ld = [&](int row, int tp, int, int) { return *s4(buf, row, tp); };
    auto st = [&](int row, int tp, int, int, float4 v) { *s4(buf, row, tp) = v; };
    if constexpr (P::NPASS >= 4 && DIT) {
      fpass2<P, 2, LOGT, NT, true, 1, 0, P::radix(2)>(tw, tabs + P::tab_off(2, TW0), tid, ld, st);
      __syncthreads();
    }
    if constexpr (P::NPASS >= 3) {
      fpass2<P, 1, LOGT, NT, DIT, 1, 0, P::radix(1)>(tw, tabs + P::tab_off(1, TW0), tid, ld, st);
      __syncthreads();
    }
    if constexpr (P::NPASS >= 4 && !DIT) {
      fpass2<P, 2, LOGT, NT, false, 1, 0, P::radix(2)>(tw, tabs + P::tab_off(2, TW0), tid, ld, st);
      __syncthreads();
    }
  };
  // forward transform of one stored strip (the R non-pad rows of the tile's columns) through buf into st_last
  auto forward = [&](const float2* __restrict__ src, float2* buf, auto st_last) {
    auto ld_g = [&](int row, int tp, int, int) {
      return __ldg(reinterpret_cast<const float4*>(src + woff(a.blocked_in, Cp, row - KLO * M0, col0g + 2 * tp)));
    };
    auto st_s = [&](int row, int tp, int, int, float4 v) { *s4(buf, row, tp) = v; };
    auto ld_s = [&](int row, int tp, int, int) { return *s4(buf, row, tp); };
    fpass2<P, 0, LOGT, NT, false, TW0, KLO, KHI>(tw, tabs, tid, ld_g, st_s);
    __syncthreads();
    middle(buf, std::false_type{});
    fpass2<P, LAST, LOGT, NT, false, 1, 0, RL>(tw, tabs, tid, ld_s, st_last);
  };
  // inverse transform from ld_first through buf to the R crop rows of dst (no trailing barrier)
  auto inverse = [&](auto ld_first, float2* buf, float2* __restrict__ dst) {
    auto st_s = [&](int row, int tp, int, int, float4 v) { *s4(buf, row, tp) = v; };
    auto ld_s = [&](int row, int tp, int, int) { return *s4(buf, row, tp); };
    fpass2<P, LAST, LOGT, NT, true, 1, 0, RL>(tw, tabs, tid, ld_first, st_s);
    __syncthreads();
    middle(buf, std::true_type{});
    auto st_g = [&](int row, int tp, int, int, float4 v) {
      *reinterpret_cast<float4*>(dst + woff(a.blocked_out, Cp, row - KLO * M0, col0g + 2 * tp)) = make_float4(v.y, v.x, v.w, v.z);
    };
    fpass2<P, 0, LOGT, NT, true, TW0, KLO, KHI>(tw, tabs, tid, ld_s, st_g);
  };
  // spectrum out: every thread writes the bins it owns in the last forward pass (value(it, k, row, tp) returns
  // the two bins of its column pair) to their natural rows of dst
  auto emit = [&](float2* __restrict__ dst, auto value) {
#pragma unroll
    for (int it = 0; it < ITL; ++it) {
      const int b = tid + it * NT;
      if (ITL * NT == NBL || b < NBL) {
        const int tp = b & (TP - 1), jj = b >> (LOGT - 1);
        float2* drow = dst + (size_t)P::perm(jj * RL) * Cp + col0g + 2 * tp;
#pragma unroll
        for (int k = 0; k < RL; ++k) {
          float4 v = value(it, k, jj * RL + k, tp);
          v.x *= a.out_scale, v.y *= a.out_scale, v.z *= a.out_scale, v.w *= a.out_scale;
          *reinterpret_cast<float4*>(drow + (size_t)k * (N / RL) * Cp) = v;
        }
      }
    }
  };

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int ct = (int)(tile % tiles_per_plane);
    const long long g = tile / tiles_per_plane;  // sample * n_colour + colour
    const int colour = (int)(g % a.n_colour);
    const long long s = g / a.n_colour;
    const int col0 = ct << LOGT;
    col0g = col0;

    if (masked && a.tile_active && !a.tile_active[ct]) {
      // every bin of these columns is outside the circular mask: the result is zero (and the row kernel
      // knows, when both are the compile-time planned ones)
      const int n_out = a.rows_skip_dead ? 0 : (a.reduce ? 1 : a.D);
      for (int d = 0; d < n_out; ++d) {
        const size_t plane = a.reduce ? (size_t)g : ((size_t)s * a.D + d) * a.n_colour + colour;
        float2* dst = a.out + plane * strip;
        for (int e = tid; e < R * TP; e += NT)
          *reinterpret_cast<float4*>(dst + woff(a.blocked_out, Cp, e / TP, col0 + 2 * (e % TP))) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
      continue;
    }

    // per-depth phase slopes (visible after the first barrier of the forward transform below)
    for (int d = tid; d < a.D; d += NT) {
      const int zi = a.depth_index ? a.depth_index[s * a.D + d] : d;
      sbeta[d] = use_h ? bsign * beta_of(a.z[zi]) : 0.0f;
    }
    // the strip the next tile of this CTA starts from: into L2 while this tile is transformed
    {
      const long long nt = tile + gridDim.x;
      if (nt < n_tiles && in_full) {
        const float2* nsrc = a.in + (size_t)(nt / tiles_per_plane) * full + ((int)(nt % tiles_per_plane) << LOGT);
        for (int r = tid; r < N; r += NT) {
#pragma unroll
          for (int c = 0; c < T; c += 4) prefetch_l2(nsrc + (size_t)r * Cp + c);  // one per 32-byte sector
        }
      }
      if (nt < n_tiles && !in_full) {
        const int nct = (int)(nt % tiles_per_plane);
        const long long ng = nt / tiles_per_plane;
        const size_t nplane = a.reduce ? (size_t)(ng / a.n_colour) * a.D * a.n_colour + (size_t)(ng % a.n_colour) : (size_t)ng;
        const float2* nsrc = a.in + nplane * strip;
        for (int e = tid; e < R * TP; e += NT) prefetch_l2(nsrc + woff(a.blocked_in, Cp, e / TP, (nct << LOGT) + 2 * (e % TP)));
      }
    }

    // w (sign bit = outside the mask) of the bins this thread owns in the last forward / first inverse pass:
    // element (it, k) of column pair tp sits at scrambled position jj*RL + k of the pre-tiled grid
    float2 wreg[ITL][RL];
    if (a.wmt) {
      const float2* wsrc = reinterpret_cast<const float2*>(a.wmt + ((size_t)colour * tiles_per_plane + ct) * NEL);
#pragma unroll
      for (int it = 0; it < ITL; ++it) {
        const int b = tid + it * NT;
        if (ITL * NT == NBL || b < NBL) {
          const int tp = b & (TP - 1), jj = b >> (LOGT - 1);
#pragma unroll
          for (int k = 0; k < RL; ++k) wreg[it][k] = __ldg(wsrc + (size_t)(jj * RL + k) * TP + tp);
        }
      }
    } else if (a.wm && (in_full || out_full)) {
      // natural-order columns: the plain w/mask grid, rows through the digit reversal
      const float* wsrc = a.wm + (size_t)colour * full + col0;
#pragma unroll
      for (int it = 0; it < ITL; ++it) {
        const int b = tid + it * NT;
        if (ITL * NT == NBL || b < NBL) {
          const int tp = b & (TP - 1), jj = b >> (LOGT - 1);
          const float* wrow = wsrc + (size_t)P::perm(jj * RL) * Cp + 2 * tp;
#pragma unroll
          for (int k = 0; k < RL; ++k) wreg[it][k] = __ldg(reinterpret_cast<const float2*>(wrow + (size_t)k * (N / RL) * Cp));
        }
      }
    } else {
#pragma unroll
      for (int it = 0; it < ITL; ++it)
#pragma unroll
        for (int k = 0; k < RL; ++k) wreg[it][k] = make_float2(0.0f, 0.0f);
    }
    auto h_of = [&](float2 x, float w, float beta) { return cmul(x, fast_cis(__fmul_rn(beta, fabsf(w)))); };

    if (!a.reduce) {
      if (in_full) {
        // the tile's columns of the given spectrum, natural rows to their scrambled slots (masked on the way
        // into the inverse transform)
        // (all of the tile requested at once, asynchronously: the pieces are 16 * TP bytes, one per row)
        const float2* src = a.in + (size_t)g * full + col0;
        constexpr int LIT = (N * TP + NT - 1) / NT;
#pragma unroll
        for (int i = 0; i < LIT; ++i) {
          const int e = tid + i * NT;
          if (LIT * NT == N * TP || e < N * TP) {
            const int pos = e >> (LOGT - 1), tp = e & (TP - 1);
            cp_async16(s4(bufX, pos, tp), src + (size_t)P::perm(pos) * Cp + 2 * tp);
          }
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
      } else {
        const float2* src = a.in + (size_t)g * strip;
        forward(src, bufA, [&](int row, int tp, int k, int it, float4 v) {
          if (masked) {
            if (signbit(wreg[it][k].x)) v.x = v.y = 0.0f;
            if (signbit(wreg[it][k].y)) v.z = v.w = 0.0f;
          }
          *s4(bufX, row, tp) = v;
        });
      }
      // the last forward pass and the first inverse pass touch the same bufX slots from the same thread:
      // no barrier in between.  Depth d goes through bufB, bufA, bufB, ... (bufA is still being read by
      // slower warps of the forward pass when depth 0 starts).
      for (int d = 0; d < a.D; ++d) {
        const size_t out_plane = ((size_t)s * a.D + d) * a.n_colour + colour;
        const float beta = sbeta[d];
        if (out_full) {
          emit(a.out + out_plane * full, [&](int it, int k, int row, int tp) {
            const float4 x = *s4(bufX, row, tp);
            if (!use_h) return x;
            const float2 p0 = h_of(make_float2(x.x, x.y), wreg[it][k].x, beta);
            const float2 p1 = h_of(make_float2(x.z, x.w), wreg[it][k].y, beta);
            return make_float4(p0.x, p0.y, p1.x, p1.y);
          });
          continue;
        }
        float2* dst = a.out + out_plane * strip;
        float2* buf = (d & 1) ? bufA : bufB;
        if (use_h) {
          inverse([&](int row, int tp, int k, int it) {
            float4 x = *s4(bufX, row, tp);
            if (late_mask) {
              if (signbit(wreg[it][k].x)) x.x = x.y = 0.0f;
              if (signbit(wreg[it][k].y)) x.z = x.w = 0.0f;
            }
            const float2 p0 = h_of(make_float2(x.x, x.y), wreg[it][k].x, beta);
            const float2 p1 = h_of(make_float2(x.z, x.w), wreg[it][k].y, beta);
            return make_float4(p0.y, p0.x, p1.y, p1.x);
          }, buf, dst);
        } else {
          inverse([&](int row, int tp, int k, int it) {
            float4 x = *s4(bufX, row, tp);
            if (late_mask) {
              if (signbit(wreg[it][k].x)) x.x = x.y = 0.0f;
              if (signbit(wreg[it][k].y)) x.z = x.w = 0.0f;
            }
            return make_float4(x.y, x.x, x.w, x.z);
          }, buf, dst);
        }
      }
    } else {
      for (int d = 0; d < a.D; ++d) {
        const size_t in_plane = ((size_t)s * a.D + d) * a.n_colour + colour;
        const float2* src = a.in + in_plane * strip;
        if (d + 1 < a.D) {  // pull the next depth's strip into L2 while this one is transformed
          const float2* nxt = src + (size_t)a.n_colour * strip;
          for (int e = tid; e < R * TP; e += NT) prefetch_l2(nxt + woff(a.blocked_in, Cp, e / TP, col0 + 2 * (e % TP)));
        }
        const bool first = d == 0;
        float2* buf = (d & 1) ? bufB : bufA;
        forward(src, buf, [&](int row, int tp, int k, int it, float4 v) {
          float2 p0 = make_float2(v.x, v.y), p1 = make_float2(v.z, v.w);
          if (use_h) {
            const float beta = sbeta[d];
            p0 = h_of(p0, wreg[it][k].x, beta);
            p1 = h_of(p1, wreg[it][k].y, beta);
          }
          float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          if (!first) acc = *s4(bufX, row, tp);
          *s4(bufX, row, tp) = make_float4(acc.x + p0.x, acc.y + p0.y, acc.z + p1.x, acc.w + p1.y);
        });
      }
      if (out_full) {
        emit(a.out + (size_t)g * full, [&](int it, int k, int row, int tp) {
          float4 x = *s4(bufX, row, tp);
          if (masked) {
            if (signbit(wreg[it][k].x)) x.x = x.y = 0.0f;
            if (signbit(wreg[it][k].y)) x.z = x.w = 0.0f;
          }
          return x;
        });
      } else
      // the buffer NOT used by the last forward transform (slower warps may still be reading that one)
      inverse([&](int row, int tp, int k, int it) {
        float4 x = *s4(bufX, row, tp);
        if (masked) {
          if (signbit(wreg[it][k].x)) x.x = x.y = 0.0f;
          if (signbit(wreg[it][k].y)) x.z = x.w = 0.0f;
        }
        return make_float4(x.y, x.x, x.w, x.z);
      }, (a.D & 1) ? bufB : bufA, a.out + (size_t)g * strip);
    }
    __syncthreads();  // the next tile's first pass rewrites bufA / bufX
  }
}

// scatter the natural-order w/mask grid into the tile order of col_fast_kernel:
//   wmt[((colour*tiles + c/T)*Rp + pos)*T + c%T] = wm[colour][row_perm[pos]][col_perm[c]]
// (wm == nullptr: IEEE-rounded device values), and flag the column tiles that hold at least one
// bin inside the mask.
__global__ void wm_tiled_kernel(Phys ph, const float* __restrict__ wm, int n_colour, int logT,
                                const int* __restrict__ row_perm, const int* __restrict__ col_perm,
                                float* __restrict__ wmt, int* __restrict__ tile_active) {
  const int Rp = ph.Rp, Cp = ph.Cp, T = 1 << logT;
  const size_t plane = (size_t)Rp * Cp;
  const size_t total = plane * n_colour;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int colour = (int)(i / plane);
    size_t rem = i - (size_t)colour * plane;
    const int tile = (int)(rem / ((size_t)Rp * T));
    rem -= (size_t)tile * Rp * T;
    const int pos = (int)(rem >> logT), t = (int)(rem & (T - 1));
    const int kr = row_perm[pos];
    const int c = (tile << logT) + t;
    const int kc = col_perm ? col_perm[c] : c;
    float w;
    if (wm) {
      w = wm[(size_t)colour * plane + (size_t)kr * Cp + kc];
    } else {
      w = w_value(ph, kr, kc, colour);
      if (radial_value(ph, kr, kc) > ph.radius) w = -w;
    }
    wmt[i] = w;
    if (colour == 0 && !signbit(w)) tile_active[tile] = 1;
  }
}

// ------------------------------------------------------------------------------------------------
// row kernels: T rows per CTA, planar in shared memory ([t][N])
// ---- row plans whose last radix is 32 (7680 = 16 * 15 * 32) ---------------------------------------------------
// Three passes instead of four (8 * 8 * 8 * 15): the row kernels are bound by shared-memory bandwidth, every pass is
// one read and one write of the row.  The radix-32 butterfly of the last pass takes 32 CONTIGUOUS samples per thread
// (no twiddles: 64 data registers and nothing else), which as 8-byte accesses would put all lanes on one bank; it
// reads 16-byte pairs instead, and the row is stored pair-swizzled: pair p (samples 2p, 2p+1) lives at
// p ^ ((p >> 4) & 7).  Lane t then touches pairs 16 t + (q ^ (t & 7)): the 8 lanes of a quarter-warp hit 8 different
// 16-byte banks.  The other passes (stride 480 and 32, both multiples of 32 samples) see their lanes' consecutive
// samples XOR-ed by one constant per warp: still conflict-free.  The 16-byte global <-> shared copies move whole
// pairs, so they only need the pair index swizzled.
// fused row kernel: 0 = never bulk copies, 1 = compiled in, on with LHG_ROWS_TMA=1, 2 = on unless LHG_ROWS_TMA=0
#ifndef LHG_ROWS_TMA_WL
#define LHG_ROWS_TMA_WL 2
#endif
// fused row kernel on W2 with 16-byte row pieces (2-column tiles): clusters of this many CTAs whose warps meet before
// their gathers.  2 = the two rows of a 32-byte sector (C4 fused rows 2.57-2.74 -> 2.28 ms, DRAM read 7.05 -> 5.0 GB per
// launch); 4 and 8 (rows of a 64 / 128-byte line) measured 3.24 / 3.59 ms: the wait for the slowest of 4 or 8 CTAs and
// the cluster placement cost more than the lines save.  LHG_ROWS_PAIR=0/2/4/8 overrides at run time.
#ifndef LHG_ROWS_PAIR_DEFAULT
#define LHG_ROWS_PAIR_DEFAULT 2
#endif
// warp-local row passes (A/B knobs: -DLHG_ROWS_WL=0 restores the CTA-synchronous passes everywhere,
// -DLHG_ROWS_WL_K13=0 in the separate forward / inverse row kernels only)
#ifndef LHG_ROWS_WL
#define LHG_ROWS_WL 1
#endif
#ifndef LHG_ROWS_WL_K13
#define LHG_ROWS_WL_K13 LHG_ROWS_WL
#endif

template <class P>
struct RowSwz {
  // mode 1: last radix 32 (see above).  mode 2: the 1024-point rows (16 * 16 * 4): un-swizzled, the radix-16 pass
  // with stride 4 and the radix-4 pass with stride 1 put the 16 lanes of a half-warp on 4 of the 16 8-byte banks
  // (ncu: 52 % of the fused row kernel's shared-memory wavefronts were bank conflicts, the pipe 74 % busy).  XOR-ing
  // sample-index bits 1, 2, 3 with bits 4, 6, 7 makes the first two passes conflict-free and the last one 2-way
  // (found by exhaustive search over such XOR maps; bit 0 stays so that 16-byte pairs stay together).
  static constexpr int mode = P::radix(P::NPASS - 1) == 32 ? 1 : (P::N == 1024 ? 2 : 0);
  static constexpr bool on = mode != 0;
  __device__ __forceinline__ static int el(int i) {
    if constexpr (mode == 1) return i ^ (((i >> 5) & 7) << 1);
    else if constexpr (mode == 2) return i ^ (((i >> 4) & 1) << 1) ^ (((i >> 6) & 3) << 2);
    else return i;
  }
  __device__ __forceinline__ static int pair(int e) {
    if constexpr (mode == 1) return e ^ ((e >> 4) & 7);
    else if constexpr (mode == 2) return e ^ ((e >> 3) & 1) ^ (((e >> 5) & 3) << 1);
    else return e;
  }
};

// the radix-32 pass over one row (M = 1: no twiddles; the same code serves DIF-last and DIT-first); SWAP: the
// re/im-swapped representation of the inverse transform is taken on the way in
template <class P, int NT, bool SWAP>
__device__ __forceinline__ void row_pass32(float2* buf, int tid) {
  constexpr int NB = P::N / 32;
#pragma unroll 1
  for (int b = tid; b < NB; b += NT) {
    float4* p = reinterpret_cast<float4*>(buf) + 16 * b;
    const int x = b & 7;
    float2 v[32];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float4 u = p[q ^ x];
      v[2 * q] = SWAP ? make_float2(u.y, u.x) : make_float2(u.x, u.y);
      v[2 * q + 1] = SWAP ? make_float2(u.w, u.z) : make_float2(u.z, u.w);
    }
    Dft<32>::run(v);
#pragma unroll
    for (int q = 0; q < 16; ++q) p[q ^ x] = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
  }
}

// ------------------------------------------------------------------------------------------------
// Global memory is touched only by cooperative, fully coalesced 16-byte accesses whose loads are all
// issued before anything is computed (prologue / epilogue over groups of 4 samples); the butterflies
// work in place in shared memory.
template <class P, int LOGT, int NT, int TW0>
struct RowSeq {
  static constexpr int N = P::N;
  template <int PASS, bool DIT>
  __device__ __forceinline__ static void one(float2* buf, const float2* tw, const float2* tabs, int tid) {
    using Sw = RowSwz<P>;
    auto ld = [&](int row, int t, int, int) { return buf[t * N + Sw::el(row)]; };
    auto st = [&](int row, int t, int, int, float2 v) { buf[t * N + Sw::el(row)] = v; };
    fpass<P, PASS, LOGT, NT, DIT, true, TW0 == 3 ? 2 : 1, 0, P::radix(PASS)>(tw, tabs + P::tab_off(PASS, TW0), tid, ld, st);
    __syncthreads();
  }
  // the same pass with the warp-local butterfly map (fpass WL): __syncwarp() instead of the CTA barrier
  template <int PASS, bool DIT>
  __device__ __forceinline__ static void one_wl(float2* buf, const float2* tw, const float2* tabs, int tid) {
    auto ld = [&](int row, int t, int, int) { return buf[t * N + row]; };
    auto st = [&](int row, int t, int, int, float2 v) { buf[t * N + row] = v; };
    fpass<P, PASS, LOGT, NT, DIT, true, TW0 == 3 ? 2 : 1, 0, P::radix(PASS), true>(tw, tabs + P::tab_off(PASS, TW0), tid, ld, st);
    __syncwarp();
  }
  __device__ __forceinline__ static void dif_middle_wl(float2* buf, const float2* tw, const float2* tabs, int tid) {
    if constexpr (P::NPASS >= 3) one_wl<1, false>(buf, tw, tabs, tid);
    if constexpr (P::NPASS >= 4) one_wl<2, false>(buf, tw, tabs, tid);
  }
  __device__ __forceinline__ static void dit_middle_wl(float2* buf, const float2* tw, const float2* tabs, int tid) {
    if constexpr (P::NPASS >= 4) one_wl<2, true>(buf, tw, tabs, tid);
    if constexpr (P::NPASS >= 3) one_wl<1, true>(buf, tw, tabs, tid);
  }
  __device__ __forceinline__ static void dif_middle(float2* buf, const float2* tw, const float2* tabs, int tid) {
    if constexpr (P::NPASS >= 3) one<1, false>(buf, tw, tabs, tid);
    if constexpr (P::NPASS >= 4) one<2, false>(buf, tw, tabs, tid);
  }
  __device__ __forceinline__ static void dit_middle(float2* buf, const float2* tw, const float2* tabs, int tid) {
    if constexpr (P::NPASS >= 4) one<2, true>(buf, tw, tabs, tid);
    if constexpr (P::NPASS >= 3) one<1, true>(buf, tw, tabs, tid);
  }
};

// Two row buffers per CTA for the inverse row kernel: only where MINB CTAs of them still fit one SM (227 KB, 1 KB of
// it reserved per CTA) and the plan has the warp-local passes (each warp gathers its own blocks)
#ifndef LHG_ROWS_DB
#define LHG_ROWS_DB 1
#endif
template <class P, int LOGT, int NT, int TW0, int MINB>
__host__ __device__ constexpr bool row_db_fits() {
  return LHG_ROWS_DB && LHG_ROWS_WL_K13 && LOGT == 0 && RowSwz<P>::mode == 0 && P::NPASS >= 3 && (P::R0 % (NT / 32)) == 0 &&
         (size_t)MINB * (sizeof(float2) * (2 * (size_t)(P::N << LOGT) + P::tab_total(TW0)) + 2048) <= 227 * 1024;
}

template <class P, int LOGT, int NT, int KLO, int KHI, int TW0, int MINB>
__global__ void __launch_bounds__(NT, MINB) row_fwd_fast_kernel(RowIn in, long long n_rows, float2* __restrict__ w1,
                                                          const float2* __restrict__ tw, int blocked, DeadCols dead,
                                                          int natural) {
  extern __shared__ float2 smem[];
  constexpr int N = P::N, T = 1 << LOGT, LAST = P::NPASS - 1, M0 = N / P::R0;
  constexpr int C = (KHI - KLO) * M0, PAD = KLO * M0;   // non-pad samples per row, zeros on each side
  constexpr int G = T * C / 4;                          // groups of 4 input samples per tile
  constexpr int GIT = (G + NT - 1) / NT;
  constexpr int GB = 2;                                 // groups whose loads are in flight together
  float2* const buf = smem;
  float2* const tabs = buf + (N << LOGT);
  using Sq = RowSeq<P, LOGT, NT, TW0>;
  const int tid = threadIdx.x;
  const long long n_groups = (n_rows + T - 1) >> LOGT;
  fill_tables<P, TW0>(tabs, tw, tid, NT);
  // which of this thread's 16-byte pieces of a row lie in column tiles inside the mask (the same for every row)
  static_assert((N / 2 + NT - 1) / NT <= 32, "one bit per piece");
  // WL (see row_inv_fwd_fused_kernel): the passes after the first one and the store of the row are warp-local
  constexpr int NW = NT / 32, PW = N / 2 / NW;
  constexpr bool WLC = LHG_ROWS_WL_K13 && LOGT == 0 && RowSwz<P>::mode == 0 && P::NPASS >= 3 && (P::R0 % NW) == 0;
  static_assert(!WLC || (PW + 31) / 32 <= 32, "one bit per piece");
  const bool wl = WLC && !natural;
  const int piece0 = wl ? (tid >> 5) * PW + (tid & 31) : tid, pstep = wl ? 32 : NT;
  const int piece_end = wl ? ((tid >> 5) + 1) * PW : N / 2;
  unsigned piece_live = 0xffffffffu;
  if (dead.active) {
    piece_live = 0;
    for (int i = 0, e = piece0; e < piece_end; ++i, e += pstep)
      if (dead.active[(2 * e) >> dead.logt]) piece_live |= 1u << i;
  }
  using Sw = RowSwz<P>;
  static_assert(Sw::mode != 1 || LOGT == 0, "the radix-32 row plan holds one row per CTA");
  auto ld_s = [&](int row, int t, int, int) { return buf[t * N + Sw::el(row)]; };
  auto st_s = [&](int row, int t, int, int, float2 v) { buf[t * N + Sw::el(row)] = v; };
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long row0 = grp << LOGT;
    // prologue: raw operands -> complex samples at their padded positions, GB groups of 4 at a time
    auto prologue = [&](auto kind_tag) {
      constexpr int KIND = decltype(kind_tag)::value;
#pragma unroll
      for (int i0 = 0; i0 < GIT; i0 += GB) {
        In4 raw[GB];
#pragma unroll
        for (int i = i0; i < i0 + GB && i < GIT; ++i) {
          const int e = tid + i * NT;
          const int t = e / (C / 4), c4 = e - t * (C / 4);
          if ((GIT * NT == G || e < G) && (T == 1 || row0 + t < n_rows))
            raw[i - i0] = fetch_input4<KIND>(in, (size_t)(row0 + t) * C + 4 * c4);
        }
#pragma unroll
        for (int i = i0; i < i0 + GB && i < GIT; ++i) {
          const int e = tid + i * NT;
          const int t = e / (C / 4), c4 = e - t * (C / 4);
          if (GIT * NT == G || e < G) {
            float2 x[4];
            if (T == 1 || row0 + t < n_rows) {
              make_input4<KIND>(in, raw[i - i0], x);
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) x[q] = make_float2(0.0f, 0.0f);
            }
            float4* dst = reinterpret_cast<float4*>(buf + t * N);
            constexpr int P0 = PAD / 2;  // first pair of the non-pad samples (PAD is even)
            dst[Sw::pair(P0 + 2 * c4)] = make_float4(x[0].x, x[0].y, x[1].x, x[1].y);
            dst[Sw::pair(P0 + 2 * c4 + 1)] = make_float4(x[2].x, x[2].y, x[3].x, x[3].y);
          }
        }
      }
    };
    switch (in.kind) {
      case ASM_IN_PHASE: prologue(std::integral_constant<int, ASM_IN_PHASE>{}); break;
      case ASM_IN_AMP_PHASE: prologue(std::integral_constant<int, ASM_IN_AMP_PHASE>{}); break;
      case ASM_IN_COMPLEX: prologue(std::integral_constant<int, ASM_IN_COMPLEX>{}); break;
      default: prologue(std::integral_constant<int, ASM_IN_COTANGENT>{}); break;
    }
    __syncthreads();
    fpass<P, 0, LOGT, NT, false, true, TW0 == 3 ? 2 : TW0, KLO, KHI>(tw, tabs, tid, ld_s, st_s);
    __syncthreads();
    if (WLC && wl) {
      if constexpr (WLC) {
        Sq::dif_middle_wl(buf, tw, tabs, tid);
        fpass<P, LAST, LOGT, NT, false, true, 1, 0, P::radix(LAST), true>(tw, tabs, tid, ld_s, st_s);
        __syncwarp();
      }
    } else {
      Sq::dif_middle(buf, tw, tabs, tid);
      if constexpr (Sw::mode == 1) row_pass32<P, NT, false>(buf, tid);
      else fpass<P, LAST, LOGT, NT, false, true, 1, 0, P::radix(LAST)>(tw, tabs, tid, ld_s, st_s);
      __syncthreads();
    }
    if (natural) {
      // spectrum-out calls: the columns leave in natural order (plain [row][N] layout), gathered from their
      // scrambled shared-memory slots
#pragma unroll
      for (int t = 0; t < T; ++t) {
        if (T > 1 && row0 + t >= n_rows) break;
        float4* gp = reinterpret_cast<float4*>(w1 + (size_t)(row0 + t) * N);
        const float2* sp = buf + t * N;
#pragma unroll 4
        for (int e = tid; e < N / 2; e += NT) {
          const float2 x0 = sp[Sw::el(P::iperm(2 * e))], x1 = sp[Sw::el(P::iperm(2 * e + 1))];
          gp[e] = make_float4(x0.x, x0.y, x1.x, x1.y);
        }
      }
    } else {
      // scrambled order straight out (the column kernel never needs the natural column order)
      // 2*NT columns further is a whole number of blocks further: the pointer advances by a constant
      const int gstep = woff_in_row(blocked, 2 * pstep);
#pragma unroll
      for (int t = 0; t < T; ++t) {
        if (T > 1 && row0 + t >= n_rows) break;
        float2* gp = w1 + woff(blocked, N, row0 + t, 0) + woff_in_row(blocked, 2 * piece0);
        const float4* sp = reinterpret_cast<const float4*>(buf + t * N);
#pragma unroll 5
        for (int e = piece0, i = 0; e < piece_end; e += pstep, gp += gstep, ++i) {
          if (!((piece_live >> i) & 1u)) continue;  // a column tile outside the mask: the column kernel never reads it
          *reinterpret_cast<float4*>(gp) = sp[Sw::pair(e)];
        }
      }
    }
    __syncthreads();
  }
}

template <class P, int LOGT, int NT, int KLO, int KHI, int TW0, int MINB>
__global__ void __launch_bounds__(NT, MINB) row_inv_fast_kernel(RowOut o, long long n_rows, const float2* __restrict__ w2,
                                                          const float2* __restrict__ tw, int blocked, DeadCols dead,
                                                          const __grid_constant__ CUtensorMap tmap, int use_tma,
                                                          int natural) {
  extern __shared__ __align__(128) float2 smem[];
  __shared__ float red[32];
  __shared__ __align__(8) unsigned long long tma_bar;
  __shared__ __align__(8) unsigned long long wbar[2][NT / 32];  // use_tma == 2: [row buffer][warp] "my blocks have landed"
  unsigned tma_phase = 0, wphase = 0;
  constexpr int N = P::N, T = 1 << LOGT, LAST = P::NPASS - 1, M0 = N / P::R0;
  constexpr int C = (KHI - KLO) * M0, PAD = KLO * M0;
  // DB: two row buffers when MINB CTAs of them fit (row_db_fits): the next row's copies are in flight while this row
  // is transformed.  One 3840-point row is 30 KB, three CTAs per SM kept only ~90 KB in flight during a third of the
  // time: latency-bound at a quarter of the copy peak.
  constexpr bool DB = row_db_fits<P, LOGT, NT, TW0, MINB>();
  float2* buf = smem;
  float2* const tabs = smem + (DB ? 2 : 1) * (N << LOGT);
  using Sq = RowSeq<P, LOGT, NT, TW0>;
  const int tid = threadIdx.x;
  const long long n_groups = (n_rows + T - 1) >> LOGT;
  float loss_acc = 0.0f;
  fill_tables<P, TW0>(tabs, tw, tid, NT);
  // which of this thread's 16-byte pieces of a row lie in column tiles inside the mask (the same for every row)
  static_assert((N / 2 + NT - 1) / NT <= 32, "one bit per piece");
  // WL (see row_inv_fwd_fused_kernel): the gather of the row and the passes before the last one are warp-local
  constexpr int NW = NT / 32, PW = N / 2 / NW;
  constexpr bool WLC = LHG_ROWS_WL_K13 && LOGT == 0 && RowSwz<P>::mode == 0 && P::NPASS >= 3 && (P::R0 % NW) == 0;
  static_assert(!WLC || (PW + 31) / 32 <= 32, "one bit per piece");
  // use_tma: 1 = the whole row by one thread (LHG_TMA=1, round 1), 2 = every warp its own blocks as two boxes on its
  // own mbarrier (see row_inv_fwd_fused_kernel)
  const bool wl = WLC && !natural && use_tma != 1;
  const bool wtma = wl && use_tma == 2;
  constexpr int BOXS = N / NW / 2;  // complex samples per box
  const int piece0 = wl ? (tid >> 5) * PW + (tid & 31) : tid, pstep = wl ? 32 : NT;
  const int piece_end = wl ? ((tid >> 5) + 1) * PW : N / 2;
  unsigned piece_live = 0xffffffffu;
  if (dead.active) {
    piece_live = 0;
    for (int i = 0, e = piece0; e < piece_end; ++i, e += pstep)
      if (dead.active[(2 * e) >> dead.logt]) piece_live |= 1u << i;
  }
  if (use_tma == 1 && tid == 0) mbar_init(&tma_bar, 1);
  if (use_tma == 2 && tid < NW) {
    mbar_init(&wbar[0][tid], 1);
    mbar_init(&wbar[1][tid], 1);
  }
  __syncthreads();  // the twiddle tables (and the mbarrier): the warp-local passes reach them before any other CTA barrier
  using Sw = RowSwz<P>;
  static_assert(Sw::mode != 1 || LOGT == 0, "the radix-32 row plan holds one row per CTA");
  auto ld_s = [&](int row, int t, int, int) { return buf[t * N + Sw::el(row)]; };
  auto st_s = [&](int row, int t, int, int, float2 v) { buf[t * N + Sw::el(row)] = v; };
  // warp-local gather of row `row` into dst (this warp's pieces), one cp.async group
  auto gather_wl = [&](long long row, float2* dst) {
    if (row < n_rows) {
      const int gstep = woff_in_row(blocked, 2 * 32);
      float4* sp = reinterpret_cast<float4*>(dst);
      const float2* gp = w2 + woff(blocked, N, row, 0) + woff_in_row(blocked, 2 * piece0);
#pragma unroll 5
      for (int e = piece0, i = 0; e < piece_end; e += 32, gp += gstep, ++i) {
        if (!((piece_live >> i) & 1u)) sp[e] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // never written by the column kernel
        else cp_async16(sp + e, gp);
      }
    }
    cp_async_commit();
  };
  // the same as two bulk copies per warp (lane 0), completing on wbar[which][warp]
  auto gather_tma = [&](long long row, float2* dst, int which) {
    if ((tid & 31) == 0 && row < n_rows) {
      const int w = tid >> 5;
      fence_proxy_async();  // the buffer's earlier generic-proxy accesses (ordered by the barrier that ended its last row)
      mbar_expect_tx(&wbar[which][w], 2 * BOXS * (unsigned)sizeof(float2));
#pragma unroll
      for (int h = 0; h < 2; ++h)
        tma_load_4d(dst + w * 2 * BOXS + h * BOXS, &tmap, 0, (int)(row & 7), (w * 2 * BOXS + h * BOXS) >> blocked,
                    (int)(row >> 3), &wbar[which][w]);
    }
  };
  const bool db = DB && wl;
  if (db) {
    if (wtma) gather_tma(blockIdx.x, smem, 0);
    else gather_wl(blockIdx.x, smem);
  }
  int parity = 0;
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x, parity ^= 1) {
    const long long row0 = grp << LOGT;
    if (DB && db) {
      // this row arrived (or is arriving) in buf; the other buffer's last readers passed the barrier that ended the
      // previous row: the next row starts its way into it now
      buf = smem + parity * (N << LOGT);
      if (wtma) gather_tma(grp + gridDim.x, smem + (parity ^ 1) * (N << LOGT), parity ^ 1);
      else gather_wl(grp + gridDim.x, smem + (parity ^ 1) * (N << LOGT));
    } else if (wtma) {
      gather_tma(row0, buf, 0);
    } else if (T == 1 && use_tma == 1) {
      // one elected thread gathers the row: 4-KB boxes of 16/32-byte pieces, landing densely in buf
      if (tid == 0) {
        fence_proxy_async();
        mbar_expect_tx(&tma_bar, N * (unsigned)sizeof(float2));
        constexpr int BOX_ELEMS = 512;  // complex samples per box (4 KB)
        const int pieces_per_box = BOX_ELEMS >> blocked;
#pragma unroll 1
        for (int b = 0; b < N / BOX_ELEMS; ++b)
          tma_load_4d(buf + b * BOX_ELEMS, &tmap, 0, (int)(row0 & 7), b * pieces_per_box, (int)(row0 >> 3), &tma_bar);
      }
    } else if (natural) {
      // spectrum-in calls: W2 rows hold natural-order columns (plain layout), scattered to their scrambled slots
#pragma unroll
      for (int t = 0; t < T; ++t) {
        float2* sp = buf + t * N;
        const bool live = T == 1 || row0 + t < n_rows;
        const float4* gp = reinterpret_cast<const float4*>(w2 + (size_t)(row0 + t) * N);
#pragma unroll 4
        for (int e = tid; e < N / 2; e += NT) {
          const float4 v = live ? __ldg(gp + e) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          sp[Sw::el(P::iperm(2 * e))] = make_float2(v.x, v.y);
          sp[Sw::el(P::iperm(2 * e + 1))] = make_float2(v.z, v.w);
        }
      }
    } else {
      const int gstep = woff_in_row(blocked, 2 * pstep);
#pragma unroll
      for (int t = 0; t < T; ++t) {
        float4* sp = reinterpret_cast<float4*>(buf + t * N);
        if (T > 1 && row0 + t >= n_rows) {
          for (int e = tid; e < N / 2; e += NT) sp[e] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          continue;
        }
        const float2* gp = w2 + woff(blocked, N, row0 + t, 0) + woff_in_row(blocked, 2 * piece0);
#pragma unroll 5
        for (int e = piece0, i = 0; e < piece_end; e += pstep, gp += gstep, ++i) {
#ifndef LHG_ROWS_DEAD_LOADS
          if (!((piece_live >> i) & 1u))  // never written by the column kernel: zero
#else
          if (dead.active && !dead.active[(2 * e) >> dead.logt])
#endif
            sp[Sw::pair(e)] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          else
            cp_async16(sp + Sw::pair(e), gp);
        }
      }
    }
    if (!(DB && db) && !wtma) cp_async_commit();
    // what the epilogue reads back (loss target / forward phase) starts its way into L2 now
    {
      const float* auxp = o.kind == ASM_OUT_ABS ? o.loss_target : (o.kind == ASM_OUT_GRAD_PHASE ? o.aux_phase : nullptr);
      if (auxp) {
        for (int e = tid; e < T * C / 32; e += NT) {
          const int t = e / (C / 32), c32 = e - t * (C / 32);
          if (T == 1 || row0 + t < n_rows) prefetch_l2(auxp + (size_t)(row0 + t) * C + 32 * c32);
        }
      }
    }
    if (wtma) {
      const int which = (DB && db) ? parity : 0;
      mbar_wait_bounded(&wbar[which][tid >> 5], (wphase >> which) & 1u);
      wphase ^= 1u << which;
      if (piece_live != 0xffffffffu) {  // column tiles outside the mask were never written by the column kernel: zero
        float4* sp = reinterpret_cast<float4*>(buf);
        for (int e = piece0, i = 0; e < piece_end; e += 32, ++i)
          if (!((piece_live >> i) & 1u)) sp[e] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
      __syncwarp();
    } else if (T == 1 && use_tma == 1) {
      mbar_wait(&tma_bar, tma_phase);
      tma_phase ^= 1u;
    } else if (DB && db) {
      cp_async_wait_group1();  // everything but the newest group (the next row)
      __syncwarp();
    } else {
      cp_async_wait_all();
      if (WLC && wl) __syncwarp();
      else __syncthreads();
    }
    auto ld_first = [&](int row, int t, int, int) { return cswap(buf[t * N + Sw::el(row)]); };
    if (WLC && wl) {
      if constexpr (WLC) {
        fpass<P, LAST, LOGT, NT, true, true, 1, 0, P::radix(LAST), true>(tw, tabs, tid, ld_first, st_s);
        __syncwarp();
        Sq::dit_middle_wl(buf, tw, tabs, tid);
        __syncthreads();
      }
    } else {
      if constexpr (Sw::mode == 1) row_pass32<P, NT, true>(buf, tid);
      else fpass<P, LAST, LOGT, NT, true, true, 1, 0, P::radix(LAST)>(tw, tabs, tid, ld_first, st_s);
      __syncthreads();
      Sq::dit_middle(buf, tw, tabs, tid);
    }
    // only the crop survives the last butterfly; the epilogue runs on its outputs in registers (lane j holds
    // sample j + (k - KLO) * M0 of the row: coalesced 4/8-byte stores, operands read back were prefetched to L2)
    auto last = [&](auto kind_tag) {
      constexpr int KIND = decltype(kind_tag)::value;
      auto st_out = [&](int row, int t, int, int, float2 v) {
        if (T == 1 || row0 + t < n_rows)
          store_output1<KIND>(o, (size_t)(row0 + t) * C + (row - PAD), make_float2(v.y, v.x), loss_acc);
      };
      fpass<P, 0, LOGT, NT, true, true, TW0 == 3 ? 2 : TW0, KLO, KHI>(tw, tabs, tid, ld_s, st_out);
    };
    switch (o.kind) {
      case ASM_OUT_ABS: last(std::integral_constant<int, ASM_OUT_ABS>{}); break;
      case ASM_OUT_ANGLE: last(std::integral_constant<int, ASM_OUT_ANGLE>{}); break;
      case ASM_OUT_ABS_ANGLE: last(std::integral_constant<int, ASM_OUT_ABS_ANGLE>{}); break;
      case ASM_OUT_COMPLEX: last(std::integral_constant<int, ASM_OUT_COMPLEX>{}); break;
      case ASM_OUT_ABS2: last(std::integral_constant<int, ASM_OUT_ABS2>{}); break;
      default: last(std::integral_constant<int, ASM_OUT_GRAD_PHASE>{}); break;
    }
    __syncthreads();
  }
  if (o.loss_partial) block_loss_reduce(loss_acc, o.loss_partial, red);
}

// ------------------------------------------------------------------------------------------------
// fused row kernel of the forward + amplitude-L2 + adjoint step: K3 of the forward call and K1 of the adjoint
// call on the same row without leaving shared memory
// ------------------------------------------------------------------------------------------------
// row of W2 -> inverse passes -> crop -> y = scale * field -> loss += (|y| - target)^2, optional |y| out ->
// cotangent cot_scale * (|y| - target) * y / |y| written over the crop -> forward passes (zero pad pruned) ->
// row of the adjoint's W1.  Against the two separate kernels this saves, per output sample, the 12 bytes of
// |y| and the saved field written, the 12 bytes read back and the second read of the target.
template <class P, int LOGT, int NT, int KLO, int KHI, int TW0, int MINB>
__global__ void __launch_bounds__(NT, MINB) row_inv_fwd_fused_kernel(FusedRows f, long long n_rows, const float2* __restrict__ w2,
                                                               float2* __restrict__ w1, const float2* __restrict__ tw,
                                                               int blocked_in, int blocked_out, DeadCols dead,
                                                               const __grid_constant__ CUtensorMap tm_in,
                                                               const __grid_constant__ CUtensorMap tm_out, int use_tma,
                                                               int pair) {
  extern __shared__ __align__(128) float2 smem[];
  __shared__ float red[32];
  __shared__ __align__(8) unsigned long long wbar[NT / 32];  // use_tma: "this warp's blocks of the row have landed"
  constexpr int N = P::N, T = 1 << LOGT, LAST = P::NPASS - 1, M0 = N / P::R0;
  constexpr int C = (KHI - KLO) * M0;
  float2* const buf = smem;
  float2* const tabs = buf + (N << LOGT);
  using Sq = RowSeq<P, LOGT, NT, TW0>;
  const int tid = threadIdx.x;
  const long long n_groups = (n_rows + T - 1) >> LOGT;
  float loss_acc = 0.0f;
  RowIn cot{};  // the cotangent arithmetic of the adjoint's prologue (common.cuh cot_value), target term only
  cot.cot_target = f.target ? f.target : reinterpret_cast<const float*>(f.target_u8);  // non-null: the term is on
  cot.cot_scale = f.cot_scale;
  fill_tables<P, TW0>(tabs, tw, tid, NT);
#ifndef LHG_ROWS_DEAD_LOADS
  // which of this thread's 16-byte pieces of a row lie in column tiles inside the mask: the same for every row, so
  // the per-piece lookups of dead.active (one global load in front of every cp.async and every store) happen once
  static_assert((N / 2 + NT - 1) / NT <= 32, "one bit per piece");
  // WL: after the first pass of the plan the row is R0 independent blocks; warp w gathers, transforms (all passes but
  // the turn) and stores blocks [w*R0/NW, (w+1)*R0/NW) on its own, so a row costs two CTA barriers (either side of the
  // turn) instead of nine and the warps of a CTA are in different phases (copy, butterflies) most of the time.
  constexpr int NW = NT / 32, PW = N / 2 / NW;  // 16-byte pieces per warp
  constexpr bool WL = LHG_ROWS_WL && LOGT == 0 && RowSwz<P>::mode == 0 && P::NPASS >= 3 && (P::R0 % NW) == 0 && (PW % 2) == 0;
  static_assert(!WL || (PW + 31) / 32 <= 32, "one bit per piece");
  const int piece0 = WL ? (tid >> 5) * PW + (tid & 31) : tid;  // this thread's first piece; the next is PSTEP further
  constexpr int PSTEP = WL ? 32 : NT;
  const int piece_end = WL ? ((tid >> 5) + 1) * PW : N / 2;
  unsigned live = 0xffffffffu;
  if (dead.active) {
    live = 0;
    for (int i = 0, e = piece0; e < piece_end; ++i, e += PSTEP)
      if (dead.active[(2 * e) >> dead.logt]) live |= 1u << i;
  }
#define LHG_PIECE_DEAD(i, e) (!((live >> (i)) & 1u))
#else
#define LHG_PIECE_DEAD(i, e) (dead.active && !dead.active[(2 * (e)) >> dead.logt])
#endif
  // use_tma (warp-local plans): a warp's blocks travel as TMA boxes of half a warp's share of the row -- W2 -> shared
  // memory on the warp's own mbarrier, shared memory -> W1' as a bulk store -- issued by lane 0: 4 bulk copies per warp
  // and row instead of 15 cp.async + 15 LDS/STG per thread in the LSU queue (`mio_throttle` is this kernel's top stall)
  constexpr int BOXS = N / NW / 2;  // complex samples per box
  unsigned wphase = 0;
  if (WL && use_tma && tid < NW) mbar_init(&wbar[tid], 1);
  // pair (launched as clusters of two CTAs): rows 2k and 2k+1 share every 32-byte sector of W2 (16 bytes each), and the
  // two CTAs that gather them drift apart by more than the sectors live in L2 (ncu: 7 GB read for 4 GB).  Warp w of one
  // CTA and warp w of the other -- same pieces, neighbouring rows -- therefore meet before they send their copies: an
  // mbarrier in each CTA with two arrivals, the partner's through distributed shared memory.
  __shared__ __align__(8) unsigned long long pbar[NT / 32];
  unsigned pphase = 0;
  if (WL && pair && tid < NW) mbar_init(&pbar[tid], pair);  // pair = CTAs per cluster (2, 4 or 8 rows that share lines)
  __syncthreads();  // the twiddle tables: the warp-local passes reach them before any other CTA barrier
  if (WL && pair) cluster_sync_all();  // the partner's barriers exist before anything arrives on them
  using Sw = RowSwz<P>;
  static_assert(Sw::mode != 1 || LOGT == 0, "the radix-32 row plan holds one row per CTA");
  auto ld_s = [&](int row, int t, int, int) { return buf[t * N + Sw::el(row)]; };
  auto st_s = [&](int row, int t, int, int, float2 v) { buf[t * N + Sw::el(row)] = v; };
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long row0 = grp << LOGT;
    if (WL && use_tma) {
      if ((tid & 31) == 0) {
        const int w = tid >> 5;
        if (pair && (row0 | (long long)(pair - 1)) < n_rows) {  // the cluster holds rows [row0 & ~(pair-1), +pair) now
          const unsigned me = cluster_rank();
          for (unsigned r = 0; r < (unsigned)pair; ++r) {
            if (r == me) mbar_arrive(&pbar[w]);
            else mbar_arrive_peer(&pbar[w], r);
          }
          mbar_wait_bounded(&pbar[w], pphase);  // (traps instead of hanging if a partner never comes)
          pphase ^= 1u;
        }
        tma_store_wait_read();  // the previous row's bulk store has read this warp's blocks
        fence_proxy_async();
        mbar_expect_tx(&wbar[w], 2 * BOXS * (unsigned)sizeof(float2));
#pragma unroll
        for (int h = 0; h < 2; ++h)
          tma_load_4d(buf + w * 2 * BOXS + h * BOXS, &tm_in, 0, (int)(row0 & 7), (w * 2 * BOXS + h * BOXS) >> blocked_in,
                      (int)(row0 >> 3), &wbar[w]);
      }
    } else {
      const int gstep = woff_in_row(blocked_in, 2 * PSTEP);
#pragma unroll
      for (int t = 0; t < T; ++t) {
        float4* sp = reinterpret_cast<float4*>(buf + t * N);
        if (T > 1 && row0 + t >= n_rows) {
          for (int e = tid; e < N / 2; e += NT) sp[e] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          continue;
        }
        const float2* gp = w2 + woff(blocked_in, N, row0 + t, 0) + woff_in_row(blocked_in, 2 * piece0);
#pragma unroll 5
        for (int e = piece0, i = 0; e < piece_end; e += PSTEP, gp += gstep, ++i) {
          if (LHG_PIECE_DEAD(i, e))
            sp[Sw::pair(e)] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          else
            cp_async16(sp + Sw::pair(e), gp);
        }
      }
    }
    cp_async_commit();
    for (int e = tid; e < T * C / 32; e += NT) {
      const int t = e / (C / 32), c32 = e - t * (C / 32);
      if (T == 1 || row0 + t < n_rows) {
        if (f.target) prefetch_l2(f.target + (size_t)(row0 + t) * C + 32 * c32);
        else if ((c32 & 3) == 0) prefetch_l2(f.target_u8 + (size_t)(row0 + t) * C + 32 * c32);
      }
    }
    cp_async_wait_all();
    auto ld_first = [&](int row, int t, int, int) { return cswap(buf[t * N + Sw::el(row)]); };
    if constexpr (WL) {
      if (use_tma) {
        mbar_wait_bounded(&wbar[tid >> 5], wphase);  // (traps instead of hanging if a copy never completes)
        wphase ^= 1u;
        if (live != 0xffffffffu) {  // column tiles outside the mask were never written by the column kernel: zero
          float4* sp = reinterpret_cast<float4*>(buf);
          for (int e = piece0, i = 0; e < piece_end; e += PSTEP, ++i)
            if (LHG_PIECE_DEAD(i, e)) sp[e] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
      }
      __syncwarp();
      fpass<P, LAST, LOGT, NT, true, true, 1, 0, P::radix(LAST), true>(tw, tabs, tid, ld_first, st_s);
      __syncwarp();
      Sq::dit_middle_wl(buf, tw, tabs, tid);
      __syncthreads();
    } else {
      __syncthreads();
      if constexpr (Sw::mode == 1) row_pass32<P, NT, true>(buf, tid);
      else fpass<P, LAST, LOGT, NT, true, true, 1, 0, P::radix(LAST)>(tw, tabs, tid, ld_first, st_s);
      __syncthreads();
      Sq::dit_middle(buf, tw, tabs, tid);
    }
    // last inverse pass, loss term + cotangent, first forward pass: one butterfly, in registers.  Output k of
    // butterfly j is crop sample j + (k - KLO) * M0 of the row; the targets are fetched before the butterfly.
    struct Tgt { float v[KHI - KLO]; };
    fturn<P, LOGT, NT, TW0 == 3 ? 2 : TW0, KLO, KHI>(
        tw, tabs, tid,
        [&](int j, int t, int) {
          Tgt a;
          const bool live = T == 1 || row0 + t < n_rows;
          if (f.target) {
            const float* tp = f.target + (size_t)(row0 + t) * C + j;
#pragma unroll
            for (int k = 0; k < KHI - KLO; ++k) a.v[k] = live ? __ldg(tp + k * M0) : 0.0f;
          } else {  // 8-bit targets: fl(v / 255), the IEEE quotient torch's .div(255) gives
            const unsigned char* tp = f.target_u8 + (size_t)(row0 + t) * C + j;
#pragma unroll
            for (int k = 0; k < KHI - KLO; ++k) a.v[k] = live ? __fdiv_rn((float)__ldg(tp + k * M0), 255.0f) : 0.0f;
          }
          return a;
        },
        ld_s,
        [&](int j, int t, int, float2 (&v)[P::R0], const Tgt& a) {
          const bool live = T == 1 || row0 + t < n_rows;
#pragma unroll
          for (int k = KLO; k < KHI; ++k) {
            float2 y = make_float2(v[k].y * f.scale, v[k].x * f.scale);
            const float r = cabs_fast(y);
            const float d = r - a.v[k - KLO];
            if (live) {
              loss_acc += d * d;
              if (f.amp_out) f.amp_out[(size_t)(row0 + t) * C + j + (k - KLO) * M0] = r;
            }
            v[k] = cot_value(cot, y, 0.0f, a.v[k - KLO], 0.0f, 0.0f);
          }
        },
        st_s);
    __syncthreads();
    if constexpr (WL) {
      Sq::dif_middle_wl(buf, tw, tabs, tid);
      fpass<P, LAST, LOGT, NT, false, true, 1, 0, P::radix(LAST), true>(tw, tabs, tid, ld_s, st_s);
      __syncwarp();
    } else {
      Sq::dif_middle(buf, tw, tabs, tid);
      if constexpr (Sw::mode == 1) row_pass32<P, NT, false>(buf, tid);
      else fpass<P, LAST, LOGT, NT, false, true, 1, 0, P::radix(LAST)>(tw, tabs, tid, ld_s, st_s);
      __syncthreads();
    }
    if (WL && use_tma) {
      fence_proxy_async();  // every lane's writes of the last pass, before the bulk store reads them
      __syncwarp();
      if ((tid & 31) == 0) {
        const int w = tid >> 5;
#pragma unroll
        for (int h = 0; h < 2; ++h)
          tma_store_4d(&tm_out, 0, (int)(row0 & 7), (w * 2 * BOXS + h * BOXS) >> blocked_out, (int)(row0 >> 3),
                       buf + w * 2 * BOXS + h * BOXS);
        tma_store_commit();
      }
    } else {
      const int gstep = woff_in_row(blocked_out, 2 * PSTEP);
#pragma unroll
      for (int t = 0; t < T; ++t) {
        if (T > 1 && row0 + t >= n_rows) break;
        float2* gp = w1 + woff(blocked_out, N, row0 + t, 0) + woff_in_row(blocked_out, 2 * piece0);
        const float4* sp = reinterpret_cast<const float4*>(buf + t * N);
#pragma unroll 5
        for (int e = piece0, i = 0; e < piece_end; e += PSTEP, gp += gstep, ++i) {
          if (LHG_PIECE_DEAD(i, e)) continue;
          *reinterpret_cast<float4*>(gp) = sp[Sw::pair(e)];
        }
      }
    }
    // the next row's copies land in the pieces this warp (WL) / this CTA has just read
    if constexpr (WL) __syncwarp();
    else __syncthreads();
  }
  if (WL && use_tma && (tid & 31) == 0) tma_store_wait_all();  // the last rows' bulk stores, before the CTA retires
  if (WL && pair) cluster_sync_all();  // nobody retires while its partner may still arrive on its barriers
  if (f.loss_partial) block_loss_reduce(loss_acc, f.loss_partial, red);
#undef LHG_PIECE_DEAD
}

// ------------------------------------------------------------------------------------------------
// plans and dispatch
// ------------------------------------------------------------------------------------------------
// A plan applies to a geometry when the un-padded extent is (KHI-KLO)*N/R0 and the pad is KLO*N/R0.
// TW0 = where the twiddles come from (FastPlan::tab_len: 1 = full shared-memory tables, 3 = product trees
// from shared-memory tables of first powers).
// MINB = CTAs per SM the register allocation is held to.
//             N     R0  R1  R2  R3  LOGT  NT  KLO KHI TW0 MINB
// 7680 = 8*8*8*15 (four passes, 80 registers, 3 CTAs per SM) is the default.  The three-pass plan 16*15*32 with the
// pair-swizzled row and the 128-bit radix-32 pass (RowSwz, row_pass32) was measured against it on C4
// (profiles/r02_summary.md): fused row kernel 3.39 ms at 3 CTAs per SM (400 bytes of spills), 3.13 ms at 2 CTAs per SM,
// against 3.14 ms for the four passes -- a quarter fewer shared-memory wavefronts bought nothing, i.e. the kernel is
// not bound by shared-memory bandwidth.  Kept behind LHG_ROWS_7680_3PASS (parity-tested) for that record.
#if defined(LHG_ROWS_7680_3PASS)
#define ROW_PLAN_7680(X) X(7680, 16, 15, 32, 1, 0, 256, 4, 12, 3, 3)
#elif defined(LHG_ROWS_7680_3PASS_MINB2)
#define ROW_PLAN_7680(X) X(7680, 16, 15, 32, 1, 0, 256, 4, 12, 3, 2)
#elif defined(LHG_ROWS_7680_3PASS_240)
#define ROW_PLAN_7680(X) X(7680, 16, 15, 32, 1, 0, 240, 4, 12, 3, 3)
#elif defined(LHG_ROWS_7680_NT320)
#define ROW_PLAN_7680(X) X(7680, 8, 8, 8, 15, 0, 320, 2, 6, 3, 3)
#elif defined(LHG_ROWS_7680_NT192)
#define ROW_PLAN_7680(X) X(7680, 8, 8, 8, 15, 0, 192, 2, 6, 3, 3)
#else
#define ROW_PLAN_7680(X) X(7680, 8, 8, 8, 15, 0, 256, 2, 6, 3, 3)
#endif
#define FAST_ROW_PLANS(X)                     \
  ROW_PLAN_7680(X)                            \
  X(3840, 16, 16, 15, 1, 0, 256, 4, 12, 3, 3) \
  X(3840, 16, 16, 15, 1, 0, 256, 0, 16, 3, 3) \
  X(1920, 8, 16, 15, 1, 1, 256, 0, 8, 3, 3)   \
  X(1024, 16, 16, 4, 1, 2, 256, 5, 11, 1, 3)  \
  X(384, 8, 16, 3, 1, 3, 256, 0, 8, 1, 3)

//             N     R0  R1  R2  R3  LOGT  NT  KLO KHI
#define FAST_COL_PLANS(X)               \
  X(4320, 16, 18, 15, 1, 1, 288, 4, 12) \
  X(2160, 16, 9, 15, 1, 2, 288, 4, 12)  \
  X(2160, 16, 9, 15, 1, 2, 288, 0, 16)  \
  X(1080, 8, 9, 15, 1, 3, 288, 0, 8)    \
  X(1024, 16, 16, 4, 1, 3, 256, 5, 11)  \
  X(384, 8, 16, 3, 1, 4, 192, 0, 8)

#define PLAN_MATCH(N, R0, KLO, KHI, n, ext, pad) \
  ((n) == N && (ext) == (KHI - KLO) * (N / R0) && (pad) == KLO * (N / R0))

// 2x padded 4320-point columns: warp-local kernel (col_warp.cuh), plan 18 x 16 x 15, two columns per tile.
// LHG_COL_WARP=0 selects the CTA-synchronous kernel above instead (decided once, at plan creation: the
// two kernels scramble the rows differently).
using WarpPlan4320 = FastPlan<4320, 18, 16, 15>;
using WarpPlan2160 = FastPlan<2160, 18, 8, 15>;  // 1080 rows padded by 540: the same kernel with 4-column tiles
static bool warp_cols_enabled() {
  static const bool on = [] {
    const char* e = getenv("LHG_COL_WARP");
    return !(e && e[0] == '0');
  }();
  return on;
}
static bool warp_cols_match(int n, int rows, int pad) {
  return warp_cols_enabled() && ((n == 4320 && rows == 2160 && pad == 1080) || (n == 2160 && rows == 1080 && pad == 540));
}
// 384 rows padded by 320 (BASELINE configs 2/3): col_warp16.cuh, plan 16 x 8 x 8, four columns per tile
using WarpPlan1024 = FastPlan<1024, 16, 8, 8>;
static bool warp16_cols_match(int n, int rows, int pad) { return warp_cols_enabled() && n == 1024 && rows == 384 && pad == 320; }

bool fast_rows_supported(int n, int cols, int pad) {
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI, TW0, MINB) \
  if (PLAN_MATCH(N, R0, KLO, KHI, n, cols, pad)) return true;
  FAST_ROW_PLANS(X)
#undef X
  return false;
}

// the same for the CTA-synchronous kernel alone (the one spectrum-in / spectrum-out calls run on)
int fast_cols_sync_logt(int n, int rows, int pad) {
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI) \
  if (PLAN_MATCH(N, R0, KLO, KHI, n, rows, pad)) return LT;
  FAST_COL_PLANS(X)
#undef X
  return -1;
}

// log2 of the columns per tile of the fast column kernel, or -1
int fast_cols_logt(int n, int rows, int pad) {
  if (warp_cols_match(n, rows, pad)) return n == 4320 ? 1 : 2;
  if (warp16_cols_match(n, rows, pad)) return 2;
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI) \
  if (PLAN_MATCH(N, R0, KLO, KHI, n, rows, pad)) return LT;
  FAST_COL_PLANS(X)
#undef X
  return -1;
}

void fast_rows_perm(int n, int* perm_out) {
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI, TW0, MINB)                           \
  if (n == N) {                                                                     \
    for (int p = 0; p < N; ++p) perm_out[p] = FastPlan<N, R0, R1, R2, R3>::perm(p); \
    return;                                                                         \
  }
  FAST_ROW_PLANS(X)
#undef X
}

void fast_cols_perm(int n, int rows, int pad, int* perm_out) {
  if (warp_cols_match(n, rows, pad)) {
    for (int p = 0; p < n; ++p) perm_out[p] = n == 4320 ? WarpPlan4320::perm(p) : WarpPlan2160::perm(p);
    return;
  }
  if (warp16_cols_match(n, rows, pad)) {
    for (int p = 0; p < n; ++p) perm_out[p] = WarpPlan1024::perm(p);
    return;
  }
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI)                                      \
  if (n == N) {                                                                     \
    for (int p = 0; p < N; ++p) perm_out[p] = FastPlan<N, R0, R1, R2, R3>::perm(p); \
    return;                                                                         \
  }
  FAST_COL_PLANS(X)
#undef X
}

// Opt a kernel in to the device's maximum dynamic shared memory ONCE per (kernel, device): asm_propagate is
// re-entrant and autograd runs backward on another thread, so a per-call set-then-launch with per-plan sizes could
// interleave between two plans.
static int optin_smem_once(const void* kernel) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({kernel, dev})) return 0;
  int max_optin = 0;
  e = cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return (int)e;
  cudaFuncAttributes fa{};
  e = cudaFuncGetAttributes(&fa, kernel);  // the opt-in limit covers static + dynamic shared memory
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - (int)fa.sharedSizeBytes);
  if (e != cudaSuccess) return (int)e;
  done.insert({kernel, dev});
  return 0;
}

template <class K>
static int grid_for(K kernel, int threads, size_t smem, int sm_count, long long work, int* grid_out) {
  int rc = optin_smem_once(reinterpret_cast<const void*>(kernel));
  if (rc) return rc;
  int occ = 1;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
  if (e != cudaSuccess) return (int)e;
  if (occ < 1) return -1;  // does not fit (e.g. an n_depth so large that the per-depth table overflows shared memory)
  long long grid = (long long)sm_count * occ;
  if (grid > work) grid = work;
  if (grid < 1) grid = 1;
  *grid_out = (int)grid;
  return 0;
}

// ---- tensor maps over the blocked W layouts (row kernels, one row per CTA) ---------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static const EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return (EncodeTiledFn) nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
// LHG_TMA=1: the inverse row kernel gathers its row with the TMA unit instead of LDGSTS.  Off by default:
// measured equal (2.32 vs 2.28 ms per C4 step) -- the gather is bound by the 16-byte granules in L2/DRAM, not by
// the L1 tag stage, and the TMA path has to read the column tiles outside the mask as well.
bool fast_rows_tma_enabled() {
  static const bool on = [] {
    const char* e = getenv("LHG_TMA");
    return e && e[0] == '1' && encode_tiled() != nullptr;
  }();
  return on;
}
// W [rows/8][Cp >> b][8][2^b] complex64 as a 4-D fp32 tensor (complex-in-piece, row-in-block, piece, row-block);
// one box = 512 complex samples (4 KB) of one row.  Returns false when the geometry does not fit.
static bool make_row_tmap(CUtensorMap* map, const float2* w, long long n_rows, int Cp, int b) {
  if (!fast_rows_tma_enabled() || b < 1 || b > 2 || (n_rows & 7) || (Cp & 511)) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)(2u << b), 8, (cuuint64_t)(Cp >> b), (cuuint64_t)(n_rows >> 3)};
  const cuuint64_t strides[3] = {(cuuint64_t)(8u << b), (cuuint64_t)(64u << b), (cuuint64_t)(64u << b) * (cuuint64_t)(Cp >> b)};
  const cuuint32_t box[4] = {(cuuint32_t)(2u << b), 1, (cuuint32_t)(512 >> b), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return encode_tiled()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)w, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// The same tensor with boxes of `box_samples` complex samples of one row (the fused row kernel's per-warp copies)
// promote: L2 promotion of the loads.  NONE for the gathers: a row takes 16 or 32 bytes of every 128/256-byte block of
// 8 rows, and with 128-byte promotion every piece pulled its whole block from DRAM (ncu: 7.67 GB read for 3.98 GB).
static bool make_row_tmap_box(CUtensorMap* map, const float2* w, long long n_rows, int Cp, int b, int box_samples,
                              bool promote = false) {
  if (!encode_tiled() || b < 1 || b > 2 || (n_rows & 7) || (box_samples >> b) > 256 || (box_samples & ((1 << b) - 1))) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)(2u << b), 8, (cuuint64_t)(Cp >> b), (cuuint64_t)(n_rows >> 3)};
  const cuuint64_t strides[3] = {(cuuint64_t)(8u << b), (cuuint64_t)(64u << b), (cuuint64_t)(64u << b) * (cuuint64_t)(Cp >> b)};
  const cuuint32_t box[4] = {(cuuint32_t)(2u << b), 1, (cuuint32_t)(box_samples >> b), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  static const bool force_promote = [] { const char* e = getenv("LHG_ROWS_TMA_PROMOTE"); return e && e[0] == '1'; }();
  return encode_tiled()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)w, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        (promote || force_promote) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// LHG_ROWS_TMA=0/1 at run time (A/B of the fused row kernel's bulk copies without a rebuild)
static bool rows_tma_wl_enabled() {
  static const bool on = [] {
    const char* e = getenv("LHG_ROWS_TMA");
    return e ? e[0] == '1' : (LHG_ROWS_TMA_WL == 2);
  }();
  return on;
}
// LHG_ROWS_PAIR=0/2/4/8 at run time: the fused row kernel as clusters of that many CTAs (rows that share sectors /
// lines of W2) whose warps meet before their gathers
static int rows_pair_size() {
  static const int cs = [] {
    const char* e = getenv("LHG_ROWS_PAIR");
    int v = e ? atoi(e) : LHG_ROWS_PAIR_DEFAULT;
    if (v == 1) v = 2;
    return (v == 2 || v == 4 || v == 8) ? v : 0;
  }();
  return cs;
}
// does the inverse row kernel gather W2 with the TMA unit (then it reads every column, also those the column
// kernel would otherwise leave unwritten)?
bool fast_row_inverse_uses_tma(int n, int C, int pad_c, long long n_rows, int blocked) {
  if (!fast_rows_tma_enabled() || blocked < 1 || blocked > 2 || (n_rows & 7) || (n & 511)) return false;
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI, TW0, MINB) \
  if (PLAN_MATCH(N, R0, KLO, KHI, n, C, pad_c)) return LT == 0 && (R3 > 1 ? R3 : R2) != 32;  /* swizzled rows: no TMA */
  FAST_ROW_PLANS(X)
#undef X
  return false;
}

int fast_row_forward(int n, const float2* tw, const RowIn& in, long long n_rows, int C, int pad_c, float2* w1,
                     int blocked, DeadCols dead, int natural, int sm_count, cudaStream_t stream) {
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI, TW0, MINB)                                   \
  if (PLAN_MATCH(N, R0, KLO, KHI, n, C, pad_c)) {                                           \
    using Pl = FastPlan<N, R0, R1, R2, R3>;                                                 \
    auto k = row_fwd_fast_kernel<Pl, LT, NT, KLO, KHI, TW0, MINB>;                                \
    const size_t smem = sizeof(float2) * ((size_t)(N << LT) + Pl::tab_total(TW0));          \
    int grid = 1;                                                                           \
    int rc = grid_for(k, NT, smem, sm_count, (n_rows + (1 << LT) - 1) >> LT, &grid);        \
    if (rc) return rc;                                                                      \
    k<<<grid, NT, smem, stream>>>(in, n_rows, w1, tw, blocked, dead, natural);                           \
    return (int)cudaPeekAtLastError();                                                      \
  }
  FAST_ROW_PLANS(X)
#undef X
  return -1;
}

int fast_row_inverse(int n, const float2* tw, const RowOut& out, long long n_rows, int C, int pad_c,
                     const float2* w2, int blocked, DeadCols dead, int natural, int sm_count, int max_blocks,
                     cudaStream_t stream) {
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI, TW0, MINB)                                   \
  if (PLAN_MATCH(N, R0, KLO, KHI, n, C, pad_c)) {                                           \
    using Pl = FastPlan<N, R0, R1, R2, R3>;                                                 \
    auto k = row_inv_fast_kernel<Pl, LT, NT, KLO, KHI, TW0, MINB>;                                \
    const size_t smem = sizeof(float2) * ((row_db_fits<Pl, LT, NT, TW0, MINB>() ? 2 : 1) * (size_t)(N << LT) + Pl::tab_total(TW0)); \
    int grid = 1;                                                                           \
    int rc = grid_for(k, NT, smem, sm_count, (n_rows + (1 << LT) - 1) >> LT, &grid);        \
    if (rc) return rc;                                                                      \
    if (max_blocks > 0 && grid > max_blocks) grid = max_blocks;                             \
    CUtensorMap tmap{};                                                                     \
    int use_tma = (LT == 0 && (R3 > 1 ? R3 : R2) != 32 && !natural &&                       \
                   make_row_tmap(&tmap, w2, n_rows, N, blocked)) ? 1 : 0;                   \
    constexpr int kNW = NT / 32;                                                            \
    constexpr bool kWL = LHG_ROWS_WL_K13 && LT == 0 && (R3 > 1 ? R3 : R2) != 32 && N != 1024 && (R0 % kNW) == 0 && (R2 > 1); \
    if (!use_tma && LHG_ROWS_TMA_WL && kWL && !natural && rows_tma_wl_enabled() &&          \
        make_row_tmap_box(&tmap, w2, n_rows, N, blocked, N / kNW / 2)) use_tma = 2;         \
    k<<<grid, NT, smem, stream>>>(out, n_rows, w2, tw, blocked, use_tma == 1 ? DeadCols{nullptr, 0} : dead, tmap, use_tma, natural); \
    return (int)cudaPeekAtLastError();                                                      \
  }
  FAST_ROW_PLANS(X)
#undef X
  return -1;
}

int fast_row_inverse_forward(int n, const float2* tw, const FusedRows& f, long long n_rows, int C, int pad_c,
                             const float2* w2, int blocked_in, float2* w1, int blocked_out, DeadCols dead,
                             int sm_count, int max_blocks, cudaStream_t stream) {
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI, TW0, MINB)                                   \
  if (PLAN_MATCH(N, R0, KLO, KHI, n, C, pad_c)) {                                           \
    using Pl = FastPlan<N, R0, R1, R2, R3>;                                                 \
    auto k = row_inv_fwd_fused_kernel<Pl, LT, NT, KLO, KHI, TW0, MINB>;                     \
    const size_t smem = sizeof(float2) * ((size_t)(N << LT) + Pl::tab_total(TW0));          \
    int grid = 1;                                                                           \
    int rc = grid_for(k, NT, smem, sm_count, (n_rows + (1 << LT) - 1) >> LT, &grid);        \
    if (rc) return rc;                                                                      \
    if (max_blocks > 0 && grid > max_blocks) grid = max_blocks;                             \
    CUtensorMap tm_in{}, tm_out{};                                                          \
    constexpr int kNW = NT / 32;                                                            \
    constexpr bool kWL = LHG_ROWS_WL && LT == 0 && (R3 > 1 ? R3 : R2) != 32 && N != 1024 && (R0 % kNW) == 0 && (R2 > 1); \
    const int use_tma = (LHG_ROWS_TMA_WL && kWL && rows_tma_wl_enabled() &&                 \
                         make_row_tmap_box(&tm_in, w2, n_rows, N, blocked_in, N / kNW / 2) &&  \
                         make_row_tmap_box(&tm_out, w1, n_rows, N, blocked_out, N / kNW / 2)) ? 1 : 0; \
    const int cs = (use_tma && blocked_in == 1) ? rows_pair_size() : 0;                     \
    if (cs >= 2 && grid >= cs) {                                                            \
      grid -= grid % (cs > 0 ? cs : 1);                                                     \
      cudaLaunchConfig_t cfg{};                                                             \
      cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = stream; \
      cudaLaunchAttribute at[1];                                                            \
      at[0].id = cudaLaunchAttributeClusterDimension;                                       \
      at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; \
      cfg.attrs = at; cfg.numAttrs = 1;                                                     \
      return (int)cudaLaunchKernelEx(&cfg, k, f, n_rows, w2, w1, tw, blocked_in, blocked_out, dead, tm_in, tm_out, use_tma, cs); \
    }                                                                                       \
    k<<<grid, NT, smem, stream>>>(f, n_rows, w2, w1, tw, blocked_in, blocked_out, dead, tm_in, tm_out, use_tma, 0); \
    return (int)cudaPeekAtLastError();                                                      \
  }
  FAST_ROW_PLANS(X)
#undef X
  return -1;
}

// LHG_COL_TMA=0 sends the adjoint column launch back to cp.async staging (A/B; default on)
static bool cols_tma_enabled() {
  static const bool on = [] {
    const char* e = getenv("LHG_COL_TMA");
    return !(e && e[0] == '0') && encode_tiled() != nullptr;
  }();
  return on;
}
// W [rows/8][Cp >> b][8][2^b] complex64 as a 4-D fp32 tensor (float-in-piece, row-in-block, piece, row-block); one box =
// the tile's `tile_floats / 2` columns x 8 rows x `blocks` row blocks of one piece column.
static bool make_col_tmap(CUtensorMap* map, const float2* w, long long n_rows, int Cp, int b, int tile_floats, int blocks) {
  if (!cols_tma_enabled() || b < 1 || b > 2 || (n_rows & 7) || tile_floats > (2 << b) || blocks > 256) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)(2u << b), 8, (cuuint64_t)(Cp >> b), (cuuint64_t)(n_rows >> 3)};
  const cuuint64_t strides[3] = {(cuuint64_t)(8u << b), (cuuint64_t)(64u << b), (cuuint64_t)(64u << b) * (cuuint64_t)(Cp >> b)};
  const cuuint32_t box[4] = {(cuuint32_t)tile_floats, 8, 1, (cuuint32_t)blocks};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return encode_tiled()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)w, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int fast_columns(const ColParams& p, int sm_count, cudaStream_t stream) {
  const bool spectrum = p.in_full || p.out_full;  // only col_fast_kernel knows natural-order spectra
  if (spectrum && (p.in_full && (p.out_full || p.reduce))) return -1;
  if (!spectrum && warp16_cols_match(p.f.n, p.R, p.pad_r) && (p.Cp & 3) == 0) {
    auto k = col_warp16_kernel<256>;
    const size_t smem = sizeof(float2) * (3 * (size_t)((1024 + 128) * 4) + 7 * 8 + 64) + sizeof(float) * (size_t)p.D;
    int grid = 1;
    const long long tiles = (long long)p.S * p.n_colour * (p.Cp >> 2);
    int rc = grid_for(k, 256, smem, sm_count, tiles, &grid);
    if (rc) return rc;
    k<<<grid, 256, smem, stream>>>(p);
    return (int)cudaPeekAtLastError();
  }
  if (!spectrum && warp_cols_match(p.f.n, p.R, p.pad_r) && (p.Cp & 3) == 0) {
    const bool big = p.f.n == 4320;
    void (*k)(ColParams, const CUtensorMap);
    if (p.reduce) k = big ? col_warp_kernel<4320, 16, 15, 1, 576, false, true> : col_warp_kernel<2160, 8, 15, 2, 576, false, true>;
    else k = big ? col_warp_kernel<4320, 16, 15, 1, 576, false, false> : col_warp_kernel<2160, 8, 15, 2, 576, false, false>;
    const size_t nel = big ? 2 * 4320 : 4 * 2160;
    const size_t tabs = big ? 15 * 15 + LHG_COL_TABQ * 240 : 7 * 15 + LHG_COL_TABQ * 120;
    const size_t smem = sizeof(float2) * (3 * nel + tabs) + sizeof(float) * (size_t)p.D;
    int grid = 1;
    const long long tiles = (long long)p.S * p.n_colour * (p.Cp >> (big ? 1 : 2));
    // The strips are staged by the TMA unit (col_warp.cuh): every strip of the adjoint launch, the one strip per tile
    // of the forward launch when D is even.  Measured on C4: column kernel 5.68 -> 5.51 ms per step with the adjoint's
    // strips alone.  The 2160-point kernel (4-column tiles, 32-byte pieces) measured 1-3 % slower with TMA under the
    // one-barrier-per-depth loop; with the paired loop (which the adjoint launch only has with TMA staging) it is
    // 3 % faster (C5: 134.8 -> 130.5 ms), so TMA is on for both (LHG_COL_TMA=0: off everywhere).
    CUtensorMap tmap{};
    const int nrb = p.R / 8;  // row blocks of a strip; split over two boxes when a box dimension (256) cannot hold them
    static const int tma_mode = [] { const char* e = getenv("LHG_COL_TMA"); return e ? atoi(e) : 1; }();
    const bool want = tma_mode != 0 && (p.reduce ? p.D > 1 : (p.D & 1) == 0);
    if (want && make_col_tmap(&tmap, p.in, (long long)p.S * (p.reduce ? p.D : 1) * p.n_colour * p.R, p.Cp, p.blocked_in,
                              big ? 4 : 8, nrb > 256 ? nrb / 2 : nrb))
    {
      if (p.reduce) k = big ? col_warp_kernel<4320, 16, 15, 1, 576, true, true> : col_warp_kernel<2160, 8, 15, 2, 576, true, true>;
      else k = big ? col_warp_kernel<4320, 16, 15, 1, 576, true, false> : col_warp_kernel<2160, 8, 15, 2, 576, true, false>;
    }
    int rc = grid_for(k, 576, smem, sm_count, tiles, &grid);
    if (rc) return rc;
    k<<<grid, 576, smem, stream>>>(p, tmap);
    return (int)cudaPeekAtLastError();
  }
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI)                                              \
  if (PLAN_MATCH(N, R0, KLO, KHI, p.f.n, p.R, p.pad_r) && (p.Cp & ((1 << LT) - 1)) == 0) {  \
    using Pl = FastPlan<N, R0, R1, R2, R3>;                                                 \
    auto k = col_fast_kernel<Pl, LT, NT, KLO, KHI>;                                         \
    const size_t smem = sizeof(float2) * (3 * (size_t)(N << LT) + Pl::tab_total(2)) +       \
                        sizeof(float) * (size_t)p.D;                                        \
    int grid = 1;                                                                           \
    const long long tiles = (long long)p.S * p.n_colour * (p.Cp >> LT);                     \
    int rc = grid_for(k, NT, smem, sm_count, tiles, &grid);                                 \
    if (rc) return rc;                                                                      \
    k<<<grid, NT, smem, stream>>>(p);                                                       \
    return (int)cudaPeekAtLastError();                                                      \
  }
  FAST_COL_PLANS(X)
#undef X
  return -1;
}

int fast_wm_tiled(const Phys& ph, const float* wm, int n_colour, int logT, const int* row_perm, const int* col_perm,
                  float* wmt, int* tile_active, int sm_count, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(tile_active, 0, sizeof(int) * (size_t)(ph.Cp >> logT), stream);
  if (e != cudaSuccess) return (int)e;
  wm_tiled_kernel<<<sm_count * 8, 256, 0, stream>>>(ph, wm, n_colour, logT, row_perm, col_perm, wmt, tile_active);
  return (int)cudaPeekAtLastError();
}

}  // namespace asmb
