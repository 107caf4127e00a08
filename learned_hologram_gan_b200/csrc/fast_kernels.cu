// fast_kernels.cu -- compile-time planned kernels for the transform lengths of the BASELINE configs.
//
// Same three-kernel pipeline as asm_b200.cu, specialised per (length, pad):
//   * radix 8..18 butterflies in registers, 3 (or 4) passes per transform instead of 5-6, twiddles from
//     shared-memory tables;
//   * the first pass of every transform reads global memory directly and the last one writes global
//     memory directly; the zero-pad rows/columns are never loaded or added (pruned first butterfly) and
//     the cropped-away outputs of the last inverse butterfly are never computed;
//   * column kernel: the (masked) forward spectrum of a tile stays in shared memory across the depth
//     loop, the transfer function is generated per depth from the tile's w values with the SFU sin/cos
//     and multiplied in as the load stage of the first inverse pass; the adjoint accumulates the depth
//     sum in the same buffer.  Tiles that lie completely outside the circular mask are not transformed.
//   * W1/W2 keep the scrambled column order of the row transform (no reordering pass); the w/mask grid
//     is pre-permuted into the tile order of the column kernel once per geometry (asm_build_wm_tiled).
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "fft_fast.cuh"

namespace asmb {

template <bool PLANAR, int LOGT, int N>
__device__ __forceinline__ int sidx(int row, int t) {
  if constexpr (PLANAR) return t * N + row;
  else return (row << LOGT) + t;
}

// middle passes (1 .. NPASS-2) of a transform held in shared memory
template <class P, int LOGT, int NT, bool PLANAR, bool TAB0>
struct Seq {
  static constexpr int N = P::N;
  template <int PASS, bool DIT>
  __device__ __forceinline__ static void one(float2* buf, const float2* tabs, int tid) {
    auto ld = [&](int row, int t, int) { return buf[sidx<PLANAR, LOGT, N>(row, t)]; };
    auto st = [&](int row, int t, int, float2 v) { buf[sidx<PLANAR, LOGT, N>(row, t)] = v; };
    fpass<P, PASS, LOGT, NT, DIT, PLANAR, true, 0, P::radix(PASS)>(nullptr, tabs + P::tab_off(PASS, TAB0), tid, ld, st);
    __syncthreads();
  }
  __device__ __forceinline__ static void dif_middle(float2* buf, const float2* tabs, int tid) {
    if constexpr (P::NPASS >= 3) one<1, false>(buf, tabs, tid);
    if constexpr (P::NPASS >= 4) one<2, false>(buf, tabs, tid);
  }
  __device__ __forceinline__ static void dit_middle(float2* buf, const float2* tabs, int tid) {
    if constexpr (P::NPASS >= 4) one<2, true>(buf, tabs, tid);
    if constexpr (P::NPASS >= 3) one<1, true>(buf, tabs, tid);
  }
};

// ------------------------------------------------------------------------------------------------
// column kernel
// ------------------------------------------------------------------------------------------------
// Shared memory per CTA: buf (exchange space of the passes), bufX (forward mode: masked spectrum of the
// tile, kept across the depth loop; reduce mode: the depth-sum accumulator), bufW (w of every bin of the
// tile, sign bit = outside the mask) and the twiddle tables.  Only the butterfly in flight lives in
// registers, so the CTA can be large.
template <class P, int LOGT, int NT, int KLO, int KHI, bool TAB0>
__global__ void __launch_bounds__(NT, 1) col_fast_kernel(ColParams a) {
  extern __shared__ float2 smem[];
  constexpr int N = P::N, T = 1 << LOGT, NEL = N << LOGT, LAST = P::NPASS - 1, R0 = P::R0, M0 = N / R0;
  float2* const buf = smem;
  float2* const bufX = buf + NEL;
  float* const bufW = reinterpret_cast<float*>(bufX + NEL);
  float2* const tabs = reinterpret_cast<float2*>(bufW + NEL);
  using Sq = Seq<P, LOGT, NT, false, TAB0>;
  const int tid = threadIdx.x;
  const float2* __restrict__ tw = a.f.tw;
  const int tiles_per_plane = a.Cp >> LOGT;
  const long long n_tiles = (long long)a.S * a.n_colour * tiles_per_plane;
  const int R = a.R, Cp = a.Cp;
  const int use_h = a.use_h;
  const bool masked = (a.flags & kFilterMask) != 0;
  const float bsign = (a.flags & kFilterConj) ? -1.0f : 1.0f;
  const size_t strip = (size_t)R * Cp;

  fill_tables<P, TAB0>(tabs, tw, tid, NT);
  __syncthreads();

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int ct = (int)(tile % tiles_per_plane);
    const long long g = tile / tiles_per_plane;  // sample * n_colour + colour
    const int colour = (int)(g % a.n_colour);
    const long long s = g / a.n_colour;
    const int col0 = ct << LOGT;

    if (masked && a.tile_active && !a.tile_active[ct]) {
      // every bin of these columns is outside the circular mask: the result is zero
      const int n_out = a.reduce ? 1 : a.D;
      for (int d = 0; d < n_out; ++d) {
        const size_t plane = a.reduce ? (size_t)g : ((size_t)s * a.D + d) * a.n_colour + colour;
        float2* dst = a.out + plane * strip + col0;
        for (int e = tid; e < (R << LOGT); e += NT) dst[(size_t)(e >> LOGT) * Cp + (e & (T - 1))] = make_float2(0.0f, 0.0f);
      }
      continue;
    }

    // w of every (scrambled position, column) of the tile: one contiguous run of the pre-tiled grid
    if (a.wmt) {
      const float4* src = reinterpret_cast<const float4*>(a.wmt + ((size_t)colour * tiles_per_plane + ct) * NEL);
      for (int e = tid; e < NEL / 4; e += NT) cp_async16(bufW + 4 * e, src + e);
      cp_async_commit();
    }

    // forward transform of one stored strip (the R non-pad rows of T columns) into st_last
    auto forward = [&](const float2* __restrict__ src, auto st_last) {
      auto ld_g = [&](int row, int t, int k) { return __ldg(src + (size_t)(row - KLO * M0) * Cp + t); };
      auto st_s = [&](int row, int t, int, float2 v) { buf[(row << LOGT) + t] = v; };
      auto ld_s = [&](int row, int t, int) { return buf[(row << LOGT) + t]; };
      fpass<P, 0, LOGT, NT, false, false, TAB0, KLO, KHI>(tw, tabs, tid, ld_g, st_s);
      __syncthreads();
      Sq::dif_middle(buf, tabs, tid);
      fpass<P, LAST, LOGT, NT, false, false, false, 0, P::radix(LAST)>(tw, tabs, tid, ld_s, st_last);
    };
    // inverse transform from ld_first to the R crop rows of dst
    auto inverse = [&](auto ld_first, float2* __restrict__ dst) {
      auto st_s = [&](int row, int t, int, float2 v) { buf[(row << LOGT) + t] = v; };
      auto ld_s = [&](int row, int t, int) { return buf[(row << LOGT) + t]; };
      fpass<P, LAST, LOGT, NT, true, false, false, 0, P::radix(LAST)>(tw, tabs, tid, ld_first, st_s);
      __syncthreads();
      Sq::dit_middle(buf, tabs, tid);
      auto st_g = [&](int row, int t, int, float2 v) { dst[(size_t)(row - KLO * M0) * Cp + t] = cswap(v); };
      fpass<P, 0, LOGT, NT, true, false, TAB0, KLO, KHI>(tw, tabs, tid, ld_s, st_g);
      __syncthreads();
    };

    if (!a.reduce) {
      const float2* src = a.in + (size_t)g * strip + col0;
      cp_async_wait_all();  // bufW (visible to the other threads after the barriers inside forward())
      forward(src, [&](int row, int t, int, float2 v) {
        const int e = (row << LOGT) + t;
        if (masked && signbit(bufW[e])) v = make_float2(0.0f, 0.0f);
        bufX[e] = v;
      });
      // the last forward pass and the first inverse pass touch the same slots from the same thread:
      // no barrier needed in between.
      for (int d = 0; d < a.D; ++d) {
        const size_t out_plane = ((size_t)s * a.D + d) * a.n_colour + colour;
        const int zi = a.depth_index ? a.depth_index[s * a.D + d] : d;
        const float beta = use_h ? bsign * beta_of(a.z[zi]) : 0.0f;
        float2* dst = a.out + out_plane * strip + col0;
        if (use_h) {
          inverse([&](int row, int t, int) {
            const int e = (row << LOGT) + t;
            return cswap(cmul(bufX[e], fast_cis(__fmul_rn(beta, fabsf(bufW[e])))));
          }, dst);
        } else {
          inverse([&](int row, int t, int) { return cswap(bufX[(row << LOGT) + t]); }, dst);
        }
      }
    } else {
      for (int e = tid; e < NEL; e += NT) bufX[e] = make_float2(0.0f, 0.0f);
      cp_async_wait_all();
      for (int d = 0; d < a.D; ++d) {
        const size_t in_plane = ((size_t)s * a.D + d) * a.n_colour + colour;
        const int zi = a.depth_index ? a.depth_index[s * a.D + d] : d;
        const float beta = use_h ? bsign * beta_of(a.z[zi]) : 0.0f;
        const float2* src = a.in + in_plane * strip + col0;
        if (d + 1 < a.D) {  // pull the next depth's strip into L2 while this one is transformed
          const float2* nxt = src + (size_t)a.n_colour * strip;
          for (int r = tid; r < R; r += NT) prefetch_l2(nxt + (size_t)r * Cp);
        }
        if (use_h) {
          forward(src, [&](int row, int t, int, float2 v) {
            const int e = (row << LOGT) + t;
            const float2 p = cmul(v, fast_cis(__fmul_rn(beta, fabsf(bufW[e]))));
            const float2 acc = bufX[e];
            bufX[e] = make_float2(acc.x + p.x, acc.y + p.y);
          });
        } else {
          forward(src, [&](int row, int t, int, float2 v) {
            const int e = (row << LOGT) + t;
            const float2 acc = bufX[e];
            bufX[e] = make_float2(acc.x + v.x, acc.y + v.y);
          });
        }
        __syncthreads();  // buf is rewritten by the next depth's first pass
      }
      inverse([&](int row, int t, int) {
        const int e = (row << LOGT) + t;
        float2 v = bufX[e];
        if (masked && signbit(bufW[e])) v = make_float2(0.0f, 0.0f);
        return cswap(v);
      }, a.out + (size_t)g * strip + col0);
    }
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
// row kernels: T rows per CTA, planar in shared memory
// ------------------------------------------------------------------------------------------------
template <class P, int LOGT, int NT, int KLO, int KHI, bool TAB0>
__global__ void __launch_bounds__(NT) row_fwd_fast_kernel(RowIn in, long long n_rows, int C, float2* __restrict__ w1,
                                                          const float2* __restrict__ tw) {
  extern __shared__ float2 smem[];
  constexpr int N = P::N, T = 1 << LOGT, LAST = P::NPASS - 1, M0 = N / P::R0;
  float2* const buf = smem;
  float2* const tabs = buf + (N << LOGT);
  using Sq = Seq<P, LOGT, NT, true, TAB0>;
  const int tid = threadIdx.x;
  const long long n_groups = (n_rows + T - 1) >> LOGT;
  fill_tables<P, TAB0>(tabs, tw, tid, NT);
  __syncthreads();
  auto ld_s = [&](int row, int t, int) { return buf[t * N + row]; };
  auto st_s = [&](int row, int t, int, float2 v) { buf[t * N + row] = v; };
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long row0 = grp << LOGT;
    auto ld_in = [&](int pos, int t, int) {
      const long long row = row0 + t;
      if (T > 1 && row >= n_rows) return make_float2(0.0f, 0.0f);
      return load_input(in, (size_t)row * C + (pos - KLO * M0));
    };
    fpass<P, 0, LOGT, NT, false, true, TAB0, KLO, KHI>(tw, tabs, tid, ld_in, st_s);
    __syncthreads();
    Sq::dif_middle(buf, tabs, tid);
    fpass<P, LAST, LOGT, NT, false, true, false, 0, P::radix(LAST)>(tw, tabs, tid, ld_s, st_s);
    __syncthreads();
    // scrambled order straight out (the column kernel never needs the natural column order)
    for (int e = tid; e < (N << LOGT) / 2; e += NT) {
      const int t = (2 * e) / N;
      const long long row = row0 + t;
      if (T == 1 || row < n_rows)
        reinterpret_cast<float4*>(w1 + (size_t)row0 * N)[e] = reinterpret_cast<const float4*>(buf)[e];
    }
    __syncthreads();
  }
}

template <class P, int LOGT, int NT, int KLO, int KHI, bool TAB0>
__global__ void __launch_bounds__(NT) row_inv_fast_kernel(RowOut o, long long n_rows, int C,
                                                          const float2* __restrict__ w2,
                                                          const float2* __restrict__ tw) {
  extern __shared__ float2 smem[];
  __shared__ float red[32];
  constexpr int N = P::N, T = 1 << LOGT, LAST = P::NPASS - 1, M0 = N / P::R0;
  float2* const buf = smem;
  float2* const tabs = buf + (N << LOGT);
  using Sq = Seq<P, LOGT, NT, true, TAB0>;
  const int tid = threadIdx.x;
  const long long n_groups = (n_rows + T - 1) >> LOGT;
  float loss_acc = 0.0f;
  fill_tables<P, TAB0>(tabs, tw, tid, NT);
  auto ld_s = [&](int row, int t, int) { return buf[t * N + row]; };
  auto st_s = [&](int row, int t, int, float2 v) { buf[t * N + row] = v; };
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long row0 = grp << LOGT;
    for (int e = tid; e < (N << LOGT) / 2; e += NT) {
      const int t = (2 * e) / N;
      if (T == 1 || row0 + t < n_rows)
        cp_async16(reinterpret_cast<float4*>(buf) + e, reinterpret_cast<const float4*>(w2 + (size_t)row0 * N) + e);
      else
        reinterpret_cast<float4*>(buf)[e] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    auto ld_first = [&](int row, int t, int) { return cswap(buf[t * N + row]); };
    fpass<P, LAST, LOGT, NT, true, true, false, 0, P::radix(LAST)>(tw, tabs, tid, ld_first, st_s);
    __syncthreads();
    Sq::dit_middle(buf, tabs, tid);
    auto st_out = [&](int pos, int t, int, float2 v) {
      const long long row = row0 + t;
      if (T == 1 || row < n_rows) store_output(o, (size_t)row * C + (pos - KLO * M0), cswap(v), loss_acc);
    };
    fpass<P, 0, LOGT, NT, true, true, TAB0, KLO, KHI>(tw, tabs, tid, ld_s, st_out);
    __syncthreads();
  }
  if (o.loss_partial) block_loss_reduce(loss_acc, o.loss_partial, red);
}

// ------------------------------------------------------------------------------------------------
// plans and dispatch
// ------------------------------------------------------------------------------------------------
// A plan applies to a geometry when the un-padded extent is (KHI-KLO)*N/R0 and the pad is KLO*N/R0.
//             N     R0  R1  R2  R3  LOGT  NT  KLO KHI TAB0
#define FAST_ROW_PLANS(X)                      \
  X(7680, 8, 8, 8, 15, 0, 256, 2, 6, false)    \
  X(3840, 16, 16, 15, 1, 0, 256, 4, 12, false) \
  X(3840, 16, 16, 15, 1, 0, 256, 0, 16, false) \
  X(1920, 8, 16, 15, 1, 1, 256, 0, 8, false)   \
  X(1024, 16, 16, 4, 1, 2, 256, 5, 11, true)   \
  X(384, 8, 16, 3, 1, 3, 256, 0, 8, true)

//             N     R0  R1  R2  R3  LOGT  NT  KLO KHI TAB0
#define FAST_COL_PLANS(X)                      \
  X(4320, 16, 18, 15, 1, 1, 576, 4, 12, true)  \
  X(2160, 16, 9, 15, 1, 2, 576, 4, 12, true)   \
  X(2160, 16, 9, 15, 1, 2, 576, 0, 16, true)   \
  X(1080, 8, 9, 15, 1, 3, 576, 0, 8, true)     \
  X(1024, 16, 16, 4, 1, 3, 512, 5, 11, true)   \
  X(384, 8, 16, 3, 1, 4, 384, 0, 8, true)

#define PLAN_MATCH(N, R0, KLO, KHI, n, ext, pad) \
  ((n) == N && (ext) == (KHI - KLO) * (N / R0) && (pad) == KLO * (N / R0))

bool fast_rows_supported(int n, int cols, int pad) {
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI, TAB0) \
  if (PLAN_MATCH(N, R0, KLO, KHI, n, cols, pad)) return true;
  FAST_ROW_PLANS(X)
#undef X
  return false;
}

// log2 of the columns per tile of the fast column kernel, or -1
int fast_cols_logt(int n, int rows, int pad) {
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI, TAB0) \
  if (PLAN_MATCH(N, R0, KLO, KHI, n, rows, pad)) return LT;
  FAST_COL_PLANS(X)
#undef X
  return -1;
}

void fast_rows_perm(int n, int* perm_out) {
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI, TAB0)                                \
  if (n == N) {                                                                     \
    for (int p = 0; p < N; ++p) perm_out[p] = FastPlan<N, R0, R1, R2, R3>::perm(p); \
    return;                                                                         \
  }
  FAST_ROW_PLANS(X)
#undef X
}

void fast_cols_perm(int n, int* perm_out) {
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI, TAB0)                                \
  if (n == N) {                                                                     \
    for (int p = 0; p < N; ++p) perm_out[p] = FastPlan<N, R0, R1, R2, R3>::perm(p); \
    return;                                                                         \
  }
  FAST_COL_PLANS(X)
#undef X
}

template <class K>
static int grid_for(K kernel, int threads, size_t smem, int sm_count, long long work, int* grid_out) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int occ = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
  if (e != cudaSuccess) return (int)e;
  if (occ < 1) occ = 1;
  long long grid = (long long)sm_count * occ;
  if (grid > work) grid = work;
  if (grid < 1) grid = 1;
  *grid_out = (int)grid;
  return 0;
}

int fast_row_forward(int n, const float2* tw, const RowIn& in, long long n_rows, int C, int pad_c, float2* w1,
                     int sm_count, cudaStream_t stream) {
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI, TAB0)                                        \
  if (PLAN_MATCH(N, R0, KLO, KHI, n, C, pad_c)) {                                           \
    using Pl = FastPlan<N, R0, R1, R2, R3>;                                                 \
    auto k = row_fwd_fast_kernel<Pl, LT, NT, KLO, KHI, TAB0>;                               \
    const size_t smem = sizeof(float2) * ((size_t)(N << LT) + Pl::tab_total(TAB0));         \
    int grid = 1;                                                                           \
    int rc = grid_for(k, NT, smem, sm_count, (n_rows + (1 << LT) - 1) >> LT, &grid);        \
    if (rc) return rc;                                                                      \
    k<<<grid, NT, smem, stream>>>(in, n_rows, C, w1, tw);                                   \
    return (int)cudaPeekAtLastError();                                                      \
  }
  FAST_ROW_PLANS(X)
#undef X
  return -1;
}

int fast_row_inverse(int n, const float2* tw, const RowOut& out, long long n_rows, int C, int pad_c,
                     const float2* w2, int sm_count, int max_blocks, cudaStream_t stream) {
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI, TAB0)                                        \
  if (PLAN_MATCH(N, R0, KLO, KHI, n, C, pad_c)) {                                           \
    using Pl = FastPlan<N, R0, R1, R2, R3>;                                                 \
    auto k = row_inv_fast_kernel<Pl, LT, NT, KLO, KHI, TAB0>;                               \
    const size_t smem = sizeof(float2) * ((size_t)(N << LT) + Pl::tab_total(TAB0));         \
    int grid = 1;                                                                           \
    int rc = grid_for(k, NT, smem, sm_count, (n_rows + (1 << LT) - 1) >> LT, &grid);        \
    if (rc) return rc;                                                                      \
    if (max_blocks > 0 && grid > max_blocks) grid = max_blocks;                             \
    k<<<grid, NT, smem, stream>>>(out, n_rows, C, w2, tw);                                  \
    return (int)cudaPeekAtLastError();                                                      \
  }
  FAST_ROW_PLANS(X)
#undef X
  return -1;
}

int fast_columns(const ColParams& p, int sm_count, cudaStream_t stream) {
#define X(N, R0, R1, R2, R3, LT, NT, KLO, KHI, TAB0)                                        \
  if (PLAN_MATCH(N, R0, KLO, KHI, p.f.n, p.R, p.pad_r) && (p.Cp & ((1 << LT) - 1)) == 0) {  \
    using Pl = FastPlan<N, R0, R1, R2, R3>;                                                 \
    auto k = col_fast_kernel<Pl, LT, NT, KLO, KHI, TAB0>;                                   \
    const size_t smem = (size_t)(N << LT) * (2 * sizeof(float2) + sizeof(float)) +          \
                        sizeof(float2) * Pl::tab_total(TAB0);                               \
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
