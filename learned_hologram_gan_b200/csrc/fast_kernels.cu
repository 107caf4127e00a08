// fast_kernels.cu -- compile-time planned kernels for the transform lengths of the BASELINE configs.
//
// Same three-kernel pipeline as asm_b200.cu, specialised per length:
//   * radix 8..18 butterflies in registers, 3 (or 4) passes per transform instead of 5-6;
//   * the first pass of every transform reads global memory (or registers) directly and the last
//     one writes global memory (or registers) directly: no staging copy;
//   * column kernel: the forward spectrum of a tile stays in REGISTERS across the depth loop
//     (each thread owns the elements of its last-pass butterflies), the transfer function is
//     generated per depth from per-thread w registers with the SFU sin/cos, multiplied in as the
//     load stage of the first inverse pass; the adjoint accumulates the depth sum in the same
//     registers.  One shared-memory buffer per tile is the only exchange space.
//   * W1/W2 keep the scrambled column order of the row transform (no reordering pass); the column
//     kernel looks the frequency bin of a stored column up through col_perm.
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "fft_fast.cuh"

namespace asmb {

template <class P, int LOGT, int NT>
struct Seq {
  static constexpr int N = P::N;
  static constexpr int NC1 = N / P::R0, NC2 = NC1 / P::R1;
  __device__ __forceinline__ static void dif_middle(float2* buf, const float2* tw, int tid) {
    auto ld = [&](int, int row, int t, int) { return buf[(row << LOGT) + t]; };
    auto st = [&](int, int row, int t, int, float2 v) { buf[(row << LOGT) + t] = v; };
    if constexpr (P::NPASS >= 3) {
      fpass<N, NC1, P::R1, LOGT, NT, false>(tw, tid, ld, st);
      __syncthreads();
    }
    if constexpr (P::NPASS >= 4) {
      fpass<N, NC2, P::R2, LOGT, NT, false>(tw, tid, ld, st);
      __syncthreads();
    }
  }
  __device__ __forceinline__ static void dit_middle(float2* buf, const float2* tw, int tid) {
    auto ld = [&](int, int row, int t, int) { return buf[(row << LOGT) + t]; };
    auto st = [&](int, int row, int t, int, float2 v) { buf[(row << LOGT) + t] = v; };
    if constexpr (P::NPASS >= 4) {
      fpass<N, NC2, P::R2, LOGT, NT, true>(tw, tid, ld, st);
      __syncthreads();
    }
    if constexpr (P::NPASS >= 3) {
      fpass<N, NC1, P::R1, LOGT, NT, true>(tw, tid, ld, st);
      __syncthreads();
    }
  }
};

// ------------------------------------------------------------------------------------------------
// column kernel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 spectral_factor(float w, float beta, int use_h, int flags) {
  if ((flags & kFilterMask) && signbit(w)) return make_float2(0.0f, 0.0f);
  if (!use_h) return make_float2(1.0f, 0.0f);
  float2 f = fast_cis(__fmul_rn(beta, fabsf(w)));
  if (flags & kFilterConj) f.y = -f.y;
  return f;
}

// Shared memory per CTA: buf (exchange space of the passes), bufX (forward mode: spectrum of the tile,
// kept across the depth loop; reduce mode: the depth-sum accumulator) and bufW (w of every bin of the
// tile, sign bit = outside the mask).  Nothing but the butterfly in flight lives in registers, so the
// CTA can be large.
template <class P, int LOGT, int NT>
__global__ void __launch_bounds__(NT, 1) col_fast_kernel(ColParams a) {
  extern __shared__ float2 buf[];
  constexpr int N = P::N, T = 1 << LOGT, RL = P::RL, NEL = N << LOGT;
  float2* const bufX = buf + NEL;
  float* const bufW = reinterpret_cast<float*>(buf + 2 * NEL);
  using Sq = Seq<P, LOGT, NT>;
  const int tid = threadIdx.x;
  const float2* __restrict__ tw = a.f.tw;
  const int tiles_per_plane = a.Cp >> LOGT;
  const long long n_tiles = (long long)a.S * a.n_colour * tiles_per_plane;
  const int R = a.R, pad_r = a.pad_r, Cp = a.Cp;
  const int use_h = a.use_h, flags = a.flags;
  auto ld_s = [&](int, int row, int t, int) { return buf[(row << LOGT) + t]; };
  auto st_s = [&](int, int row, int t, int, float2 v) { buf[(row << LOGT) + t] = v; };

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int ct = (int)(tile % tiles_per_plane);
    const long long g = tile / tiles_per_plane;  // sample * n_colour + colour
    const int colour = (int)(g % a.n_colour);
    const long long s = g / a.n_colour;
    const int col0 = ct << LOGT;

    // w of every (scrambled position, column) of the tile
    for (int e = tid; e < NEL; e += NT) {
      const int t = e & (T - 1);
      const int kr = P::perm(e >> LOGT);
      const int kc = a.col_perm ? __ldg(a.col_perm + col0 + t) : col0 + t;
      float w;
      if (a.wm) {
        w = __ldg(a.wm + ((size_t)colour * N + kr) * Cp + kc);
      } else {
        w = w_value(a.ph, kr, kc, colour);
        if (radial_value(a.ph, kr, kc) > a.ph.radius) w = -w;
      }
      bufW[e] = w;
    }

    // forward transform of one stored strip (R crop rows of T columns) into st_last
    auto forward = [&](const float2* __restrict__ src, auto st_last) {
      auto ld_g = [&](int, int row, int t, int) {
        const int r = row - pad_r;
        return (r >= 0 && r < R) ? __ldg(src + (size_t)r * Cp + t) : make_float2(0.0f, 0.0f);
      };
      fpass<N, N, P::R0, LOGT, NT, false>(tw, tid, ld_g, st_s);
      __syncthreads();
      Sq::dif_middle(buf, tw, tid);
      fpass<N, RL, RL, LOGT, NT, false>(tw, tid, ld_s, st_last);
    };
    // inverse transform from ld_first to the R crop rows of dst
    auto inverse = [&](auto ld_first, float2* __restrict__ dst) {
      fpass<N, RL, RL, LOGT, NT, true>(tw, tid, ld_first, st_s);
      __syncthreads();
      Sq::dit_middle(buf, tw, tid);
      auto st_g = [&](int, int row, int t, int, float2 v) {
        const int r = row - pad_r;
        if (r >= 0 && r < R) dst[(size_t)r * Cp + t] = cswap(v);
      };
      fpass<N, N, P::R0, LOGT, NT, true>(tw, tid, ld_s, st_g);
      __syncthreads();
    };

    if (!a.reduce) {
      forward(a.in + (size_t)g * R * Cp + col0,
              [&](int, int row, int t, int, float2 v) { bufX[(row << LOGT) + t] = v; });
      // the last forward pass and the first inverse pass touch the same slots from the same thread:
      // no barrier needed in between (bufW was filled before the barriers inside forward()).
      for (int d = 0; d < a.D; ++d) {
        const size_t out_plane = ((size_t)s * a.D + d) * a.n_colour + colour;
        const int zi = a.depth_index ? a.depth_index[s * a.D + d] : d;
        const float beta = use_h ? beta_of(a.z[zi]) : 0.0f;
        inverse([&](int, int row, int t, int) {
          const int e = (row << LOGT) + t;
          return cswap(cmul(bufX[e], spectral_factor(bufW[e], beta, use_h, flags)));
        }, a.out + out_plane * (size_t)R * Cp + col0);
      }
    } else {
      for (int e = tid; e < NEL; e += NT) bufX[e] = make_float2(0.0f, 0.0f);
      for (int d = 0; d < a.D; ++d) {
        const size_t in_plane = ((size_t)s * a.D + d) * a.n_colour + colour;
        const int zi = a.depth_index ? a.depth_index[s * a.D + d] : d;
        const float beta = use_h ? beta_of(a.z[zi]) : 0.0f;
        forward(a.in + in_plane * (size_t)R * Cp + col0, [&](int, int row, int t, int, float2 v) {
          const int e = (row << LOGT) + t;
          const float2 p = cmul(v, spectral_factor(bufW[e], beta, use_h, flags));
          const float2 acc = bufX[e];
          bufX[e] = make_float2(acc.x + p.x, acc.y + p.y);
        });
        __syncthreads();  // buf is rewritten by the next depth's first pass
      }
      inverse([&](int, int row, int t, int) { return cswap(bufX[(row << LOGT) + t]); },
              a.out + (size_t)g * R * Cp + col0);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// row kernels: T rows interleaved in shared memory
// ------------------------------------------------------------------------------------------------
template <class P, int LOGT, int NT>
__global__ void __launch_bounds__(NT) row_fwd_fast_kernel(RowIn in, long long n_rows, int C, int pad_c,
                                                          float2* __restrict__ w1, const float2* __restrict__ tw) {
  extern __shared__ float2 buf[];
  constexpr int N = P::N, T = 1 << LOGT, RL = P::RL;
  using Sq = Seq<P, LOGT, NT>;
  const int tid = threadIdx.x;
  const long long n_groups = (n_rows + T - 1) >> LOGT;
  auto ld_s = [&](int, int row, int t, int) { return buf[(row << LOGT) + t]; };
  auto st_s = [&](int, int row, int t, int, float2 v) { buf[(row << LOGT) + t] = v; };
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long row0 = grp << LOGT;
    auto ld_in = [&](int, int pos, int t, int) {
      const long long row = row0 + t;
      const int c = pos - pad_c;
      return (row < n_rows && c >= 0 && c < C) ? load_input(in, (size_t)row * C + c) : make_float2(0.0f, 0.0f);
    };
    fpass<N, N, P::R0, LOGT, NT, false>(tw, tid, ld_in, st_s);
    __syncthreads();
    Sq::dif_middle(buf, tw, tid);
    fpass<N, RL, RL, LOGT, NT, false>(tw, tid, ld_s, st_s);
    __syncthreads();
    // scrambled order straight out (the column kernel looks the bin up through col_perm)
    for (int e = tid; e < N * T; e += NT) {
      const int t = e / N, i = e - t * N;
      const long long row = row0 + t;
      if (row < n_rows) w1[(size_t)row * N + i] = buf[(i << LOGT) + t];
    }
    __syncthreads();
  }
}

template <class P, int LOGT, int NT>
__global__ void __launch_bounds__(NT) row_inv_fast_kernel(RowOut o, long long n_rows, int C, int pad_c,
                                                          const float2* __restrict__ w2,
                                                          const float2* __restrict__ tw) {
  extern __shared__ float2 buf[];
  __shared__ float red[32];
  constexpr int N = P::N, T = 1 << LOGT, RL = P::RL;
  using Sq = Seq<P, LOGT, NT>;
  const int tid = threadIdx.x;
  const long long n_groups = (n_rows + T - 1) >> LOGT;
  float loss_acc = 0.0f;
  auto ld_s = [&](int, int row, int t, int) { return buf[(row << LOGT) + t]; };
  auto st_s = [&](int, int row, int t, int, float2 v) { buf[(row << LOGT) + t] = v; };
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long row0 = grp << LOGT;
    for (int e = tid; e < N * T; e += NT) {
      const int t = e / N, i = e - t * N;
      const long long row = row0 + t;
      buf[(i << LOGT) + t] = row < n_rows ? cswap(__ldg(w2 + (size_t)row * N + i)) : make_float2(0.0f, 0.0f);
    }
    __syncthreads();
    fpass<N, RL, RL, LOGT, NT, true>(tw, tid, ld_s, st_s);
    __syncthreads();
    Sq::dit_middle(buf, tw, tid);
    auto st_out = [&](int, int pos, int t, int, float2 v) {
      const long long row = row0 + t;
      const int c = pos - pad_c;
      if (row < n_rows && c >= 0 && c < C) store_output(o, (size_t)row * C + c, cswap(v), loss_acc);
    };
    fpass<N, N, P::R0, LOGT, NT, true>(tw, tid, ld_s, st_out);
    __syncthreads();
  }
  if (o.loss_partial) block_loss_reduce(loss_acc, o.loss_partial, red);
}

// ------------------------------------------------------------------------------------------------
// plans and dispatch
// ------------------------------------------------------------------------------------------------
//            N     R0  R1  R2  R3  LOGT  NT
#define FAST_ROW_PLANS(X)        \
  X(7680, 8, 8, 8, 15, 0, 512)   \
  X(3840, 16, 16, 15, 1, 0, 256) \
  X(1920, 8, 16, 15, 1, 1, 256)  \
  X(1024, 16, 16, 4, 1, 2, 256)  \
  X(384, 8, 16, 3, 1, 3, 256)

//            N     R0  R1  R2  R3  LOGT  NT  variant (LHG_COL_VARIANT, 0 = default)
#define FAST_COL_PLANS(X)              \
  X(4320, 16, 18, 15, 1, 1, 576, 0)    \
  X(4320, 16, 18, 15, 1, 1, 288, 1)    \
  X(2160, 16, 9, 15, 1, 2, 576, 0)     \
  X(1080, 8, 9, 15, 1, 3, 576, 0)      \
  X(1024, 16, 16, 4, 1, 3, 512, 0)     \
  X(384, 8, 16, 3, 1, 4, 384, 0)

bool fast_rows_supported(int n) {
#define X(N, R0, R1, R2, R3, LT, NT) \
  if (n == N) return true;
  FAST_ROW_PLANS(X)
#undef X
  return false;
}

bool fast_cols_supported(int n) {
#define X(N, R0, R1, R2, R3, LT, NT, VAR) \
  if (n == N) return true;
  FAST_COL_PLANS(X)
#undef X
  return false;
}

void fast_rows_perm(int n, int* perm_out) {
#define X(N, R0, R1, R2, R3, LT, NT)                                           \
  if (n == N) {                                                                \
    for (int p = 0; p < N; ++p) perm_out[p] = FastPlan<N, R0, R1, R2, R3>::perm(p); \
    return;                                                                    \
  }
  FAST_ROW_PLANS(X)
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
#define X(N, R0, R1, R2, R3, LT, NT)                                                        \
  if (n == N) {                                                                             \
    auto k = row_fwd_fast_kernel<FastPlan<N, R0, R1, R2, R3>, LT, NT>;                      \
    const size_t smem = sizeof(float2) * N << LT;                                           \
    int grid = 1;                                                                           \
    int rc = grid_for(k, NT, smem, sm_count, (n_rows + (1 << LT) - 1) >> LT, &grid);        \
    if (rc) return rc;                                                                      \
    k<<<grid, NT, smem, stream>>>(in, n_rows, C, pad_c, w1, tw);                            \
    return (int)cudaPeekAtLastError();                                                      \
  }
  FAST_ROW_PLANS(X)
#undef X
  return -1;
}

int fast_row_inverse(int n, const float2* tw, const RowOut& out, long long n_rows, int C, int pad_c,
                     const float2* w2, int sm_count, int max_blocks, cudaStream_t stream) {
#define X(N, R0, R1, R2, R3, LT, NT)                                                        \
  if (n == N) {                                                                             \
    auto k = row_inv_fast_kernel<FastPlan<N, R0, R1, R2, R3>, LT, NT>;                      \
    const size_t smem = sizeof(float2) * N << LT;                                           \
    int grid = 1;                                                                           \
    int rc = grid_for(k, NT, smem, sm_count, (n_rows + (1 << LT) - 1) >> LT, &grid);        \
    if (rc) return rc;                                                                      \
    if (max_blocks > 0 && grid > max_blocks) grid = max_blocks;                             \
    k<<<grid, NT, smem, stream>>>(out, n_rows, C, pad_c, w2, tw);                           \
    return (int)cudaPeekAtLastError();                                                      \
  }
  FAST_ROW_PLANS(X)
#undef X
  return -1;
}

static int col_variant() {
  static const int v = [] {
    const char* e = getenv("LHG_COL_VARIANT");
    return e ? atoi(e) : 0;
  }();
  return v;
}

int fast_columns(const ColParams& p, int sm_count, cudaStream_t stream) {
  const int var = col_variant();
#define X(N, R0, R1, R2, R3, LT, NT, VAR)                                                   \
  if (p.f.n == N && (p.Cp & ((1 << LT) - 1)) == 0 && (VAR == var || (VAR == 0 && N != 4320))) { \
    auto k = col_fast_kernel<FastPlan<N, R0, R1, R2, R3>, LT, NT>;                          \
    const size_t smem = (size_t)(N << LT) * (2 * sizeof(float2) + sizeof(float));           \
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

}  // namespace asmb
