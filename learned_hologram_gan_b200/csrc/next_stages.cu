// next_stages.cu -- SURVEY.md 8(f), rows N1-N4: the stages either side of the propagation path
// (include/lhg_next_b200.h).  All of them are HBM-bound streaming passes over [plane][row][col] fp32 data:
// 16-byte accesses, every input element read once per pass (a strip of rows walks down the plane with the
// previous / next row kept in registers, horizontal neighbours through warp shuffles), fixed-order two-stage
// reductions (per-block partials, one finishing block, double accumulation in the finisher), no float atomics.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/lhg_next_b200.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int launched(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) return fail(LHG_ECUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
  return LHG_OK;
}

constexpr int kThreads = 128;      // 4 warps; a block spans kThreads * V columns
constexpr int kRowsPerStrip = 16;  // rows one block walks down (one halo row above / below); 32 measured slower
constexpr int kFinishThreads = 256;

// ---- strip geometry ------------------------------------------------------------------------------------
struct Strip {
  long long plane;
  int r0, r1, c0;
  bool active, produces, has_left, has_right;  // has_right: column c0+V exists; has_left: column c0-1 exists
};

__host__ __device__ inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// HALO = 0: the horizontal neighbours of a warp's edge lanes are read from memory (cheap when a neighbour is a plain
// load).  HALO = 1 / 2: lane 31 (and lane 0) of every warp only PROVIDE values to their neighbour lanes and produce
// nothing, so a warp covers 31*V (30*V) columns and no lane ever recomputes a neighbour: for the focal loss a
// neighbour is two sin/cos pairs, and the divergent edge-lane path cost 25 % more SFU instructions per row.
template <int HALO>
__host__ __device__ constexpr int strip_block_cols(int v) { return (kThreads / 32) * (32 - HALO) * v; }

template <int V, int HALO = 0>
__device__ __forceinline__ Strip decode_strip(int rows, int cols) {
  const int ncb = div_up(cols, strip_block_cols<HALO>(V));
  const int nrs = div_up(rows, kRowsPerStrip);
  long long b = blockIdx.x;
  const int cb = (int)(b % ncb);
  b /= ncb;
  const int rs = (int)(b % nrs);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Strip s;
  s.plane = b / nrs;
  s.r0 = rs * kRowsPerStrip;
  s.r1 = min(rows, s.r0 + kRowsPerStrip);
  s.c0 = ((cb * (kThreads / 32) + warp) * (32 - HALO) + lane - (HALO == 2 ? 1 : 0)) * V;
  s.active = s.c0 >= 0 && s.c0 < cols;
  s.produces = s.active && (HALO == 0 || lane < 31) && (HALO < 2 || lane > 0);
  s.has_left = s.active && s.c0 > 0;
  s.has_right = s.active && s.c0 + V < cols;
  return s;
}

long long strip_blocks(long long planes, int rows, int cols, int v, int halo = 0) {
  const int bc = halo == 0 ? strip_block_cols<0>(v) : (halo == 1 ? strip_block_cols<1>(v) : strip_block_cols<2>(v));
  return planes * div_up(rows, kRowsPerStrip) * div_up(cols, bc);
}

template <int V>
struct Row {
  float v[V];
};

template <int V>
__device__ __forceinline__ Row<V> load_row(const float* p, bool active) {
  Row<V> r;
#pragma unroll
  for (int k = 0; k < V; ++k) r.v[k] = 0.0f;
  if (active) {
    if constexpr (V == 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p));
      r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) r.v[k] = __ldg(p + k);
    }
  }
  return r;
}

template <int V>
__device__ __forceinline__ void store_row(float* p, const Row<V>& r, bool active) {
  if (!active) return;
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else {
#pragma unroll
    for (int k = 0; k < V; ++k) p[k] = r.v[k];
  }
}

// value at column c0+V (first element of the next lane; the last lane of a warp reads it from memory)
template <int V>
__device__ __forceinline__ float right_of(float first, const float* rowp, const Strip& s) {
  float x = __shfl_down_sync(0xffffffffu, first, 1);
  if ((threadIdx.x & 31) == 31 && s.has_right) x = __ldg(rowp + V);
  return x;
}
// value at column c0-1
template <int V>
__device__ __forceinline__ float left_of(float last, const float* rowp, const Strip& s) {
  float x = __shfl_up_sync(0xffffffffu, last, 1);
  if ((threadIdx.x & 31) == 0 && s.has_left) x = __ldg(rowp - 1);
  return x;
}

__device__ __forceinline__ float sgnf(float x) { return (float)(x > 0.0f) - (float)(x < 0.0f); }

// ---- block reductions (fixed order) --------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void block_sum_store(float (&acc)[N], float* dst) {
  __shared__ float sm[kThreads / 32][N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float x = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    acc[i] = x;
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) sm[threadIdx.x >> 5][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float x = 0.0f;
    for (int w = 0; w < kThreads / 32; ++w) x += sm[w][threadIdx.x];
    dst[threadIdx.x] = x;
  }
}

__device__ __forceinline__ float nanmax(float a, float b) { return (a != a) ? a : ((b != b) ? b : fmaxf(a, b)); }
__device__ __forceinline__ float nanmin(float a, float b) { return (a != a) ? a : ((b != b) ? b : fminf(a, b)); }

template <int N>
__device__ __forceinline__ void block_max_store(float (&acc)[N], float* dst) {
  __shared__ float smx[kThreads / 32][N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float x = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = nanmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    acc[i] = x;
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) smx[threadIdx.x >> 5][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float x = smx[0][threadIdx.x];
    for (int w = 1; w < kThreads / 32; ++w) x = nanmax(x, smx[w][threadIdx.x]);
    dst[threadIdx.x] = x;
  }
}

// ---- N1a: mse + total variation of hat and target in one pass -------------------------------------------
// partial[block][5] = { sum (h-t)^2, sum |dx h|, sum |dy h|, sum |dx t|, sum |dy t| }
template <int V, bool HAS_T>
__global__ void __launch_bounds__(kThreads) amp_terms_kernel(const float* __restrict__ hat,
                                                             const float* __restrict__ tgt, int rows, int cols,
                                                             float* __restrict__ partial) {
  const Strip s = decode_strip<V>(rows, cols);
  const size_t base = (size_t)s.plane * rows * cols + s.c0;
  float acc[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  bool have_prev = s.r0 > 0;
  Row<V> ph = load_row<V>(hat + base + (size_t)max(s.r0 - 1, 0) * cols, s.active && have_prev);
  Row<V> pt = ph;
  if (HAS_T) pt = load_row<V>(tgt + base + (size_t)max(s.r0 - 1, 0) * cols, s.active && have_prev);
  for (int r = s.r0; r < s.r1; ++r) {
    const float* hp = hat + base + (size_t)r * cols;
    const Row<V> h = load_row<V>(hp, s.active);
    const float hr = right_of<V>(h.v[0], hp, s);
    Row<V> t = h;
    float tr = 0.0f;
    if (HAS_T) {
      const float* tp = tgt + base + (size_t)r * cols;
      t = load_row<V>(tp, s.active);
      tr = right_of<V>(t.v[0], tp, s);
    }
    if (s.active) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        if (HAS_T) {
          const float d = h.v[k] - t.v[k];
          acc[0] = fmaf(d, d, acc[0]);
        }
        if (k + 1 < V) {
          acc[1] += fabsf(h.v[k + 1] - h.v[k]);
          if (HAS_T) acc[3] += fabsf(t.v[k + 1] - t.v[k]);
        }
        if (have_prev) {
          acc[2] += fabsf(h.v[k] - ph.v[k]);
          if (HAS_T) acc[4] += fabsf(t.v[k] - pt.v[k]);
        }
      }
      if (s.has_right) {
        acc[1] += fabsf(hr - h.v[V - 1]);
        if (HAS_T) acc[3] += fabsf(tr - t.v[V - 1]);
      }
    }
    ph = h;
    pt = t;
    have_prev = true;
  }
  block_sum_store<5>(acc, partial + (size_t)blockIdx.x * 5);
}

// Middle stage of the reductions: CTA b folds its contiguous slice of the per-CTA float partials ([n][W], the first
// NSUM columns are sums, the rest maxima) into one row of doubles, stage[b][W].  One finishing CTA alone took 42 us
// for the 77 760 x 5 partials of the config-4 stack (13 % of the whole loss pass); <= 148 CTAs do it in a few us.
constexpr int kStageMaxBlocks = 148;
constexpr int kStageSlice = 2048;                 // partial rows per middle-stage CTA (at least)
constexpr int kStageFloats = kStageMaxBlocks * 8 * 2;  // room for stage[148][<= 8] doubles at the head of `partial`

int stage_blocks(long long n) {
  long long g = (n + kStageSlice - 1) / kStageSlice;
  return (int)(g < 1 ? 1 : (g > kStageMaxBlocks ? kStageMaxBlocks : g));
}

template <int W, int NSUM>
__global__ void __launch_bounds__(kFinishThreads) stage_partials_kernel(const float* __restrict__ partial,
                                                                       long long n, double* __restrict__ stage) {
  __shared__ double sm[kFinishThreads][W];
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long i0 = per * blockIdx.x, i1 = min(n, i0 + per);
  double acc[W];
#pragma unroll
  for (int k = 0; k < W; ++k) acc[k] = 0.0;
  for (long long i = i0 + threadIdx.x; i < i1; i += kFinishThreads) {
#pragma unroll
    for (int k = 0; k < W; ++k) {
      const float x = __ldg(partial + i * W + k);
      if (k < NSUM) acc[k] += (double)x;
      else acc[k] = (double)nanmax((float)acc[k], x);
    }
  }
#pragma unroll
  for (int k = 0; k < W; ++k) sm[threadIdx.x][k] = acc[k];
  __syncthreads();
  for (int o = kFinishThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
#pragma unroll
      for (int k = 0; k < W; ++k) {
        if (k < NSUM) sm[threadIdx.x][k] += sm[threadIdx.x + o][k];
        else sm[threadIdx.x][k] = (double)nanmax((float)sm[threadIdx.x][k], (float)sm[threadIdx.x + o][k]);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < W) stage[(size_t)blockIdx.x * W + threadIdx.x] = sm[0][threadIdx.x];
}

// one block: sums the partials in a fixed order in double, writes the five loss terms
__global__ void __launch_bounds__(kFinishThreads) amp_terms_finish_kernel(const double* __restrict__ partial,
                                                                         long long nblocks, double n, double n1,
                                                                         double n2, float alpha, int has_t,
                                                                         float* __restrict__ terms) {
  __shared__ double sm[kFinishThreads][5];
  double s[5] = {0, 0, 0, 0, 0};
  for (long long i = threadIdx.x; i < nblocks; i += kFinishThreads) {
#pragma unroll
    for (int k = 0; k < 5; ++k) s[k] += partial[i * 5 + k];
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) sm[threadIdx.x][k] = s[k];
  __syncthreads();
  for (int o = kFinishThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
#pragma unroll
      for (int k = 0; k < 5; ++k) sm[threadIdx.x][k] += sm[threadIdx.x + o][k];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float nanv = __int_as_float(0x7fc00000);
    const float mse = has_t ? (float)(sm[0][0] / n) : nanv;
    const float tvh = (float)(sm[0][1] / n1) + (float)(sm[0][2] / n2);
    const float tvt = has_t ? (float)(sm[0][3] / n1) + (float)(sm[0][4] / n2) : nanv;
    const float tvl = fabsf(tvh - tvt);
    terms[0] = mse;
    terms[1] = tvh;
    terms[2] = tvt;
    terms[3] = tvl;
    terms[4] = mse + alpha * tvl;
  }
}

// ---- N1b: gradient of g0*mse + g1*TV(hat) with respect to hat -------------------------------------------
template <int V, bool HAS_T>
__global__ void __launch_bounds__(kThreads) amp_backward_kernel(const float* __restrict__ hat,
                                                                const float* __restrict__ tgt,
                                                                const float* __restrict__ g,
                                                                const float* __restrict__ terms, float alpha,
                                                                int rows, int cols, float two_over_n, float inv_n1,
                                                                float inv_n2, float* __restrict__ grad) {
  const Strip s = decode_strip<V>(rows, cols);
  const size_t base = (size_t)s.plane * rows * cols + s.c0;
  // upstream cotangents of the five terms folded into the two the stencil needs:
  // d/d mse = g0 + g4,  d/d TV(hat) = g1 + sgn(TV(hat) - TV(target)) * (g3 + alpha * g4)
  float a = 0.0f, g1 = __ldg(g + 1);
  if (HAS_T) {
    const float g4 = __ldg(g + 4);
    a = (__ldg(g) + g4) * two_over_n;
    g1 = fmaf(sgnf(__ldg(terms + 1) - __ldg(terms + 2)), fmaf(alpha, g4, __ldg(g + 3)), g1);
  }
  const float bx = g1 * inv_n1, by = g1 * inv_n2;
  bool have_prev = s.r0 > 0;
  Row<V> prev = load_row<V>(hat + base + (size_t)max(s.r0 - 1, 0) * cols, s.active && have_prev);
  Row<V> cur = load_row<V>(hat + base + (size_t)s.r0 * cols, s.active);
  for (int r = s.r0; r < s.r1; ++r) {
    const bool have_next = r + 1 < rows;
    const size_t off = base + (size_t)r * cols;
    const Row<V> next = load_row<V>(hat + off + (have_next ? cols : 0), s.active && have_next);
    Row<V> t = cur;
    if (HAS_T) t = load_row<V>(tgt + off, s.active);
    const float left = left_of<V>(cur.v[V - 1], hat + off, s);
    const float right = right_of<V>(cur.v[0], hat + off, s);
    Row<V> out;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float c = cur.v[k];
      const float l = k > 0 ? cur.v[k > 0 ? k - 1 : 0] : left;
      const float rr = k + 1 < V ? cur.v[k + 1 < V ? k + 1 : 0] : right;
      const bool hl = k > 0 || s.has_left, hr = k + 1 < V || s.has_right;
      const float gx = (hl ? sgnf(c - l) : 0.0f) - (hr ? sgnf(rr - c) : 0.0f);
      const float gy = (have_prev ? sgnf(c - prev.v[k]) : 0.0f) - (have_next ? sgnf(next.v[k] - c) : 0.0f);
      float o = bx * gx + by * gy;
      if (HAS_T) o = fmaf(a, c - t.v[k], o);
      out.v[k] = o;
    }
    store_row<V>(grad + off, out, s.active);
    prev = cur;
    cur = next;
    have_prev = true;
  }
}

// ---- N1c: focal sin/cos phase-gradient loss --------------------------------------------------------------
// u = sin f - sin r, v = cos f - cos r; SFU sin/cos after a two-term Cody-Waite reduction (abs error < 5e-7).
__device__ __forceinline__ void sincos_turns(float x, float* s, float* c) {
  // nearest turn by the 1.5 * 2^23 trick (FP32 pipe; FRND runs at the SFU's rate)
  const float k = __fadd_rn(__fmaf_rn(x, 0.15915494309189535f, 12582912.0f), -12582912.0f);
  float y = fmaf(k, -6.2831854820251465f, x);
  y = fmaf(k, 1.7484556e-7f, y);
  *s = __sinf(y);
  *c = __cosf(y);
}

template <int V>
struct RowUV {
  float u[V], v[V], sf[V], cf[V];
};

template <int V>
__device__ __forceinline__ RowUV<V> load_uv(const float* f, const float* r, bool active) {
  const Row<V> a = load_row<V>(f, active), b = load_row<V>(r, active);
  RowUV<V> o;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    float sr, cr;
    sincos_turns(a.v[k], &o.sf[k], &o.cf[k]);
    sincos_turns(b.v[k], &sr, &cr);
    o.u[k] = o.sf[k] - sr;
    o.v[k] = o.cf[k] - cr;
  }
  return o;
}

template <int V>
__device__ __forceinline__ float2 right_uv(const RowUV<V>& x, const float* f, const float* r, const Strip& s) {
  float2 o;
  o.x = __shfl_down_sync(0xffffffffu, x.u[0], 1);
  o.y = __shfl_down_sync(0xffffffffu, x.v[0], 1);
  return o;  // lane 31 is a halo lane (decode_strip<V, HALO >= 1>): every producing lane has its neighbour in the warp
}
template <int V>
__device__ __forceinline__ float2 left_uv(const RowUV<V>& x, const float* f, const float* r, const Strip& s) {
  float2 o;
  o.x = __shfl_up_sync(0xffffffffu, x.u[V - 1], 1);
  o.y = __shfl_up_sync(0xffffffffu, x.v[V - 1], 1);
  return o;  // lane 0 is a halo lane (decode_strip<V, 2>)
}

// partial[block][6] = { sum d1^2, sum d2^2, sum d1, sum d2, max d1, max d2 } over both (sin, cos) channels
template <int V>
__global__ void __launch_bounds__(kThreads) focal_terms_kernel(const float* __restrict__ fake,
                                                               const float* __restrict__ real, int rows, int cols,
                                                               float* __restrict__ partial) {
  const Strip s = decode_strip<V, 1>(rows, cols);
  const size_t base = (size_t)s.plane * rows * cols + s.c0;
  float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f}, mx[2] = {0.0f, 0.0f};
  bool have_prev = s.r0 > 0;
  const size_t poff = base + (size_t)max(s.r0 - 1, 0) * cols;
  RowUV<V> prev = load_uv<V>(fake + poff, real + poff, s.active && have_prev);
  for (int r = s.r0; r < s.r1; ++r) {
    const size_t off = base + (size_t)r * cols;
    const RowUV<V> cur = load_uv<V>(fake + off, real + off, s.active);
    const float2 rt = right_uv<V>(cur, fake + off, real + off, s);
    if (s.produces) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        if (k + 1 < V || s.has_right) {
          const float du = fabsf((k + 1 < V ? cur.u[k + 1 < V ? k + 1 : 0] : rt.x) - cur.u[k]);
          const float dv = fabsf((k + 1 < V ? cur.v[k + 1 < V ? k + 1 : 0] : rt.y) - cur.v[k]);
          sum[0] = fmaf(du, du, sum[0]);
          sum[0] = fmaf(dv, dv, sum[0]);
          sum[2] += du + dv;
          mx[0] = fmaxf(mx[0], fmaxf(du, dv));  // a NaN input reaches the loss through the sums
        }
        if (have_prev) {
          const float du = fabsf(cur.u[k] - prev.u[k]);
          const float dv = fabsf(cur.v[k] - prev.v[k]);
          sum[1] = fmaf(du, du, sum[1]);
          sum[1] = fmaf(dv, dv, sum[1]);
          sum[3] += du + dv;
          mx[1] = fmaxf(mx[1], fmaxf(du, dv));
        }
      }
    }
    prev = cur;
    have_prev = true;
  }
  block_sum_store<4>(sum, partial + (size_t)blockIdx.x * 6);
  block_max_store<2>(mx, partial + (size_t)blockIdx.x * 6 + 4);
}

__global__ void __launch_bounds__(kFinishThreads) focal_terms_finish_kernel(const double* __restrict__ partial,
                                                                           long long nblocks, double n1, double n2,
                                                                           float* __restrict__ terms) {
  __shared__ double sm[kFinishThreads][4];
  __shared__ float mm[kFinishThreads][2];
  double s[4] = {0, 0, 0, 0};
  float m[2] = {0.0f, 0.0f};
  for (long long i = threadIdx.x; i < nblocks; i += kFinishThreads) {
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] += partial[i * 6 + k];
    m[0] = nanmax(m[0], (float)partial[i * 6 + 4]);
    m[1] = nanmax(m[1], (float)partial[i * 6 + 5]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) sm[threadIdx.x][k] = s[k];
  mm[threadIdx.x][0] = m[0]; mm[threadIdx.x][1] = m[1];
  __syncthreads();
  for (int o = kFinishThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
#pragma unroll
      for (int k = 0; k < 4; ++k) sm[threadIdx.x][k] += sm[threadIdx.x + o][k];
      mm[threadIdx.x][0] = nanmax(mm[threadIdx.x][0], mm[threadIdx.x + o][0]);
      mm[threadIdx.x][1] = nanmax(mm[threadIdx.x][1], mm[threadIdx.x + o][1]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    terms[0] = mm[0][0];
    terms[1] = mm[0][1];
    // mean(d * (d / max)) = sum d^2 / (max * count); 0/0 = NaN for identical inputs, as in the reference
    terms[2] = (float)(sm[0][0] / ((double)mm[0][0] * n1)) + (float)(sm[0][1] / ((double)mm[0][1] * n2));
    terms[3] = (float)(sm[0][2] / n1) + (float)(sm[0][3] / n2);  // phase_sincos_gradient_loss: plain means
  }
}

// d loss / d fake = cos f * G_u - sin f * G_v,  G_u[i] = cx * sum_{row nbrs}(u_i - u_n) + cy * sum_{col nbrs}(u_i - u_n)
// L1 = true: the un-weighted loss (phase_sincos_gradient_loss): the stencil takes sgn(u_i - u_n) and no maxima
template <int V, bool L1 = false>
__global__ void __launch_bounds__(kThreads) focal_backward_kernel(const float* __restrict__ fake,
                                                                  const float* __restrict__ real,
                                                                  const float* __restrict__ terms,
                                                                  const float* __restrict__ g, int rows, int cols,
                                                                  float inv_n1, float inv_n2,
                                                                  float* __restrict__ grad) {
  const Strip s = decode_strip<V, 2>(rows, cols);
  const size_t base = (size_t)s.plane * rows * cols + s.c0;
  const float up = __ldg(g);
  const float cx = L1 ? up * inv_n1 : up * inv_n1 / __ldg(terms);
  const float cy = L1 ? up * inv_n2 : up * inv_n2 / __ldg(terms + 1);
  auto df = [](float a, float b) { return L1 ? sgnf(a - b) : a - b; };
  bool have_prev = s.r0 > 0;
  const size_t poff = base + (size_t)max(s.r0 - 1, 0) * cols;
  RowUV<V> prev = load_uv<V>(fake + poff, real + poff, s.active && have_prev);
  RowUV<V> cur = load_uv<V>(fake + base + (size_t)s.r0 * cols, real + base + (size_t)s.r0 * cols, s.active);
  for (int r = s.r0; r < s.r1; ++r) {
    const bool have_next = r + 1 < rows;
    const size_t off = base + (size_t)r * cols;
    const size_t noff = off + (have_next ? cols : 0);
    const RowUV<V> next = load_uv<V>(fake + noff, real + noff, s.active && have_next);
    const float2 lt = left_uv<V>(cur, fake + off, real + off, s);
    const float2 rt = right_uv<V>(cur, fake + off, real + off, s);
    Row<V> out;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const bool hl = k > 0 || s.has_left, hr = k + 1 < V || s.has_right;
      const float ul = k > 0 ? cur.u[k > 0 ? k - 1 : 0] : lt.x, vl = k > 0 ? cur.v[k > 0 ? k - 1 : 0] : lt.y;
      const float ur = k + 1 < V ? cur.u[k + 1 < V ? k + 1 : 0] : rt.x;
      const float vr = k + 1 < V ? cur.v[k + 1 < V ? k + 1 : 0] : rt.y;
      const float u = cur.u[k], v = cur.v[k];
      const float gux = (hl ? df(u, ul) : 0.0f) + (hr ? df(u, ur) : 0.0f);
      const float gvx = (hl ? df(v, vl) : 0.0f) + (hr ? df(v, vr) : 0.0f);
      const float guy = (have_prev ? df(u, prev.u[k]) : 0.0f) + (have_next ? df(u, next.u[k]) : 0.0f);
      const float gvy = (have_prev ? df(v, prev.v[k]) : 0.0f) + (have_next ? df(v, next.v[k]) : 0.0f);
      const float gu = cx * gux + cy * guy, gv = cx * gvx + cy * gvy;
      out.v[k] = cur.cf[k] * gu - cur.sf[k] * gv;
    }
    store_row<V>(grad + off, out, s.produces);
    prev = cur;
    cur = next;
    have_prev = true;
  }
}

// ---- N1c: point-wise phase losses (loss.py:186-208) -------------------------------------------------------
// focal_sincos_phase_loss: d = |sin f - sin r|, |cos f - cos r| (two channels), loss = mean(d * d / max d)
//                          = sum d^2 / (max d * count) with the focal weight a constant of the graph;
// plain_phase_loss:        mean |f - r|.  One pass gives both: partial[block][3] = { sum d^2, sum |f - r|, max d }.
template <int V>
__global__ void __launch_bounds__(kThreads) phase_point_terms_kernel(const float* __restrict__ fake,
                                                                     const float* __restrict__ real, int rows, int cols,
                                                                     float* __restrict__ partial) {
  const Strip s = decode_strip<V, 0>(rows, cols);
  const size_t base = (size_t)s.plane * rows * cols + s.c0;
  float sum[2] = {0.0f, 0.0f}, mx[1] = {0.0f};
  for (int r = s.r0; r < s.r1; ++r) {
    const size_t off = base + (size_t)r * cols;
    const RowUV<V> cur = load_uv<V>(fake + off, real + off, s.active);
    const Row<V> a = load_row<V>(fake + off, s.active), b = load_row<V>(real + off, s.active);
    if (s.active) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float du = fabsf(cur.u[k]), dv = fabsf(cur.v[k]);
        sum[0] = fmaf(du, du, sum[0]);
        sum[0] = fmaf(dv, dv, sum[0]);
        sum[1] += fabsf(a.v[k] - b.v[k]);
        mx[0] = fmaxf(mx[0], fmaxf(du, dv));
      }
    }
  }
  block_sum_store<2>(sum, partial + (size_t)blockIdx.x * 3);
  block_max_store<1>(mx, partial + (size_t)blockIdx.x * 3 + 2);
}

// terms = { max d, focal_sincos_phase_loss, plain_phase_loss }
__global__ void __launch_bounds__(kFinishThreads) phase_point_finish_kernel(const double* __restrict__ partial,
                                                                           long long nblocks, double n,
                                                                           float* __restrict__ terms) {
  __shared__ double sm[kFinishThreads][2];
  __shared__ float mm[kFinishThreads];
  double s0 = 0, s1 = 0;
  float m = 0.0f;
  for (long long i = threadIdx.x; i < nblocks; i += kFinishThreads) {
    s0 += partial[i * 3];
    s1 += partial[i * 3 + 1];
    m = nanmax(m, (float)partial[i * 3 + 2]);
  }
  sm[threadIdx.x][0] = s0; sm[threadIdx.x][1] = s1; mm[threadIdx.x] = m;
  __syncthreads();
  for (int o = kFinishThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sm[threadIdx.x][0] += sm[threadIdx.x + o][0];
      sm[threadIdx.x][1] += sm[threadIdx.x + o][1];
      mm[threadIdx.x] = nanmax(mm[threadIdx.x], mm[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    terms[0] = mm[0];
    terms[1] = (float)(sm[0][0] / ((double)mm[0] * 2.0 * n));  // both channels count (loss.py:187-192); 0/0 = NaN as there
    terms[2] = (float)(sm[0][1] / n);
  }
}

// FOCAL: d loss / d fake = g * (u cos f - v sin f) / (max d * 2n)   (the weight d / max d is detached)
// else : d loss / d fake = g * sgn(f - r) / n
template <int V, bool FOCAL>
__global__ void __launch_bounds__(kThreads) phase_point_backward_kernel(const float* __restrict__ fake,
                                                                        const float* __restrict__ real,
                                                                        const float* __restrict__ terms,
                                                                        const float* __restrict__ g, int rows, int cols,
                                                                        float inv_n, float* __restrict__ grad) {
  const Strip s = decode_strip<V, 0>(rows, cols);
  const size_t base = (size_t)s.plane * rows * cols + s.c0;
  const float c = FOCAL ? __ldg(g) * 0.5f * inv_n / __ldg(terms) : __ldg(g) * inv_n;
  for (int r = s.r0; r < s.r1; ++r) {
    const size_t off = base + (size_t)r * cols;
    Row<V> out;
    if constexpr (FOCAL) {
      const RowUV<V> cur = load_uv<V>(fake + off, real + off, s.active);
#pragma unroll
      for (int k = 0; k < V; ++k) out.v[k] = c * (cur.u[k] * cur.cf[k] - cur.v[k] * cur.sf[k]);
    } else {
      const Row<V> a = load_row<V>(fake + off, s.active), b = load_row<V>(real + off, s.active);
#pragma unroll
      for (int k = 0; k < V; ++k) out.v[k] = c * sgnf(a.v[k] - b.v[k]);
    }
    store_row<V>(grad + off, out, s.active);
  }
}

// ---- N4: per-plane min/max, normalise, 8-bit pack -------------------------------------------------------
constexpr int kMinMaxChunk = 8192;  // elements per block (256 threads x 8 float4... = 2 float4 per thread x 4)
constexpr int kMinMaxMaxBlocks = 256;

int minmax_blocks(long long plane_elems) {
  long long nb = (plane_elems + kMinMaxChunk - 1) / kMinMaxChunk;
  return (int)(nb < 1 ? 1 : (nb > kMinMaxMaxBlocks ? kMinMaxMaxBlocks : nb));
}

// grid = planes * nb; partial[plane][block][2]
__global__ void __launch_bounds__(256) minmax_partial_kernel(const float* __restrict__ x, long long plane_elems,
                                                             int nb, int vec, float* __restrict__ partial) {
  const long long plane = blockIdx.x / nb;
  const int b = blockIdx.x % nb;
  const float* p = x + plane * plane_elems;
  float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
  if (vec) {
    const long long n4 = plane_elems >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(p);
    for (long long i = (long long)b * 256 + threadIdx.x; i < n4; i += (long long)nb * 256) {
      const float4 t = __ldg(p4 + i);
      lo = nanmin(nanmin(lo, t.x), nanmin(nanmin(t.y, t.z), t.w));
      hi = nanmax(nanmax(hi, t.x), nanmax(nanmax(t.y, t.z), t.w));
    }
  } else {
    for (long long i = (long long)b * 256 + threadIdx.x; i < plane_elems; i += (long long)nb * 256) {
      const float t = __ldg(p + i);
      lo = nanmin(lo, t);
      hi = nanmax(hi, t);
    }
  }
  __shared__ float slo[8], shi[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = nanmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = nanmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    slo[threadIdx.x >> 5] = lo;
    shi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      lo = nanmin(lo, slo[w]);
      hi = nanmax(hi, shi[w]);
    }
    partial[((size_t)plane * nb + b) * 2 + 0] = lo;
    partial[((size_t)plane * nb + b) * 2 + 1] = hi;
  }
}

// one warp per plane
__global__ void __launch_bounds__(32) minmax_finish_kernel(const float* __restrict__ partial, int nb,
                                                           float* __restrict__ minmax) {
  const size_t plane = blockIdx.x;
  float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
  for (int i = threadIdx.x; i < nb; i += 32) {
    lo = nanmin(lo, partial[(plane * nb + i) * 2 + 0]);
    hi = nanmax(hi, partial[(plane * nb + i) * 2 + 1]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = nanmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = nanmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (threadIdx.x == 0) {
    minmax[plane * 2 + 0] = lo;
    minmax[plane * 2 + 1] = hi;
  }
}

__device__ __forceinline__ float normalise(float x, float lo, float range) {
  return __fdiv_rn(__fsub_rn(x, lo), range);
}

// by_max == 0: (x - min) / (max - min)   (util.py:83);   by_max == 1: x / (max * 1.01)   (util.py:64-65)
__global__ void __launch_bounds__(256) normalize_kernel(const float* __restrict__ x, const float* __restrict__ minmax,
                                                        long long plane_elems, int blocks_per_plane, int vec,
                                                        int by_max, float* __restrict__ out) {
  const long long plane = blockIdx.x / blocks_per_plane;
  const int b = blockIdx.x % blocks_per_plane;
  const float lo = by_max ? 0.0f : __ldg(minmax + plane * 2);
  const float range = by_max ? __fmul_rn(__ldg(minmax + plane * 2 + 1), 1.01f)
                             : __fsub_rn(__ldg(minmax + plane * 2 + 1), lo);
  const float* p = x + plane * plane_elems;
  float* q = out + plane * plane_elems;
  if (vec) {
    const long long n4 = plane_elems >> 2;
    for (long long i = (long long)b * 256 + threadIdx.x; i < n4; i += (long long)blocks_per_plane * 256) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
      t.x = normalise(t.x, lo, range); t.y = normalise(t.y, lo, range);
      t.z = normalise(t.z, lo, range); t.w = normalise(t.w, lo, range);
      reinterpret_cast<float4*>(q)[i] = t;
    }
  } else {
    for (long long i = (long long)b * 256 + threadIdx.x; i < plane_elems; i += (long long)blocks_per_plane * 256)
      q[i] = normalise(__ldg(p + i), lo, range);
  }
}

__device__ __forceinline__ unsigned to_u8(float x, float lo, float range, bool norm) {
  const float v = __fmul_rn(norm ? normalise(x, lo, range) : x, 255.0f);
  // numpy's float -> uint8 cast truncates toward zero; values are in [0, 255] here (NaN -> 0)
  return (unsigned)(__float2int_rz(v)) & 0xffu;
}

// one thread = 4 consecutive pixels of one image (PIX4) or one pixel; x [images,3,pix] -> out [images,pix,OC]
template <int OC, bool PIX4>
__global__ void __launch_bounds__(256) pack_u8_kernel(const float* __restrict__ x, const float* __restrict__ minmax,
                                                      long long pix, int blocks_per_image, uint8_t* __restrict__ out) {
  const long long img = blockIdx.x / blocks_per_image;
  const int b = blockIdx.x % blocks_per_image;
  const bool norm = minmax != nullptr;
  float lo[3] = {0.0f, 0.0f, 0.0f}, range[3] = {1.0f, 1.0f, 1.0f};
  if (norm) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      lo[c] = __ldg(minmax + (img * 3 + c) * 2);
      range[c] = __fsub_rn(__ldg(minmax + (img * 3 + c) * 2 + 1), lo[c]);
    }
  }
  const float* p = x + img * 3 * pix;
  uint8_t* q = out + img * pix * OC;
  if (PIX4) {
    const long long n4 = pix >> 2;
    for (long long i = (long long)b * 256 + threadIdx.x; i < n4; i += (long long)blocks_per_image * 256) {
      float ch[3][4];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p + c * pix) + i);
        ch[c][0] = t.x; ch[c][1] = t.y; ch[c][2] = t.z; ch[c][3] = t.w;
      }
      unsigned bytes[4 * OC];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) bytes[k * OC + c] = to_u8(ch[c][k], lo[c], range[c], norm);
        if (OC == 4) bytes[k * OC + 3] = 255u;
      }
      unsigned words[OC];
#pragma unroll
      for (int w = 0; w < OC; ++w)
        words[w] = bytes[4 * w] | (bytes[4 * w + 1] << 8) | (bytes[4 * w + 2] << 16) | (bytes[4 * w + 3] << 24);
      unsigned* dst = reinterpret_cast<unsigned*>(q + i * 4 * OC);  // 4*OC bytes per thread, 4-byte aligned
      if (OC == 4) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(words[0], words[1], words[2], words[3 % OC]);
      } else {
#pragma unroll
        for (int w = 0; w < OC; ++w) dst[w] = words[w];
      }
    }
  } else {
    for (long long i = (long long)b * 256 + threadIdx.x; i < pix; i += (long long)blocks_per_image * 256) {
#pragma unroll
      for (int c = 0; c < 3; ++c) q[i * OC + c] = (uint8_t)to_u8(__ldg(p + c * pix + i), lo[c], range[c], norm);
      if (OC == 4) q[i * OC + 3] = 255;
    }
  }
}

// ---- N2: AP2POH tail -----------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ float2 conv_at(const float2* __restrict__ f, int rows, int cols, int r, int c,
                                          const float* __restrict__ w, float bias) {
  constexpr int H = K / 2;
  float re = 0.0f, im = 0.0f;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const int rr = r + i - H;
    if (rr < 0 || rr >= rows) continue;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int cc = c + j - H;
      if (cc < 0 || cc >= cols) continue;
      const float2 t = __ldg(f + (size_t)rr * cols + cc);
      re = fmaf(w[i * K + j], t.x, re);
      im = fmaf(w[i * K + j], t.y, im);
    }
  }
  return make_float2(re + bias, im + bias);
}

// grid = planes * blocks_per_plane, block b walks rows b, b + blocks_per_plane, ... (no per-pixel divisions);
// PASS 0: partial[plane][block] = max |m|;  PASS 1: writes the POH
constexpr int kTailThreads = 128;
constexpr int kTailMaxBlocks = 1024;
template <int K, int PASS>
__global__ void __launch_bounds__(kTailThreads) ap2poh_tail_kernel(const float2* __restrict__ field,
                                                                   const float* __restrict__ weights,
                                                                   const float* __restrict__ bias, int rows, int cols,
                                                                   int blocks_per_plane, float* __restrict__ partial,
                                                                   const float* __restrict__ plane_max,
                                                                   float* __restrict__ poh) {
  __shared__ float w[K * K];
  __shared__ float red[kTailThreads / 32];
  const long long plane = blockIdx.x / blocks_per_plane;
  const int b = blockIdx.x % blocks_per_plane;
  const int colour = (int)(plane % 3);
  if (threadIdx.x < K * K) w[threadIdx.x] = __ldg(weights + colour * K * K + threadIdx.x);
  __syncthreads();
  const float bs = __ldg(bias + colour);
  const size_t pix = (size_t)rows * cols;
  const float2* f = field + (size_t)plane * pix;
  float scale = 0.0f;
  if (PASS == 1) scale = __fmul_rn(__ldg(plane_max + plane), 1.01f);
  float mx = 0.0f;
  for (int r = b; r < rows; r += blocks_per_plane) {
    for (int c = threadIdx.x; c < cols; c += kTailThreads) {
      const float2 m = conv_at<K>(f, rows, cols, r, c, w, bs);
      const float a = hypotf(m.x, m.y);
      if (PASS == 0) {
        mx = nanmax(mx, a);
      } else {
        const float ac = acosf(__fdiv_rn(a, scale));
        const float p = atan2f(m.y, m.x);
        poh[(size_t)plane * pix + (size_t)r * cols + c] = ((r + c) & 1) ? p - ac : p + ac;
      }
    }
  }
  if (PASS == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = nanmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < kTailThreads / 32; ++k) mx = nanmax(mx, red[k]);
      partial[(size_t)plane * blocks_per_plane + b] = mx;
    }
  }
}

// k = 3 (the reference's kernel_size) on even widths.  The generic kernel above spends its time in L1 wavefronts
// (9 x 8 B per pixel and pass), a row-by-row walk in load latency.  Here a thread owns TWO adjacent columns of a
// 16-row strip and issues all 18 of its 16-byte loads (16 rows + one halo row above and below) before anything is
// computed; horizontal neighbours come from the adjacent lanes by shuffle.  Lanes 0 and 31 of every warp only carry
// the halo columns (a warp loads 64 columns and produces 60), so no lane ever reads a neighbour from memory.
constexpr int kT3Rows = 16, kT3WarpCols = 60, kT3BlockCols = kT3WarpCols * (kThreads / 32);

long long tail3_blocks_per_plane(int rows, int cols) {
  return (long long)div_up(rows, kT3Rows) * div_up(cols, kT3BlockCols);
}

template <int PASS>
__global__ void __launch_bounds__(kThreads) ap2poh_tail3_kernel(const float2* __restrict__ field,
                                                                const float* __restrict__ weights,
                                                                const float* __restrict__ bias, int rows, int cols,
                                                                float* __restrict__ partial,
                                                                const float* __restrict__ plane_max,
                                                                float* __restrict__ poh) {
  const int ncb = div_up(cols, kT3BlockCols), nrs = div_up(rows, kT3Rows);
  long long bb = blockIdx.x;
  const int cb = (int)(bb % ncb);
  bb /= ncb;
  const int rs = (int)(bb % nrs);
  const long long plane = bb / nrs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (cb * (kThreads / 32) + warp) * kT3WarpCols + 2 * (lane - 1);
  const bool in = c0 >= 0 && c0 < cols;  // cols is even: c0 + 1 is inside as well
  const bool produces = in && lane >= 1 && lane <= 30;
  const int r0 = rs * kT3Rows;
  const int colour = (int)(plane % 3);
  float w[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w[i] = __ldg(weights + colour * 9 + i);
  const float bs = __ldg(bias + colour);
  const size_t pix = (size_t)rows * cols;
  const float2* f = field + (size_t)plane * pix;
  float scale = 0.0f;
  if (PASS == 1) scale = __fmul_rn(__ldg(plane_max + plane), 1.01f);

  float4 x[kT3Rows + 2];  // (re, im) of columns c0 and c0 + 1, rows r0 - 1 .. r0 + 16
#pragma unroll
  for (int i = 0; i < kT3Rows + 2; ++i) {
    const int r = r0 - 1 + i;
    x[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (in && r >= 0 && r < rows) x[i] = __ldg(reinterpret_cast<const float4*>(f + (size_t)r * cols + c0));
  }
  float mx[1] = {0.0f};
  float2 lft[3], rgt[3];  // column c0 - 1 and column c0 + 2 of the three rows in the window
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    lft[i + 1] = make_float2(__shfl_up_sync(0xffffffffu, x[i].z, 1), __shfl_up_sync(0xffffffffu, x[i].w, 1));
    rgt[i + 1] = make_float2(__shfl_down_sync(0xffffffffu, x[i].x, 1), __shfl_down_sync(0xffffffffu, x[i].y, 1));
  }
#pragma unroll
  for (int i = 1; i <= kT3Rows; ++i) {
    lft[0] = lft[1]; lft[1] = lft[2];
    rgt[0] = rgt[1]; rgt[1] = rgt[2];
    lft[2] = make_float2(__shfl_up_sync(0xffffffffu, x[i + 1].z, 1), __shfl_up_sync(0xffffffffu, x[i + 1].w, 1));
    rgt[2] = make_float2(__shfl_down_sync(0xffffffffu, x[i + 1].x, 1), __shfl_down_sync(0xffffffffu, x[i + 1].y, 1));
    const int r = r0 + i - 1;
    if (r >= rows) continue;  // uniform over the CTA
    float re0 = 0.0f, im0 = 0.0f, re1 = 0.0f, im1 = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float4 v = x[i - 1 + k];
      const float wl = w[3 * k], wc = w[3 * k + 1], wr = w[3 * k + 2];
      re0 = fmaf(wl, lft[k].x, re0); im0 = fmaf(wl, lft[k].y, im0);
      re0 = fmaf(wc, v.x, re0);      im0 = fmaf(wc, v.y, im0);
      re0 = fmaf(wr, v.z, re0);      im0 = fmaf(wr, v.w, im0);
      re1 = fmaf(wl, v.x, re1);      im1 = fmaf(wl, v.y, im1);
      re1 = fmaf(wc, v.z, re1);      im1 = fmaf(wc, v.w, im1);
      re1 = fmaf(wr, rgt[k].x, re1); im1 = fmaf(wr, rgt[k].y, im1);
    }
    re0 += bs; im0 += bs; re1 += bs; im1 += bs;
    if (produces) {
      const float a0 = hypotf(re0, im0), a1 = hypotf(re1, im1);
      if (PASS == 0) {
        mx[0] = nanmax(mx[0], nanmax(a0, a1));
      } else {
        const float ac0 = acosf(__fdiv_rn(a0, scale)), ac1 = acosf(__fdiv_rn(a1, scale));
        const float p0 = atan2f(im0, re0), p1 = atan2f(im1, re1);
        const bool odd = (r + c0) & 1;  // c0 is even: column c0 + 1 has the other parity
        *reinterpret_cast<float2*>(poh + (size_t)plane * pix + (size_t)r * cols + c0) =
            make_float2(odd ? p0 - ac0 : p0 + ac0, odd ? p1 + ac1 : p1 - ac1);
      }
    }
  }
  if (PASS == 0) block_max_store<1>(mx, partial + blockIdx.x);
}

// one CTA per plane: the per-CTA maxima of the plane -> plane_max[plane]
__global__ void __launch_bounds__(kFinishThreads) plane_max_finish_kernel(const float* __restrict__ partial, int nb,
                                                                         float* __restrict__ plane_max) {
  __shared__ float red[kFinishThreads / 32];
  const size_t plane = blockIdx.x;
  float mx = 0.0f;
  for (int i = threadIdx.x; i < nb; i += kFinishThreads) mx = nanmax(mx, __ldg(partial + plane * nb + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = nanmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < kFinishThreads / 32; ++k) mx = nanmax(mx, red[k]);
    plane_max[plane] = mx;
  }
}

// ---- N2 backward: d POH / d (field, kernels, biases) ----------------------------------------------------------
// POH = p + sigma*acos(a), a = A/(1.01*M), A = |m|, M = max over the plane, p = angle(m), m = conv(x) + b.
// With g = dL/dPOH:  ga = -sigma*g/sqrt(1-a^2);  S = sum ga*A over the plane;
//   gA = ga/(1.01 M) - [A == M] * S/(1.01 M^2)      (torch.max sends the gradient of M to the arg-max pixel)
//   gm = gA*m/A + g*(i m)/A^2                        (abs and angle backward, 0 at m = 0)
//   dL/dx = corr(gm, w),  dL/dw[d] = sum_q Re(conj(gm(q - d)) x(q)),  dL/db = sum Re(gm) + Im(gm).
// The passes recompute m with conv_at<K> (also for the maximum), so A == M is decided on identical bits.
template <int K>
struct TailPix {
  float2 m;
  float A, ga;
};

template <int K>
__device__ __forceinline__ TailPix<K> tail_pixel(const float2* __restrict__ f, int rows, int cols, int r, int c,
                                                 const float* __restrict__ w, float bs, float scale, float g) {
  TailPix<K> o;
  o.m = conv_at<K>(f, rows, cols, r, c, w, bs);
  o.A = hypotf(o.m.x, o.m.y);
  const float a = __fdiv_rn(o.A, scale);
  const float sg = ((r + c) & 1) ? -g : g;
  o.ga = -sg * rsqrtf(fmaxf(1.0f - a * a, 1e-30f));
  return o;
}

// PASS 0: partial[plane][block] = sum ga*A;  PASS 1: gm[plane][r][c] (complex)
template <int K, int PASS>
__global__ void __launch_bounds__(kTailThreads) ap2poh_tail_bwd_kernel(
    const float2* __restrict__ field, const float* __restrict__ weights, const float* __restrict__ bias,
    const float* __restrict__ g_poh, int rows, int cols, int blocks_per_plane, const float* __restrict__ plane_max,
    const float* __restrict__ plane_sum, float* __restrict__ partial, float2* __restrict__ gm) {
  __shared__ float w[K * K];
  __shared__ float red[kTailThreads / 32];
  const long long plane = blockIdx.x / blocks_per_plane;
  const int b = blockIdx.x % blocks_per_plane;
  const int colour = (int)(plane % 3);
  if (threadIdx.x < K * K) w[threadIdx.x] = __ldg(weights + colour * K * K + threadIdx.x);
  __syncthreads();
  const float bs = __ldg(bias + colour);
  const size_t pix = (size_t)rows * cols;
  const float2* f = field + (size_t)plane * pix;
  const float M = __ldg(plane_max + plane);
  const float scale = __fmul_rn(M, 1.01f);
  float peak_term = 0.0f;
  if (PASS == 1) peak_term = __ldg(plane_sum + plane) / (scale * M);
  float acc = 0.0f;
  for (int r = b; r < rows; r += blocks_per_plane) {
    for (int c = threadIdx.x; c < cols; c += kTailThreads) {
      const size_t i = (size_t)plane * pix + (size_t)r * cols + c;
      const float g = __ldg(g_poh + i);
      const TailPix<K> t = tail_pixel<K>(f, rows, cols, r, c, w, bs, scale, g);
      if (PASS == 0) {
        acc = fmaf(t.ga, t.A, acc);
      } else {
        float2 o = make_float2(0.0f, 0.0f);
        if (t.A > 0.0f) {
          float gA = t.ga / scale;
          if (t.A == M) gA -= peak_term;
          const float ia = 1.0f / t.A, gp = g * ia * ia;
          o.x = gA * t.m.x * ia - gp * t.m.y;
          o.y = gA * t.m.y * ia + gp * t.m.x;
        }
        gm[i] = o;
      }
    }
  }
  if (PASS == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < kTailThreads / 32; ++k) acc += red[k];
      partial[(size_t)plane * blocks_per_plane + b] = acc;
    }
  }
}

// one CTA per plane: sum of the plane's per-CTA partials (double accumulation, fixed order)
__global__ void __launch_bounds__(kFinishThreads) plane_sum_finish_kernel(const float* __restrict__ partial, int nb,
                                                                         float* __restrict__ plane_sum) {
  __shared__ double sm[kFinishThreads];
  const size_t plane = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += kFinishThreads) acc += (double)__ldg(partial + plane * nb + i);
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kFinishThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) plane_sum[plane] = (float)sm[0];
}

// grad_field(q) = sum_d w[d] gm(q - d);  partial[plane][block][K*K+1] = { sum_q Re(conj(gm(q - d)) x(q)) ..., sum gm }
template <int K>
__global__ void __launch_bounds__(kTailThreads) ap2poh_tail_bwd_conv_kernel(
    const float2* __restrict__ field, const float* __restrict__ weights, const float2* __restrict__ gm, int rows,
    int cols, int blocks_per_plane, float* __restrict__ partial, float2* __restrict__ grad_field) {
  constexpr int H = K / 2, NW = K * K + 1;
  __shared__ float w[K * K];
  __shared__ float red[kTailThreads / 32][NW];
  const long long plane = blockIdx.x / blocks_per_plane;
  const int b = blockIdx.x % blocks_per_plane;
  const int colour = (int)(plane % 3);
  if (threadIdx.x < K * K) w[threadIdx.x] = __ldg(weights + colour * K * K + threadIdx.x);
  __syncthreads();
  const size_t pix = (size_t)rows * cols;
  const float2* f = field + (size_t)plane * pix;
  const float2* gp = gm + (size_t)plane * pix;
  float acc[NW];
#pragma unroll
  for (int k = 0; k < NW; ++k) acc[k] = 0.0f;
  for (int r = b; r < rows; r += blocks_per_plane) {
    for (int c = threadIdx.x; c < cols; c += kTailThreads) {
      const float2 x = __ldg(f + (size_t)r * cols + c);
      float gre = 0.0f, gim = 0.0f;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const int rr = r - (i - H);  // output pixel p = q - d whose tap d = (i - H, j - H) reads x(q)
        if (rr < 0 || rr >= rows) continue;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int cc = c - (j - H);
          if (cc < 0 || cc >= cols) continue;
          const float2 t = __ldg(gp + (size_t)rr * cols + cc);
          gre = fmaf(w[i * K + j], t.x, gre);
          gim = fmaf(w[i * K + j], t.y, gim);
          acc[i * K + j] = fmaf(t.x, x.x, fmaf(t.y, x.y, acc[i * K + j]));
          if (i == H && j == H) acc[K * K] += t.x + t.y;
        }
      }
      grad_field[(size_t)plane * pix + (size_t)r * cols + c] = make_float2(gre, gim);
    }
  }
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    float x = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = x;
  }
  __syncthreads();
  if (threadIdx.x < NW) {
    float x = 0.0f;
    for (int wv = 0; wv < kTailThreads / 32; ++wv) x += red[wv][threadIdx.x];
    partial[((size_t)plane * blocks_per_plane + b) * NW + threadIdx.x] = x;
  }
}

// one CTA per colour: the partials of all planes of that colour -> grad_weights[colour][K*K], grad_bias[colour]
__global__ void __launch_bounds__(kFinishThreads) tail_weight_finish_kernel(const float* __restrict__ partial,
                                                                           long long planes, int blocks_per_plane,
                                                                           int nw, float* __restrict__ grad_weights,
                                                                           float* __restrict__ grad_bias) {
  __shared__ double sm[kFinishThreads];
  const int colour = blockIdx.x;
  for (int k = 0; k < nw; ++k) {
    double acc = 0.0;
    for (long long plane = colour; plane < planes; plane += 3)
      for (int i = threadIdx.x; i < blocks_per_plane; i += kFinishThreads)
        acc += (double)__ldg(partial + ((size_t)plane * blocks_per_plane + i) * nw + k);
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = kFinishThreads / 2; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      if (k < nw - 1) grad_weights[colour * (nw - 1) + k] = (float)sm[0];
      else grad_bias[colour] = (float)sm[0];
    }
    __syncthreads();
  }
}

// ---- N3 device side ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) assemble_rgbd_kernel(const float* __restrict__ img,
                                                            const float* __restrict__ depth, int depth_planes,
                                                            long long plane_elems, long long total, int vec,
                                                            float* __restrict__ out) {
  // total = n * 4 * plane_elems (in units of `vec` floats when vec == 4)
  const long long pe = vec == 4 ? plane_elems >> 2 : plane_elems;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long item = i / (4 * pe), rem = i % (4 * pe);
    const long long ch = rem / pe, e = rem % pe;
    const float* src = ch < 3 ? img + (item * 3 + ch) * plane_elems : depth + item * depth_planes * plane_elems;
    if (vec == 4)
      reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(src) + e);
    else
      out[i] = __ldg(src + e);
  }
}

__global__ void __launch_bounds__(256) scale_two_pi_kernel(const float* __restrict__ x, long long n,
                                                           float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    out[i] = __fmul_rn(6.2831854820251465f, __ldg(x + i));
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int grid_for(long long work_items, int per_block) {
  long long nb = (work_items + per_block - 1) / per_block;
  const long long cap = 148LL * 32;  // a few waves of the 148 SMs; grid-stride loops cover the rest
  if (nb > cap) nb = cap;
  return (int)(nb < 1 ? 1 : nb);
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" int lhg_next_version(void) { return LHG_NEXT_VERSION; }
extern "C" const char* lhg_next_last_error(void) { return g_err; }
extern "C" long long lhg_next_launch_count(void) { return g_launches.load(); }

extern "C" size_t lhg_next_partial_floats(long long planes, int rows, int cols) {
  if (planes <= 0 || rows <= 0 || cols <= 0) return 0;
  const size_t amp = (size_t)strip_blocks(planes, rows, cols, 1, 0) * 5, focal = (size_t)strip_blocks(planes, rows, cols, 1, 1) * 6;
  const size_t strips = (amp > focal ? amp : focal) + kStageFloats;
  const size_t mm = (size_t)planes * minmax_blocks((long long)rows * cols) * 2;
  const size_t tail = (size_t)planes * (rows < kTailMaxBlocks ? rows : kTailMaxBlocks);
  return strips > mm ? (strips > tail ? strips : tail) : (mm > tail ? mm : tail);
}

static int check_planes(const char* what, long long planes, int rows, int cols) {
  if (planes < 0 || rows <= 0 || cols <= 0) return fail(LHG_EINVAL, "%s: bad shape [%lld,%d,%d]", what, planes, rows, cols);
  if (planes > 0 && strip_blocks(planes, rows, cols, 1) > 0x7fffffffLL)
    return fail(LHG_EINVAL, "%s: too many strips for one launch", what);
  return LHG_OK;
}

extern "C" int lhg_amp_loss_terms(const float* hat, const float* target, long long planes, int rows, int cols,
                                  float alpha, float* partial, size_t partial_floats, float* terms,
                                  lhg_stream stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_planes("lhg_amp_loss_terms", planes, rows, cols)) return rc;
  if (!hat || !partial || !terms) return fail(LHG_EINVAL, "lhg_amp_loss_terms: null pointer");
  const bool v4 = cols % 4 == 0 && aligned16(hat) && (!target || aligned16(target));
  const long long nblocks = strip_blocks(planes, rows, cols, v4 ? 4 : 1);
  if ((size_t)nblocks * 5 + kStageFloats > partial_floats)
    return fail(LHG_EWORKSPACE, "lhg_amp_loss_terms: partial buffer holds %zu floats, need %lld", partial_floats,
                nblocks * 5 + kStageFloats);
  if (reinterpret_cast<uintptr_t>(partial) & 7u) return fail(LHG_EINVAL, "lhg_amp_loss_terms: partial must be 8-byte aligned");
  double* stage = reinterpret_cast<double*>(partial);
  partial += kStageFloats;
  if (nblocks > 0) {
    if (v4) {
      if (target) amp_terms_kernel<4, true><<<(unsigned)nblocks, kThreads, 0, stream>>>(hat, target, rows, cols, partial);
      else amp_terms_kernel<4, false><<<(unsigned)nblocks, kThreads, 0, stream>>>(hat, target, rows, cols, partial);
    } else {
      if (target) amp_terms_kernel<1, true><<<(unsigned)nblocks, kThreads, 0, stream>>>(hat, target, rows, cols, partial);
      else amp_terms_kernel<1, false><<<(unsigned)nblocks, kThreads, 0, stream>>>(hat, target, rows, cols, partial);
    }
    if (int rc = launched("amp_terms_kernel")) return rc;
  }
  const double n = (double)planes * rows * cols, n1 = (double)planes * rows * (cols - 1),
               n2 = (double)planes * (rows - 1) * cols;
  const int sb = stage_blocks(nblocks);
  stage_partials_kernel<5, 5><<<sb, kFinishThreads, 0, stream>>>(partial, nblocks, stage);
  if (int rc = launched("stage_partials_kernel")) return rc;
  amp_terms_finish_kernel<<<1, kFinishThreads, 0, stream>>>(stage, sb, n, n1, n2, alpha, target ? 1 : 0, terms);
  return launched("amp_terms_finish_kernel");
}

extern "C" int lhg_amp_loss_backward(const float* hat, const float* target, const float* g, const float* terms,
                                     float alpha, long long planes, int rows, int cols, float* grad_hat,
                                     lhg_stream stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_planes("lhg_amp_loss_backward", planes, rows, cols)) return rc;
  if (!hat || !g || !grad_hat || (target && !terms)) return fail(LHG_EINVAL, "lhg_amp_loss_backward: null pointer");
  if (planes == 0) return LHG_OK;
  const bool v4 = cols % 4 == 0 && aligned16(hat) && aligned16(grad_hat) && (!target || aligned16(target));
  const long long nblocks = strip_blocks(planes, rows, cols, v4 ? 4 : 1);
  const double n = (double)planes * rows * cols, n1 = (double)planes * rows * (cols - 1),
               n2 = (double)planes * (rows - 1) * cols;
  const float a = (float)(2.0 / n), i1 = n1 > 0 ? (float)(1.0 / n1) : 0.0f, i2 = n2 > 0 ? (float)(1.0 / n2) : 0.0f;
  if (v4) {
    if (target) amp_backward_kernel<4, true><<<(unsigned)nblocks, kThreads, 0, stream>>>(hat, target, g, terms, alpha, rows, cols, a, i1, i2, grad_hat);
    else amp_backward_kernel<4, false><<<(unsigned)nblocks, kThreads, 0, stream>>>(hat, target, g, terms, alpha, rows, cols, a, i1, i2, grad_hat);
  } else {
    if (target) amp_backward_kernel<1, true><<<(unsigned)nblocks, kThreads, 0, stream>>>(hat, target, g, terms, alpha, rows, cols, a, i1, i2, grad_hat);
    else amp_backward_kernel<1, false><<<(unsigned)nblocks, kThreads, 0, stream>>>(hat, target, g, terms, alpha, rows, cols, a, i1, i2, grad_hat);
  }
  return launched("amp_backward_kernel");
}

extern "C" int lhg_focal_phase_loss_terms(const float* fake_phase, const float* real_phase, long long planes,
                                          int rows, int cols, float* partial, size_t partial_floats, float* terms,
                                          lhg_stream stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_planes("lhg_focal_phase_loss_terms", planes, rows, cols)) return rc;
  if (!fake_phase || !real_phase || !partial || !terms) return fail(LHG_EINVAL, "lhg_focal_phase_loss_terms: null pointer");
  const bool v4 = cols % 4 == 0 && aligned16(fake_phase) && aligned16(real_phase);
  const long long nblocks = strip_blocks(planes, rows, cols, v4 ? 4 : 1, 1);
  if ((size_t)nblocks * 6 + kStageFloats > partial_floats)
    return fail(LHG_EWORKSPACE, "lhg_focal_phase_loss_terms: partial buffer holds %zu floats, need %lld",
                partial_floats, nblocks * 6 + kStageFloats);
  if (reinterpret_cast<uintptr_t>(partial) & 7u) return fail(LHG_EINVAL, "lhg_focal_phase_loss_terms: partial must be 8-byte aligned");
  double* stage = reinterpret_cast<double*>(partial);
  partial += kStageFloats;
  if (nblocks > 0) {
    if (v4) focal_terms_kernel<4><<<(unsigned)nblocks, kThreads, 0, stream>>>(fake_phase, real_phase, rows, cols, partial);
    else focal_terms_kernel<1><<<(unsigned)nblocks, kThreads, 0, stream>>>(fake_phase, real_phase, rows, cols, partial);
    if (int rc = launched("focal_terms_kernel")) return rc;
  }
  // both channels (sin, cos) count: the reference concatenates them along dim 1 (loss.py:136-141)
  const double n1 = 2.0 * (double)planes * rows * (cols - 1), n2 = 2.0 * (double)planes * (rows - 1) * cols;
  const int sb = stage_blocks(nblocks);
  stage_partials_kernel<6, 4><<<sb, kFinishThreads, 0, stream>>>(partial, nblocks, stage);
  if (int rc = launched("stage_partials_kernel")) return rc;
  focal_terms_finish_kernel<<<1, kFinishThreads, 0, stream>>>(stage, sb, n1, n2, terms);
  return launched("focal_terms_finish_kernel");
}

extern "C" int lhg_focal_phase_loss_backward(const float* fake_phase, const float* real_phase, const float* terms,
                                             const float* g, long long planes, int rows, int cols,
                                             float* grad_fake, lhg_stream stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_planes("lhg_focal_phase_loss_backward", planes, rows, cols)) return rc;
  if (!fake_phase || !real_phase || !terms || !g || !grad_fake)
    return fail(LHG_EINVAL, "lhg_focal_phase_loss_backward: null pointer");
  if (planes == 0) return LHG_OK;
  const bool v4 = cols % 4 == 0 && aligned16(fake_phase) && aligned16(real_phase) && aligned16(grad_fake);
  const long long nblocks = strip_blocks(planes, rows, cols, v4 ? 4 : 1, 2);
  const double n1 = 2.0 * (double)planes * rows * (cols - 1), n2 = 2.0 * (double)planes * (rows - 1) * cols;
  const float i1 = n1 > 0 ? (float)(1.0 / n1) : 0.0f, i2 = n2 > 0 ? (float)(1.0 / n2) : 0.0f;
  if (v4) focal_backward_kernel<4><<<(unsigned)nblocks, kThreads, 0, stream>>>(fake_phase, real_phase, terms, g, rows, cols, i1, i2, grad_fake);
  else focal_backward_kernel<1><<<(unsigned)nblocks, kThreads, 0, stream>>>(fake_phase, real_phase, terms, g, rows, cols, i1, i2, grad_fake);
  return launched("focal_backward_kernel");
}

extern "C" int lhg_phase_gradient_loss_backward(const float* fake_phase, const float* real_phase, const float* g,
                                                long long planes, int rows, int cols, float* grad_fake,
                                                lhg_stream stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_planes("lhg_phase_gradient_loss_backward", planes, rows, cols)) return rc;
  if (!fake_phase || !real_phase || !g || !grad_fake)
    return fail(LHG_EINVAL, "lhg_phase_gradient_loss_backward: null pointer");
  if (planes == 0) return LHG_OK;
  const bool v4 = cols % 4 == 0 && aligned16(fake_phase) && aligned16(real_phase) && aligned16(grad_fake);
  const long long nblocks = strip_blocks(planes, rows, cols, v4 ? 4 : 1, 2);
  const double n1 = 2.0 * (double)planes * rows * (cols - 1), n2 = 2.0 * (double)planes * (rows - 1) * cols;
  const float i1 = n1 > 0 ? (float)(1.0 / n1) : 0.0f, i2 = n2 > 0 ? (float)(1.0 / n2) : 0.0f;
  if (v4) focal_backward_kernel<4, true><<<(unsigned)nblocks, kThreads, 0, stream>>>(fake_phase, real_phase, nullptr, g, rows, cols, i1, i2, grad_fake);
  else focal_backward_kernel<1, true><<<(unsigned)nblocks, kThreads, 0, stream>>>(fake_phase, real_phase, nullptr, g, rows, cols, i1, i2, grad_fake);
  return launched("focal_backward_kernel<L1>");
}

extern "C" int lhg_phase_point_loss_terms(const float* fake_phase, const float* real_phase, long long planes,
                                         int rows, int cols, float* partial, size_t partial_floats, float* terms,
                                         lhg_stream stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_planes("lhg_phase_point_loss_terms", planes, rows, cols)) return rc;
  if (!fake_phase || !real_phase || !partial || !terms) return fail(LHG_EINVAL, "lhg_phase_point_loss_terms: null pointer");
  const bool v4 = cols % 4 == 0 && aligned16(fake_phase) && aligned16(real_phase);
  const long long nblocks = strip_blocks(planes, rows, cols, v4 ? 4 : 1, 0);
  if ((size_t)nblocks * 3 + kStageFloats > partial_floats)
    return fail(LHG_EWORKSPACE, "lhg_phase_point_loss_terms: partial buffer holds %zu floats, need %lld",
                partial_floats, nblocks * 3 + kStageFloats);
  if (reinterpret_cast<uintptr_t>(partial) & 7u) return fail(LHG_EINVAL, "lhg_phase_point_loss_terms: partial must be 8-byte aligned");
  double* stage = reinterpret_cast<double*>(partial);
  partial += kStageFloats;
  if (nblocks > 0) {
    if (v4) phase_point_terms_kernel<4><<<(unsigned)nblocks, kThreads, 0, stream>>>(fake_phase, real_phase, rows, cols, partial);
    else phase_point_terms_kernel<1><<<(unsigned)nblocks, kThreads, 0, stream>>>(fake_phase, real_phase, rows, cols, partial);
    if (int rc = launched("phase_point_terms_kernel")) return rc;
  }
  const int sb = stage_blocks(nblocks);
  stage_partials_kernel<3, 2><<<sb, kFinishThreads, 0, stream>>>(partial, nblocks, stage);
  if (int rc = launched("stage_partials_kernel")) return rc;
  phase_point_finish_kernel<<<1, kFinishThreads, 0, stream>>>(stage, sb, (double)planes * rows * cols, terms);
  return launched("phase_point_finish_kernel");
}

extern "C" int lhg_phase_point_loss_backward(const float* fake_phase, const float* real_phase, const float* terms,
                                             const float* g, int focal, long long planes, int rows, int cols,
                                             float* grad_fake, lhg_stream stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_planes("lhg_phase_point_loss_backward", planes, rows, cols)) return rc;
  if (!fake_phase || !real_phase || !g || !grad_fake || (focal && !terms))
    return fail(LHG_EINVAL, "lhg_phase_point_loss_backward: null pointer");
  if (planes == 0) return LHG_OK;
  const bool v4 = cols % 4 == 0 && aligned16(fake_phase) && aligned16(real_phase) && aligned16(grad_fake);
  const long long nblocks = strip_blocks(planes, rows, cols, v4 ? 4 : 1, 0);
  const float inv_n = (float)(1.0 / ((double)planes * rows * cols));
  if (focal) {
    if (v4) phase_point_backward_kernel<4, true><<<(unsigned)nblocks, kThreads, 0, stream>>>(fake_phase, real_phase, terms, g, rows, cols, inv_n, grad_fake);
    else phase_point_backward_kernel<1, true><<<(unsigned)nblocks, kThreads, 0, stream>>>(fake_phase, real_phase, terms, g, rows, cols, inv_n, grad_fake);
  } else {
    if (v4) phase_point_backward_kernel<4, false><<<(unsigned)nblocks, kThreads, 0, stream>>>(fake_phase, real_phase, terms, g, rows, cols, inv_n, grad_fake);
    else phase_point_backward_kernel<1, false><<<(unsigned)nblocks, kThreads, 0, stream>>>(fake_phase, real_phase, terms, g, rows, cols, inv_n, grad_fake);
  }
  return launched("phase_point_backward_kernel");
}

extern "C" int lhg_plane_minmax(const float* x, long long planes, long long plane_elems, float* partial,
                                size_t partial_floats, float* minmax, lhg_stream stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (planes < 0 || plane_elems <= 0) return fail(LHG_EINVAL, "lhg_plane_minmax: bad shape [%lld,%lld]", planes, plane_elems);
  if (planes == 0) return LHG_OK;
  if (!x || !partial || !minmax) return fail(LHG_EINVAL, "lhg_plane_minmax: null pointer");
  const int nb = minmax_blocks(plane_elems);
  if ((size_t)planes * nb * 2 > partial_floats)
    return fail(LHG_EWORKSPACE, "lhg_plane_minmax: partial buffer holds %zu floats, need %lld", partial_floats,
                planes * nb * 2);
  if (planes * nb > 0x7fffffffLL) return fail(LHG_EINVAL, "lhg_plane_minmax: too many blocks");
  const int vec = plane_elems % 4 == 0 && aligned16(x);
  minmax_partial_kernel<<<(unsigned)(planes * nb), 256, 0, stream>>>(x, plane_elems, nb, vec, partial);
  if (int rc = launched("minmax_partial_kernel")) return rc;
  minmax_finish_kernel<<<(unsigned)planes, 32, 0, stream>>>(partial, nb, minmax);
  return launched("minmax_finish_kernel");
}

static int normalize_planes(const char* what, const float* x, const float* minmax, long long planes,
                            long long plane_elems, int by_max, float* out, cudaStream_t stream) {
  if (planes < 0 || plane_elems <= 0) return fail(LHG_EINVAL, "%s: bad shape", what);
  if (planes == 0) return LHG_OK;
  if (!x || !minmax || !out) return fail(LHG_EINVAL, "%s: null pointer", what);
  const int nb = minmax_blocks(plane_elems);
  const int vec = plane_elems % 4 == 0 && aligned16(x) && aligned16(out);
  normalize_kernel<<<(unsigned)(planes * nb), 256, 0, stream>>>(x, minmax, plane_elems, nb, vec, by_max, out);
  return launched("normalize_kernel");
}

extern "C" int lhg_normalize_planes(const float* x, const float* minmax, long long planes, long long plane_elems,
                                    float* out, lhg_stream stream) {
  return normalize_planes("lhg_normalize_planes", x, minmax, planes, plane_elems, 0, out, (cudaStream_t)stream);
}

extern "C" int lhg_amplitude_normalize(const float* x, const float* minmax, long long planes, long long plane_elems,
                                       float* out, lhg_stream stream) {
  return normalize_planes("lhg_amplitude_normalize", x, minmax, planes, plane_elems, 1, out, (cudaStream_t)stream);
}

extern "C" int lhg_pack_rgb_u8(const float* x, const float* minmax, long long images, int rows, int cols,
                               int out_channels, uint8_t* out, lhg_stream stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (images < 0 || rows <= 0 || cols <= 0) return fail(LHG_EINVAL, "lhg_pack_rgb_u8: bad shape");
  if (out_channels != 3 && out_channels != 4) return fail(LHG_EINVAL, "lhg_pack_rgb_u8: out_channels must be 3 or 4");
  if (images == 0) return LHG_OK;
  if (!x || !out) return fail(LHG_EINVAL, "lhg_pack_rgb_u8: null pointer");
  const long long pix = (long long)rows * cols;
  const int nb = minmax_blocks(pix);
  const bool p4 = pix % 4 == 0 && aligned16(x) && aligned16(out);
  const unsigned grid = (unsigned)(images * nb);
  if (out_channels == 4) {
    if (p4) pack_u8_kernel<4, true><<<grid, 256, 0, stream>>>(x, minmax, pix, nb, out);
    else pack_u8_kernel<4, false><<<grid, 256, 0, stream>>>(x, minmax, pix, nb, out);
  } else {
    if (p4) pack_u8_kernel<3, true><<<grid, 256, 0, stream>>>(x, minmax, pix, nb, out);
    else pack_u8_kernel<3, false><<<grid, 256, 0, stream>>>(x, minmax, pix, nb, out);
  }
  return launched("pack_u8_kernel");
}

template <int K>
static int launch_tail(const float2* field, const float* weights, const float* bias, long long planes, int rows,
                       int cols, int nb, float* partial, float* plane_max, float* poh, cudaStream_t stream) {
  ap2poh_tail_kernel<K, 0><<<(unsigned)(planes * nb), kTailThreads, 0, stream>>>(field, weights, bias, rows, cols, nb, partial, nullptr, nullptr);
  if (int rc = launched("ap2poh_tail_kernel<max>")) return rc;
  plane_max_finish_kernel<<<(unsigned)planes, kFinishThreads, 0, stream>>>(partial, nb, plane_max);
  if (int rc = launched("plane_max_finish_kernel")) return rc;
  ap2poh_tail_kernel<K, 1><<<(unsigned)(planes * nb), kTailThreads, 0, stream>>>(field, weights, bias, rows, cols, nb, nullptr, plane_max, poh);
  return launched("ap2poh_tail_kernel<poh>");
}

extern "C" int lhg_ap2poh_tail(const void* field, const float* weights, const float* bias, int ksize,
                               long long planes, int rows, int cols, float* partial, size_t partial_floats,
                               float* plane_max, float* poh, lhg_stream stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_planes("lhg_ap2poh_tail", planes, rows, cols)) return rc;
  if (planes % 3 != 0) return fail(LHG_EINVAL, "lhg_ap2poh_tail: planes must be a multiple of 3 colours");
  if (ksize != 1 && ksize != 3 && ksize != 5 && ksize != 7)
    return fail(LHG_EINVAL, "lhg_ap2poh_tail: kernel size %d not in {1,3,5,7}", ksize);
  if (planes == 0) return LHG_OK;
  if (!field || !weights || !bias || !partial || !plane_max || !poh)
    return fail(LHG_EINVAL, "lhg_ap2poh_tail: null pointer");
  const int nb = rows < kTailMaxBlocks ? rows : kTailMaxBlocks;
  if ((size_t)planes * nb > partial_floats)
    return fail(LHG_EWORKSPACE, "lhg_ap2poh_tail: partial buffer holds %zu floats, need %lld", partial_floats,
                planes * nb);
  const float2* f = (const float2*)field;
  if (ksize == 3 && cols % 2 == 0 && aligned16(field) && (reinterpret_cast<uintptr_t>(poh) & 7u) == 0) {
    // one partial per CTA, the CTAs of a plane are consecutive
    const long long per_plane = tail3_blocks_per_plane(rows, cols);
    if ((size_t)(planes * per_plane) > partial_floats)
      return fail(LHG_EWORKSPACE, "lhg_ap2poh_tail: partial buffer holds %zu floats, need %lld", partial_floats,
                  planes * per_plane);
    const unsigned grid = (unsigned)(planes * per_plane);
    ap2poh_tail3_kernel<0><<<grid, kThreads, 0, stream>>>(f, weights, bias, rows, cols, partial, nullptr, nullptr);
    if (int rc = launched("ap2poh_tail3_kernel<max>")) return rc;
    plane_max_finish_kernel<<<(unsigned)planes, kFinishThreads, 0, stream>>>(partial, (int)per_plane, plane_max);
    if (int rc = launched("plane_max_finish_kernel")) return rc;
    ap2poh_tail3_kernel<1><<<grid, kThreads, 0, stream>>>(f, weights, bias, rows, cols, nullptr, plane_max, poh);
    return launched("ap2poh_tail3_kernel<poh>");
  }
  switch (ksize) {
    case 1: return launch_tail<1>(f, weights, bias, planes, rows, cols, nb, partial, plane_max, poh, stream);
    case 3: return launch_tail<3>(f, weights, bias, planes, rows, cols, nb, partial, plane_max, poh, stream);
    case 5: return launch_tail<5>(f, weights, bias, planes, rows, cols, nb, partial, plane_max, poh, stream);
    default: return launch_tail<7>(f, weights, bias, planes, rows, cols, nb, partial, plane_max, poh, stream);
  }
}

template <int K>
static int launch_tail_bwd(const float2* field, const float* weights, const float* bias, const float* g_poh,
                           long long planes, int rows, int cols, int nb, float* partial, float* plane_max,
                           float* plane_sum, float2* gm, float2* grad_field, float* grad_weights, float* grad_bias,
                           cudaStream_t stream) {
  const unsigned grid = (unsigned)(planes * nb);
  ap2poh_tail_kernel<K, 0><<<grid, kTailThreads, 0, stream>>>(field, weights, bias, rows, cols, nb, partial, nullptr, nullptr);
  if (int rc = launched("ap2poh_tail_kernel<max>")) return rc;
  plane_max_finish_kernel<<<(unsigned)planes, kFinishThreads, 0, stream>>>(partial, nb, plane_max);
  if (int rc = launched("plane_max_finish_kernel")) return rc;
  ap2poh_tail_bwd_kernel<K, 0><<<grid, kTailThreads, 0, stream>>>(field, weights, bias, g_poh, rows, cols, nb, plane_max, nullptr, partial, nullptr);
  if (int rc = launched("ap2poh_tail_bwd_kernel<sum>")) return rc;
  plane_sum_finish_kernel<<<(unsigned)planes, kFinishThreads, 0, stream>>>(partial, nb, plane_sum);
  if (int rc = launched("plane_sum_finish_kernel")) return rc;
  ap2poh_tail_bwd_kernel<K, 1><<<grid, kTailThreads, 0, stream>>>(field, weights, bias, g_poh, rows, cols, nb, plane_max, plane_sum, nullptr, gm);
  if (int rc = launched("ap2poh_tail_bwd_kernel<gm>")) return rc;
  ap2poh_tail_bwd_conv_kernel<K><<<grid, kTailThreads, 0, stream>>>(field, weights, gm, rows, cols, nb, partial, grad_field);
  if (int rc = launched("ap2poh_tail_bwd_conv_kernel")) return rc;
  tail_weight_finish_kernel<<<3, kFinishThreads, 0, stream>>>(partial, planes, nb, K * K + 1, grad_weights, grad_bias);
  return launched("tail_weight_finish_kernel");
}

extern "C" size_t lhg_ap2poh_tail_backward_floats(int ksize, long long planes, int rows, int cols) {
  if (planes <= 0 || rows <= 0 || cols <= 0 || ksize < 1) return 0;
  const size_t nb = rows < kTailMaxBlocks ? rows : kTailMaxBlocks;
  // per-CTA partials (K*K+1 each) + plane_max + plane_sum + the complex gm plane stack
  return (size_t)planes * nb * (ksize * ksize + 1) + 2 * (size_t)planes + 2 + 2 * (size_t)planes * rows * cols;
}

extern "C" int lhg_ap2poh_tail_backward(const void* field, const float* weights, const float* bias, int ksize,
                                        const float* g_poh, long long planes, int rows, int cols, float* scratch,
                                        size_t scratch_floats, void* grad_field, float* grad_weights,
                                        float* grad_bias, lhg_stream stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_planes("lhg_ap2poh_tail_backward", planes, rows, cols)) return rc;
  if (planes % 3 != 0) return fail(LHG_EINVAL, "lhg_ap2poh_tail_backward: planes must be a multiple of 3 colours");
  if (ksize != 1 && ksize != 3 && ksize != 5 && ksize != 7)
    return fail(LHG_EINVAL, "lhg_ap2poh_tail_backward: kernel size %d not in {1,3,5,7}", ksize);
  if (!field || !weights || !bias || !g_poh || !scratch || !grad_field || !grad_weights || !grad_bias)
    return fail(LHG_EINVAL, "lhg_ap2poh_tail_backward: null pointer");
  if (planes == 0) {
    cudaMemsetAsync(grad_weights, 0, sizeof(float) * 3 * ksize * ksize, stream);
    cudaMemsetAsync(grad_bias, 0, sizeof(float) * 3, stream);
    return LHG_OK;
  }
  const size_t need = lhg_ap2poh_tail_backward_floats(ksize, planes, rows, cols);
  if (scratch_floats < need)
    return fail(LHG_EWORKSPACE, "lhg_ap2poh_tail_backward: scratch holds %zu floats, need %zu", scratch_floats, need);
  if (reinterpret_cast<uintptr_t>(scratch) & 7u) return fail(LHG_EINVAL, "lhg_ap2poh_tail_backward: scratch must be 8-byte aligned");
  const int nb = rows < kTailMaxBlocks ? rows : kTailMaxBlocks;
  float2* gm = reinterpret_cast<float2*>(scratch);  // first: keeps the complex stack 8-byte aligned
  float* plane_max = scratch + 2 * (size_t)planes * rows * cols;
  float* plane_sum = plane_max + planes;
  float* partial = plane_sum + planes + (planes & 1 ? 1 : 0);
  const float2* f = (const float2*)field;
  float2* gf = (float2*)grad_field;
  switch (ksize) {
    case 1: return launch_tail_bwd<1>(f, weights, bias, g_poh, planes, rows, cols, nb, partial, plane_max, plane_sum, gm, gf, grad_weights, grad_bias, stream);
    case 3: return launch_tail_bwd<3>(f, weights, bias, g_poh, planes, rows, cols, nb, partial, plane_max, plane_sum, gm, gf, grad_weights, grad_bias, stream);
    case 5: return launch_tail_bwd<5>(f, weights, bias, g_poh, planes, rows, cols, nb, partial, plane_max, plane_sum, gm, gf, grad_weights, grad_bias, stream);
    default: return launch_tail_bwd<7>(f, weights, bias, g_poh, planes, rows, cols, nb, partial, plane_max, plane_sum, gm, gf, grad_weights, grad_bias, stream);
  }
}

extern "C" int lhg_bin_gather(const void* base, long long n_items, size_t item_bytes, size_t copy_bytes,
                              const int64_t* idx, int n, void* dst, int threads) {
  if (n < 0 || n_items < 0 || copy_bytes > item_bytes) return fail(LHG_EINVAL, "lhg_bin_gather: bad sizes");
  if (n == 0 || copy_bytes == 0) return LHG_OK;
  if (!base || !idx || !dst) return fail(LHG_EINVAL, "lhg_bin_gather: null pointer");
  for (int i = 0; i < n; ++i)
    if (idx[i] < 0 || idx[i] >= n_items)
      return fail(LHG_EINVAL, "lhg_bin_gather: index %lld out of range [0,%lld)", (long long)idx[i], n_items);
  if (threads <= 0) {
    const unsigned hc = std::thread::hardware_concurrency();
    threads = (int)(hc ? (hc > 16 ? 16 : hc) : 4);
  }
  // one item is one contiguous memcpy (a page-cache read of the memory-mapped file); small jobs stay on the caller
  if ((size_t)n * copy_bytes < (1u << 20) || threads == 1 || n == 1) threads = 1;
  if (threads > n) threads = n;
  auto work = [=](int t) {
    for (int i = t; i < n; i += threads)
      memcpy((char*)dst + (size_t)i * copy_bytes, (const char*)base + (size_t)idx[i] * item_bytes, copy_bytes);
  };
  if (threads == 1) {
    work(0);
    return LHG_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(threads - 1);
  for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  return LHG_OK;
}

extern "C" int lhg_assemble_rgbd(const float* img, const float* depth, int depth_planes, long long n,
                                 long long plane_elems, float* out, lhg_stream stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || plane_elems <= 0 || depth_planes < 1) return fail(LHG_EINVAL, "lhg_assemble_rgbd: bad shape");
  if (n == 0) return LHG_OK;
  if (!img || !depth || !out) return fail(LHG_EINVAL, "lhg_assemble_rgbd: null pointer");
  const int vec = (plane_elems % 4 == 0 && aligned16(img) && aligned16(depth) && aligned16(out)) ? 4 : 1;
  const long long total = n * 4 * (plane_elems / vec);
  assemble_rgbd_kernel<<<grid_for(total, 256 * 4), 256, 0, stream>>>(img, depth, depth_planes, plane_elems, total, vec, out);
  return launched("assemble_rgbd_kernel");
}

extern "C" int lhg_scale_two_pi(const float* x, long long n, float* out, lhg_stream stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0) return fail(LHG_EINVAL, "lhg_scale_two_pi: bad size");
  if (n == 0) return LHG_OK;
  if (!x || !out) return fail(LHG_EINVAL, "lhg_scale_two_pi: null pointer");
  scale_two_pi_kernel<<<grid_for(n, 256 * 4), 256, 0, stream>>>(x, n, out);
  return launched("scale_two_pi_kernel");
}
