// asm_b200.cu -- C ABI + generic kernels of the B200-native angular-spectrum propagation path.
//
// Pipeline of one asm_propagate call (DESIGN.md "Kernels"):
//   K1 row_forward_kernel : prologue (phase -> phasor, cotangent, ...) + implicit zero-pad +
//                           row FFT of the R non-pad rows only          -> W1 [P_in, R, Cp]
//   K2 column_kernel      : T adjacent columns per CTA in shared memory: column FFT, then per
//                           depth  x filter (H generated on the fly, fp32-faithful) + column IFFT,
//                           writing only the R crop rows                -> W2 [P_out, R, Cp]
//                           (or the reverse loop for the adjoint: sum over depth, one IFFT)
//   K3 row_inverse_kernel : row IFFT + crop + epilogue (abs / angle / complex / |.|^2 /
//                           phase-gradient, optional fused amplitude-L2 partial sums)
// Spectrum-out methods stop after K2, spectrum-in methods start at K2.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <new>
#include <vector>

#include "common.cuh"

using namespace asmb;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return fail(ASM_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                  __LINE__);                                                                 \
  } while (0)

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
struct FftHost {
  Fft1d dev{};  // device-visible copy (pointers are device pointers)
  void* tw = nullptr;
  void* perm = nullptr;
  void* iperm = nullptr;
  void* chirp = nullptr;
  void* hf = nullptr;
};

struct asm_plan {
  int device = 0;
  int R = 0, C = 0, pad_r = 0, pad_c = 0, Rp = 0, Cp = 0, n_colour = 3;
  Phys phys{};
  FftHost fft_rows;  // length Cp: transforms ALONG a row
  FftHost fft_cols;  // length Rp: transforms ALONG a column
  int sm_count = 148;
  int max_smem = 48 * 1024;
  bool rows_fast = false, cols_fast = false;  // compile-time planned kernels exist for (Cp, C, pad_c) / (Rp, R, pad_r)
  int col_logt = 0;                           // log2(columns per tile) of the fast column kernel
  bool spec_fast = false;                     // spectrum-in / spectrum-out calls run on compile-time planned kernels
  int* col_perm = nullptr;                    // device: frequency bin of stored column c (fast rows)
  int* row_perm = nullptr;                    // device: frequency bin of scrambled row position (fast columns)
};

static bool factorize(int n, std::vector<int>& radices) {
  int a = 0, b = 0, c = 0, m = n;
  while (m % 2 == 0) { m /= 2; ++a; }
  while (m % 3 == 0) { m /= 3; ++b; }
  while (m % 5 == 0) { m /= 5; ++c; }
  if (m != 1) return false;
  radices.clear();
  // even radices first (long contiguous strides), odd radices last (odd strides avoid the
  // shared-memory bank conflicts of the short-stride tail passes)
  while (a >= 3) { radices.push_back(8); a -= 3; }
  if (a == 2) radices.push_back(4);
  if (a == 1) radices.push_back(2);
  for (int i = 0; i < c; ++i) radices.push_back(5);
  for (int i = 0; i < b; ++i) radices.push_back(3);
  if (radices.empty()) radices.push_back(1);
  return (int)radices.size() <= kMaxPass;
}

// perm[pos] = natural index stored at position pos after the DIF passes
static void build_perm(const std::vector<int>& radices, size_t pass, int n, int base_pos, int base_idx,
                       int idx_stride, std::vector<int>& perm) {
  if (pass == radices.size() || n == 1) {
    perm[base_pos] = base_idx;
    return;
  }
  const int r = radices[pass];
  const int m = n / r;
  for (int q = 0; q < r; ++q)
    build_perm(radices, pass + 1, m, base_pos + q * m, base_idx + q * idx_stride, idx_stride * r, perm);
}

static int upload(void** dst, const void* src, size_t bytes) {
  CUDA_TRY(cudaMalloc(dst, bytes));
  CUDA_TRY(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
  return ASM_OK;
}

// in-place radix-2 FFT in double (host, plan time only; length a power of two)
static void host_fft_pow2(std::vector<double>& re, std::vector<double>& im) {
  const size_t n = re.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const double ang = -2.0 * M_PI / (double)len;
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; ++k) {
        const double wr = cos(ang * (double)k), wi = sin(ang * (double)k);
        const size_t a = i + k, b = i + k + len / 2;
        const double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
        re[b] = re[a] - xr; im[b] = im[a] - xi;
        re[a] += xr; im[a] += xi;
      }
  }
}

static int make_fft(int n, FftHost& out) {
  std::vector<int> radices;
  if (n < 1) return fail(ASM_EINVAL, "transform length %d", n);
  const bool smooth = factorize(n, radices);
  int len = n;  // length the radix passes run at
  if (!smooth) {
    // Bluestein: convolution length m = 2^k >= 2n-1
    len = 1;
    while (len < 2 * n - 1) len <<= 1;
    if (!factorize(len, radices)) return fail(ASM_EUNSUPPORTED_SIZE, "cannot plan length %d", n);
  }
  if (radices.size() == 1 && radices[0] == 1) radices.clear();
  out.dev.n = n;
  out.dev.blue = smooth ? 0 : 1;
  out.dev.m = smooth ? 0 : len;
  out.dev.buf_len = len;
  out.dev.npass = (int)radices.size();
  for (size_t i = 0; i < radices.size(); ++i) out.dev.radix[i] = radices[i];
  std::vector<float2> tw(len);
  for (int k = 0; k < len; ++k) {
    const double a = -2.0 * M_PI * (double)k / (double)len;
    tw[k] = make_float2((float)cos(a), (float)sin(a));
  }
  std::vector<int> pass_perm(len);
  build_perm(radices, 0, len, 0, 0, 1, pass_perm);
  std::vector<int> perm(n), iperm(n);
  if (smooth) {
    perm = pass_perm;
    for (int p = 0; p < n; ++p) iperm[perm[p]] = p;
  } else {
    for (int p = 0; p < n; ++p) perm[p] = iperm[p] = p;  // Bluestein output is in natural order
    std::vector<float2> chirp(n);
    std::vector<double> hr(len, 0.0), hi(len, 0.0);
    for (long long k = 0; k < n; ++k) {
      const long long k2 = (k * k) % (2LL * n);  // reduce before the multiply by pi/n
      const double a = M_PI * (double)k2 / (double)n;
      chirp[k] = make_float2((float)cos(a), (float)-sin(a));
      hr[k] = cos(a); hi[k] = sin(a);            // conj(c)[k], wrapped to negative lags as well
      if (k > 0) { hr[len - k] = cos(a); hi[len - k] = sin(a); }
    }
    host_fft_pow2(hr, hi);
    std::vector<float2> hf(len);
    for (int p = 0; p < len; ++p)
      hf[p] = make_float2((float)(hr[pass_perm[p]] / len), (float)(hi[pass_perm[p]] / len));
    int rc = upload(&out.chirp, chirp.data(), sizeof(float2) * n);
    if (rc == ASM_OK) rc = upload(&out.hf, hf.data(), sizeof(float2) * len);
    if (rc != ASM_OK) return rc;
    out.dev.chirp = (const float2*)out.chirp;
    out.dev.hf = (const float2*)out.hf;
  }
  int rc = upload(&out.tw, tw.data(), sizeof(float2) * len);
  if (rc == ASM_OK) rc = upload(&out.perm, perm.data(), sizeof(int) * n);
  if (rc == ASM_OK) rc = upload(&out.iperm, iperm.data(), sizeof(int) * n);
  if (rc != ASM_OK) return rc;
  out.dev.tw = (const float2*)out.tw;
  out.dev.perm = (const int*)out.perm;
  out.dev.iperm = (const int*)out.iperm;
  return ASM_OK;
}

static void free_fft(FftHost& f) {
  if (f.tw) cudaFree(f.tw);
  if (f.perm) cudaFree(f.perm);
  if (f.iperm) cudaFree(f.iperm);
  if (f.chirp) cudaFree(f.chirp);
  if (f.hf) cudaFree(f.hf);
  f = FftHost{};
}

// ------------------------------------------------------------------------------------------
// grid builders (attributes of the reference classes)
// ------------------------------------------------------------------------------------------
__global__ void build_grid_kernel(Phys ph, int kind, int n_colour, const float* __restrict__ wm,
                                  const float* __restrict__ z, int n_depth,
                                  int flags, void* __restrict__ out) {
  const size_t plane = (size_t)ph.Rp * ph.Cp;
  size_t total = plane;
  if (kind == ASM_GRID_W) total = plane * n_colour;
  if (kind == ASM_GRID_H || kind == ASM_GRID_BAND_LIMIT) total = plane * n_colour * n_depth;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t pl = i / plane;
    const size_t rem = i - pl * plane;
    const int kr = (int)(rem / ph.Cp);
    const int kc = (int)(rem - (size_t)kr * ph.Cp);
    if (kind == ASM_GRID_W) {
      ((float*)out)[i] = w_value(ph, kr, kc, (int)pl);
    } else if (kind == ASM_GRID_CIRC_MASK) {
      ((float*)out)[i] = radial_value(ph, kr, kc) > ph.radius ? 0.0f : 1.0f;
    } else if (kind == ASM_GRID_RADIAL) {
      ((float*)out)[i] = radial_value(ph, kr, kc);
    } else if (kind == ASM_GRID_H) {
      const int d = (int)(pl / n_colour);
      const int c = (int)(pl - (size_t)d * n_colour);
      ((float2*)out)[i] = filter_value(ph, wm, 1, flags, kr, kc, c, beta_of(z[d]));
    } else {  // band limit, asm.py:173-193
      const int d = (int)(pl / n_colour);
      const int c = (int)(pl - (size_t)d * n_colour);
      const float zz = z[d];
      const float ar = __fmul_rn(ph.two_d_r, zz);
      const float ac = __fmul_rn(ph.two_d_c, zz);
      const float lr = __fdiv_rn(1.0f, __fmul_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(ar, ar), 1.0f)), ph.lambda[c]));
      const float lc = __fdiv_rn(1.0f, __fmul_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(ac, ac), 1.0f)), ph.lambda[c]));
      const float fx = fabsf(__fmul_rn(signed_bin(kr, ph.Rp), ph.fscale_r));
      const float fy = fabsf(__fmul_rn(signed_bin(kc, ph.Cp), ph.fscale_c));
      ((unsigned char*)out)[i] = (fx < lr && fy < lc) ? 1 : 0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// K1: prologue + row forward FFT (generic)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
row_forward_kernel(Fft1d f, RowIn in, long long n_rows, int C, int pad_c, float2* __restrict__ w1) {
  extern __shared__ float2 buf[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int n = f.n;
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    for (int i = tid; i < n; i += nthr) {
      const int c = i - pad_c;
      float2 v = make_float2(0.0f, 0.0f);
      if (c >= 0 && c < C) v = load_input(in, (size_t)row * C + c);
      buf[i] = v;
    }
    __syncthreads();
    fft_dif(buf, f, 0, tid, nthr);
    float2* dst = w1 + (size_t)row * n;
    for (int k = tid; k < n; k += nthr) dst[k] = buf[__ldg(f.iperm + k)];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// K3: row inverse FFT + crop + epilogue (generic)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
row_inverse_kernel(Fft1d f, RowOut o, long long n_rows, int C, int pad_c, const float2* __restrict__ w2) {
  extern __shared__ float2 buf[];
  __shared__ float red[32];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int n = f.n;
  float loss_acc = 0.0f;
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const float2* src = w2 + (size_t)row * n;
    for (int k = tid; k < n; k += nthr) buf[__ldg(f.iperm + k)] = cswap(src[k]);
    __syncthreads();
    fft_dit(buf, f, 0, tid, nthr);
    for (int c = tid; c < C; c += nthr)
      store_output(o, (size_t)row * C + c, cswap(buf[pad_c + c]), loss_acc);
    __syncthreads();
  }
  if (o.loss_partial) block_loss_reduce(loss_acc, o.loss_partial, red);
}

// ------------------------------------------------------------------------------------------
// K2: column pass
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ void col_load(const ColParams& P, float2* __restrict__ buf, size_t in_plane,
                                         int col0, int tid, int nthr) {
  const int T = 1 << P.logT, tmask = T - 1;
  const int n = P.f.n;
  if (P.in_full) {
    // already a spectrum: scatter natural rows to their scrambled slots, the DIT consumes them
    const float2* src = P.in + in_plane * (size_t)n * P.Cp + col0;
    for (int e = tid; e < (n << P.logT); e += nthr) {
      const int i = e >> P.logT, t = e & tmask;
      buf[((size_t)__ldg(P.f.iperm + i) << P.logT) + t] = src[(size_t)i * P.Cp + t];
    }
  } else {
    const float2* src = P.in + in_plane * (size_t)P.R * P.Cp + col0;
    for (int e = tid; e < (n << P.logT); e += nthr) {
      const int i = e >> P.logT, t = e & tmask;
      const int r = i - P.pad_r;
      float2 v = make_float2(0.0f, 0.0f);
      if (r >= 0 && r < P.R) v = src[(size_t)r * P.Cp + t];
      buf[e] = v;
    }
  }
}

__global__ void __launch_bounds__(512)
column_kernel(ColParams P) {
  extern __shared__ float2 smem[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int n = P.f.n;
  const int T = 1 << P.logT, tmask = T - 1;
  const int nel = n << P.logT;
  float2* bufS = smem;
  float2* bufB = P.two_buf ? smem + ((size_t)P.f.buf_len << P.logT) : smem;
  const int tiles_per_plane = P.Cp >> P.logT;
  const long long n_tiles = (long long)P.S * P.n_colour * tiles_per_plane;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int ct = (int)(tile % tiles_per_plane);
    const long long g = tile / tiles_per_plane;  // s*n_colour + c
    const int colour = (int)(g % P.n_colour);
    const long long s = g / P.n_colour;
    const int col0 = ct << P.logT;

    if (!P.reduce) {
      col_load(P, bufS, (size_t)g, col0, tid, nthr);
      __syncthreads();
      if (!P.in_full) fft_dif(bufS, P.f, P.logT, tid, nthr);
      for (int d = 0; d < P.D; ++d) {
        const size_t out_plane = ((size_t)s * P.D + d) * P.n_colour + colour;
        const int zi = P.depth_index ? P.depth_index[s * P.D + d] : d;
        const float beta = P.use_h ? beta_of(P.z[zi]) : 0.0f;
        if (P.out_full) {
          float2* dst = P.out + out_plane * (size_t)n * P.Cp + col0;
          for (int e = tid; e < nel; e += nthr) {
            const int pos = e >> P.logT, t = e & tmask;
            const int kr = __ldg(P.f.perm + pos);
            float2 v = cmul(bufS[e], filter_value(P.ph, P.wm, P.use_h, P.flags, kr, P.col_perm ? __ldg(P.col_perm + col0 + t) : col0 + t, colour, beta));
            v.x *= P.out_scale;
            v.y *= P.out_scale;
            dst[(size_t)kr * P.Cp + t] = v;
          }
        } else {
          for (int e = tid; e < nel; e += nthr) {
            const int pos = e >> P.logT, t = e & tmask;
            const int kr = __ldg(P.f.perm + pos);
            const float2 v = cmul(bufS[e], filter_value(P.ph, P.wm, P.use_h, P.flags, kr, P.col_perm ? __ldg(P.col_perm + col0 + t) : col0 + t, colour, beta));
            bufB[e] = cswap(v);
          }
          __syncthreads();
          fft_dit(bufB, P.f, P.logT, tid, nthr);
          float2* dst = P.out + out_plane * (size_t)P.R * P.Cp + col0;
          for (int e = tid; e < (P.R << P.logT); e += nthr) {
            const int r = e >> P.logT, t = e & tmask;
            dst[(size_t)r * P.Cp + t] = cswap(bufB[((size_t)(r + P.pad_r) << P.logT) + t]);
          }
        }
        __syncthreads();
      }
    } else {
      // adjoint shape: sum over the depth planes in the spectral domain, one inverse transform
      for (int e = tid; e < nel; e += nthr) bufS[e] = make_float2(0.0f, 0.0f);
      for (int d = 0; d < P.D; ++d) {
        const size_t in_plane = ((size_t)s * P.D + d) * P.n_colour + colour;
        const int zi = P.depth_index ? P.depth_index[s * P.D + d] : d;
        const float beta = P.use_h ? beta_of(P.z[zi]) : 0.0f;
        __syncthreads();
        col_load(P, bufB, in_plane, col0, tid, nthr);
        __syncthreads();
        if (!P.in_full) fft_dif(bufB, P.f, P.logT, tid, nthr);
        for (int e = tid; e < nel; e += nthr) {
          const int pos = e >> P.logT, t = e & tmask;
          const int kr = __ldg(P.f.perm + pos);
          const float2 v = cmul(bufB[e], filter_value(P.ph, P.wm, P.use_h, P.flags, kr, P.col_perm ? __ldg(P.col_perm + col0 + t) : col0 + t, colour, beta));
          float2 a = bufS[e];
          a.x += v.x;
          a.y += v.y;
          bufS[e] = a;
        }
      }
      __syncthreads();
      const size_t out_plane = (size_t)g;
      if (P.out_full) {
        float2* dst = P.out + out_plane * (size_t)n * P.Cp + col0;
        for (int e = tid; e < nel; e += nthr) {
          const int pos = e >> P.logT, t = e & tmask;
          float2 v = bufS[e];
          v.x *= P.out_scale;
          v.y *= P.out_scale;
          dst[(size_t)__ldg(P.f.perm + pos) * P.Cp + t] = v;
        }
      } else {
        for (int e = tid; e < nel; e += nthr) bufS[e] = cswap(bufS[e]);
        __syncthreads();
        fft_dit(bufS, P.f, P.logT, tid, nthr);
        float2* dst = P.out + out_plane * (size_t)P.R * P.Cp + col0;
        for (int e = tid; e < (P.R << P.logT); e += nthr) {
          const int r = e >> P.logT, t = e & tmask;
          dst[(size_t)r * P.Cp + t] = cswap(bufS[((size_t)(r + P.pad_r) << P.logT) + t]);
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------
#include <atomic>
#include <mutex>
namespace {

std::atomic<long long> g_launches{0};
std::atomic<int> g_profile{0};
struct ProfRec { int kernel; cudaEvent_t a, b; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
// Events come from a pool that asm_profile_enable(1) fills BEFORE the region it measures: cudaEventCreate between two
// launches of a timed step was seen to stall the stream for tens of milliseconds once in a few runs.
std::vector<cudaEvent_t> g_event_pool;
bool pool_take(cudaEvent_t* a, cudaEvent_t* b) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_event_pool.size() < 2) return false;
  *a = g_event_pool.back();
  g_event_pool.pop_back();
  *b = g_event_pool.back();
  g_event_pool.pop_back();
  return true;
}

struct LaunchScope {  // counts a launch and, when profiling, brackets it with events on its stream
  cudaStream_t stream;
  ProfRec rec{};
  bool on = false;
  LaunchScope(int kernel, cudaStream_t s) : stream(s) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (g_profile.load(std::memory_order_relaxed)) {
      rec.kernel = kernel;
      if (pool_take(&rec.a, &rec.b) ||
          (cudaEventCreate(&rec.a) == cudaSuccess && cudaEventCreate(&rec.b) == cudaSuccess)) {
        on = true;
        cudaEventRecord(rec.a, stream);
      }
    }
  }
  ~LaunchScope() {
    if (on) {
      cudaEventRecord(rec.b, stream);
      std::lock_guard<std::mutex> lk(g_prof_mu);
      g_prof.push_back(rec);
    }
  }
};

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

bool in_is_spatial(int k) { return k != ASM_IN_SPECTRUM; }
bool out_is_spatial(int k) { return k != ASM_OUT_SPECTRUM; }

struct Shape {
  long long p_in_per_sample, p_out_per_sample;  // planes
  size_t w1_per_sample, w2_per_sample;          // bytes
  size_t w3_per_sample;                         // fused step: W1 of the adjoint (the adjoint's W2 reuses W1)
};

int shape_of(const asm_plan* p, const asm_io* io, Shape& s) {
  if (!p || !io) return fail(ASM_EINVAL, "null plan or descriptor");
  if (io->struct_bytes != (int32_t)sizeof(asm_io))
    return fail(ASM_EINVAL, "asm_io.struct_bytes=%d, library expects %d", io->struct_bytes, (int)sizeof(asm_io));
  if (io->n_samples < 0 || io->n_depth < 1) return fail(ASM_EINVAL, "n_samples=%d n_depth=%d", io->n_samples, io->n_depth);
  if (io->in_kind < 0 || io->in_kind > ASM_IN_COTANGENT) return fail(ASM_EINVAL, "in_kind=%d", io->in_kind);
  if (io->out_kind < 0 || io->out_kind > ASM_OUT_GRAD_PHASE) return fail(ASM_EINVAL, "out_kind=%d", io->out_kind);
  if (io->filter_kind != ASM_FILTER_NONE && io->filter_kind != ASM_FILTER_H)
    return fail(ASM_EINVAL, "filter_kind=%d", io->filter_kind);
  if (!in_is_spatial(io->in_kind) && !out_is_spatial(io->out_kind))
    return fail(ASM_EINVAL, "spectrum in and spectrum out in one call is not a propagation");
  const long long din = io->reduce_depth ? io->n_depth : 1;
  const long long dout = io->reduce_depth ? 1 : io->n_depth;
  s.p_in_per_sample = din * p->n_colour;
  s.p_out_per_sample = dout * p->n_colour;
  const size_t strip = (size_t)p->R * p->Cp * sizeof(float2);
  s.w1_per_sample = in_is_spatial(io->in_kind) ? s.p_in_per_sample * strip : 0;
  s.w2_per_sample = out_is_spatial(io->out_kind) ? s.p_out_per_sample * strip : 0;
  s.w3_per_sample = 0;
  if (io->adj_grad_phase) {
    if (!p->rows_fast || !p->cols_fast)
      return fail(ASM_EUNSUPPORTED_SIZE, "fused step: this geometry has no compile-time planned kernels");
    if (io->out_kind != ASM_OUT_ABS || io->reduce_depth ||
        (io->in_kind != ASM_IN_PHASE && io->in_kind != ASM_IN_AMP_PHASE))
      return fail(ASM_EINVAL, "fused step needs phase (or amplitude+phase) in, ASM_OUT_ABS, reduce_depth = 0");
    s.w3_per_sample = s.w2_per_sample;
  }
  return ASM_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" int asm_version(void) { return ASM_B200_VERSION; }
extern "C" int asm_sizeof_io(void) { return (int)sizeof(asm_io); }
extern "C" long long asm_launch_count(void) { return g_launches.load(); }
extern "C" int asm_profile_enable(int on) {
  if (on) {  // enough events for 1024 launches; more are created on demand
    std::lock_guard<std::mutex> lk(g_prof_mu);
    while (g_event_pool.size() < 2048) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) break;
      g_event_pool.push_back(e);
    }
  }
  g_profile.store(on ? 1 : 0);
  return ASM_OK;
}
extern "C" int asm_profile_collect(double* out_ms, long long* out_launches, int n) {
  if (!out_ms || !out_launches || n < 3) return fail(ASM_EINVAL, "asm_profile_collect: need 3 slots");
  std::vector<ProfRec> recs;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    recs.swap(g_prof);
  }
  int rc = ASM_OK;
  for (auto& r : recs) {
    float ms = 0.0f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      if (r.kernel >= 0 && r.kernel < n) {
        out_ms[r.kernel] += ms;
        out_launches[r.kernel] += 1;
      }
    } else {
      rc = fail(ASM_ECUDA, "profile event failed");
    }
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_event_pool.push_back(r.a);  // back to the pool
    g_event_pool.push_back(r.b);
  }
  return rc;
}
extern "C" const char* asm_last_error(void) { return g_err; }

extern "C" int asm_plan_create(asm_plan** out, int device, int rows, int cols, int pad_rows, int pad_cols,
                               double pitch, const float* wavelengths, int n_colour, double mask_radius) {
  if (!out) return fail(ASM_EINVAL, "out is null");
  *out = nullptr;
  if (rows < 1 || cols < 1 || pad_rows < 0 || pad_cols < 0)
    return fail(ASM_EINVAL, "rows=%d cols=%d pad=%d,%d", rows, cols, pad_rows, pad_cols);
  if (n_colour < 1 || n_colour > kMaxColour || !wavelengths)
    return fail(ASM_EINVAL, "n_colour=%d (max %d)", n_colour, kMaxColour);
  if (!(pitch > 0.0)) return fail(ASM_EINVAL, "pitch=%g", pitch);
  const int Rp = rows + 2 * pad_rows, Cp = cols + 2 * pad_cols;
  const int shorter = Rp < Cp ? Rp : Cp;
  if (mask_radius > shorter / 2.0)  // util.py:225-229
    return fail(ASM_EINVAL, "The radius %g is larger than the half of the sample size %g", mask_radius,
                shorter / 2.0);
  DeviceGuard guard(device);
  if (!guard.ok) return fail(ASM_ECUDA, "cannot select CUDA device %d", device);
  asm_plan* p = new (std::nothrow) asm_plan();
  if (!p) return fail(ASM_EINVAL, "out of host memory");
  p->device = device;
  p->R = rows; p->C = cols; p->pad_r = pad_rows; p->pad_c = pad_cols; p->Rp = Rp; p->Cp = Cp;
  p->n_colour = n_colour;
  Phys& ph = p->phys;
  ph.Rp = Rp; ph.Cp = Cp;
  ph.fscale_r = (float)(1.0 / ((double)Rp * pitch));
  ph.fscale_c = (float)(1.0 / ((double)Cp * pitch));
  ph.uscale_r = (float)(1.0 / (double)Rp);
  ph.uscale_c = (float)(1.0 / (double)Cp);
  ph.short_edge = (float)shorter;
  ph.radius = (float)mask_radius;
  ph.two_d_r = (float)(2.0 * (1.0 / ((double)Rp * pitch)));
  ph.two_d_c = (float)(2.0 * (1.0 / ((double)Cp * pitch)));
  for (int c = 0; c < n_colour; ++c) {
    volatile float l = wavelengths[c];
    volatile float l2 = l * l;       // wave_length**2, fp32 (asm.py:166)
    volatile float inv = 1.0f / l2;  // fp32 division
    ph.inv_l2[c] = inv;
    ph.lambda[c] = l;
  }
  int rc = make_fft(Cp, p->fft_rows);
  if (rc == ASM_OK) rc = make_fft(Rp, p->fft_cols);
  if (rc != ASM_OK) {
    asm_plan_destroy(p);
    return rc;
  }
  const char* no_fast = getenv("LHG_DISABLE_FAST");
  if (!(no_fast && no_fast[0] == '1')) {
    p->rows_fast = fast_rows_supported(Cp, cols, pad_cols);
    p->col_logt = fast_cols_logt(Rp, rows, pad_rows);
    p->cols_fast = p->col_logt >= 0 && (Cp & ((1 << p->col_logt) - 1)) == 0;
    const int sync_logt = fast_cols_sync_logt(Rp, rows, pad_rows);
    p->spec_fast = p->rows_fast && sync_logt >= 1 && (Cp & ((1 << sync_logt) - 1)) == 0;
  }
  if (p->rows_fast) {
    std::vector<int> perm(Cp);
    fast_rows_perm(Cp, perm.data());
    if (cudaMalloc((void**)&p->col_perm, sizeof(int) * Cp) != cudaSuccess ||
        cudaMemcpy(p->col_perm, perm.data(), sizeof(int) * Cp, cudaMemcpyHostToDevice) != cudaSuccess) {
      asm_plan_destroy(p);
      return fail(ASM_ECUDA, "cannot upload the column permutation");
    }
  }
  if (p->cols_fast) {
    std::vector<int> perm(Rp);
    fast_cols_perm(Rp, rows, pad_rows, perm.data());
    if (cudaMalloc((void**)&p->row_perm, sizeof(int) * Rp) != cudaSuccess ||
        cudaMemcpy(p->row_perm, perm.data(), sizeof(int) * Rp, cudaMemcpyHostToDevice) != cudaSuccess) {
      asm_plan_destroy(p);
      return fail(ASM_ECUDA, "cannot upload the row permutation");
    }
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
    p->sm_count = prop.multiProcessorCount;
    p->max_smem = (int)prop.sharedMemPerBlockOptin;
  }
  *out = p;
  return ASM_OK;
}

extern "C" int asm_plan_destroy(asm_plan* p) {
  if (!p) return ASM_OK;
  DeviceGuard guard(p->device);
  free_fft(p->fft_rows);
  free_fft(p->fft_cols);
  if (p->col_perm) cudaFree(p->col_perm);
  if (p->row_perm) cudaFree(p->row_perm);
  delete p;
  return ASM_OK;
}

extern "C" int asm_plan_info(const asm_plan* p, int32_t* out, int n) {
  if (!p || !out || n < 3) return fail(ASM_EINVAL, "asm_plan_info: need 3 slots");
  out[0] = p->Rp;
  out[1] = p->Cp;
  out[2] = (p->fft_rows.dev.blue || p->fft_cols.dev.blue) ? 0 : 1;
  return ASM_OK;
}

extern "C" size_t asm_workspace_bytes(const asm_plan* p, const asm_io* io) {
  Shape s;
  if (shape_of(p, io, s) != ASM_OK) return 0;
  return align_up(s.w1_per_sample * (size_t)io->n_samples, 256) +
         align_up(s.w2_per_sample * (size_t)io->n_samples, 256) +
         align_up(s.w3_per_sample * (size_t)io->n_samples, 256) + 256;
}

extern "C" int asm_build_grid(const asm_plan* p, int kind, const float* wm_grid, const float* z_dev,
                              int n_depth, int flags, void* out_dev, asm_stream stream) {
  if (!p || !out_dev) return fail(ASM_EINVAL, "null plan or output");
  if (kind < ASM_GRID_W || kind > ASM_GRID_BAND_LIMIT) return fail(ASM_EINVAL, "grid kind %d", kind);
  if ((kind == ASM_GRID_H || kind == ASM_GRID_BAND_LIMIT) && (!z_dev || n_depth < 1))
    return fail(ASM_EINVAL, "grid kind %d needs distances", kind);
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail(ASM_ECUDA, "cannot select CUDA device %d", p->device);
  build_grid_kernel<<<p->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(p->phys, kind, p->n_colour, wm_grid, z_dev,
                                                                      n_depth, flags, out_dev);
  CUDA_TRY(cudaPeekAtLastError());
  return ASM_OK;
}

extern "C" size_t asm_wm_tiled_bytes(const asm_plan* p) {
  if (!p || !p->cols_fast) return 0;
  return align_up(sizeof(float) * (size_t)p->n_colour * p->Rp * p->Cp, 256) + sizeof(int) * (size_t)(p->Cp >> p->col_logt);
}

extern "C" int asm_build_wm_tiled(const asm_plan* p, const float* wm_grid, void* out_dev, asm_stream stream) {
  if (!p || !out_dev) return fail(ASM_EINVAL, "null plan or output");
  if (!p->cols_fast) return fail(ASM_EINVAL, "this geometry has no tiled column kernel (asm_wm_tiled_bytes == 0)");
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail(ASM_ECUDA, "cannot select CUDA device %d", p->device);
  float* wmt = (float*)out_dev;
  int* active = (int*)((char*)out_dev + align_up(sizeof(float) * (size_t)p->n_colour * p->Rp * p->Cp, 256));
  const int rc = fast_wm_tiled(p->phys, wm_grid, p->n_colour, p->col_logt, p->row_perm,
                               p->rows_fast ? p->col_perm : nullptr, wmt, active, p->sm_count, (cudaStream_t)stream);
  if (rc != 0) return fail(ASM_ECUDA, "wm tiling launch failed: %s", cudaGetErrorString((cudaError_t)rc));
  return ASM_OK;
}

static int pick_threads(int n_elems_per_pass) {
  // smallest power of two in [128, 512] that gives each thread at most ~8 radix-4 butterflies
  int t = 128;
  while (t < 512 && t * 32 < n_elems_per_pass) t <<= 1;
  return t;
}

namespace {
// the one finishing block of the fused loss: partial sums in index order, double accumulator, one thread per
// strided slice then a fixed-order tree -- bit-identical from run to run (SURVEY 8(b) "Determinism")
__global__ void __launch_bounds__(256, 1) loss_finish_kernel(const float* __restrict__ partial, int len, double scale,
                                                             float* __restrict__ out) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < len; i += 256) acc += (double)partial[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(red[0] * scale);
}
}  // namespace

extern "C" int asm_loss_finish(const float* loss_partial, int len, double scale, float* out_dev, asm_stream stream) {
  if (!loss_partial || !out_dev || len < 1) return fail(ASM_EINVAL, "asm_loss_finish: null pointer or len < 1");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  loss_finish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(loss_partial, len, scale, out_dev);
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) return fail(ASM_ECUDA, "asm_loss_finish: %s", cudaGetErrorString(e));
  return ASM_OK;
}

extern "C" int asm_fused_step_supported(const asm_plan* p) { return (p && p->rows_fast && p->cols_fast) ? 1 : 0; }

extern "C" int asm_propagate(const asm_plan* p, const asm_io* io, asm_stream stream_) {
  Shape sh;
  int rc = shape_of(p, io, sh);
  if (rc != ASM_OK) return rc;
  if (io->n_samples == 0) return ASM_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool sin = in_is_spatial(io->in_kind), sout = out_is_spatial(io->out_kind);
  if (!io->in0 && io->in_kind != ASM_IN_PHASE) return fail(ASM_EINVAL, "in0 is null");
  if ((io->in_kind == ASM_IN_PHASE || io->in_kind == ASM_IN_AMP_PHASE) && !io->in1)
    return fail(ASM_EINVAL, "in1 (phase) is null");
  const bool fused_step = io->adj_grad_phase != nullptr;
  if (!io->out0 && !fused_step) return fail(ASM_EINVAL, "out0 is null");
  if (fused_step && (!io->loss_target || !io->loss_partial))
    return fail(ASM_EINVAL, "fused step needs loss_target and loss_partial");
  if (io->loss_target_u8 && !fused_step)
    return fail(ASM_EINVAL, "loss_target_u8 is honoured by the fused step only (adj_grad_phase)");
  if (io->out_kind == ASM_OUT_ABS_ANGLE && !io->out1) return fail(ASM_EINVAL, "out1 is null");
  if (io->out_kind == ASM_OUT_GRAD_PHASE && !io->aux_phase) return fail(ASM_EINVAL, "aux_phase is null");
  if (io->filter_kind == ASM_FILTER_H && (!io->z_dev || io->n_z < 1)) return fail(ASM_EINVAL, "z_dev is null");
  if (io->filter_kind == ASM_FILTER_H && !io->depth_index && io->n_z < io->n_depth)
    return fail(ASM_EINVAL, "n_z=%d < n_depth=%d", io->n_z, io->n_depth);
  if (io->loss_partial && (io->out_kind != ASM_OUT_ABS || !io->loss_target || io->loss_partial_len < 1))
    return fail(ASM_EINVAL, "fused loss needs ASM_OUT_ABS, a target and loss_partial_len >= 1");
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail(ASM_ECUDA, "cannot select CUDA device %d", p->device);

  // ---- split the batch so W1+W2 of a chunk fit the scratch buffer ----
  const size_t per_sample = align_up(sh.w1_per_sample, 256) + align_up(sh.w2_per_sample, 256) +
                            align_up(sh.w3_per_sample, 256);
  long long chunk = io->n_samples;
  if (per_sample > 0) {
    if (!io->workspace) return fail(ASM_EWORKSPACE, "workspace is null");
    const size_t usable = io->workspace_bytes > 256 ? io->workspace_bytes - 256 : 0;
    chunk = (long long)(usable / per_sample);
    if (chunk < 1)
      return fail(ASM_EWORKSPACE, "workspace of %zu bytes cannot hold one sample (%zu bytes)",
                  io->workspace_bytes, per_sample);
    if (chunk > io->n_samples) chunk = io->n_samples;
  }
  char* ws = (char*)(((uintptr_t)io->workspace + 255) & ~(uintptr_t)255);

  // ---- column pass configuration ----
  const int two_buf = io->n_depth > 1 ? 1 : 0;
  const size_t col_len = (size_t)p->fft_cols.dev.buf_len, row_len = (size_t)p->fft_rows.dev.buf_len;
  int logT = 4;
  while (logT > 0 && ((p->Cp & ((1 << logT) - 1)) != 0 ||
                      (size_t)(1 + two_buf) * col_len * sizeof(float2) * (1u << logT) > (size_t)p->max_smem))
    --logT;
  const size_t col_smem = (size_t)(1 + two_buf) * col_len * sizeof(float2) << logT;
  // prefer >= 2 resident CTAs per SM when the tile is still at least 4 columns wide
  int logT_use = logT;
  while (logT_use > 2 && ((size_t)(1 + two_buf) * col_len * sizeof(float2) << logT_use) * 2 > (size_t)p->max_smem)
    --logT_use;
  const size_t col_smem_use = (size_t)(1 + two_buf) * col_len * sizeof(float2) << logT_use;
  const size_t row_smem = row_len * sizeof(float2);
  // the run-time planned kernels opt in to the device's maximum dynamic shared memory ONCE per device (plans with
  // different needs, and autograd's backward thread, then never race between a set and a launch)
  {
    static std::mutex mu;
    static std::vector<int> done;
    std::lock_guard<std::mutex> lk(mu);
    if (std::find(done.begin(), done.end(), p->device) == done.end()) {
      // the opt-in limit covers static + dynamic shared memory of a kernel
      auto optin = [&](const void* k) -> cudaError_t {
        cudaFuncAttributes fa{};
        cudaError_t e = cudaFuncGetAttributes(&fa, k);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, p->max_smem - (int)fa.sharedSizeBytes);
      };
      CUDA_TRY(optin((const void*)column_kernel));
      CUDA_TRY(optin((const void*)row_forward_kernel));
      CUDA_TRY(optin((const void*)row_inverse_kernel));
      done.push_back(p->device);
    }
  }
  const int row_threads = pick_threads(p->Cp);
  const int col_threads = pick_threads(p->Rp << logT_use);

  if (io->loss_partial)
    CUDA_TRY(cudaMemsetAsync(io->loss_partial, 0, sizeof(float) * io->loss_partial_len, stream));

  // compile-time planned kernels: only when both ends are spatial (W1/W2 then keep the scrambled
  // column order of the fast row transform; natural-order spectra in global memory need the generic rows)
  // Their prologue / epilogue use 16-byte accesses: any tensor that is not 16-byte aligned (a view with an odd
  // storage offset) sends the whole call to the run-time planned kernels.  Rows and columns switch together:
  // wm_tiled is laid out for the column order the compile-time planned ROW kernel produces.
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool aligned = al16(io->in0) && al16(io->in1) && al16(io->cot_abs) && al16(io->cot_angle) &&
                       al16(io->cot_abs2) && al16(io->cot_target) && al16(io->out0) && al16(io->out1) &&
                       al16(io->save_field) && al16(io->aux_phase) && al16(io->aux_amp) && al16(io->loss_target) &&
                       al16(io->wm_tiled) && al16(io->adj_grad_phase);
  const bool needs_w = io->filter_kind == ASM_FILTER_H || (io->filter_flags & ASM_FILTER_CIRC_MASK);
  // one side a natural-order spectrum (F-7, F-11, F-12, F-13 and their adjoints): the row kernels keep W1 / W2 in
  // natural column order, the CTA-synchronous column kernel reads / writes the spectrum and the plain w/mask grid
  static const bool spec_enabled = [] { const char* e = getenv("LHG_SPECTRUM_FAST"); return !(e && e[0] == '0'); }();
  const bool fast_spec = spec_enabled && p->spec_fast && (sin != sout) && aligned && (io->wm_grid || !needs_w) &&
                         !(!sin && io->reduce_depth && io->n_depth > 1);
  const bool fast_both = p->rows_fast && p->cols_fast && sin && sout && aligned && (io->wm_tiled || !needs_w);
  const bool fast_rows = fast_spec || fast_both || (p->rows_fast && !p->cols_fast && sin && sout && aligned);
  const bool fast_cols = fast_spec || fast_both || (p->cols_fast && !p->rows_fast && sin && sout && (io->wm_tiled || !needs_w));
  if (fused_step && !fast_both)
    return fail(ASM_EUNSUPPORTED_SIZE, "fused step: needs 16-byte aligned tensors and the tiled w/mask grid");
  // shared-memory limits and occupancy of the run-time planned kernels: only where they are the ones that run
  int row_occ_f = 1, row_occ_i = 1, col_occ = 1;
  if (!fast_rows) {
    if (row_smem > (size_t)p->max_smem)
      return fail(ASM_EUNSUPPORTED_SIZE, "row of %d samples does not fit shared memory", p->Cp);
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&row_occ_f, row_forward_kernel, row_threads, row_smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&row_occ_i, row_inverse_kernel, row_threads, row_smem));
  }
  if (!fast_cols) {
    if (col_smem > (size_t)p->max_smem)
      return fail(ASM_EUNSUPPORTED_SIZE, "column of %d samples does not fit shared memory", p->Rp);
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&col_occ, column_kernel, col_threads, col_smem_use));
  }
  if (row_occ_f < 1) row_occ_f = 1;
  if (row_occ_i < 1) row_occ_i = 1;
  if (col_occ < 1) col_occ = 1;
  const int* col_perm = (fast_rows && !fast_spec) ? p->col_perm : nullptr;
  const int natural = fast_spec ? 1 : 0;
  // blocked W1/W2 (common.cuh woff): only between the compile-time planned kernels, and only when a column
  // tile is 2 or 4 columns wide (16- or 32-byte pieces in the plain layout).  W1 (written by
  // the row kernel) uses 4-column blocks, W2 (written by the column kernel) 2-column blocks.
  static const int blk_in = [] { const char* e = getenv("LHG_BLOCK_W1"); return e ? atoi(e) : 2; }();
  static const int blk_out = [] { const char* e = getenv("LHG_BLOCK_W2"); return e ? atoi(e) : -1; }();
  const bool can_block = fast_rows && fast_cols && !fast_spec && p->col_logt <= 2 && (p->R % 8) == 0 && (p->Cp % 4) == 0;
  // W2 blocks are as wide as a column tile (its writer then stores whole blocks): 8x2 for 2-column tiles
  // (128-byte lines), 8x4 for 4-column tiles
  const int blocked_in = can_block ? blk_in : 0;
  const int blocked_out = can_block ? (blk_out >= 0 ? blk_out : p->col_logt) : 0;
  // column tiles entirely outside the circular mask: skipped by all three compile-time planned kernels
  DeadCols dead{nullptr, 0};
  if (fast_rows && fast_cols && !fast_spec && io->wm_tiled && (io->filter_flags & ASM_FILTER_CIRC_MASK)) {
    dead.active = (const int*)((const char*)io->wm_tiled +
                               align_up(sizeof(float) * (size_t)p->n_colour * p->Rp * p->Cp, 256));
    dead.logt = p->col_logt;
  }

  const size_t in_elem = io->in_kind == ASM_IN_SPECTRUM ? (size_t)p->Rp * p->Cp : (size_t)p->R * p->C;
  const size_t out_elem = io->out_kind == ASM_OUT_SPECTRUM ? (size_t)p->Rp * p->Cp : (size_t)p->R * p->C;
  const size_t rc_elem = (size_t)p->R * p->C;

  for (long long s0 = 0; s0 < io->n_samples; s0 += chunk) {
    const long long ns = (io->n_samples - s0 < chunk) ? io->n_samples - s0 : chunk;
    const size_t pin0 = (size_t)s0 * sh.p_in_per_sample;    // first input plane of the chunk
    const size_t pout0 = (size_t)s0 * sh.p_out_per_sample;  // first output plane of the chunk
    float2* w1 = (float2*)ws;
    float2* w2 = (float2*)(ws + align_up(sh.w1_per_sample * (size_t)ns, 256));
    float2* w3 = (float2*)((char*)w2 + align_up(sh.w2_per_sample * (size_t)ns, 256));

    if (sin) {
      RowIn ri{};
      ri.kind = io->in_kind;
      const size_t esz0 = (io->in_kind == ASM_IN_COMPLEX || io->in_kind == ASM_IN_COTANGENT) ? sizeof(float2) : sizeof(float);
      ri.in0 = io->in0 ? (const char*)io->in0 + pin0 * rc_elem * esz0 : nullptr;
      ri.in1 = io->in1 ? (const char*)io->in1 + pin0 * rc_elem * sizeof(float) : nullptr;
      ri.cot_abs = io->cot_abs ? io->cot_abs + pin0 * rc_elem : nullptr;
      ri.cot_angle = io->cot_angle ? io->cot_angle + pin0 * rc_elem : nullptr;
      ri.cot_abs2 = io->cot_abs2 ? io->cot_abs2 + pin0 * rc_elem : nullptr;
      ri.cot_target = io->cot_target ? io->cot_target + pin0 * rc_elem : nullptr;
      ri.cot_scale = io->cot_scale;
      ri.phase_scale = io->phase_scale;
      const long long n_rows = ns * sh.p_in_per_sample * p->R;
      long long grid = (long long)p->sm_count * row_occ_f;
      if (grid > n_rows) grid = n_rows;
      if (fast_rows) {
        LaunchScope ls(0, stream);
        const int frc = fast_row_forward(p->Cp, p->fft_rows.dev.tw, ri, n_rows, p->C, p->pad_c, w1, blocked_in, dead, natural, p->sm_count, stream);
        if (frc != 0) return fail(ASM_ECUDA, "fast row-forward launch failed (%d: %s)", frc,
                                  frc > 0 ? cudaGetErrorString((cudaError_t)frc) : "no plan");
      } else {
        LaunchScope ls(0, stream);
        row_forward_kernel<<<(unsigned)grid, row_threads, row_smem, stream>>>(p->fft_rows.dev, ri, n_rows, p->C,
                                                                              p->pad_c, w1);
      }
      CUDA_TRY(cudaPeekAtLastError());
    }
    {
      ColParams cp{};
      cp.f = p->fft_cols.dev;
      cp.ph = p->phys;
      cp.logT = logT_use;
      cp.S = (int)ns;
      cp.D = io->n_depth;
      cp.n_colour = p->n_colour;
      cp.reduce = (io->reduce_depth && io->n_depth > 1) ? 1 : 0;
      cp.in_full = sin ? 0 : 1;
      cp.out_full = sout ? 0 : 1;
      cp.R = p->R;
      cp.pad_r = p->pad_r;
      cp.Cp = p->Cp;
      cp.use_h = io->filter_kind == ASM_FILTER_H ? 1 : 0;
      cp.flags = io->filter_flags;
      cp.two_buf = two_buf;
      cp.in = sin ? w1 : (const float2*)io->in0 + pin0 * in_elem;
      cp.out = sout ? w2 : (float2*)io->out0 + pout0 * out_elem;
      cp.z = io->z_dev;
      cp.wm = io->wm_grid;
      cp.depth_index = io->depth_index ? io->depth_index + s0 * io->n_depth : nullptr;
      cp.out_scale = io->out_scale;
      const long long n_tiles = ns * p->n_colour * (p->Cp >> logT_use);
      long long grid = (long long)p->sm_count * col_occ;
      if (grid > n_tiles) grid = n_tiles;
      cp.col_perm = col_perm;
      // W2 columns outside the mask stay unwritten only if the inverse row kernel will not read them
      cp.rows_skip_dead = (dead.active && sout &&
                           (fused_step ||  // the fused row kernel never gathers with the TMA unit
                            !fast_row_inverse_uses_tma(p->Cp, p->C, p->pad_c, ns * sh.p_out_per_sample * p->R, blocked_out)))
                              ? 1 : 0;
      cp.blocked_in = blocked_in;
      cp.blocked_out = blocked_out;
      if (fast_cols && io->wm_tiled && !fast_spec) {
        cp.wmt = (const float*)io->wm_tiled;
        cp.tile_active = (const int*)((const char*)io->wm_tiled +
                                      align_up(sizeof(float) * (size_t)p->n_colour * p->Rp * p->Cp, 256));
      }
      int frc = -1;
      if (fast_cols) {
        LaunchScope ls(1, stream);
        frc = fast_columns(cp, p->sm_count, stream);
        if (frc > 0) return fail(ASM_ECUDA, "fast column launch failed: %s", cudaGetErrorString((cudaError_t)frc));
      }
      // the compile-time planned row kernel has already written W1 in ITS layout (scrambled columns, blocked,
      // dead tiles skipped): the run-time planned column kernel cannot take over from there
      if (frc != 0 && fast_cols)
        return fail(ASM_EUNSUPPORTED_SIZE, "no compile-time planned column kernel for this call (n_depth = %d)", io->n_depth);
      if (frc != 0) {
        LaunchScope ls(1, stream);
        column_kernel<<<(unsigned)grid, col_threads, col_smem_use, stream>>>(cp);
      }
      CUDA_TRY(cudaPeekAtLastError());
      if (fused_step) {
        // row inverse of the forward + row forward of the adjoint in one kernel: W2 -> W1' (w3)
        FusedRows fr{};
        if (io->loss_target_u8) fr.target_u8 = (const unsigned char*)io->loss_target + pout0 * rc_elem;
        else fr.target = io->loss_target + pout0 * rc_elem;
        fr.amp_out = io->out0 ? (float*)io->out0 + pout0 * rc_elem : nullptr;
        fr.scale = io->out_scale;
        fr.cot_scale = io->adj_cot_scale;
        fr.loss_partial = io->loss_partial;
        const long long n_rows_out = ns * sh.p_out_per_sample * p->R;
        {
          LaunchScope ls(3, stream);
          const int rrc = fast_row_inverse_forward(p->Cp, p->fft_rows.dev.tw, fr, n_rows_out, p->C, p->pad_c, w2, blocked_out,
                                                   w3, blocked_in, dead, p->sm_count, io->loss_partial_len, stream);
          if (rrc != 0) return fail(ASM_ECUDA, "fused row launch failed (%d: %s)", rrc,
                                    rrc > 0 ? cudaGetErrorString((cudaError_t)rrc) : "no plan");
        }
        // adjoint column pass: conj(filter), summed over depth, W1' -> W2' (= the W1 region, free by now)
        ColParams ca = cp;
        ca.reduce = io->n_depth > 1 ? 1 : 0;
        ca.flags = io->filter_flags ^ ASM_FILTER_CONJ;
        ca.in = w3;
        ca.out = w1;
        ca.rows_skip_dead = (dead.active && !fast_row_inverse_uses_tma(p->Cp, p->C, p->pad_c,
                                                                       ns * sh.p_in_per_sample * p->R, blocked_out)) ? 1 : 0;
        {
          LaunchScope ls(1, stream);
          const int crc = fast_columns(ca, p->sm_count, stream);
          if (crc != 0) return fail(ASM_ECUDA, "fused step: adjoint column launch failed (%d)", crc);
        }
        // adjoint row inverse: d/dphase = s * a * Im(conj(e^{i s phase}) * xbar)
        RowOut ra{};
        ra.kind = ASM_OUT_GRAD_PHASE;
        ra.out0 = io->adj_grad_phase + pin0 * rc_elem;
        ra.aux_phase = (const float*)io->in1 + pin0 * rc_elem;
        ra.aux_amp = io->in_kind == ASM_IN_AMP_PHASE ? (const float*)io->in0 + pin0 * rc_elem : nullptr;
        ra.phase_scale = io->phase_scale;
        ra.scale = io->out_scale;
        const long long n_rows_in = ns * sh.p_in_per_sample * p->R;
        {
          LaunchScope ls(2, stream);
          const int rrc = fast_row_inverse(p->Cp, p->fft_rows.dev.tw, ra, n_rows_in, p->C, p->pad_c, w1, blocked_out, dead,
                                           0, p->sm_count, 0, stream);
          if (rrc != 0) return fail(ASM_ECUDA, "fused step: adjoint row launch failed (%d)", rrc);
        }
        CUDA_TRY(cudaPeekAtLastError());
        continue;
      }
    }
    if (sout) {
      RowOut ro{};
      ro.kind = io->out_kind;
      const size_t esz = io->out_kind == ASM_OUT_COMPLEX ? sizeof(float2) : sizeof(float);
      ro.out0 = (char*)io->out0 + pout0 * rc_elem * esz;
      ro.out1 = io->out1 ? (char*)io->out1 + pout0 * rc_elem * sizeof(float) : nullptr;
      ro.save_field = io->save_field ? (float2*)io->save_field + pout0 * rc_elem : nullptr;
      ro.aux_phase = io->aux_phase ? io->aux_phase + pout0 * rc_elem : nullptr;
      ro.aux_amp = io->aux_amp ? io->aux_amp + pout0 * rc_elem : nullptr;
      ro.phase_scale = io->phase_scale;
      ro.scale = io->out_scale;
      ro.loss_target = io->loss_target ? io->loss_target + pout0 * rc_elem : nullptr;
      ro.loss_partial = io->loss_partial;
      const long long n_rows = ns * sh.p_out_per_sample * p->R;
      long long grid = (long long)p->sm_count * row_occ_i;
      if (grid > n_rows) grid = n_rows;
      if (io->loss_partial && grid > io->loss_partial_len) grid = io->loss_partial_len;
      if (fast_rows) {
        LaunchScope ls(2, stream);
        const int frc = fast_row_inverse(p->Cp, p->fft_rows.dev.tw, ro, n_rows, p->C, p->pad_c, w2, blocked_out, dead, natural, p->sm_count,
                                         io->loss_partial ? io->loss_partial_len : 0, stream);
        if (frc != 0) return fail(ASM_ECUDA, "fast row-inverse launch failed (%d: %s)", frc,
                                  frc > 0 ? cudaGetErrorString((cudaError_t)frc) : "no plan");
      } else {
        LaunchScope ls(2, stream);
        row_inverse_kernel<<<(unsigned)grid, row_threads, row_smem, stream>>>(p->fft_rows.dev, ro, n_rows, p->C,
                                                                              p->pad_c, w2);
      }
      CUDA_TRY(cudaPeekAtLastError());
    }
  }
  return ASM_OK;
}
