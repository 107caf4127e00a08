// common.cuh -- structs and device helpers shared by the generic and the compile-time planned kernels.
#pragma once
#include <cuda_runtime.h>

#include "../../include/asm_b200.h"
#include "fft_core.cuh"
#include "physics.cuh"

namespace asmb {

// ---- what the row-forward pass reads (prologue) ---------------------------------------------
struct RowIn {
  int kind;
  const void* in0;
  const void* in1;
  const float* cot_abs;
  const float* cot_angle;
  const float* cot_abs2;
  const float* cot_target;
  float cot_scale;
  float phase_scale;
};

__device__ __forceinline__ float2 load_input(const RowIn& in, size_t idx) {
  switch (in.kind) {
    case ASM_IN_PHASE: {
      float s, c;
      sincosf(__fmul_rn(in.phase_scale, ((const float*)in.in1)[idx]), &s, &c);
      return make_float2(c, s);
    }
    case ASM_IN_AMP_PHASE: {
      float s, c;
      sincosf(__fmul_rn(in.phase_scale, ((const float*)in.in1)[idx]), &s, &c);
      const float a = ((const float*)in.in0)[idx];
      return make_float2(a * c, a * s);
    }
    case ASM_IN_COMPLEX:
      return ((const float2*)in.in0)[idx];
    case ASM_IN_COTANGENT: {
      const float2 y = ((const float2*)in.in0)[idx];
      const float r2 = y.x * y.x + y.y * y.y;
      float2 acc = make_float2(0.0f, 0.0f);
      if (r2 > 0.0f) {
        const float r = sqrtf(r2);
        float g = 0.0f;
        if (in.cot_abs) g += in.cot_abs[idx];
        if (in.cot_target) g += in.cot_scale * (r - in.cot_target[idx]);
        const float gr = g / r;
        acc.x = gr * y.x;
        acc.y = gr * y.y;
        if (in.cot_angle) {
          const float ga = in.cot_angle[idx] / r2;
          acc.x -= ga * y.y;
          acc.y += ga * y.x;
        }
      }
      if (in.cot_abs2) {
        const float g2 = 2.0f * in.cot_abs2[idx];
        acc.x += g2 * y.x;
        acc.y += g2 * y.y;
      }
      return acc;
    }
    default:
      return make_float2(0.0f, 0.0f);
  }
}


// ---- what the row-inverse pass writes (epilogue) ---------------------------------------------
struct RowOut {
  int kind;
  void* out0;
  void* out1;
  float2* save_field;
  const float* aux_phase;
  const float* aux_amp;
  float phase_scale;
  float scale;
  const float* loss_target;
  float* loss_partial;
};

// v = un-normalised cropped field sample; idx = flat [plane,row,col] index of the output tensors
__device__ __forceinline__ void store_output(const RowOut& o, size_t idx, float2 v, float& loss_acc) {
  v.x *= o.scale;
  v.y *= o.scale;
  if (o.save_field) o.save_field[idx] = v;
  switch (o.kind) {
    case ASM_OUT_ABS: {
      const float a = sqrtf(v.x * v.x + v.y * v.y);
      ((float*)o.out0)[idx] = a;
      if (o.loss_target) {
        const float d = a - o.loss_target[idx];
        loss_acc += d * d;
      }
      break;
    }
    case ASM_OUT_ANGLE:
      ((float*)o.out0)[idx] = atan2f(v.y, v.x);
      break;
    case ASM_OUT_ABS_ANGLE:
      ((float*)o.out0)[idx] = sqrtf(v.x * v.x + v.y * v.y);
      ((float*)o.out1)[idx] = atan2f(v.y, v.x);
      break;
    case ASM_OUT_COMPLEX:
      ((float2*)o.out0)[idx] = v;
      break;
    case ASM_OUT_ABS2:
      ((float*)o.out0)[idx] = v.x * v.x + v.y * v.y;
      break;
    case ASM_OUT_GRAD_PHASE: {
      float s, cs;
      sincosf(__fmul_rn(o.phase_scale, o.aux_phase[idx]), &s, &cs);
      const float a = o.aux_amp ? o.aux_amp[idx] : 1.0f;
      ((float*)o.out0)[idx] = o.phase_scale * a * (v.y * cs - v.x * s);
      if (o.out1) ((float*)o.out1)[idx] = v.x * cs + v.y * s;
      break;
    }
    default:
      break;
  }
}

// ---- 4-wide versions (compile-time planned kernels): every global access is 16 bytes per lane and all
// loads of a group are issued before anything is computed or stored ---------------------------------
struct In4 {  // raw operands of 4 consecutive input samples
  float4 a, b, c, d, e, f;
};

__device__ __forceinline__ float4 ldg4(const float* p, size_t idx) { return __ldg(reinterpret_cast<const float4*>(p + idx)); }

template <int KIND>
__device__ __forceinline__ In4 fetch_input4(const RowIn& in, size_t idx) {
  In4 r;
  r.a = r.b = r.c = r.d = r.e = r.f = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if constexpr (KIND == ASM_IN_PHASE) {
    r.a = ldg4((const float*)in.in1, idx);
  } else if constexpr (KIND == ASM_IN_AMP_PHASE) {
    r.a = ldg4((const float*)in.in1, idx);
    r.b = ldg4((const float*)in.in0, idx);
  } else if constexpr (KIND == ASM_IN_COMPLEX) {
    r.a = ldg4((const float*)in.in0, 2 * idx);
    r.b = ldg4((const float*)in.in0, 2 * idx + 4);
  } else if constexpr (KIND == ASM_IN_COTANGENT) {
    r.a = ldg4((const float*)in.in0, 2 * idx);
    r.b = ldg4((const float*)in.in0, 2 * idx + 4);
    if (in.cot_abs) r.c = ldg4(in.cot_abs, idx);
    if (in.cot_target) r.d = ldg4(in.cot_target, idx);
    if (in.cot_angle) r.e = ldg4(in.cot_angle, idx);
    if (in.cot_abs2) r.f = ldg4(in.cot_abs2, idx);
  }
  return r;
}

// cotangent of |y| / angle(y) / |y|^2 / the fused amplitude-L2 term (see load_input); 1/|y| from the SFU
// reciprocal square root (2 ulp; the gradient gate is 1e-4)
__device__ __forceinline__ float2 cot_value(const RowIn& in, float2 y, float g_abs, float tgt, float g_angle,
                                            float g_abs2) {
  const float r2 = y.x * y.x + y.y * y.y;
  float2 acc = make_float2(0.0f, 0.0f);
  if (r2 > 0.0f) {
    const float inv_r = rsqrtf(r2);
    float g = 0.0f;
    if (in.cot_abs) g += g_abs;
    if (in.cot_target) g += in.cot_scale * (r2 * inv_r - tgt);
    const float gr = g * inv_r;
    acc.x = gr * y.x;
    acc.y = gr * y.y;
    if (in.cot_angle) {
      const float ga = g_angle * inv_r * inv_r;
      acc.x -= ga * y.y;
      acc.y += ga * y.x;
    }
  }
  if (in.cot_abs2) {
    const float g2 = 2.0f * g_abs2;
    acc.x += g2 * y.x;
    acc.y += g2 * y.y;
  }
  return acc;
}

template <int KIND>
__device__ __forceinline__ void make_input4(const RowIn& in, const In4& r, float2 (&x)[4]) {
  const float pa[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
  const float pb[4] = {r.b.x, r.b.y, r.b.z, r.b.w};
  if constexpr (KIND == ASM_IN_PHASE) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s, c;
      sincosf(__fmul_rn(in.phase_scale, pa[i]), &s, &c);
      x[i] = make_float2(c, s);
    }
  } else if constexpr (KIND == ASM_IN_AMP_PHASE) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s, c;
      sincosf(__fmul_rn(in.phase_scale, pa[i]), &s, &c);
      x[i] = make_float2(pb[i] * c, pb[i] * s);
    }
  } else if constexpr (KIND == ASM_IN_COMPLEX) {
    x[0] = make_float2(r.a.x, r.a.y);
    x[1] = make_float2(r.a.z, r.a.w);
    x[2] = make_float2(r.b.x, r.b.y);
    x[3] = make_float2(r.b.z, r.b.w);
  } else if constexpr (KIND == ASM_IN_COTANGENT) {
    const float2 y[4] = {make_float2(r.a.x, r.a.y), make_float2(r.a.z, r.a.w), make_float2(r.b.x, r.b.y),
                         make_float2(r.b.z, r.b.w)};
    const float ga[4] = {r.c.x, r.c.y, r.c.z, r.c.w}, tg[4] = {r.d.x, r.d.y, r.d.z, r.d.w};
    const float gg[4] = {r.e.x, r.e.y, r.e.z, r.e.w}, g2[4] = {r.f.x, r.f.y, r.f.z, r.f.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = cot_value(in, y[i], ga[i], tg[i], gg[i], g2[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = make_float2(0.0f, 0.0f);
  }
}

struct Aux4 {  // operands the epilogue reads back: loss target or forward phase (a), forward amplitude (b)
  float4 a, b;
};

template <int KIND>
__device__ __forceinline__ Aux4 fetch_aux4(const RowOut& o, size_t idx) {
  Aux4 r;
  r.a = r.b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if constexpr (KIND == ASM_OUT_ABS) {
    if (o.loss_target) r.a = ldg4(o.loss_target, idx);
  } else if constexpr (KIND == ASM_OUT_GRAD_PHASE) {
    r.a = ldg4(o.aux_phase, idx);
    if (o.aux_amp) r.b = ldg4(o.aux_amp, idx);
  }
  return r;
}

__device__ __forceinline__ void st4(void* p, size_t idx, float4 v) { reinterpret_cast<float4*>((float*)p + idx)[0] = v; }
// |v| with the SFU square root (relative error ~1e-7; the amplitude gate is 1e-5)
__device__ __forceinline__ float cabs_fast(float2 v) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v.x * v.x + v.y * v.y));
  return r;
}

// v[0..3] = un-normalised cropped field samples idx .. idx+3 (idx a multiple of 4)
template <int KIND>
__device__ __forceinline__ void store_output4(const RowOut& o, size_t idx, float2 (&v)[4], const Aux4& aux,
                                              float& loss_acc) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i].x *= o.scale;
    v[i].y *= o.scale;
  }
  if (o.save_field) {
    st4(o.save_field, 2 * idx, make_float4(v[0].x, v[0].y, v[1].x, v[1].y));
    st4(o.save_field, 2 * idx + 4, make_float4(v[2].x, v[2].y, v[3].x, v[3].y));
  }
  float r[4], q[4];
  if constexpr (KIND == ASM_OUT_ABS) {
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = cabs_fast(v[i]);
    st4(o.out0, idx, make_float4(r[0], r[1], r[2], r[3]));
    if (o.loss_target) {
      const float t[4] = {aux.a.x, aux.a.y, aux.a.z, aux.a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float d = r[i] - t[i];
        loss_acc += d * d;
      }
    }
  } else if constexpr (KIND == ASM_OUT_ANGLE) {
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = atan2f(v[i].y, v[i].x);
    st4(o.out0, idx, make_float4(r[0], r[1], r[2], r[3]));
  } else if constexpr (KIND == ASM_OUT_ABS_ANGLE) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r[i] = cabs_fast(v[i]);
      q[i] = atan2f(v[i].y, v[i].x);
    }
    st4(o.out0, idx, make_float4(r[0], r[1], r[2], r[3]));
    st4(o.out1, idx, make_float4(q[0], q[1], q[2], q[3]));
  } else if constexpr (KIND == ASM_OUT_COMPLEX) {
    st4(o.out0, 2 * idx, make_float4(v[0].x, v[0].y, v[1].x, v[1].y));
    st4(o.out0, 2 * idx + 4, make_float4(v[2].x, v[2].y, v[3].x, v[3].y));
  } else if constexpr (KIND == ASM_OUT_ABS2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = v[i].x * v[i].x + v[i].y * v[i].y;
    st4(o.out0, idx, make_float4(r[0], r[1], r[2], r[3]));
  } else if constexpr (KIND == ASM_OUT_GRAD_PHASE) {
    const float ph[4] = {aux.a.x, aux.a.y, aux.a.z, aux.a.w};
    const float am[4] = {aux.b.x, aux.b.y, aux.b.z, aux.b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s, cs;
      sincosf(__fmul_rn(o.phase_scale, ph[i]), &s, &cs);
      const float a = o.aux_amp ? am[i] : 1.0f;
      r[i] = o.phase_scale * a * (v[i].y * cs - v[i].x * s);
      q[i] = v[i].x * cs + v[i].y * s;
    }
    st4(o.out0, idx, make_float4(r[0], r[1], r[2], r[3]));
    if (o.out1) st4(o.out1, idx, make_float4(q[0], q[1], q[2], q[3]));
  }
}

// one sample at a time, straight from the registers of the last inverse butterfly (lanes hold consecutive
// samples: 4- and 8-byte accesses, coalesced); the same arithmetic as store_output4
template <int KIND>
__device__ __forceinline__ void store_output1(const RowOut& o, size_t idx, float2 v, float& loss_acc) {
  v.x *= o.scale;
  v.y *= o.scale;
  if (o.save_field) o.save_field[idx] = v;
  if constexpr (KIND == ASM_OUT_ABS) {
    const float r = cabs_fast(v);
    ((float*)o.out0)[idx] = r;
    if (o.loss_target) {
      const float d = r - __ldg(o.loss_target + idx);
      loss_acc += d * d;
    }
  } else if constexpr (KIND == ASM_OUT_ANGLE) {
    ((float*)o.out0)[idx] = atan2f(v.y, v.x);
  } else if constexpr (KIND == ASM_OUT_ABS_ANGLE) {
    ((float*)o.out0)[idx] = cabs_fast(v);
    ((float*)o.out1)[idx] = atan2f(v.y, v.x);
  } else if constexpr (KIND == ASM_OUT_COMPLEX) {
    ((float2*)o.out0)[idx] = v;
  } else if constexpr (KIND == ASM_OUT_ABS2) {
    ((float*)o.out0)[idx] = v.x * v.x + v.y * v.y;
  } else if constexpr (KIND == ASM_OUT_GRAD_PHASE) {
    float s, cs;
    sincosf(__fmul_rn(o.phase_scale, __ldg(o.aux_phase + idx)), &s, &cs);
    const float a = o.aux_amp ? __ldg(o.aux_amp + idx) : 1.0f;
    ((float*)o.out0)[idx] = o.phase_scale * a * (v.y * cs - v.x * s);
    if (o.out1) ((float*)o.out1)[idx] = v.x * cs + v.y * s;
  }
}

// fixed-order block reduction of the fused L2 partial sum; adds into loss_partial[blockIdx.x]
__device__ __forceinline__ void block_loss_reduce(float loss_acc, float* loss_partial, float* red /*[32]*/) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int off = 16; off > 0; off >>= 1) loss_acc += __shfl_down_sync(0xffffffffu, loss_acc, off);
  if ((tid & 31) == 0) red[tid >> 5] = loss_acc;
  __syncthreads();
  if (tid < 32) {
    float v = tid < ((nthr + 31) >> 5) ? red[tid] : 0.0f;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (tid == 0) loss_partial[blockIdx.x] += v;
  }
}

// ---- column pass arguments (generic and fast kernels) ---------------------------------------------
struct ColParams {
  Fft1d f;  // length Rp
  Phys ph;
  int logT;
  int S, D, n_colour, reduce;
  int in_full, out_full;  // 1: natural-order padded spectrum in global memory
  int R, pad_r, Cp;
  int use_h, flags;
  int two_buf;
  const float2* in;
  float2* out;
  const float* z;
  const float* wm;
  const int* depth_index;
  const int* col_perm;  // frequency bin of stored column c (NULL = natural order)
  float out_scale;
  // compile-time planned kernel only: w/mask grid in tile order and per-tile "inside the mask" flags
  const float* wmt;
  const int* tile_active;
  int blocked_in, blocked_out;  // layouts of the strip read (W1) and written (W2), see woff()
  int rows_skip_dead;           // the row kernels neither write nor read column tiles outside the mask
};

// Offset (in complex samples) of strip element (row r, stored column c) of the W1/W2 intermediates.
//   blocked == 0  plain [row][Cp]: the row kernels stream it, the column kernel touches 16 bytes per row
//                 per 2-column tile (measured: that pattern cost the column kernel 3.5 ms per C4 step);
//   blocked == b  [row/8][Cp/2^b][8 rows][2^b columns]: 8-row x 2^b-column blocks.  A column tile then
//                 reads/writes runs of 8 rows inside 128 * 2^(b-1) bytes, a row kernel pieces of 8 * 2^b
//                 bytes, and a row still lands in one contiguous 8-row band of DRAM.
// Rows are global (plane * R + r): R is a multiple of 8 whenever a blocked layout is selected.
// The WRITER of an intermediate picks the width so that it stores whole 32-byte sectors (a half-written
// sector costs a DRAM fill read plus two write-backs, measured 3x the algorithmic traffic): the row kernels
// write W1 with b = 2 (32 bytes of one row), the column kernel writes W2 with b = 1 (whole 128-byte lines);
// the readers take the 16-byte pieces that result.
__device__ __forceinline__ size_t woff(int blocked, int Cp, long long r, int c) {
  if (blocked) {
    const int b = blocked;
    return ((((size_t)(r >> 3) * (size_t)(Cp >> b) + (size_t)(c >> b)) << (3 + b)) + (size_t)(((int)r & 7) << b) +
            (size_t)(c & ((1 << b) - 1)));
  }
  return (size_t)r * Cp + c;
}

// offset of column c inside its row (woff(blocked, Cp, r, c) = woff(blocked, Cp, r, 0) + woff_in_row(blocked, c))
__device__ __forceinline__ int woff_in_row(int blocked, int c) {
  return blocked ? (((c >> blocked) << (3 + blocked)) + (c & ((1 << blocked) - 1))) : c;
}

// ---- compile-time planned kernels (fast_kernels.cu) ---------------------------------------------
// a plan exists for transform length n with `ext` non-pad samples and `pad` zeros on each side
bool fast_rows_supported(int n, int cols, int pad);
int fast_cols_logt(int n, int rows, int pad);  // log2(columns per tile) of the fast column kernel, -1 = none
// the same for the CTA-synchronous column kernel, which also takes / leaves natural-order spectra (the natural = 1
// variants of the row kernels below then keep W1 / W2 in natural column order)
int fast_cols_sync_logt(int n, int rows, int pad);
// scrambled position -> natural bin of the fast row / column transform (host copy)
void fast_rows_perm(int n, int* perm_out);
void fast_cols_perm(int n, int rows, int pad, int* perm_out);
int fast_wm_tiled(const Phys& ph, const float* wm, int n_colour, int logT, const int* row_perm, const int* col_perm,
                  float* wmt, int* tile_active, int sm_count, cudaStream_t stream);
// dead: per column tile of 2^dead_logt stored columns, 0 = every bin outside the mask (NULL = none known);
// such columns are not written by the forward row kernel and read as zeros by the inverse one
struct DeadCols {
  const int* active;
  int logt;
};
int fast_row_forward(int n, const float2* tw, const RowIn& in, long long n_rows, int C, int pad_c, float2* w1,
                     int blocked, DeadCols dead, int natural, int sm_count, cudaStream_t stream);
int fast_row_inverse(int n, const float2* tw, const RowOut& out, long long n_rows, int C, int pad_c,
                     const float2* w2, int blocked, DeadCols dead, int natural, int sm_count, int max_blocks,
                     cudaStream_t stream);
// fused K3 (forward call) + K1 (adjoint call) of the amplitude-L2 step, see fast_kernels.cu
struct FusedRows {
  const float* target;   // f32 [rows, C] amplitude target, or
  const unsigned char* target_u8;  // ... u8 [rows, C]: target = fl(v / 255) (exactly one of the two is set)
  float* amp_out;        // optional |y| output, f32 [rows, C]
  float scale;           // out_scale of the forward call
  float cot_scale;       // cotangent = cot_scale * (|y| - target) * y / |y|
  float* loss_partial;   // block partials of sum((|y| - target)^2)
};
int fast_row_inverse_forward(int n, const float2* tw, const FusedRows& f, long long n_rows, int C, int pad_c,
                             const float2* w2, int blocked_in, float2* w1, int blocked_out, DeadCols dead,
                             int sm_count, int max_blocks, cudaStream_t stream);
bool fast_row_inverse_uses_tma(int n, int C, int pad_c, long long n_rows, int blocked);
int fast_columns(const ColParams& p, int sm_count, cudaStream_t stream);

}  // namespace asmb
