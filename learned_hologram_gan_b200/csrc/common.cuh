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
};

// ---- compile-time planned kernels (fast_kernels.cu) ---------------------------------------------
// a plan exists for transform length n with `ext` non-pad samples and `pad` zeros on each side
bool fast_rows_supported(int n, int cols, int pad);
int fast_cols_logt(int n, int rows, int pad);  // log2(columns per tile) of the fast column kernel, -1 = none
// scrambled position -> natural bin of the fast row / column transform (host copy)
void fast_rows_perm(int n, int* perm_out);
void fast_cols_perm(int n, int* perm_out);
int fast_wm_tiled(const Phys& ph, const float* wm, int n_colour, int logT, const int* row_perm, const int* col_perm,
                  float* wmt, int* tile_active, int sm_count, cudaStream_t stream);
int fast_row_forward(int n, const float2* tw, const RowIn& in, long long n_rows, int C, int pad_c, float2* w1,
                     int sm_count, cudaStream_t stream);
int fast_row_inverse(int n, const float2* tw, const RowOut& out, long long n_rows, int C, int pad_c,
                     const float2* w2, int sm_count, int max_blocks, cudaStream_t stream);
int fast_columns(const ColParams& p, int sm_count, cudaStream_t stream);

}  // namespace asmb
