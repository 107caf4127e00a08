// physics.cuh -- fp32-faithful frequency grids, circular mask and transfer function.
//
// The reference evaluates these with separately rounded fp32 torch ops; its fp32 H is
// 4e-4 (relative L2) away from the exact H, so "more accurate" arithmetic FAILS parity.
// Every operation below is therefore an explicitly rounded intrinsic (__fmul_rn, __fadd_rn,
// __fsqrt_rn, __fdiv_rn: never contracted into FMA) in the reference's evaluation order:
//   fftfreq          k * fl32(1/(n*d))                         asm.py:56-57, util.py:232-233
//   w                sqrt(max(1/l^2 - (fx^2 + fy^2), 0))       asm.py:163-171
//   theta            fl(fl(fl32(-2 pi) * z) * w)               asm.py:206-211
//   H                (cos theta, sin theta)
//   radial           sqrt(u^2 + v^2) * min(Rp,Cp)  > radius    util.py:234-241
#pragma once
#include <cuda_runtime.h>

namespace asmb {

constexpr int kMaxColour = 4;

struct Phys {
  int Rp, Cp;
  float fscale_r, fscale_c;  // fl32(1/(Rp*pitch)), fl32(1/(Cp*pitch))
  float uscale_r, uscale_c;  // fl32(1/Rp), fl32(1/Cp)
  float short_edge;          // (float)min(Rp,Cp)
  float radius;              // (float)mask_radius
  float inv_l2[kMaxColour];  // fl(1/fl(lambda^2))
  float lambda[kMaxColour];
  float two_d_r, two_d_c;    // fl32(2/(Rp*pitch)), fl32(2/(Cp*pitch))  (band limit, asm.py:174-183)
};

__device__ __forceinline__ float signed_bin(int k, int n) {
  return (float)(k < ((n + 1) >> 1) ? k : k - n);
}

__device__ __forceinline__ float w_value(const Phys& p, int kr, int kc, int colour) {
  const float fx = __fmul_rn(signed_bin(kr, p.Rp), p.fscale_r);
  const float fy = __fmul_rn(signed_bin(kc, p.Cp), p.fscale_c);
  const float sq = __fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy));
  const float d = __fsub_rn(p.inv_l2[colour], sq);
  return __fsqrt_rn(fmaxf(d, 0.0f));
}

__device__ __forceinline__ float radial_value(const Phys& p, int kr, int kc) {
  const float u = __fmul_rn(signed_bin(kr, p.Rp), p.uscale_r);
  const float v = __fmul_rn(signed_bin(kc, p.Cp), p.uscale_c);
  return __fmul_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v))), p.short_edge);
}

// beta = fl(fl32(-2 pi) * z), computed once per depth
__device__ __forceinline__ float beta_of(float z) { return __fmul_rn(-6.2831854820251465f, z); }

__device__ __forceinline__ float2 h_value(const Phys& p, int kr, int kc, int colour, float beta) {
  const float theta = __fmul_rn(beta, w_value(p, kr, kc, colour));
  float s, c;
  sincosf(theta, &s, &c);
  return make_float2(c, s);
}

enum { kFilterConj = 1, kFilterMask = 2 };

// complete spectral filter value at natural frequency bin (kr, kc).
// wm != nullptr: caller-provided grid, |wm| = w, sign bit = outside the circular mask.
__device__ __forceinline__ float2 filter_value(const Phys& p, const float* __restrict__ wm, int use_h,
                                               int flags, int kr, int kc, int colour, float beta) {
  float2 f = make_float2(1.0f, 0.0f);
  if (wm) {
    const float v = __ldg(wm + ((size_t)colour * p.Rp + kr) * p.Cp + kc);
    if ((flags & kFilterMask) && signbit(v)) return make_float2(0.0f, 0.0f);
    if (use_h) {
      float s, c;
      sincosf(__fmul_rn(beta, fabsf(v)), &s, &c);
      f = make_float2(c, (flags & kFilterConj) ? -s : s);
    }
    return f;
  }
  if (use_h) {
    f = h_value(p, kr, kc, colour, beta);
    if (flags & kFilterConj) f.y = -f.y;
  }
  if (flags & kFilterMask) {
    if (radial_value(p, kr, kc) > p.radius) f = make_float2(0.0f, 0.0f);
  }
  return f;
}

}  // namespace asmb
