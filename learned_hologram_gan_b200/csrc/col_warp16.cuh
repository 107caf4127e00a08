// col_warp16.cuh -- warp-local column kernel for N = 16 * 8 * 8 = 1024 (BASELINE configs 2/3: 384 rows padded
// by 320), four columns per tile, two CTAs per SM.
//
// Same structure as col_warp.cuh: the radix-16 pass runs across the CTA, after it every block of 64
// consecutive positions is an independent 64-point transform (8 x 8).  A warp owns two blocks of all four
// columns and runs both radix-8 passes, the transfer-function multiply and the matching inverse passes on
// them with __syncwarp() only; one CTA barrier per transform, two exchange buffers used alternately.
// pad = 320 = 5 * 64: butterfly j of the radix-16 pass sees its non-pad samples at k in [5, 11).
// Shared-memory layout [position][4 columns] with one padding position after every 8, so that the stride-8
// accesses of the last radix-8 pass fall into different banks.
#pragma once
#include "common.cuh"
#include "fft_fast.cuh"

namespace asmb {

template <int NT>
__global__ void __launch_bounds__(NT, 2) col_warp16_kernel(ColParams a) {
  constexpr int N = 1024, R0 = 16, M0 = 64, L = 64, R1 = 8, R2 = 8, T = 4, LOGT = 2;
  constexpr int KLO = 5, KHI = 11;  // pad = KLO * M0 = 320
  constexpr int NWARP = NT / 32, BPW = R0 / NWARP;  // blocks per warp
  static_assert(NT == M0 * T, "one radix-16 butterfly per thread");
  static_assert(R0 % NWARP == 0, "blocks per warp");
  constexpr int NELP = (N + N / 8) * T;             // padded elements per buffer
  constexpr int TAB1 = (R1 - 1) * R2;               // W_64^(j q), q = 1..7, j < 8
  extern __shared__ float2 smem[];
  float2* const bufA = smem;
  float2* const bufB = bufA + NELP;
  float2* const bufX = bufB + NELP;
  float2* const tab1 = bufX + NELP;
  float2* const tab0 = tab1 + TAB1;                 // W_N^j, j < M0
  float* const sbeta = reinterpret_cast<float*>(tab0 + M0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float2* __restrict__ tw = a.f.tw;
  const int tiles_per_plane = a.Cp >> LOGT;
  const long long n_tiles = (long long)a.S * a.n_colour * tiles_per_plane;
  const int R = a.R, Cp = a.Cp;
  const int use_h = a.use_h;
  const bool masked = (a.flags & kFilterMask) != 0;
  const float bsign = (a.flags & kFilterConj) ? -1.0f : 1.0f;
  const size_t strip = (size_t)R * Cp;
  auto phys = [](int pos, int t) { return ((pos + (pos >> 3)) << LOGT) + t; };

  for (int e = tid; e < TAB1; e += NT) {
    const int q = e / R2 + 1, j = e - (q - 1) * R2;
    tab1[e] = __ldg(tw + (size_t)(j * q) * R0);
  }
  for (int e = tid; e < M0; e += NT) tab0[e] = __ldg(tw + e);
  __syncthreads();

  // ---- radix-16 pass across the CTA: column t0 = tid & 3, butterfly j0 = tid >> 2 ----------------------
  const int j0 = tid >> LOGT, t0 = tid & (T - 1);
  auto kstride = [&](int b) { return b ? (((long long)(M0 / 8) * (Cp >> b)) << (3 + b)) : (long long)M0 * Cp; };
  const long long kstr_in = kstride(a.blocked_in), kstr_out = kstride(a.blocked_out);
  long long off_in = 0, off_out = 0;  // offset of k = KLO of the current tile
  auto twiddles16 = [&](float2 (&w)[16]) {
    w[0] = make_float2(1.0f, 0.0f);
    w[1] = tab0[j0];
    tw_chain_step<16, 2>(w);
  };
  auto stage_inputs = [&](const float2* __restrict__ src, float2* buf) {
#pragma unroll
    for (int k = KLO; k < KHI; ++k) cp_async8(buf + phys(j0 + k * M0, t0), src + (off_in + (k - KLO) * kstr_in));
    cp_async_commit();
  };
  auto pass0_forward = [&](const float2* __restrict__ src, float2* buf, bool staged) {
    float2 v[16], w[16];
    if (staged) cp_async_wait_all();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k < KLO || k >= KHI) v[k] = make_float2(0.0f, 0.0f);
      else if (staged) v[k] = buf[phys(j0 + k * M0, t0)];
      else v[k] = __ldg(src + (off_in + (k - KLO) * kstr_in));
    }
    DftPruned<16, KLO, KHI>::run(v);
    twiddles16(w);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      if (q > 0) v[q] = cmul(v[q], w[q]);
      buf[phys(j0 + q * M0, t0)] = v[q];
    }
  };
  auto pass0_inverse = [&](const float2* buf, float2* __restrict__ dst) {
    float2 v[16], w[16];
    twiddles16(w);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      v[q] = buf[phys(j0 + q * M0, t0)];
      if (q > 0) v[q] = cmul(v[q], w[q]);
    }
    Dft<16>::run(v);
    dst += off_out;
#pragma unroll
    for (int k = KLO; k < KHI; ++k, dst += kstr_out) *dst = cswap(v[k]);  // the other outputs are dead code
  };

  // ---- warp-local passes: lane -> column lt = lane & 3, butterfly lj = lane >> 2 of block warp*BPW + i ----
  const int lt = lane & (T - 1), lj = lane >> LOGT;
  auto pass1 = [&](float2* buf, int bbase, auto dit_tag) {
    constexpr bool DIT = decltype(dit_tag)::value;
    float2 v[R1];
#pragma unroll
    for (int k = 0; k < R1; ++k) v[k] = buf[phys(bbase + lj + k * R2, lt)];
    if (!DIT) Dft<R1>::run(v);
#pragma unroll
    for (int q = 1; q < R1; ++q) v[q] = cmul(v[q], tab1[(q - 1) * R2 + lj]);
    if (DIT) Dft<R1>::run(v);
#pragma unroll
    for (int k = 0; k < R1; ++k) buf[phys(bbase + lj + k * R2, lt)] = v[k];
  };

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int ct = (int)(tile % tiles_per_plane);
    const long long g = tile / tiles_per_plane;  // sample * n_colour + colour
    const int colour = (int)(g % a.n_colour);
    const long long s = g / a.n_colour;
    const int col0 = ct << LOGT;
    off_in = (long long)woff(a.blocked_in, Cp, j0, col0 + t0);
    off_out = (long long)woff(a.blocked_out, Cp, j0, col0 + t0);

    if (masked && a.tile_active && !a.tile_active[ct]) {
      const int n_out = a.rows_skip_dead ? 0 : (a.reduce ? 1 : a.D);
      for (int d = 0; d < n_out; ++d) {
        const size_t plane = a.reduce ? (size_t)g : ((size_t)s * a.D + d) * a.n_colour + colour;
        float2* dst = a.out + plane * strip;
        for (int e = tid; e < R * T; e += NT) dst[woff(a.blocked_out, Cp, e >> LOGT, col0 + (e & (T - 1)))] = make_float2(0.0f, 0.0f);
      }
      continue;
    }

    for (int d = tid; d < a.D; d += NT) {
      const int zi = a.depth_index ? a.depth_index[s * a.D + d] : d;
      sbeta[d] = use_h ? bsign * beta_of(a.z[zi]) : 0.0f;
    }
    // w (sign bit = outside the mask) of the 2 x 8 bins this lane owns in the last radix-8 pass
    float wreg[BPW][R2];
#pragma unroll
    for (int i = 0; i < BPW; ++i) {
      const int bbase = (warp * BPW + i) * L;
#pragma unroll
      for (int k = 0; k < R2; ++k)
        wreg[i][k] = a.wmt ? __ldg(a.wmt + ((size_t)colour * tiles_per_plane + ct) * (N * T) + ((bbase + lj * R2 + k) << LOGT) + lt)
                           : 0.0f;
    }

    // radix-8 (M = 1) butterfly of block i on this lane's bins; ld/st functors see (i, k)
    auto pass2_dif = [&](float2* buf, auto&& sink) {
#pragma unroll
      for (int i = 0; i < BPW; ++i) {
        const int bbase = (warp * BPW + i) * L;
        float2 v[R2];
#pragma unroll
        for (int k = 0; k < R2; ++k) v[k] = buf[phys(bbase + lj * R2 + k, lt)];
        Dft<R2>::run(v);
#pragma unroll
        for (int k = 0; k < R2; ++k) sink(i, k, bbase + lj * R2 + k, v[k]);
      }
    };
    auto pass2_dit = [&](float2* buf, auto&& source) {
#pragma unroll
      for (int i = 0; i < BPW; ++i) {
        const int bbase = (warp * BPW + i) * L;
        float2 v[R2];
#pragma unroll
        for (int k = 0; k < R2; ++k) v[k] = source(i, k, bbase + lj * R2 + k);
        Dft<R2>::run(v);
#pragma unroll
        for (int k = 0; k < R2; ++k) buf[phys(bbase + lj * R2 + k, lt)] = v[k];
      }
    };
    auto both_blocks_pass1 = [&](float2* buf, auto dit_tag) {
#pragma unroll
      for (int i = 0; i < BPW; ++i) pass1(buf, (warp * BPW + i) * L, dit_tag);
      __syncwarp();
    };

    if (!a.reduce) {
      pass0_forward(a.in + (size_t)g * strip, bufA, false);
      __syncthreads();
      both_blocks_pass1(bufA, std::false_type{});
      pass2_dif(bufA, [&](int i, int k, int pos, float2 v) {
        if (masked && signbit(wreg[i][k])) v = make_float2(0.0f, 0.0f);
        bufX[phys(pos, lt)] = v;
      });
      for (int d = 0; d < a.D; ++d) {
        const size_t out_plane = ((size_t)s * a.D + d) * a.n_colour + colour;
        float2* buf = (d & 1) ? bufA : bufB;
        const float beta = sbeta[d], beta_t = beta * 0.15915494309189535f;
        pass2_dit(buf, [&](int i, int k, int pos) {
          float2 x = bufX[phys(pos, lt)];
          if (use_h) x = cmul(x, fast_cis_bw(beta, beta_t, fabsf(wreg[i][k])));
          return cswap(x);
        });
        __syncwarp();
        both_blocks_pass1(buf, std::true_type{});
        __syncthreads();
        pass0_inverse(buf, a.out + out_plane * strip);
      }
    } else {
      for (int d = 0; d < a.D; ++d) {
        const size_t in_plane = ((size_t)s * a.D + d) * a.n_colour + colour;
        const float2* src = a.in + in_plane * strip;
        float2* buf = (d & 1) ? bufB : bufA;
        pass0_forward(src, buf, d > 0);
        __syncthreads();
        if (d + 1 < a.D) stage_inputs(src + (size_t)a.n_colour * strip, (d & 1) ? bufA : bufB);
        both_blocks_pass1(buf, std::false_type{});
        const float beta = sbeta[d], beta_t = beta * 0.15915494309189535f;
        pass2_dif(buf, [&](int i, int k, int pos, float2 v) {
          if (use_h) v = cmul(v, fast_cis_bw(beta, beta_t, fabsf(wreg[i][k])));
          if (d > 0) v = cadd(v, bufX[phys(pos, lt)]);
          bufX[phys(pos, lt)] = v;
        });
      }
      float2* buf = ((a.D - 1) & 1) ? bufB : bufA;
      pass2_dit(buf, [&](int i, int k, int pos) {
        float2 x = bufX[phys(pos, lt)];
        if (masked && signbit(wreg[i][k])) x = make_float2(0.0f, 0.0f);
        return cswap(x);
      });
      __syncwarp();
      both_blocks_pass1(buf, std::true_type{});
      __syncthreads();
      pass0_inverse(buf, a.out + (size_t)g * strip);
    }
    __syncthreads();  // the next tile's radix-16 pass rewrites bufA
  }
}

}  // namespace asmb
