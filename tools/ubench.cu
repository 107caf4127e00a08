// ubench.cu -- pipe-throughput microbenchmarks that decide how the FFT butterflies are written
// (scalar FADD/FFMA vs the packed FADD2/FFMA2 of sm_100, shared-memory and SFU rates).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

typedef unsigned long long u64;
#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* __restrict__ in, int iters) {
  __shared__ float2 sm[2048];
  const int tid = threadIdx.x;
  for (int i = tid; i < 2048; i += 256) sm[i] = make_float2(i * 1e-3f, 1.0f);
  __syncthreads();
  float a[16], x[8];
  u64 p[8], px[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = in[tid + i];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = in[tid + 32 + i];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = ((u64)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
#pragma unroll
  for (int i = 0; i < 4; ++i) px[i] = ((u64)__float_as_uint(x[2 * i]) << 32) | __float_as_uint(x[2 * i + 1]);
  unsigned saddr = (unsigned)__cvta_generic_to_shared(sm) + tid * 8;
  unsigned saddr16 = (unsigned)__cvta_generic_to_shared(sm) + (tid & 31) * 16;
  int q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = tid + i;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // scalar FADD reg,reg
#define X(i) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(x[i])); asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i + 8]) : "f"(x[(i + 3) & 7]));
      REP8(X)
#undef X
    } else if (MODE == 1) {  // packed FADD2
#define X(i) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(px[i & 3])); asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p[(i + 4) & 7]) : "l"(px[(i + 1) & 3]));
      REP8(X)
#undef X
    } else if (MODE == 2) {  // scalar FFMA, three distinct registers
#define X(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(x[i]), "f"(x[(i + 1) & 7])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i + 8]) : "f"(x[(i + 2) & 7]), "f"(x[(i + 5) & 7]));
      REP8(X)
#undef X
    } else if (MODE == 3) {  // packed FFMA2
#define X(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(px[i & 3]), "l"(px[(i + 1) & 3])); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[(i + 4) & 7]) : "l"(px[(i + 2) & 3]), "l"(px[(i + 3) & 3]));
      REP8(X)
#undef X
    } else if (MODE == 4) {  // scalar FMUL reg,reg
#define X(i) asm volatile("mul.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(x[i])); asm volatile("mul.f32 %0, %0, %1;" : "+f"(a[i + 8]) : "f"(x[(i + 3) & 7]));
      REP8(X)
#undef X
    } else if (MODE == 5) {  // FADD + FFMA alternating
#define X(i) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(x[i])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i + 8]) : "f"(x[(i + 2) & 7]), "f"(x[(i + 5) & 7]));
      REP8(X)
#undef X
    } else if (MODE == 6) {  // LDS.64 only (results unused)
#define X(i) { u64 v; asm volatile("ld.shared.b64 %0, [%1+" #i "*2048];" : "=l"(v) : "r"(saddr)); } { u64 v; asm volatile("ld.shared.b64 %0, [%1+" #i "*2048];" : "=l"(v) : "r"(saddr)); }
      REP8(X)
#undef X
    } else if (MODE == 7) {  // STS.64
#define X(i) asm volatile("st.shared.b64 [%0+" #i "*2048], %1;" :: "r"(saddr), "l"(p[i])); asm volatile("st.shared.b64 [%0+" #i "*2048], %1;" :: "r"(saddr), "l"(px[i & 3]));
      REP8(X)
#undef X
    } else if (MODE == 8) {  // FADD2 + LDS.64 1:1
#define X(i) { u64 v; asm volatile("ld.shared.b64 %0, [%1+" #i "*2048];" : "=l"(v) : "r"(saddr)); asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(v)); }
      REP8(X) REP8(X)
#undef X
    } else if (MODE == 9) {  // MUFU sin + cos (approx)
#define X(i) asm volatile("sin.approx.f32 %0, %0;" : "+f"(a[i])); asm volatile("cos.approx.f32 %0, %0;" : "+f"(a[i + 8]));
      REP8(X)
#undef X
    } else if (MODE == 10) {  // FADD + IADD 1:1
#define X(i) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(x[i])); asm volatile("add.s32 %0, %0, %1;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]));
      REP8(X)
#undef X
    } else if (MODE == 11) {  // FADD + LDS.64 2:1 (scalar complex add fed from shared memory)
#define X(i) { float v0, v1; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+" #i "*2048];" : "=f"(v0), "=f"(v1) : "r"(saddr)); asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(v0)); asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i + 8]) : "f"(v1)); }
      REP8(X)
#undef X
    } else if (MODE == 12) {  // LDS.128 only
#define X(i) { unsigned v0, v1, v2, v3; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4+" #i "*512];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(saddr16)); }
      REP8(X) REP8(X)
#undef X
    } else if (MODE == 13) {  // FMUL2
#define X(i) asm volatile("mul.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(px[i & 3])); asm volatile("mul.f32x2 %0, %0, %1;" : "+l"(p[(i + 4) & 7]) : "l"(px[(i + 1) & 3]));
      REP8(X)
#undef X
    } else if (MODE == 14) {  // FADD2 + IADD 1:1
#define X(i) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(px[i & 3])); asm volatile("add.s32 %0, %0, %1;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]));
      REP8(X)
#undef X
    } else if (MODE == 15) {  // FADD2 + FADD2 + MOV-like (PRMT) 2:1
#define X(i) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(px[i & 3])); asm volatile("prmt.b32 %0, %0, %1, 0x1032;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]));
      REP8(X)
#undef X
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (float)(p[i] & 0xffff) + q[i];
  out[blockIdx.x * 256 + tid] = s;
}

template <int MODE>
void run(const char* name, int instr_per_iter, int ctas_per_sm) {
  int sms = 148, khz = 1965000;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int iters = 8192, grid = sms * ctas_per_sm;
  float *out, *in;
  cudaMalloc(&out, sizeof(float) * grid * 256);
  cudaMalloc(&in, sizeof(float) * 4096);
  cudaMemset(in, 0, sizeof(float) * 4096);
  k<MODE><<<grid, 256>>>(out, in, iters);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, in, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double total = (double)grid * 8 * iters * instr_per_iter;
  printf("%-22s ctas/SM %d  ms %.3f  warp-instr/clk/SM %.3f (at %d kHz)  err %s\n", name, ctas_per_sm, ms,
         total / (ms * 1e-3) / sms / (khz * 1e3), khz, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(in);
}

int main() {
  for (int c : {2, 4}) {
    run<0>("FADD r,r", 16, c);
    run<1>("FADD2", 16, c);
    run<2>("FFMA 3-reg", 16, c);
    run<3>("FFMA2", 16, c);
    run<4>("FMUL r,r", 16, c);
    run<13>("FMUL2", 16, c);
    run<5>("FADD+FFMA", 16, c);
    run<6>("LDS.64", 16, c);
    run<12>("LDS.128", 16, c);
    run<7>("STS.64", 16, c);
    run<8>("FADD2+LDS.64 1:1", 32, c);
    run<11>("2FADD+LDS.64", 24, c);
    run<9>("MUFU sin+cos", 16, c);
    run<10>("FADD+IADD", 16, c);
    run<14>("FADD2+IADD", 16, c);
    run<15>("FADD2+PRMT", 16, c);
  }
  return 0;
}
