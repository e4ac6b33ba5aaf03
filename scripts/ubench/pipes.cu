// Micro-benchmark of the per-SMSP issue cost (cycles per warp-instruction) of the instructions the
// attention math warps lean on: MUFU.EX2 (f32 and f16x2), F2FP packs, HADD2.F32 unpack, FFMA, LDS.
// One CTA on one SM, W warps per scheduler; each warp runs N independent chains (ILP 8).
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 2048

template <int OP>
__global__ void k(float* out, long long* cyc, int warps) {
  __shared__ uint32_t sm[4096];
  float x[8];
  uint32_t u[8];
  for (int i = 0; i < 8; ++i) { x[i] = -1.0f - threadIdx.x * 1e-3f - i; u[i] = threadIdx.x * 77 + i; }
  sm[threadIdx.x] = threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }
      if (OP == 1) { asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i])); }
      if (OP == 2) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[i]), "f"(x[(i + 1) & 7])); x[i] += __uint_as_float(u[i]); }
      if (OP == 3) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[i]), "f"(x[(i + 1) & 7])); x[i] += __uint_as_float(u[i]); }
      if (OP == 4) { __half2 h = *reinterpret_cast<__half2*>(&u[i]); float2 f = __half22float2(h); x[i] += f.x; u[i] += __float_as_uint(f.y); }
      if (OP == 5) { x[i] = fmaf(x[i], 1.0001f, 0.5f); }
      if (OP == 6) { u[i] = sm[(u[i] + threadIdx.x) & 4095]; }
      if (OP == 7) { u[i] = __funnelshift_r(u[i], u[(i + 1) & 7], 16); }
      if (OP == 8) { x[i] = fmaxf(x[i], x[(i + 3) & 7] + 1.f); }
      if (OP == 9) { u[i] = __byte_perm(u[i], u[(i + 1) & 7], 0x5432); }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(u[i]);
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int extra) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
  for (int wps : {1, 2, 4}) {
    int threads = wps * 4 * 32;
    k<OP><<<1, threads>>>(out, cyc, wps);
    k<OP><<<1, threads>>>(out, cyc, wps);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (ITERS * 8.0 * (1 + extra));
    printf("%-22s warps/SMSP %d : %7.2f cycles per warp-instr per warp, %6.2f per SMSP issue\n", name, wps, per, per / wps);
  }
}

int main() {
  run<0>("MUFU.EX2 f32", 0);
  run<1>("MUFU.EX2 f16x2", 0);
  run<2>("F2FP.F16 pack (+FADD)", 1);
  run<3>("F2FP.BF16 pack (+FADD)", 1);
  run<4>("HADD2.F32 x2 (+2 add)", 3);
  run<5>("FFMA", 0);
  run<6>("LDS.32 (dependent)", 0);
  run<7>("SHF funnel", 0);
  run<8>("FADD+FMNMX", 1);
  run<9>("PRMT", 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
