// Micro-benchmark: tcgen05.ld throughput (TMEM -> registers).  One CTA per SM, W warps (4, 8, 16) each
// reading 32 lanes x 32 columns (4 KB per instruction) ITERS times from its own lane quadrant.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../musicgeneration_b200/csrc/tc_common.cuh"

using namespace mt;
namespace mt { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

#define ITERS 512

template <int X32>
__global__ void k(long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
    if (X32) {
      uint32_t r[32];
      tc::tmem_ld_32x32(tmem + ((i * 32) & 511 & ~31u) % 480, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; ++x) acc ^= r[x];
    } else {
      uint32_t r[16];
      tc::tmem_ld_32x16(tmem + ((i * 16) & 511) % 480, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 16; ++x) acc ^= r[x];
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(slot, 512);
}

int main() {
  long long* d; uint32_t* s;
  cudaMalloc(&d, 8); cudaMalloc(&s, 148 * 1024 * 4);
  for (int warps : {4, 8, 16}) {
    for (int x32 = 1; x32 >= 0; --x32) {
      if (x32) k<1><<<148, warps * 32>>>(d, s); else k<0><<<148, warps * 32>>>(d, s);
      long long h = 0;
      cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)warps * ITERS * 32 * (x32 ? 32 : 16) * 4;
      printf("%2d warps, 32x32b.x%d : %7.1f cycles per load per warp, %6.1f B/clk/SM  %s\n", warps, x32 ? 32 : 16,
             (double)h / ITERS, bytes / h, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
