// Micro-benchmark: what does it cost to flush a [128 x 64] fp32 accumulator block per pipeline step into
// global memory with reductions?  (Design input for a relative-attention backward kernel that owns a query
// tile and adds its per-step dE block -- or an FA-style dQ block -- to a global fp32 buffer.)
//   variant 0: red.global.add.v4.f32 from registers, 512 threads x 4 per flush (hot 512 KB target, as dE)
//   variant 1: same, target spread over 64 MB (as a dQ buffer)
//   variant 2: cp.reduce.async.bulk (32 KB from shared memory, one thread), hot target
//   variant 3: cp.reduce.async.bulk, spread target
// Grid = 148 CTAs x 512 threads (one per SM), FLUSHES flushes per CTA.  Prints us per flush per SM and the
// aggregate reduce payload rate.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define FLUSHES 512

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int V>
__global__ void __launch_bounds__(512, 1) k(float* dst, size_t region_blocks, float* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* stage = reinterpret_cast<float*>(smem);
  const int a = threadIdx.x & 127, q = threadIdx.x >> 7;
  float r[16];
  for (int x = 0; x < 16; ++x) r[x] = 1e-6f * (threadIdx.x + x);
  if (V >= 2) {
    for (int x = threadIdx.x; x < 8192; x += 512) stage[x] = 1e-6f * x;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  for (int f = 0; f < FLUSHES; ++f) {
    const size_t blk = ((size_t)blockIdx.x * 7 + (size_t)f * 13) % region_blocks;
    float* base = dst + blk * 8192;
    if (V < 2) {
      float* d = base + a * 64 + 16 * q;
#pragma unroll
      for (int x = 0; x < 16; x += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + x), "f"(r[x]), "f"(r[x + 1]), "f"(r[x + 2]),
                     "f"(r[x + 3]) : "memory");
    } else {
      if (threadIdx.x == 0) {
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(base),
                     "r"(smem_u32(stage)), "r"(32768) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncthreads();
    }
  }
  if (V >= 2 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (r[3] == 123.f) sink[0] = r[0];
}

template <int V>
void run(const char* name, float* dst, size_t region_blocks, float* sink) {
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<148, 512, 65536>>>(dst, region_blocks, sink);
  cudaEventRecord(e0);
  k<V><<<148, 512, 65536>>>(dst, region_blocks, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  printf("%-44s %8.3f ms  %7.3f us/flush/SM (%6.0f cycles @1.965GHz)  %7.1f GB/s payload  [%s]\n", name, ms,
         ms * 1e3 / FLUSHES, ms * 1e3 / FLUSHES * 1965.0, 148.0 * FLUSHES * 32768 / (ms * 1e-3) / 1e9, cudaGetErrorString(e));
}

int main() {
  float* dst; float* sink;
  const size_t big = 2048;      // 2048 blocks x 32 KB = 64 MB
  cudaMalloc(&dst, big * 32768); cudaMalloc(&sink, 16);
  cudaMemset(dst, 0, big * 32768);
  run<0>("red.v4 from registers, hot 512 KB", dst, 16, sink);
  run<1>("red.v4 from registers, spread 64 MB", dst, big, sink);
  run<2>("cp.reduce.async.bulk 32 KB, hot 512 KB", dst, 16, sink);
  run<3>("cp.reduce.async.bulk 32 KB, spread 64 MB", dst, big, sink);
  return 0;
}
