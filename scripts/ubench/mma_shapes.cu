// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) as a function of N
// and of the operand majors / sources.  One CTA per SM, one thread issues REP chains of 64 MMAs on
// fixed (garbage) operands and waits for the commit; smem tiles are laid out as the attention kernels
// use them (128B swizzle, [128 x 64] 16-bit tiles).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../musicgeneration_b200/csrc/tc_common.cuh"

using namespace mt;

namespace mt { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

constexpr int TILE = 16384;

// MODE 0: A K-major, B K-major   1: A K-major, B MN-major   2: A MN-major, B MN-major   3: A TMEM, B MN-major
// 4: A TMEM, B K-major
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, int N, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  for (int i = threadIdx.x; i < 8 * TILE / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const int a_mn = MODE == 2, b_mn = (MODE == 1 || MODE == 2 || MODE == 3);
    const uint32_t idesc = tc::make_idesc(128, N, 1, 1, a_mn, b_mn);
    const uint64_t ad = a_mn ? tc::make_sdesc(tc::smem_u32(smem), TILE, 1024) : tc::make_sdesc(tc::smem_u32(smem), 16, 1024);
    const uint64_t bd = b_mn ? tc::make_sdesc(tc::smem_u32(smem + 4 * TILE), TILE, 1024)
                             : tc::make_sdesc(tc::smem_u32(smem + 4 * TILE), 16, 1024);
    long long best = 1ll << 60;
    for (int r = 0; r < reps; ++r) {
      long long t0 = clock64();
      for (int i = 0; i < 64; ++i) {
        const int k = i & 3;      // 4 k-steps of a 64-deep tile, over and over
        const uint64_t a_k = a_mn ? ad + 128 * k : ad + 2 * k;
        const uint64_t b_k = b_mn ? bd + 128 * k : bd + 2 * k;
        if (MODE >= 3) tc::umma_f16_ts(tmem, tmem + 256 + 8 * k, b_k, idesc, i != 0);
        else tc::umma_f16(tmem, a_k, b_k, idesc, i != 0);
      }
      tc::umma_commit(&bar);
      tc::mbar_wait(&bar, r & 1);
      long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    if (blockIdx.x == 0) out[0] = best;
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

template <int MODE>
void run(const char* name, long long* d) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * TILE + 1024);
  for (int N : {64, 128, 256}) {
    if (MODE == 2 && N > 128) continue;
    k<MODE><<<148, 128, 8 * TILE + 1024>>>(d, N, 8);
    long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-28s N=%3d : %6.1f cycles per MMA (64 chained, incl. commit latency)  %s\n", name, N, h / 64.0,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  run<0>("A smem K-major, B K-major", d);
  run<1>("A smem K-major, B MN-major", d);
  run<2>("A smem MN-major, B MN-major", d);
  run<3>("A TMEM, B MN-major", d);
  run<4>("A TMEM, B K-major", d);
  return 0;
}
