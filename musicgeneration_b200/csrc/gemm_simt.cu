// Reference-precision GEMM (FFMA, fp32 accumulate) for every nn.Linear on the path
// (MT/layers.py:72,77,82,108,157,158; MT/network.py:39) -- the fp32 parity mode, and the
// fallback-free path for shapes the tcgen05 kernel does not take (e.g. tiny test shapes).
// C[M,N] = epi(op(A)[M,K] . op(B)[K,N]); 128x128x16 block tile, 8x8 register tile,
// deterministic split-K through a caller workspace for skinny outputs (weight gradients).
#include "ops.cuh"

namespace mt {

constexpr int GS_BM = 128, GS_BN = 128, GS_BK = 16, GS_THREADS = 256;

struct GemmArgs {
  const void* A; const void* B; void* C;
  const float* bias; const float* addend; const void* aux;
  int64_t M, N, K, lda, ldb, ldc;
  int transA, transB, epi;
  int splits;        // >1: write raw partial sums to part[z][M][N]
  int64_t k_per_split;
  float* part;
};

template <typename TI, typename TO>
__device__ __forceinline__ void epilogue_store(const GemmArgs& g, int64_t m, int64_t n, float v) {
  if (g.epi & MT_EPI_BIAS) v += g.bias[n];
  if (g.epi & MT_EPI_ADD) v += g.addend[m * g.ldc + n];
  if (g.epi & MT_EPI_RELU) v = fmaxf(v, 0.f);
  if (g.epi & MT_EPI_RELU_MASK) {
    if (!(to_f<TI>(reinterpret_cast<const TI*>(g.aux)[m * g.ldc + n]) > 0.f)) v = 0.f;
  }
  reinterpret_cast<TO*>(g.C)[m * g.ldc + n] = from_f<TO>(v);
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(GS_THREADS) gemm_simt_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[GS_BK][GS_BM + 4];
  __shared__ __align__(16) float Bs[GS_BK][GS_BN + 4];
  const TI* A = reinterpret_cast<const TI*>(g.A);
  const TI* B = reinterpret_cast<const TI*>(g.B);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads, each 8x8 (two 4-wide halves)
  const int64_t m0 = (int64_t)blockIdx.y * GS_BM, n0 = (int64_t)blockIdx.x * GS_BN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_per_split;
  const int64_t kend = min(g.K, kbeg + g.k_per_split);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += GS_BK) {
    // ---- stage A tile: As[k][m] ----
#pragma unroll
    for (int e = 0; e < (GS_BM * GS_BK) / GS_THREADS; ++e) {
      int idx = tid + e * GS_THREADS;
      int mm, kk;
      if (g.transA) { kk = idx / GS_BM; mm = idx % GS_BM; }   // stored [K,M]: m contiguous
      else          { mm = idx / GS_BK; kk = idx % GS_BK; }   // stored [M,K]: k contiguous
      int64_t m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < g.M && k < kend) v = to_f<TI>(g.transA ? A[k * g.lda + m] : A[m * g.lda + k]);
      As[kk][mm] = v;
    }
#pragma unroll
    for (int e = 0; e < (GS_BN * GS_BK) / GS_THREADS; ++e) {
      int idx = tid + e * GS_THREADS;
      int nn, kk;
      if (g.transB) { nn = idx / GS_BK; kk = idx % GS_BK; }   // stored [N,K]: k contiguous
      else          { kk = idx / GS_BN; nn = idx % GS_BN; }   // stored [K,N]: n contiguous
      int64_t n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < g.N && k < kend) v = to_f<TI>(g.transB ? B[n * g.ldb + k] : B[k * g.ldb + n]);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GS_BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int64_t n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      if (g.splits > 1) g.part[((int64_t)blockIdx.z * g.M + m) * g.N + n] = acc[i][j];
      else epilogue_store<TI, TO>(g, m, n, acc[i][j]);
    }
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) gemm_splitk_fold_kernel(GemmArgs g) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= g.M * g.N) return;
  int64_t m = idx / g.N, n = idx - m * g.N;
  float s = 0.f;
  for (int z = 0; z < g.splits; ++z) s += g.part[((int64_t)z * g.M + m) * g.N + n];
  epilogue_store<TI, TO>(g, m, n, s);
}

static int simt_splits(int64_t M, int64_t N, int64_t K) {
  int64_t tiles = ((M + GS_BM - 1) / GS_BM) * ((N + GS_BN - 1) / GS_BN);
  if (tiles >= 64 || K < 1024) return 1;
  int64_t want = (2 * (int64_t)sm_count() + tiles - 1) / tiles;
  int64_t maxs = K / 256;
  if (want > maxs) want = maxs;
  if (want > 64) want = 64;
  return want < 1 ? 1 : (int)want;
}

size_t gemm_simt_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  int s = simt_splits(M, N, K);
  return s > 1 ? (size_t)s * M * N * sizeof(float) : 0;
}

int gemm_simt(const void* A, const void* B, void* C, const float* bias, const float* addend,
              const void* aux, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
              int64_t ldc, int transA, int transB, int in_dtype, int out_dtype, int epilogue,
              void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.bias = bias; g.addend = addend; g.aux = aux;
  g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.transA = transA; g.transB = transB; g.epi = epilogue;
  g.splits = simt_splits(M, N, K);
  if (g.splits > 1 && (!workspace || workspace_bytes < (size_t)g.splits * M * N * sizeof(float))) {
    g.splits = 1;  // legal, just slower
  }
  g.k_per_split = g.splits > 1 ? (((K + g.splits - 1) / g.splits + GS_BK - 1) / GS_BK) * GS_BK : K;
  g.part = reinterpret_cast<float*>(workspace);
  dim3 grid((unsigned)((N + GS_BN - 1) / GS_BN), (unsigned)((M + GS_BM - 1) / GS_BM), (unsigned)g.splits);
  MT_DISPATCH_DTYPE(in_dtype, TI, MT_DISPATCH_F32_BF16(out_dtype, TO, {
    gemm_simt_kernel<TI, TO><<<grid, GS_THREADS, 0, stream>>>(g);
    if (g.splits > 1) {
      int rc = check_launch("gemm_simt");
      if (rc) return rc;
      int64_t n = M * N;
      gemm_splitk_fold_kernel<TI, TO><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(g);
    }
  }));
  return check_launch("gemm_simt");
}

}  // namespace mt
