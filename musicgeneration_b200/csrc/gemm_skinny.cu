// Skinny GEMM for the KV-cached decode step (K5, M = local batch <= 64 rows):
//     C[M, N] = epi(A[M, K] . W[N, K]^T + bias),   A, W bf16, fp32 accumulate, C fp32 or bf16.
// A decode step multiplies a few dozen activation rows by every weight matrix of the stack, so
// the op is bound by reading W once (HBM / L2) and by launch + fill latency, not by FLOPs; a
// 128-row tcgen05 tile would waste 3/4 of its MMA rows and the FFMA tile kernel (gemm_simt.cu)
// needs ~80 us per call at this shape.  Here one CTA owns 8 output columns: the whole A panel and
// the 8 weight rows are staged in shared memory by coalesced 16-byte loads, the four warps split K
// four ways and run mma.sync.m16n8k16 (the legacy tensor path is the right size for an 8-column
// strip), and the partial sums are folded through shared memory with the bias / ReLU epilogue.
// N / 8 CTAs (64 .. 192 for the stack's matrices) keep every SM's memory pipe busy.
#include "ops.cuh"

namespace mt {

namespace {

constexpr int SK_BN = 8;            // output columns per CTA
constexpr int SK_WARPS = 4;         // K split
constexpr int SK_THREADS = SK_WARPS * 32;
constexpr int SK_PAD = 8;           // elements of row padding (16 B): ldmatrix rows fall on distinct banks

struct SkinnyParams {
  const __nv_bfloat16* A; const __nv_bfloat16* W; void* C; const float* bias;
  int M, N, K;
  int64_t lda, ldw, ldc;
  int epi, out_bf16;                // out_bf16: 16-bit output (bf16, or f16 when out_f16)
  int in_f16, out_f16;              // f16 operands / output: the first encoder layer of the bf16 mode
  int mtiles;                       // ceil(M / 16)
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_f16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int MT>       // MT = number of 16-row tiles (1..4)
__global__ void __launch_bounds__(SK_THREADS)
gemm_skinny_kernel(const SkinnyParams p) {
  // Programmatic dependent launch: the successor may start launching now; the weight rows (constant during a
  // decode) are requested BEFORE waiting for the predecessor grid -- only the activation panel depends on it --
  // so their HBM / L2 latency runs under the predecessor's tail.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(16) uint8_t smem[];
  const int pitch = p.K + SK_PAD;                                   // elements
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);       // [MT*16][pitch]
  __nv_bfloat16* sW = sA + MT * 16 * pitch;                         // [8][pitch]
  float* red = reinterpret_cast<float*>(sW + SK_BN * pitch);        // [SK_WARPS][MT*16][8]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * SK_BN;
  const int kv = p.K / 8;                                           // 16-byte vectors per row

  // ---- stage A (rows >= M zero-filled) and the 8 weight rows (rows >= N clamped: their columns are
  // never stored) with cp.async: every 16-byte vector of the panel is in flight at once (a
  // load-then-store loop serialised ~16 L2 round trips per thread)
  const int kvs = 31 - __clz(kv);                                   // shift instead of a division when kv is a power of two
  const bool pow2 = (kv & (kv - 1)) == 0;
  for (int v = tid; v < SK_BN * kv; v += SK_THREADS) {
    const int r = pow2 ? (v >> kvs) : v / kv, c = v - r * kv;
    const int n = min(n0 + r, p.N - 1);
    cp_async16(sW + r * pitch + c * 8, p.W + (int64_t)n * p.ldw + c * 8);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");       // the activations come from the predecessor
  for (int v = tid; v < MT * 16 * kv; v += SK_THREADS) {
    const int r = pow2 ? (v >> kvs) : v / kv, c = v - r * kv;
    __nv_bfloat16* dst = sA + r * pitch + c * 8;
    if (r < p.M) cp_async16(dst, p.A + (int64_t)r * p.lda + c * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- this warp's quarter of K
  const int kq = p.K / SK_WARPS, kbeg = warp * kq;
  float acc[MT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m) { acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f; }
  // ldmatrix.x4 of a 16x16 A tile: lanes 0-15 address rows 0-15 at k0, lanes 16-31 rows 0-15 at k0+8
  const int lrow = lane & 15, lcol = (lane >> 4) * 8;
  const uint32_t sA_u = static_cast<uint32_t>(__cvta_generic_to_shared(sA));
  const __nv_bfloat16* wrow = sW + (lane >> 2) * pitch + 2 * (lane & 3);
  for (int k0 = kbeg; k0 < kbeg + kq; k0 += 16) {
    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wrow + k0);
    const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wrow + k0 + 8);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      uint32_t a[4];
      ldmatrix_x4(a, sA_u + (uint32_t)(((m * 16 + lrow) * pitch + k0 + lcol) * 2));
      if (p.in_f16) mma_f16_16816(acc[m], a, b0, b1);
      else mma_bf16_16816(acc[m], a, b0, b1);
    }
  }
  // ---- fold the four K quarters; C fragment: c0,c1 -> row lane/4, cols 2*(lane%4)+{0,1}; c2,c3 -> row + 8
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    float* dst = red + ((warp * MT + m) * 16) * 8;
    const int r = lane >> 2, c = 2 * (lane & 3);
    dst[r * 8 + c] = acc[m][0]; dst[r * 8 + c + 1] = acc[m][1];
    dst[(r + 8) * 8 + c] = acc[m][2]; dst[(r + 8) * 8 + c + 1] = acc[m][3];
  }
  __syncthreads();
  for (int e = tid; e < MT * 16 * 8; e += SK_THREADS) {
    const int r = e >> 3, c = e & 7;
    const int n = n0 + c;
    if (r >= p.M || n >= p.N) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; ++w) s += red[w * MT * 16 * 8 + e];
    if (p.epi & MT_EPI_BIAS) s += p.bias[n];
    if (p.epi & MT_EPI_RELU) s = fmaxf(s, 0.f);
    if (p.out_f16) reinterpret_cast<__half*>(p.C)[(int64_t)r * p.ldc + n] = __float2half_rn(s);
    else if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.C)[(int64_t)r * p.ldc + n] = __float2bfloat16_rn(s);
    else reinterpret_cast<float*>(p.C)[(int64_t)r * p.ldc + n] = s;
  }
}

size_t skinny_smem(int mtiles, int K) {
  return (size_t)(mtiles * 16 + SK_BN) * (K + SK_PAD) * 2 + (size_t)SK_WARPS * mtiles * 16 * 8 * 4;
}

}  // namespace

bool gemm_skinny_supported(int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int transA, int transB,
                           int in_dtype, int out_dtype, int epilogue, const void* A, const void* B) {
  if (in_dtype != MT_BF16 && in_dtype != MT_F16) return false;
  if (out_dtype != MT_F32 && out_dtype != MT_BF16 && out_dtype != MT_F16) return false;
  if (transA || !transB) return false;                               // x . W^T only
  if (M > 64 || K % 64 || K > 2048 || N < 1) return false;
  if (epilogue & ~(MT_EPI_BIAS | MT_EPI_RELU)) return false;
  if (lda % 8 || ldb % 8 || !aligned(A, 16) || !aligned(B, 16)) return false;
  return skinny_smem((int)((M + 15) / 16), (int)K) <= 200 * 1024;
}

int gemm_skinny(const void* A, const void* B, void* C, const float* bias, int64_t M, int64_t N, int64_t K,
                int64_t lda, int64_t ldb, int64_t ldc, int in_dtype, int out_dtype, int epilogue, cudaStream_t st) {
  SkinnyParams p;
  p.A = reinterpret_cast<const __nv_bfloat16*>(A); p.W = reinterpret_cast<const __nv_bfloat16*>(B);
  p.C = C; p.bias = bias; p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.lda = lda; p.ldw = ldb; p.ldc = ldc; p.epi = epilogue; p.out_bf16 = (out_dtype != MT_F32);
  p.in_f16 = (in_dtype == MT_F16); p.out_f16 = (out_dtype == MT_F16);
  p.mtiles = (int)((M + 15) / 16);
  const size_t smem = skinny_smem(p.mtiles, (int)K);
  const unsigned grid = (unsigned)((N + SK_BN - 1) / SK_BN);
#define MT_SKINNY_LAUNCH(MTC)                                                                                   \
  {                                                                                                             \
    auto kern = gemm_skinny_kernel<MTC>;                                                                        \
    static size_t attr = 0;                                                                                     \
    if (smem > 48 * 1024 && smem > attr) {                                                                      \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
      if (e != cudaSuccess) { set_error("gemm_skinny: smem attribute: %s", cudaGetErrorString(e)); return (int)e; } \
      attr = smem;                                                                                              \
    }                                                                                                           \
    launch_chain(kern, dim3(grid), dim3(SK_THREADS), smem, st, p);                                                                    \
  }
  switch (p.mtiles) {
    case 1: MT_SKINNY_LAUNCH(1) break;
    case 2: MT_SKINNY_LAUNCH(2) break;
    case 3: MT_SKINNY_LAUNCH(3) break;
    default: MT_SKINNY_LAUNCH(4) break;
  }
#undef MT_SKINNY_LAUNCH
  return check_launch("gemm_skinny");
}

}  // namespace mt
