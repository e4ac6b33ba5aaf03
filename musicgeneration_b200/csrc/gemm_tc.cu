// tcgen05 GEMM for the nn.Linear layers of the path (K5): C[M,N] = epi(op(A) . op(B)),
// bf16 (or f16) operands, fp32 accumulation in TMEM, fused bias / ReLU / residual-add /
// ReLU-mask epilogue, fp32 or bf16 output.
//
//   * 128 x 128 output tile per CTA, K blocked by 64; operands arrive by TMA
//     (cp.async.bulk.tensor, 128B swizzle) into a 3-stage shared-memory ring;
//   * one elected thread issues tcgen05.mma (M=128, N=128, K=16) with the accumulator in
//     128 TMEM columns; tcgen05.commit releases ring slots and signals the epilogue;
//   * warp roles: 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2..5 = epilogue
//     (tcgen05.ld 32x32b, lane = output row);
//   * both operand majors are supported through the UMMA descriptors, so forward (A K-major,
//     B = W[N,K] K-major), dgrad (B = W[N,K] read as [K,N]: MN-major) and wgrad (A = dY[T,N]
//     read as [N,T]^T: MN-major, B = X[T,K]: MN-major) need no transposed copies;
//   * ~100 KB of shared memory and 128 TMEM columns per CTA -> two CTAs per SM, so one CTA's
//     epilogue overlaps the other's main loop; skinny outputs (weight gradients) use a
//     deterministic split-K through the caller's workspace.
#include "ops.cuh"
#include "tc_common.cuh"

namespace mt {

namespace {

// Tile shapes: 128 x 128 (3-stage ring) or 128 x 256 (2-stage ring, one N = 256 MMA per k-step).
// With K of only 256..512 the kernel is bound by L2 -> shared-memory traffic, not by the tensor pipe:
// a 128 x 128 x 64 block loads 32 KB for 2.1 MFLOP (64 FLOP/B), a 128 x 256 x 64 block 48 KB for
// 4.2 MFLOP (87 FLOP/B).  Both shapes fit two CTAs per SM (<= 100 KB of shared memory, <= 256 TMEM columns).
constexpr int BM = 128, BK = 64;
constexpr int TILE_BYTES = BM * BK * 2;          // 16 KB: one [128 x 64] 16-bit tile (A stage; B stage per 128 columns)
constexpr int TC_THREADS = 192;
constexpr int EPI_PITCH = 132;          // floats per staged row (== 4 mod 32: conflict-free 128-bit stores)
template <int BN_> struct Shape {
  static constexpr int STAGES = (BN_ == 128) ? 3 : 2;
  static constexpr int TILE_B = BN_ * BK * 2;
  static constexpr int RING = STAGES * (TILE_BYTES + TILE_B);
  static constexpr int ONES = RING + 1024;            // 2 KB of bf16 1.0 (the B operand of the column-sum product), 1 KB aligned
  static constexpr size_t SMEM = RING + 1024 + 2048 + 1024;
  static_assert(BM * EPI_PITCH * 4 <= RING, "epilogue staging aliases the operand ring");
};

struct TcGemmParams {
  void* C;
  const float* bias;
  const float* addend;
  const void* aux;
  int64_t M, N, K, ldc;
  int epi, out_bf16, vec_ok;     // out_bf16: the output is 16-bit (bf16, or f16 when out_f16 is set)
  int out_f16;
  int add_inplace;          // EPI_ADD with addend == C (fp32, no ReLU / mask): accumulate with red.global.add.v4.f32
  int splits;
  int64_t k_per_split;
  float* part;
  // weight-gradient launches only (A MN-major): cs_out[m] = sum_k A[k, m] -- the bias gradient -- from one extra
  // N = 16 product per k-step against a tile of ones in the CTAs of the first tile column; split-K partials
  // in cs_part[z][M], folded with the rest
  float* cs_out;
  float* cs_part;
};

// 16-bit output: bf16, or f16 (the first encoder layer's QKV projection in the mixed mode)
__device__ __forceinline__ void store4_lp(const TcGemmParams& p, int64_t off, const float4& r) {
  if (p.out_f16) store4<__half>(reinterpret_cast<__half*>(p.C) + off, r);
  else store4<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(p.C) + off, r);
}
__device__ __forceinline__ void store1_lp(const TcGemmParams& p, int64_t off, float x) {
  if (p.out_f16) reinterpret_cast<__half*>(p.C)[off] = __float2half_rn(x);
  else reinterpret_cast<__nv_bfloat16*>(p.C)[off] = __float2bfloat16_rn(x);
}

// One quad (row, n .. n+3) of the output tile: bias / residual add / ReLU / ReLU-mask, then store.
__device__ __forceinline__ void epilogue_quad(const TcGemmParams& p, int64_t row, int64_t n, float (&o)[4]) {
  if (p.splits > 1) {                     // raw partial sums; the fold kernel applies the epilogue
    float* dst = p.part + ((int64_t)blockIdx.z * p.M + row) * p.N + n;
    if (n + 3 < p.N && (p.N & 3) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (n + t < p.N) dst[t] = o[t];
    }
    return;
  }
  const int64_t off = row * p.ldc + n;
  if (n + 3 < p.N && p.vec_ok) {
    if (p.epi & MT_EPI_BIAS) {
      float4 b = *reinterpret_cast<const float4*>(p.bias + n);
      o[0] += b.x; o[1] += b.y; o[2] += b.z; o[3] += b.w;
    }
    if (p.add_inplace) {
      // C += acc (fp32, the addend IS the output: residual-path gradient): one vector reduction, done by
      // L2 -- no load and so no exposed DRAM latency in the epilogue.  One contribution per element, so the
      // result does not depend on any ordering.
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<float*>(p.C) + off),
                   "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]) : "memory");
      return;
    }
    if (p.epi & MT_EPI_ADD) {
      float4 a = *reinterpret_cast<const float4*>(p.addend + off);
      o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
    }
    if (p.epi & MT_EPI_RELU) {
#pragma unroll
      for (int t = 0; t < 4; ++t) o[t] = fmaxf(o[t], 0.f);
    }
    if (p.epi & MT_EPI_RELU_MASK) {
      float4 m = load4<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(p.aux) + off);
      if (!(m.x > 0.f)) o[0] = 0.f;
      if (!(m.y > 0.f)) o[1] = 0.f;
      if (!(m.z > 0.f)) o[2] = 0.f;
      if (!(m.w > 0.f)) o[3] = 0.f;
    }
    float4 r = make_float4(o[0], o[1], o[2], o[3]);
    if (p.out_bf16) store4_lp(p, off, r);
    else store4<float>(reinterpret_cast<float*>(p.C) + off, r);
  } else {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int64_t nn = n + t;
      if (nn >= p.N) continue;
      float x = o[t];
      if (p.epi & MT_EPI_BIAS) x += p.bias[nn];
      if (p.epi & MT_EPI_ADD) x += p.addend[off + t];
      if (p.epi & MT_EPI_RELU) x = fmaxf(x, 0.f);
      if (p.epi & MT_EPI_RELU_MASK) {
        if (!(__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.aux)[off + t]) > 0.f)) x = 0.f;
      }
      if (p.out_bf16) store1_lp(p, off + t, x);
      else reinterpret_cast<float*>(p.C)[off + t] = x;
    }
  }
}

// The same quad with its residual addend / ReLU-mask operand already in registers (vectorised case only).
__device__ __forceinline__ void epilogue_quad_pre(const TcGemmParams& p, int64_t row, int64_t n, float (&o)[4],
                                                  const float4& add, const uint2& aux, bool has_add, bool has_aux) {
  const int64_t off = row * p.ldc + n;
  if (p.epi & MT_EPI_BIAS) {
    float4 b = *reinterpret_cast<const float4*>(p.bias + n);
    o[0] += b.x; o[1] += b.y; o[2] += b.z; o[3] += b.w;
  }
  if (has_add) { o[0] += add.x; o[1] += add.y; o[2] += add.z; o[3] += add.w; }
  if (p.epi & MT_EPI_RELU) {
#pragma unroll
    for (int t = 0; t < 4; ++t) o[t] = fmaxf(o[t], 0.f);
  }
  if (has_aux) {
    const float m0 = __uint_as_float(aux.x << 16), m1 = __uint_as_float(aux.x & 0xffff0000u);
    const float m2 = __uint_as_float(aux.y << 16), m3 = __uint_as_float(aux.y & 0xffff0000u);
    if (!(m0 > 0.f)) o[0] = 0.f;
    if (!(m1 > 0.f)) o[1] = 0.f;
    if (!(m2 > 0.f)) o[2] = 0.f;
    if (!(m3 > 0.f)) o[3] = 0.f;
  }
  float4 r = make_float4(o[0], o[1], o[2], o[3]);
  if (p.out_bf16) store4_lp(p, off, r);
  else store4<float>(reinterpret_cast<float*>(p.C) + off, r);
}

template <int A_MN, int B_MN, int BN, bool CS = false>
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcGemmParams p, const int fmt) {
  constexpr int STAGES = Shape<BN>::STAGES;
  constexpr int TILE_B = Shape<BN>::TILE_B;
  constexpr bool CS_OK = CS && (A_MN == 1 && BN == 128);    // the column-sum accumulator takes TMEM columns BN .. BN+15
  constexpr int TMEM_COLS = CS_OK ? 256 : BN;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
  const bool do_cs = CS_OK && p.cs_out != nullptr && blockIdx.x == 0;
  if (do_cs) {
    for (int i = threadIdx.x; i < 2048 / 16; i += TC_THREADS)
      reinterpret_cast<uint4*>(smem + Shape<BN>::ONES)[i] = (fmt == 1) ? make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u)
                                                                      : make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    tc::fence_proxy_async();            // generic-proxy stores -> visible to the tensor core's shared-memory reads
  }
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * TILE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * TILE_B);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * p.k_per_split;
  const int64_t kend = min(p.K, kbeg + p.k_per_split);
  const int nkb = (int)((kend - kbeg + BK - 1) / BK);

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        tc::mbar_wait(&empty[s], ph ^ 1);
        tc::mbar_arrive_expect_tx(&full[s], TILE_BYTES + TILE_B);
        const int k = (int)(kbeg + (int64_t)kb * BK);
        uint8_t* a = sA + s * TILE_BYTES;
        uint8_t* b = sB + s * TILE_B;
        if (A_MN) {
          tc::tma_load_2d(a, &tmA, &full[s], m0, k);
          tc::tma_load_2d(a + TILE_BYTES / 2, &tmA, &full[s], m0 + 64, k);
        } else {
          tc::tma_load_2d(a, &tmA, &full[s], k, m0);
        }
        if (B_MN) {          // 64-column groups of the [K, N] matrix, 8 KB apart
#pragma unroll
          for (int g = 0; g < BN / 64; ++g) tc::tma_load_2d(b + g * (TILE_BYTES / 2), &tmB, &full[s], n0 + 64 * g, k);
        } else {
          tc::tma_load_2d(b, &tmB, &full[s], k, n0);      // one [BN x 64] box
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(BM, BN, fmt, fmt, A_MN, B_MN);
      const uint64_t ad_k = tc::make_sdesc(tc::smem_u32(sA), 16, 1024), ad_mn = tc::make_sdesc(tc::smem_u32(sA), TILE_BYTES / 2, 1024);
      const uint64_t bd_k = tc::make_sdesc(tc::smem_u32(sB), 16, 1024), bd_mn = tc::make_sdesc(tc::smem_u32(sB), TILE_BYTES / 2, 1024);
      const uint32_t idesc_cs = tc::make_idesc(BM, 16, fmt, fmt, A_MN, 0);
      const uint64_t ones_d = tc::make_sdesc(tc::smem_u32(smem + Shape<BN>::ONES), 16, 1024);     // K-major, 16 rows: every element is 1
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        tc::mbar_wait(&full[s], ph);
        tc::tc_fence_after();
        // K-major: 16 elements = 32 B inside the 128 B swizzle row; 8-row groups 1024 B apart.
        // MN-major: 16 k-rows of 128 B = 2048 B; the two 64-wide MN halves are 8192 B apart.
        // One descriptor per operand and stage; a k-step is an add on its 16-byte address field.
        const uint64_t ad0 = (A_MN ? ad_mn : ad_k) + (uint64_t)s * (TILE_BYTES >> 4);
        const uint64_t bd0 = (B_MN ? bd_mn : bd_k) + (uint64_t)s * (TILE_B >> 4);
#pragma unroll
        for (int k4 = 0; k4 < BK / 16; ++k4)
          tc::umma_f16(tmem_base, ad0 + (A_MN ? 128 : 2) * k4, bd0 + (B_MN ? 128 : 2) * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
        if (do_cs) {      // D2[m, 0..15] += sum over the 16 k of A[k, m] * 1
#pragma unroll
          for (int k4 = 0; k4 < BK / 16; ++k4)
            tc::umma_f16(tmem_base + BN, ad0 + 128 * k4, ones_d, idesc_cs, (kb | k4) != 0 ? 1u : 0u);
        }
        tc::umma_commit(&empty[s]);
      }
      tc::umma_commit(tmem_full);
    }
  } else {
    // Epilogue.  TMEM gives each thread one ROW of the tile, which would make every global access
    // a 32-way strided one.  So: (1) park the fp32 accumulators in shared memory (the operand ring
    // is dead by now: every TMA load has landed and every MMA has retired before tmem_full fires),
    // (2) re-read them row-contiguously so that bias / addend / aux loads and the C stores are
    // fully coalesced 128-bit accesses.
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after();
    float* stage = reinterpret_cast<float*>(smem);            // [128][EPI_PITCH]
    const int et = threadIdx.x - 64;              // 0..127
    const int col = (et & 31) * 4;                // 4 consecutive columns of a 128-column half
#pragma unroll 1
    for (int hn = 0; hn < BN / 128; ++hn) {
      if (hn > 0) tc::named_bar_sync(1, 128);     // the previous half has been read out of the staging rows
      {
        float* srow = stage + (q * 32 + lane) * EPI_PITCH;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hn * 128 + c * 32), r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(srow + c * 32 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
      }
      tc::tc_fence_before();
      tc::named_bar_sync(1, 128);
      const int64_t n = (int64_t)n0 + hn * 128 + col;
#pragma unroll 4
      for (int r = et >> 5; r < BM; r += 4) {
        const int64_t row = (int64_t)m0 + r;
        if (row >= p.M || n >= p.N) continue;
        const float4 acc = *reinterpret_cast<const float4*>(stage + r * EPI_PITCH + col);
        float o[4] = {acc.x, acc.y, acc.z, acc.w};
        epilogue_quad(p, row, n, o);
      }
    }
    if (do_cs) {                                  // column BN of row (q * 32 + lane) = this CTA's share of the column sum
      uint32_t r[32];
      tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)BN, r);
      tc::tmem_ld_wait();
      const int64_t row = (int64_t)m0 + q * 32 + lane;
      if (row < p.M) {
        if (p.splits > 1) p.cs_part[(int64_t)blockIdx.z * p.M + row] = __uint_as_float(r[0]);
        else p.cs_out[row] = __uint_as_float(r[0]);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------
// Persistent variant (A K-major: x.W^T and dgrad, no split-K).  One CTA per SM walks output tiles
// (n fastest, so concurrently running CTAs share the A panel in L2); the operand ring and the
// barriers run straight through tile boundaries, TWO accumulators live in TMEM, and eight epilogue
// warps drain tile i (TMEM -> registers -> per-warp staging rows -> 128-byte row segments in global
// memory, bias / ReLU / residual / ReLU-mask applied on the way) while the MMA warp already fills the
// other accumulator with tile i+1.  In the one-tile-per-CTA kernel above the main loop and the
// epilogue of a CTA are serial (ncu: 47 % of warp samples are warps parked at a barrier or at
// EXIT waiting for the other phase).
// ---------------------------------------------------------------------------------------
constexpr int PS_THREADS = 320;                 // TMA warp, MMA warp, 8 epilogue warps
constexpr int PS_STG_WORDS = 36;                // staging row pitch: 32 columns + 4 (16-byte stores conflict-free)
constexpr int PS_STG_BYTES = 32 * PS_STG_WORDS * 4;
template <int BN_> struct PShape {
  static constexpr int STAGES = (BN_ == 256) ? 3 : 5;
  static constexpr int TILE_B = BN_ * BK * 2;
  static constexpr int RING = STAGES * (TILE_BYTES + TILE_B);
  static constexpr size_t SMEM = RING + 8 * PS_STG_BYTES + 256 + 1024;
  static_assert(SMEM <= 232448, "shared memory budget");
};

template <int B_MN, int BN>
__global__ void __launch_bounds__(PS_THREADS, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const TcGemmParams p, const int fmt, const int tiles_n, const int num_tiles) {
  constexpr int STAGES = PShape<BN>::STAGES;
  constexpr int TILE_B = PShape<BN>::TILE_B;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * TILE_BYTES;
  uint8_t* stg = sB + STAGES * TILE_B;
  uint64_t* full = reinterpret_cast<uint64_t*>(stg + 8 * PS_STG_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;      // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (int)((p.K + BK - 1) / BK);

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&acc_full[b], 1); tc::mbar_init(&acc_empty[b], 256); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 2 * BN);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int g = 0;                                  // k-blocks issued so far (over all tiles)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % STAGES;
          tc::mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(&full[s], TILE_BYTES + TILE_B);
          const int k = kb * BK;
          uint8_t* b = sB + s * TILE_B;
          tc::tma_load_2d(sA + s * TILE_BYTES, &tmA, &full[s], k, m0);
          if (B_MN) {
#pragma unroll
            for (int gq = 0; gq < BN / 64; ++gq) tc::tma_load_2d(b + gq * (TILE_BYTES / 2), &tmB, &full[s], n0 + 64 * gq, k);
          } else {
            tc::tma_load_2d(b, &tmB, &full[s], k, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(BM, BN, fmt, fmt, 0, B_MN);
      const uint64_t ad_k = tc::make_sdesc(tc::smem_u32(sA), 16, 1024);
      const uint64_t bd_k = tc::make_sdesc(tc::smem_u32(sB), 16, 1024), bd_mn = tc::make_sdesc(tc::smem_u32(sB), TILE_BYTES / 2, 1024);
      int g = 0, i = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i) {
        const int buf = i & 1;
        tc::mbar_wait(&acc_empty[buf], ((i >> 1) & 1) ^ 1);       // the epilogue has read this accumulator
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = g % STAGES;
          tc::mbar_wait(&full[s], (g / STAGES) & 1);
          tc::tc_fence_after();
          const uint64_t ad0 = ad_k + (uint64_t)s * (TILE_BYTES >> 4);
          const uint64_t bd0 = (B_MN ? bd_mn : bd_k) + (uint64_t)s * (TILE_B >> 4);
#pragma unroll
          for (int k4 = 0; k4 < BK / 16; ++k4)
            tc::umma_f16(d_tmem, ad0 + 2 * k4, bd0 + (B_MN ? 128 : 2) * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
          tc::umma_commit(&empty[s]);
        }
        tc::umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // epilogue warps 2..9: TMEM quadrant = warp % 4 (hardware rule), column half = first / second set of four
    const int ew = warp - 2, q = warp & 3, chalf = ew >> 2;
    constexpr int CPW = BN / 2;                   // columns per warp
    float* stage = reinterpret_cast<float*>(stg + ew * PS_STG_BYTES);
    const int piece = lane & 7, rsub = lane >> 3;
    // The residual addend (fp32, 16 B per quad) and the ReLU-mask operand (bf16, 8 B per quad) are the
    // only global READS of the epilogue.  Loaded where they are used, two at a time, they are a chain
    // of exposed DRAM latencies (qkv dgrad + residual: 110 us against 59 us without the addend), so
    // the 8 quads a thread owns in a 32-column chunk are fetched one chunk ahead -- for the first
    // chunk of a tile before the accumulator is even complete.  (Plain loads: the addend may alias
    // the output; a thread reads exactly the quads it later writes.)
    const bool pre_add = (p.epi & MT_EPI_ADD) && p.vec_ok && p.splits == 1 && !p.add_inplace;
    const bool pre_aux = (p.epi & MT_EPI_RELU_MASK) && p.vec_ok && p.splits == 1;
    auto prefetch = [&](int m0, int n0, int c, float4 (&pa)[8], uint2 (&px)[8]) {
      if (!(pre_add || pre_aux)) return;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int64_t row = (int64_t)m0 + q * 32 + it * 4 + rsub;
        const int64_t n = (int64_t)n0 + chalf * CPW + c * 32 + piece * 4;
        const bool ok = row < p.M && n + 3 < p.N;
        const int64_t off = row * p.ldc + n;
        pa[it] = (pre_add && ok) ? *reinterpret_cast<const float4*>(p.addend + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        px[it] = (pre_aux && ok) ? *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.aux) + off)
                                 : make_uint2(0x3f803f80u, 0x3f803f80u);
      }
    };
    int i = 0;
    if (pre_add || pre_aux) {
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        const int buf = i & 1;
        float4 pa[8], pa_next[8];
        uint2 px[8], px_next[8];
        prefetch(m0, n0, 0, pa, px);
        const bool mask_fast = pre_aux && !pre_add && p.out_bf16 && !p.out_f16 && !(p.epi & (MT_EPI_BIAS | MT_EPI_RELU | MT_EPI_ADD)) &&
                               m0 + BM <= p.M && n0 + BN <= p.N;
        tc::mbar_wait(&acc_full[buf], (i >> 1) & 1);
        tc::tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + chalf * CPW);
#pragma unroll 1
        for (int c = 0; c < CPW / 32; ++c) {
          uint32_t r[32];
          tc::tmem_ld_32x32(t0 + (uint32_t)(c * 32), r);
          if (c + 1 < CPW / 32) prefetch(m0, n0, c + 1, pa_next, px_next);
          tc::tmem_ld_wait();
          if (c == CPW / 32 - 1) {                  // last read of this accumulator: hand it back to the MMA warp
            tc::tc_fence_before();
            tc::mbar_arrive(&acc_empty[buf]);
          }
          __syncwarp();                             // the previous chunk's rows have been read by every lane
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(stage + lane * PS_STG_WORDS + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
          __syncwarp();
          if (mask_fast) {
            // interior tile of a ReLU-masked bf16 product (the FFN_suf dgrad): 8 independent LDS -> mask -> STG
            // chains per thread, no per-quad bounds or flag tests (28.4 us against 16.7 us for the unmasked product)
            float4 acc[8];
#pragma unroll
            for (int it = 0; it < 8; ++it)
              acc[it] = *reinterpret_cast<const float4*>(stage + (it * 4 + rsub) * PS_STG_WORDS + piece * 4);
            const int64_t off0 = ((int64_t)m0 + q * 32 + rsub) * p.ldc + (int64_t)n0 + chalf * CPW + c * 32 + piece * 4;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              float4 v = acc[it];
              const uint2 a = px[it];
              if (!(__uint_as_float(a.x << 16) > 0.f)) v.x = 0.f;
              if (!(__uint_as_float(a.x & 0xffff0000u) > 0.f)) v.y = 0.f;
              if (!(__uint_as_float(a.y << 16) > 0.f)) v.z = 0.f;
              if (!(__uint_as_float(a.y & 0xffff0000u) > 0.f)) v.w = 0.f;
              store4<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(p.C) + off0 + (int64_t)(it * 4) * p.ldc, v);
            }
          } else {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rsub;
            const int64_t row = (int64_t)m0 + q * 32 + rr;
            const int64_t n = (int64_t)n0 + chalf * CPW + c * 32 + piece * 4;
            if (row >= p.M || n >= p.N) continue;
            const float4 acc = *reinterpret_cast<const float4*>(stage + rr * PS_STG_WORDS + piece * 4);
            float o[4] = {acc.x, acc.y, acc.z, acc.w};
            if (n + 3 < p.N) epilogue_quad_pre(p, row, n, o, pa[it], px[it], pre_add, pre_aux);
            else epilogue_quad(p, row, n, o);
          }
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) { pa[it] = pa_next[it]; px[it] = px_next[it]; }
        }
      }
    } else {
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        const int buf = i & 1;
        tc::mbar_wait(&acc_full[buf], (i >> 1) & 1);
        tc::tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + chalf * CPW);
        // interior tile, vectorised output, no operand to read back: the bias quad is loaded once per
        // chunk (it depends on the column only) and the 8 rows of a thread are independent LDS -> FADD ->
        // STG chains (per-quad bias loads sat in every chain: 'FADD ... stall_long_sb' in the source page)
        const bool fast = p.vec_ok && p.splits == 1 && !(p.epi & MT_EPI_RELU_MASK) &&
                          (!(p.epi & MT_EPI_ADD) || p.add_inplace) && m0 + BM <= p.M && n0 + BN <= p.N;
#pragma unroll 1
        for (int c = 0; c < CPW / 32; ++c) {
          uint32_t r[32];
          tc::tmem_ld_32x32(t0 + (uint32_t)(c * 32), r);
          const int64_t nq = (int64_t)n0 + chalf * CPW + c * 32 + piece * 4;
          float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
          if (fast && (p.epi & MT_EPI_BIAS)) bq = *reinterpret_cast<const float4*>(p.bias + nq);
          tc::tmem_ld_wait();
          if (c == CPW / 32 - 1) {                  // last read of this accumulator: hand it back to the MMA warp
            tc::tc_fence_before();
            tc::mbar_arrive(&acc_empty[buf]);
          }
          __syncwarp();                             // the previous chunk's rows have been read by every lane
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(stage + lane * PS_STG_WORDS + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
          __syncwarp();
          if (fast) {
            float4 acc[8];
#pragma unroll
            for (int it = 0; it < 8; ++it)
              acc[it] = *reinterpret_cast<const float4*>(stage + (it * 4 + rsub) * PS_STG_WORDS + piece * 4);
            const bool relu = (p.epi & MT_EPI_RELU) != 0;
            const int64_t off0 = ((int64_t)m0 + q * 32 + rsub) * p.ldc + nq;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              float4 v = make_float4(acc[it].x + bq.x, acc[it].y + bq.y, acc[it].z + bq.z, acc[it].w + bq.w);
              if (relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
              const int64_t off = off0 + (int64_t)(it * 4) * p.ldc;
              if (p.add_inplace)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<float*>(p.C) + off),
                             "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
              else if (p.out_bf16) store4_lp(p, off, v);
              else store4<float>(reinterpret_cast<float*>(p.C) + off, v);
            }
          } else {
#pragma unroll 2
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + rsub;
              const int64_t row = (int64_t)m0 + q * 32 + rr;
              const int64_t n = (int64_t)n0 + chalf * CPW + c * 32 + piece * 4;
              if (row >= p.M || n >= p.N) continue;
              const float4 acc = *reinterpret_cast<const float4*>(stage + rr * PS_STG_WORDS + piece * 4);
              float o[4] = {acc.x, acc.y, acc.z, acc.w};
              epilogue_quad(p, row, n, o);
            }
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 2 * BN);
  }
}

// blocks past the output's own fold the column-sum partials (ascending z, like everything else here)
__device__ __forceinline__ bool fold_colsum_blocks(const TcGemmParams& p, int64_t main_blocks) {
  if ((int64_t)blockIdx.x < main_blocks) return false;
  const int64_t m = ((int64_t)blockIdx.x - main_blocks) * blockDim.x + threadIdx.x;
  if (p.cs_out != nullptr && m < p.M) {
    float s = 0.f;
    for (int z = 0; z < p.splits; ++z) s += p.cs_part[(int64_t)z * p.M + m];
    p.cs_out[m] = s;
  }
  return true;
}

__global__ void __launch_bounds__(256) tc_splitk_fold_kernel(TcGemmParams p) {
  if (fold_colsum_blocks(p, (p.M * p.N + 255) / 256)) return;
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= p.M * p.N) return;
  int64_t m = idx / p.N, n = idx - m * p.N;
  float s = 0.f;
  for (int z = 0; z < p.splits; ++z) s += p.part[((int64_t)z * p.M + m) * p.N + n];
  if (p.epi & MT_EPI_BIAS) s += p.bias[n];
  if (p.epi & MT_EPI_ADD) s += p.addend[m * p.ldc + n];
  if (p.epi & MT_EPI_RELU) s = fmaxf(s, 0.f);
  if (p.epi & MT_EPI_RELU_MASK) {
    if (!(__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.aux)[m * p.ldc + n]) > 0.f)) s = 0.f;
  }
  if (p.out_bf16) store1_lp(p, m * p.ldc + n, s);
  else reinterpret_cast<float*>(p.C)[m * p.ldc + n] = s;
}

// (Measured and dropped, round 2: the fold INSIDE the GEMM launch -- an arrival counter per output tile, the tile's last
// CTA sums its partials in the same order -- is parity-green but 10.08 vs 9.42 ms per config-B step: one CTA folding a
// 128 x 128 tile over its splits is a chain of L2 round trips at the tail of every weight-gradient launch (~25 us),
// where this launch spreads the same loads over the whole chip in 8 us.)
// The weight-gradient case (fp32 output, no epilogue, N % 4 == 0): one quad per thread, the partial
// loads of eight splits in flight at a time; same ascending-z summation order as the scalar kernel.
__global__ void __launch_bounds__(256) tc_splitk_fold_vec_kernel(TcGemmParams p) {
  const int64_t quads = p.M * p.N / 4;
  if (fold_colsum_blocks(p, (quads + 255) / 256)) return;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= quads) return;
  const int64_t e = idx * 4;
  const int64_t m = e / p.N, n = e - m * p.N;
  const float4* src = reinterpret_cast<const float4*>(p.part) + idx;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int z = 0;
  for (; z + 8 <= p.splits; z += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (int64_t)(z + u) * quads);
#pragma unroll
    for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
  }
  for (; z < p.splits; ++z) {
    const float4 v = __ldcg(src + (int64_t)z * quads);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + m * p.ldc + n) = s;
}

// Split-K factor of a launch with few output tiles and a long K (the weight gradients: K = tokens).
// Two CTAs are resident per SM, so tiles * splits is kept AT OR BELOW 2 * SMs: one more CTA than
// that is a second wave of the whole K range (measured: 336 CTAs of K/7 took 65.9 us for the QKV
// gradient of config B -- two waves -- where 288 CTAs of K/6 fit one).
int tc_splits(int64_t M, int64_t N, int64_t K, int bn) {
  int64_t tiles = ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
  if (tiles >= sm_count() || K < 2048) return 1;
  static const int per_sm = [] { const char* e = getenv("MT_SPLITK_PER_SM"); int v = e ? atoi(e) : 2; return v == 1 ? 1 : 2; }();
  int64_t want = (per_sm * (int64_t)sm_count()) / tiles;
  int64_t maxs = K / 512;
  if (want > maxs) want = maxs;
  if (want > 32) want = 32;
  return want < 1 ? 1 : (int)want;
}

// 128 x 256 tiles for the MN-major-A (weight gradient) products (87 instead of 64 FLOP per byte fetched
// from L2) are available with MT_WGRAD_BN=256 but measured SLOWER on B200 at config B's shapes (QKV
// gradient 59.3 vs 56.4 us, fc 34.9 vs 26.9 us incl. the fold): the 2-stage ring of the wide shape hides
// less of the L2 latency than the 3-stage ring of the 128 x 128 one.  Default: 128 x 128.
bool wgrad_wide() {
  static const bool wide = [] {
    const char* e = getenv("MT_WGRAD_BN");
    return e && atoi(e) == 256;
  }();
  return wide;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// tensor maps (driver entry point resolved at run time)
// ---------------------------------------------------------------------------------------
namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                 int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return MT_E_UNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d rows=%ld cols=%ld ld=%ld box=%dx%d) failed: %d", (long)rows, (long)cols,
              (long)ld, box_cols, box_rows, (int)r);
    return MT_E_ARG;
  }
  return 0;
}

int make_tmap_2d_f32(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                     int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return MT_E_UNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d f32 rows=%ld cols=%ld ld=%ld box=%dx%d) failed: %d", (long)rows, (long)cols,
              (long)ld, box_cols, box_rows, (int)r);
    return MT_E_ARG;
  }
  return 0;
}

int make_tmap_3d(CUtensorMap* map, const void* base, int64_t d0, int64_t d1, int64_t d2, int64_t s1_elems,
                 int64_t s2_elems, int box0, int box1, int box2) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return MT_E_UNSUPPORTED; }
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)s1_elems * 2, (cuuint64_t)s2_elems * 2};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, (cuuint32_t)box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d %ldx%ldx%ld) failed: %d", (long)d0, (long)d1, (long)d2, (int)r);
    return MT_E_ARG;
  }
  return 0;
}

int make_tmap_blhd(CUtensorMap* map, const void* base, int64_t dh, int64_t L, int64_t h, int64_t B, int64_t sl,
                   int64_t sh, int64_t sb, int box_d, int box_l) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return MT_E_UNSUPPORTED; }
  // dimension order {dh, h, L, B}: strides increase (sh < sl < sb for the packed qkv / O layouts)
  cuuint64_t dims[4] = {(cuuint64_t)dh, (cuuint64_t)h, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sh * 2, (cuuint64_t)sl * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_d, 1, (cuuint32_t)box_l, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d dh=%ld L=%ld h=%ld B=%ld sl=%ld sh=%ld sb=%ld) failed: %d", (long)dh,
              (long)L, (long)h, (long)B, (long)sl, (long)sh, (long)sb, (int)r);
    return MT_E_ARG;
  }
  return 0;
}

}  // namespace tc

// ---------------------------------------------------------------------------------------
// launcher
// ---------------------------------------------------------------------------------------
bool gemm_tc_supported(int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int transA,
                       int transB, int in_dtype, int out_dtype, int epilogue, const void* A, const void* B,
                       const void* C) {
  (void)epilogue; (void)C;
  if (in_dtype != MT_BF16 && in_dtype != MT_F16) return false;
  if (out_dtype != MT_F32 && out_dtype != MT_BF16 && out_dtype != MT_F16) return false;
  if (transA && !(transA && !transB)) return false;                 // TN only (wgrad form)
  if (!aligned(A, 16) || !aligned(B, 16)) return false;
  if (lda % 8 || ldb % 8) return false;
  if (M < 64 || N < 8 || K < 16) return false;                      // tiny problems: SIMT path
  if (M > (1ll << 31) - 256 || N > (1ll << 31) - 256 || K > (1ll << 31) - 256) return false;
  (void)ldc;
  return mt_device_ok() != 0;
}

size_t gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  int s = max(tc_splits(M, N, K, 128), tc_splits(M, N, K, 256));
  if (s > 1) return (size_t)s * (M * N + M) * sizeof(float);      // + the column-sum partials of mt_wgrad_bias
  return s > 1 ? (size_t)s * M * N * sizeof(float) : 0;
}

int gemm_tc(const void* A, const void* B, void* C, const float* bias, const float* addend, const void* aux,
            int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int transA, int transB,
            int in_dtype, int out_dtype, int epilogue, void* workspace, size_t workspace_bytes,
            cudaStream_t stream, float* colsum_out) {
  CUtensorMap tmA, tmB;
  int rc;
  const int a_mn = transA ? 1 : 0;        // stored [K, M]: M contiguous
  const int b_mn = transB ? 0 : 1;        // stored [K, N]: N contiguous
  if (a_mn) rc = tc::make_tmap_2d(&tmA, A, K, M, lda, 64, BK);
  else rc = tc::make_tmap_2d(&tmA, A, M, K, lda, BK, BM);
  if (rc) return rc;
  // 128 x 256 tiles when the output is at least 256 columns wide and a plain (no split-K) product
  const bool wide = a_mn ? ((N % 256 == 0) && b_mn && wgrad_wide())
                         : ((N % 256 == 0) && tc_splits(M, N, K, 128) == 1);
  const int BN = wide ? 256 : 128;
  if (b_mn) rc = tc::make_tmap_2d(&tmB, B, K, N, ldb, 64, BK);
  else rc = tc::make_tmap_2d(&tmB, B, N, K, ldb, BK, BN);
  if (rc) return rc;

  TcGemmParams p;
  p.C = C; p.bias = bias; p.addend = addend; p.aux = aux;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.epi = epilogue;
  p.out_bf16 = (out_dtype != MT_F32);
  p.out_f16 = (out_dtype == MT_F16);
  const int oes = p.out_bf16 ? 2 : 4;
  p.vec_ok = (ldc % 4 == 0) && aligned(C, 4 * oes) && (!(epilogue & MT_EPI_BIAS) || aligned(bias, 16)) &&
             (!(epilogue & MT_EPI_ADD) || aligned(addend, 16)) && (!(epilogue & MT_EPI_RELU_MASK) || aligned(aux, 8));
  p.splits = tc_splits(M, N, K, BN);
  if (p.splits > 1 && (!workspace || workspace_bytes < (size_t)p.splits * M * N * sizeof(float))) p.splits = 1;
  p.add_inplace = (epilogue & MT_EPI_ADD) && addend == C && !p.out_bf16 && p.vec_ok &&
                  !(epilogue & (MT_EPI_RELU | MT_EPI_RELU_MASK)) && (N % 4 == 0);
  p.k_per_split = K;
  if (p.splits > 1) {
    int64_t kps = (((K + p.splits - 1) / p.splits + BK - 1) / BK) * BK;
    p.k_per_split = kps;
    p.splits = (int)((K + kps - 1) / kps);            // every split owns at least one k-block
  }
  p.part = reinterpret_cast<float*>(workspace);
  p.cs_out = nullptr;
  p.cs_part = nullptr;
  if (colsum_out) {
    if (!(a_mn && b_mn && BN == 128)) { set_error("gemm_tc: column sums come with the weight-gradient form only"); return MT_E_UNSUPPORTED; }
    if (p.splits > 1 && workspace_bytes < (size_t)p.splits * (M * N + M) * sizeof(float)) {
      set_error("gemm_tc: workspace too small for the column-sum partials");
      return MT_E_WORKSPACE;
    }
    p.cs_out = colsum_out;
    p.cs_part = p.part + (int64_t)p.splits * M * N;
  }
  const int fmt = (in_dtype == MT_BF16) ? 1 : 0;

  cudaError_t e = cudaSuccess;
  if (!a_mn && p.splits == 1 && M * N >= (int64_t)BM * 128 * sm_count()) {
    // persistent kernel: enough tiles to give every SM several
    const int tiles_n = (int)((N + BN - 1) / BN), tiles_m = (int)((M + BM - 1) / BM);
    const int num_tiles = tiles_n * tiles_m;
    const unsigned pgrid = (unsigned)(num_tiles < sm_count() ? num_tiles : sm_count());
#define MT_PS_LAUNCH(BMJ, BNC)                                                                        \
    {                                                                                                 \
      auto kern = gemm_tc_persist_kernel<BMJ, BNC>;                                                   \
      static unsigned long long attr_done = 0; const unsigned long long attr_bit = attr_dev_bit();                                                                  \
      if (!(attr_done & attr_bit)) {                                                                               \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PShape<BNC>::SMEM); \
        if (e == cudaSuccess) attr_done |= attr_bit;                                                               \
      }                                                                                               \
      if (e == cudaSuccess) kern<<<pgrid, PS_THREADS, PShape<BNC>::SMEM, stream>>>(tmA, tmB, p, fmt, tiles_n, num_tiles); \
    }
    if (b_mn) { if (wide) MT_PS_LAUNCH(1, 256) else MT_PS_LAUNCH(1, 128) }
    else { if (wide) MT_PS_LAUNCH(0, 256) else MT_PS_LAUNCH(0, 128) }
#undef MT_PS_LAUNCH
    if (e != cudaSuccess) { set_error("gemm_tc: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    return check_launch("gemm_tc_persist");
  }
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)p.splits);
#define MT_TC_LAUNCH(AM, BMJ, BNC)                                                                    \
  {                                                                                                   \
    auto kern = gemm_tc_kernel<AM, BMJ, BNC>;                                                         \
    static unsigned long long attr_done = 0; const unsigned long long attr_bit = attr_dev_bit();                                                                    \
    if (!(attr_done & attr_bit)) {                                                                                 \
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Shape<BNC>::SMEM); \
      if (e == cudaSuccess) attr_done |= attr_bit;                                                                 \
    }                                                                                                 \
    if (e == cudaSuccess) kern<<<grid, TC_THREADS, Shape<BNC>::SMEM, stream>>>(tmA, tmB, p, fmt);     \
  }
  if (!a_mn && !b_mn) { if (wide) MT_TC_LAUNCH(0, 0, 256) else MT_TC_LAUNCH(0, 0, 128) }
  else if (!a_mn && b_mn) { if (wide) MT_TC_LAUNCH(0, 1, 256) else MT_TC_LAUNCH(0, 1, 128) }
  else if (a_mn && b_mn) {
    if (p.cs_out) {
      auto kern = gemm_tc_kernel<1, 1, 128, true>;
      static unsigned long long attr_done = 0; const unsigned long long attr_bit = attr_dev_bit();
      if (!(attr_done & attr_bit)) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Shape<128>::SMEM);
        if (e == cudaSuccess) attr_done |= attr_bit;
      }
      if (e == cudaSuccess) kern<<<grid, TC_THREADS, Shape<128>::SMEM, stream>>>(tmA, tmB, p, fmt);
    } else if (wide) MT_TC_LAUNCH(1, 1, 256) else MT_TC_LAUNCH(1, 1, 128)
  }
  else { set_error("gemm_tc: unsupported operand majors"); return MT_E_UNSUPPORTED; }
#undef MT_TC_LAUNCH
  if (e != cudaSuccess) { set_error("gemm_tc: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  rc = check_launch("gemm_tc");
  if (rc) return rc;
  if (p.splits > 1) {
    int64_t n = M * N;
    const unsigned cs_blocks = p.cs_out ? (unsigned)((M + 255) / 256) : 0u;
    if (epilogue == 0 && !p.out_bf16 && (N % 4 == 0) && (ldc % 4 == 0) && aligned(C, 16) && aligned(workspace, 16))
      tc_splitk_fold_vec_kernel<<<(unsigned)((n / 4 + 255) / 256) + cs_blocks, 256, 0, stream>>>(p);
    else
      tc_splitk_fold_kernel<<<(unsigned)((n + 255) / 256) + cs_blocks, 256, 0, stream>>>(p);
    rc = check_launch("gemm_tc_fold");
  }
  return rc;
}

}  // namespace mt
