// Backward of the fused relative global attention, fourth generation: the dQ / dE role of a TRAINING step.
// Like the dK/dV role (rga_tc_bwd4.cu) it reads the P tiles the forward kept instead of rebuilding them, and it
// needs no spilled dS either: dP = dO V^T is one N = 128 product per step and dS = P o (dP - D) / sqrt(dh) six
// instructions per element pair, cheaper than moving a dS tile through HBM twice.
//
// One CTA owns a query tile i of up to four consecutive heads and walks the key tiles j = 0 .. i, i.e. the tile
// diagonals d = i - j = i .. 0.  With dG = dS in band coordinates (dG[a][127 - a + b] = dS[a][b], 256 band columns
// = E rows c0 - 127 .. c0 + 128, c0 = max_seq - 1 - 128 d), per step:
//     dP   = dO_i V_j^T                                   N = 128, K = 64     (TMEM, double-buffered)
//     dQ  += dS K_j                                       N = 64,  K = 128    (dS is the TMEM A operand)
//     dQ  += dG_blk E_blk ;  dE_blk = dG_blk^T Q_i          N = 64,  K = 128    (ONE band block per step, see below)
// The band of step n is E blocks {lo, hi} with lo(n + 1) = hi(n), and inside such a block the two steps write
// DISJOINT elements (row a: step n its columns 0 .. 126 - a, step n + 1 its columns 127 - a .. 127).  So the blocks
// live in a ring of three shared-memory operands, a step stores its two halves into two of them, and the block that
// is complete after step n (its lo block) is multiplied once -- 3 products of N = 64 per step where the consumers of
// spilled dS (rga_tc_bwd3.cu) issued 5 -- and its dE product is final: it leaves TMEM through the flusher warps as
// a TMA reduction into dE, no rotating accumulators.
//
//   warps 0-15  : math, two groups of 8 on alternate steps (as in rga_tc_bwd4.cu): row a = 32 (w & 3) + lane, 64 key
//                 columns; P streams from global memory through registers
//   warp 16     : TMA loader A (dO_i per head, V_j x 2: what dP needs); L2 prefetch of P;  warp 19: loader B (Q_i, K_j, E)
//   warp 17     : issues dP (a step ahead);  warp 18: issues the dQ / dE products
//   warps 20-23 : flushers -- dE block of every step (TMEM -> 128B-swizzled staging tile -> cp.reduce.async.bulk), dQ
//                 of every head (TMEM -> global)
#include "ops.cuh"
#include "rga_tc_common.cuh"

#include <stdlib.h>

namespace mt {

using namespace rga;

namespace {

constexpr int Q4_MATH = 512;
constexpr int WQ_LOAD = 16, WQ_MMA_A = 17, WQ_MMA_B = 18, WQ_LOAD2 = 19, WQ_FLUSH = 20, QFL_THREADS = 128;
constexpr int Q4_THREADS = (WQ_FLUSH + 4) * 32;
constexpr int PT_BYTES = 2 * TILE;

// shared memory (TILE = 16 KB): Q, dO (resident per head); K; V x 2; E x 2; staging [128 x 32] fp32; dG ring 3 x 2
struct LayQ {
  static constexpr int Q = 0, DO = TILE, K = 2 * TILE, V0 = 3 * TILE, E0 = 5 * TILE, STG = 7 * TILE, DG = 8 * TILE,
                       BAR = 14 * TILE;
};
constexpr int SMEMQ = LayQ::BAR + 512;
static_assert(SMEMQ <= 232448, "shared memory budget");
// TMEM columns: dP x 2 | dS operand x 2 (16-bit pairs) | dQ | dE block
constexpr uint32_t TMQ_DP = 0, TMQ_DS = 256, TMQ_DQ = 384, TMQ_DE = 448;

enum { BQ_QF = 0, BQ_QE = 1, BQ_DOF = 2, BQ_DOE = 3, BQ_VF = 4, BQ_VE = 6, BQ_KF = 8, BQ_KE = 9, BQ_EF = 10, BQ_EE = 12,
       BQ_DPF = 14, BQ_DPE = 16, BQ_RDY = 18, BQ_DGF = 20, BQ_DEF = 23, BQ_DEE = 24, BQ_ACC = 25, BQ_DQE = 26,
       BQ_TMEM = 27 };

struct BwdQParams {
  void* dq;
  int64_t sb, sl, sh;
  float* dE;
  const float* lse; const float* delta;
  const uint8_t* stash; const float* mrow;
  int B, h, L, max_seq, nT, nTri;
  int heads_per_cta;
  int qk_fmt;                                  // 1 = bf16; 0 = f16 (dO arrives as f16(g dO), g dS is packed as f16)
  float gscale, out_scale;                     // out_scale = 1 / g on dQ and dE
  float scale;
  long long* trace; int trace_z;               // MT_RGA_TRACE=z: clock64 stamps of CTA (0,0,z), [4 agents][32 steps][8 events]
};

#define TRACEQ(agent, n, ev)                                                                        \
  do {                                                                                              \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (int)blockIdx.z == p.trace_z && (n) < 32)  \
      p.trace[((agent) * 32 + (n)) * 8 + (ev)] = clock64();                                         \
  } while (0)

// predicated shared-memory stores on 32-bit shared addresses (the band stores select among several addresses per word:
// on generic pointers that is 64-bit select / add chains and generic stores)
__device__ __forceinline__ void sts32_if(uint32_t addr, uint32_t v, bool ok) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.shared.b32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"((uint32_t)ok) : "memory");
}
__device__ __forceinline__ void sts16_if(uint32_t addr, uint32_t v, bool ok) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b16 h;\n\tsetp.ne.b32 p, %2, 0;\n\tcvt.u16.u32 h, %1;\n\t@p st.shared.b16 [%0], h;\n\t}" ::"r"(addr), "r"(v), "r"((uint32_t)ok) : "memory");
}

__device__ __forceinline__ int64_t tile_index_q(const BwdQParams& p, int b, int hh, int it, int jt) {
  return ((int64_t)b * p.h + hh) * p.nTri + (it * (it + 1) / 2 + jt);
}

template <bool HF>
__global__ void __launch_bounds__(Q4_THREADS, 1)
rga_bwd4_dqe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                    const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmDE,
                    const BwdQParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LayQ::BAR);
  uint64_t* q_full = bars + BQ_QF;        // loader -> dQ/dE issuer: Q of the head
  uint64_t* q_empty = bars + BQ_QE;       // the head's last products are done with Q
  uint64_t* do_full = bars + BQ_DOF;      // loader -> dP issuer: dO of the head
  uint64_t* do_empty = bars + BQ_DOE;     // the head's last dP product is done with dO
  uint64_t* v_full = bars + BQ_VF;        // [2]
  uint64_t* v_empty = bars + BQ_VE;       // [2]
  uint64_t* k_full = bars + BQ_KF;
  uint64_t* k_empty = bars + BQ_KE;       // dS.K of the step is done with the K slot
  uint64_t* e_full = bars + BQ_EF;        // [2] the lo E block of step n is in slot n & 1
  uint64_t* e_empty = bars + BQ_EE;       // [2]
  uint64_t* dp_full = bars + BQ_DPF;      // [2] dP issuer -> math group
  uint64_t* dp_empty = bars + BQ_DPE;     // [2] math group -> dP issuer
  uint64_t* ready = bars + BQ_RDY;        // [2] math group -> dQ/dE issuer: dS in TMEM, the step's band halves in shared memory
  uint64_t* dg_free = bars + BQ_DGF;      // [3] the products of step n are done (dG ring slot n % 3, the TMEM dS slot n & 1)
  uint64_t* de_full = bars + BQ_DEF;      // dQ/dE issuer -> flushers: the step's dE block is final
  uint64_t* de_empty = bars + BQ_DEE;     // flushers -> issuer: it has left TMEM
  uint64_t* acc_done = bars + BQ_ACC;     // issuer -> flushers: the head's dQ is final
  uint64_t* dq_empty = bars + BQ_DQE;     // flushers -> issuer: it has left TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BQ_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int it = p.nT - 1 - (int)blockIdx.z, b = blockIdx.y, hh0 = (int)blockIdx.x * p.heads_per_cta;
  const int per = it + 1;                                      // steps (key tiles) per head
  const int n_items = min(p.heads_per_cta, p.h - hh0);
  const int nsteps = n_items * per;

  if (warp == WQ_LOAD && lane == 0) {
    tc::tma_prefetch_desc(&tmQ); tc::tma_prefetch_desc(&tmK); tc::tma_prefetch_desc(&tmV);
    tc::tma_prefetch_desc(&tmDO); tc::tma_prefetch_desc(&tmE); tc::tma_prefetch_desc(&tmDE);
    tc::mbar_init(q_full, 1); tc::mbar_init(q_empty, 1); tc::mbar_init(do_full, 1); tc::mbar_init(do_empty, 1);
    tc::mbar_init(k_full, 1); tc::mbar_init(k_empty, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&v_full[s], 1); tc::mbar_init(&v_empty[s], 1);
      tc::mbar_init(&e_full[s], 1); tc::mbar_init(&e_empty[s], 1);
      tc::mbar_init(&dp_full[s], 1); tc::mbar_init(&dp_empty[s], Q4_MATH / 64);
      tc::mbar_init(&ready[s], Q4_MATH / 64);
    }
    for (int s = 0; s < 3; ++s) tc::mbar_init(&dg_free[s], 1);
    tc::mbar_init(de_full, 1); tc::mbar_init(de_empty, QFL_THREADS / 32);
    tc::mbar_init(acc_done, 1); tc::mbar_init(dq_empty, QFL_THREADS / 32);
    tc::fence_barrier_init();
  }
  if (warp == WQ_MMA_A) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint64_t TS16 = TILE >> 4;

  if (warp == WQ_LOAD) {
    // ================================ loader A: what dP needs (dO per head, V_j) ==================
    // (two loader threads: the waits of the late operands -- K, E: released by the products of the previous step --
    // must not hold back the loads the dP issuer needs a step ahead)
    if (lane == 0) {
      int jt = 0, item = 0;
      for (int n = 0; n < nsteps; ++n) {
        const int hh = hh0 + item;
        if (jt == 0) {
          if (item >= 1) tc::mbar_wait(do_empty, (item - 1) & 1);
          tc::mbar_arrive_expect_tx(do_full, TILE);
          tc::tma_load_4d(smem + LayQ::DO, &tmDO, do_full, 0, hh, it * TT, b);
        }
        if (n >= 2) tc::mbar_wait(&v_empty[n & 1], ((n >> 1) - 1) & 1);       // dP of step n - 2 is done with the slot
        tc::mbar_arrive_expect_tx(&v_full[n & 1], TILE);
        tc::tma_load_4d(smem + LayQ::V0 + (n & 1) * TILE, &tmV, &v_full[n & 1], 0, hh, jt * TT, b);
        {                         // the P tile and its row references three steps ahead -> L2
          int j3 = jt + 3, i3 = item;
          while (j3 >= per) { j3 -= per; ++i3; }
          if (i3 < n_items) {
            const int64_t tix = tile_index_q(p, b, hh0 + i3, it, j3);
            tc::bulk_prefetch_l2(p.stash + tix * (int64_t)PT_BYTES, PT_BYTES);
            tc::bulk_prefetch_l2(p.mrow + tix * 2 * TT, 2 * TT * 4);
          }
        }
        if (++jt == per) { jt = 0; ++item; }
      }
    }
  } else if (warp == WQ_LOAD2) {
    // ================================ loader B: Q per head, K_j, one new E block per step ==========
    if (lane == 0) {
      int jt = 0, item = 0;
      for (int n = 0; n < nsteps; ++n) {
        const int hh = hh0 + item;
        const bool first = (jt == 0), last = (jt == it);
        if (first) {
          if (item >= 1) tc::mbar_wait(q_empty, (item - 1) & 1);
          tc::mbar_arrive_expect_tx(q_full, TILE);
          tc::tma_load_4d(smem + LayQ::Q, &tmQ, q_full, 0, hh, it * TT, b);
        }
        if (n >= 1) tc::mbar_wait(k_empty, (n - 1) & 1);
        tc::mbar_arrive_expect_tx(k_full, TILE);
        tc::tma_load_4d(smem + LayQ::K, &tmK, k_full, 0, hh, jt * TT, b);
        // E blocks: the lo block of step n lives in slot n & 1.  A head's first step loads its own; every step but the
        // head's last loads its hi block = the lo block of step n + 1 (the hi block of the diagonal tile is never used:
        // those are the j > i positions)
        const int c0 = p.max_seq - 1 - (it - jt) * TT;
        if (first) {
          if (n >= 2) tc::mbar_wait(&e_empty[n & 1], ((n >> 1) - 1) & 1);
          tc::mbar_arrive_expect_tx(&e_full[n & 1], TILE);
          tc::tma_load_2d(smem + LayQ::E0 + (n & 1) * TILE, &tmE, &e_full[n & 1], 0, c0 - (TT - 1));
        }
        if (!last) {
          const int s1 = (n + 1) & 1;
          if (n >= 1) tc::mbar_wait(&e_empty[s1], (((n + 1) >> 1) - 1) & 1);     // products of step n - 1
          tc::mbar_arrive_expect_tx(&e_full[s1], TILE);
          tc::tma_load_2d(smem + LayQ::E0 + s1 * TILE, &tmE, &e_full[s1], 0, c0 + 1);
        }
        if (++jt == per) { jt = 0; ++item; }
      }
    }
  } else if (warp == WQ_MMA_A) {
    // ================================ dP issuer =============================================
    if (lane == 0) {
      const uint32_t id_kk = tc::make_idesc(TT, TT, p.qk_fmt, p.qk_fmt, 0, 0);       // dO (K-major) x V (K-major), N = 128
      const uint64_t vd0 = tc::make_sdesc(tc::smem_u32(smem + LayQ::V0), 16, 1024);
      const uint64_t dod = tc::make_sdesc(tc::smem_u32(smem + LayQ::DO), 16, 1024);
      int jt = 0, item = 0;
      for (int n = 0; n < nsteps; ++n) {
        TRACEQ(1, n, 0);
        if (jt == 0) tc::mbar_wait(do_full, item & 1);
        tc::mbar_wait(&v_full[n & 1], (n >> 1) & 1);
        TRACEQ(1, n, 1);
        tc::mbar_wait(&dp_empty[n & 1], ((n >> 1) & 1) ^ 1);          // the math group has read dP of step n - 2
        tc::tc_fence_after();
        TRACEQ(1, n, 2);
        const uint64_t vd = vd0 + (uint64_t)(n & 1) * TS16;
#pragma unroll
        for (int k4 = 0; k4 < DHC / 16; ++k4)
          tc::umma_f16(tmem + TMQ_DP + 128 * (uint32_t)(n & 1), dod + 2 * k4, vd + 2 * k4, id_kk, k4 != 0);
        tc::umma_commit(&dp_full[n & 1]);
        tc::umma_commit(&v_empty[n & 1]);
        TRACEQ(1, n, 3);
        if (++jt == per) { tc::umma_commit(do_empty); jt = 0; ++item; }
      }
    }
  } else if (warp == WQ_MMA_B) {
    // ================================ dQ / dE issuer ========================================
    if (lane == 0) {
      const uint32_t id_kmn = tc::make_idesc(TT, DHC, p.qk_fmt, p.qk_fmt, 0, 1);     // A K-major (TMEM dS / dG block), B MN-major (K / E), N = 64
      const uint32_t id_mnmn = tc::make_idesc(TT, DHC, p.qk_fmt, p.qk_fmt, 1, 1);    // A MN-major (dG block), B MN-major (Q), N = 64
      const uint64_t kd_mn = tc::make_sdesc(tc::smem_u32(smem + LayQ::K), 1024, 1024);
      const uint64_t ed_mn0 = tc::make_sdesc(tc::smem_u32(smem + LayQ::E0), 1024, 1024);
      const uint64_t qd_mn = tc::make_sdesc(tc::smem_u32(smem + LayQ::Q), 1024, 1024);
      const uint64_t dg_k0 = tc::make_sdesc(tc::smem_u32(smem + LayQ::DG), 16, 1024);
      const uint64_t dg_mn0 = tc::make_sdesc(tc::smem_u32(smem + LayQ::DG), TILE, 1024);
      int jt = 0, item = 0, r3 = 0;                       // r3 = n % 3
      for (int n = 0; n < nsteps; ++n) {
        TRACEQ(2, n, 0);
        if (jt == 0) {
          tc::mbar_wait(q_full, item & 1);
          if (item >= 1) tc::mbar_wait(dq_empty, (item - 1) & 1);      // the previous head's dQ has left TMEM
        }
        tc::mbar_wait(&e_full[n & 1], (n >> 1) & 1);
        TRACEQ(2, n, 1);
        tc::mbar_wait(&ready[n & 1], (n >> 1) & 1);
        tc::tc_fence_after();
        TRACEQ(2, n, 2);
        // Order: the block products first, dS . K last.  The single K slot is released by the previous step's dS . K --
        // its LAST products -- so the new tile has the time of 16 products to land; and the dE product of the previous
        // step, in the middle of its sequence, has left TMEM through the flushers by the time this step's is issued.
        const uint64_t dgk = dg_k0 + (uint64_t)r3 * 2 * TS16, dgm = dg_mn0 + (uint64_t)r3 * 2 * TS16;
        const uint64_t ed = ed_mn0 + (uint64_t)(n & 1) * TS16;
#pragma unroll
        for (int k16 = 0; k16 < TT / 16; ++k16)           // dQ += dG_blk . E_blk (contraction over the block's 128 band columns)
          tc::umma_f16(tmem + TMQ_DQ, dgk + (uint64_t)(k16 >> 2) * TS16 + 2 * (k16 & 3), ed + 128 * k16, id_kmn, (jt | k16) != 0);
        if (n >= 1) { tc::mbar_wait(de_empty, (n - 1) & 1); tc::tc_fence_after(); }     // the previous block has left TMEM
#pragma unroll
        for (int k16 = 0; k16 < TT / 16; ++k16)           // dE_blk = dG_blk^T . Q_i (contraction over the 128 query rows)
          tc::umma_f16(tmem + TMQ_DE, dgm + 128 * k16, qd_mn + 128 * k16, id_mnmn, k16 != 0);
        tc::umma_commit(de_full);
        tc::umma_commit(&e_empty[n & 1]);
        tc::mbar_wait(k_full, n & 1);
        tc::tc_fence_after();
#pragma unroll
        for (int k16 = 0; k16 < TT / 16; ++k16)           // dQ += dS . K_j : dS is the TMEM A operand (8 columns per 16 keys)
          tc::umma_f16_ts(tmem + TMQ_DQ, tmem + TMQ_DS + 64 * (uint32_t)(n & 1) + 8 * k16, kd_mn + 128 * k16, id_kmn, 1);
        tc::umma_commit(k_empty);
        tc::umma_commit(&dg_free[r3]);
        TRACEQ(2, n, 3);
        if (++r3 == 3) r3 = 0;
        if (++jt == per) { tc::umma_commit(acc_done); tc::umma_commit(q_empty); jt = 0; ++item; }
      }
    }
  } else if (warp >= WQ_FLUSH) {
    // ================================ flushers ==============================================
    // warp f = warp - 20 reads TMEM lanes 32 f .. +31 (block row / query row a).  A dE block leaves in two halves of 32
    // columns through the 128B-swizzled staging tile [128 rows x 32 fp32] as TMA reductions (rows outside
    // [0, max_seq) are clipped by the tensor map); the TMEM slot is released as soon as both halves are in registers.
    const int a = (warp - WQ_FLUSH) * 32 + lane, ftid = threadIdx.x - WQ_FLUSH * 32;
    const uint32_t lane_base = (uint32_t)((warp - WQ_FLUSH) * 32) << 16;
    uint8_t* const stg = smem + LayQ::STG;
    const float osc = p.out_scale;
    int jt = 0, item = 0;
    for (int n = 0; n < nsteps; ++n) {
      const int erow0 = p.max_seq - TT * (it - jt + 1);         // first E row of the step's lo block
      tc::mbar_wait(de_full, n & 1);
      tc::tc_fence_after();
      uint32_t r0[32], r1[32];
      tc::tmem_ld_32x32(tmem + TMQ_DE + lane_base, r0);
      tc::tmem_ld_32x32(tmem + TMQ_DE + lane_base + 32, r1);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      tc::mbar_arrive_warp(de_empty);
      if (erow0 < p.max_seq && erow0 + TT > 0) {                // (uniform) else nothing of the block exists
        if (erow0 < 0) {
          // the block that straddles E row 0 (max_seq not a multiple of the tile edge; once per head): a bulk-tensor
          // reduction at a negative row coordinate faults on sm_100a, so these rows leave through vector reductions
          // (as the general flush they measured 0.66 vs 0.556 ms for the backward of a config-B layer)
          const int erow = erow0 + a;
          if (erow >= 0 && erow < p.max_seq) {
            float* dst = p.dE + (int64_t)erow * DHC;
#pragma unroll
            for (int x = 0; x < 32; x += 4) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + x), "f"(__uint_as_float(r0[x]) * osc),
                           "f"(__uint_as_float(r0[x + 1]) * osc), "f"(__uint_as_float(r0[x + 2]) * osc),
                           "f"(__uint_as_float(r0[x + 3]) * osc) : "memory");
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 32 + x), "f"(__uint_as_float(r1[x]) * osc),
                           "f"(__uint_as_float(r1[x + 1]) * osc), "f"(__uint_as_float(r1[x + 2]) * osc),
                           "f"(__uint_as_float(r1[x + 3]) * osc) : "memory");
            }
          }
        } else {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            if (ftid == 0) tc::bulk_wait_read0();               // the previous reduction has read the staging tile
            tc::named_bar_sync(9, QFL_THREADS);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t* r = hf ? r1 : r0;
              *reinterpret_cast<float4*>(stg + swz_chunk(a, c)) =
                  make_float4(__uint_as_float(r[4 * c]) * osc, __uint_as_float(r[4 * c + 1]) * osc,
                              __uint_as_float(r[4 * c + 2]) * osc, __uint_as_float(r[4 * c + 3]) * osc);
            }
            tc::fence_proxy_async();
            tc::named_bar_sync(9, QFL_THREADS);
            if (ftid == 0) {
              tc::tma_reduce_add_2d(&tmDE, stg, 32 * hf, erow0);
              tc::bulk_commit();
            }
          }
        }
      }
      if (++jt == per) {          // the head's dQ: row a, 64 columns
        tc::mbar_wait(acc_done, item & 1);
        tc::tc_fence_after();
        tc::tmem_ld_32x32(tmem + TMQ_DQ + lane_base, r0);
        tc::tmem_ld_32x32(tmem + TMQ_DQ + lane_base + 32, r1);
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        tc::mbar_arrive_warp(dq_empty);
        const int row = it * TT + a;
        if (row < p.L) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dq) + (int64_t)b * p.sb + (int64_t)row * p.sl +
                                                (int64_t)(hh0 + item) * p.sh);
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            dst[x] = make_uint4(pack_bf16x2(__uint_as_float(r0[8 * x]) * osc, __uint_as_float(r0[8 * x + 1]) * osc),
                                pack_bf16x2(__uint_as_float(r0[8 * x + 2]) * osc, __uint_as_float(r0[8 * x + 3]) * osc),
                                pack_bf16x2(__uint_as_float(r0[8 * x + 4]) * osc, __uint_as_float(r0[8 * x + 5]) * osc),
                                pack_bf16x2(__uint_as_float(r0[8 * x + 6]) * osc, __uint_as_float(r0[8 * x + 7]) * osc));
            dst[4 + x] = make_uint4(pack_bf16x2(__uint_as_float(r1[8 * x]) * osc, __uint_as_float(r1[8 * x + 1]) * osc),
                                    pack_bf16x2(__uint_as_float(r1[8 * x + 2]) * osc, __uint_as_float(r1[8 * x + 3]) * osc),
                                    pack_bf16x2(__uint_as_float(r1[8 * x + 4]) * osc, __uint_as_float(r1[8 * x + 5]) * osc),
                                    pack_bf16x2(__uint_as_float(r1[8 * x + 6]) * osc, __uint_as_float(r1[8 * x + 7]) * osc));
          }
        }
        jt = 0; ++item;
      }
    }
    if (ftid == 0) tc::bulk_wait0();
  } else if (warp < Q4_MATH / 32) {
    // ================================ math warps ============================================
    const int w4 = warp & 3, hq = (warp >> 2) & 1, grp = warp >> 3;
    const int a = w4 * 32 + lane, a7 = a & 7;
    const uint32_t lane_base = (uint32_t)(w4 * 32) << 16;
    const int64_t my_off = ((int64_t)(8 * hq + w4) * 128 + lane) * 16;      // stash layout: see rga_tc.cu / rga_tc_bwd4.cu
    const uint32_t dp_addr = tmem + TMQ_DP + 128 * (uint32_t)grp + 64 * (uint32_t)hq + lane_base;
    const uint32_t ds_addr = tmem + TMQ_DS + 64 * (uint32_t)grp + 32 * (uint32_t)hq + lane_base;
    // band coordinates of the thread's 64 values: columns sh + 64 hq + x, sh = 127 - a.  32-bit word w of the 256-column
    // band holds columns 2 w, 2 w + 1; for odd sh every output word takes its halves from two neighbouring values
    const int sh = 127 - a;
    const bool odd = sh & 1;
    const int base_w = (sh >> 1) + 32 * hq;
    const int cb = base_w >> 2, r0w = base_w & 3;
    const uint32_t sel = odd ? 0x5432u : 0x7654u;
    const uint32_t dg_row = tc::smem_u32(smem + LayQ::DG) + a * 128;
    // word y of a chunk iteration: which of the iteration's three chunk addresses, and the offset inside the chunk
    int woff[4]; bool wcarry[4];
#pragma unroll
    for (int y = 0; y < 4; ++y) { wcarry[y] = (r0w + y) >= 4; woff[y] = ((r0w + y) & 3) << 2; }

    int jt = grp, item = 0;
    while (jt >= per) { jt -= per; ++item; }
    int fj = jt, fitem = item;
    // tile (head hh0 + x, it, j) of the stash = tile0 + x nTri + j; the row statistics of head hh0 + x = row0 + x L
    const int64_t tile0 = tile_index_q(p, b, hh0, it, 0);
    const uint8_t* const stash0 = p.stash + tile0 * (int64_t)PT_BYTES + my_off;
    const float* const mrow0 = p.mrow + (tile0 * 2 + hq) * TT + a;        // [tile][key half][row]: this thread's half
    const bool row_ok = it * TT + a < p.L;
    const int64_t row0 = ((int64_t)b * p.h + hh0) * p.L + (row_ok ? it * TT + a : 0);
    auto fetch = [&](uint32_t (&R)[32], float& lse_raw, float& d_raw, float& mref) {
      if (fitem < n_items) {
        const int trel = fitem * p.nTri + fj;
        const uint8_t* src = stash0 + (int64_t)trel * PT_BYTES;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          tc::ldg128_stream(src + 8192 * (c >> 2) + 512 * (c & 3), R[4 * c], R[4 * c + 1], R[4 * c + 2], R[4 * c + 3]);
        // (raw values only: any arithmetic here would wait for the loads inside the fetch)
        const int64_t ro = row0 + (int64_t)fitem * p.L;
        lse_raw = __ldg(p.lse + ro); d_raw = __ldg(p.delta + ro);
        mref = __ldg(mrow0 + (int64_t)trel * 2 * TT);
        fj += 2;
        while (fj >= per) { fj -= per; ++fitem; }
      }
    };
    uint32_t R[32];
    float nl = 0.f, nd = 0.f, nm = 0.f;
    fetch(R, nl, nd, nm);
    int r3 = grp;                                   // n % 3
    for (int n = grp, m = 0; n < nsteps; n += 2, ++m) {
      const bool tr = (lane == 0 && w4 == 0 && hq == 0);
      if (tr) TRACEQ(0, n, 0);
      // dS = P_stored f (dP - D) / sqrt(dh) = P_stored (dP sf - Dsf); (f16 mode: dP and D carry the loss scale g)
      const float f = tc::fast_exp2(nm - (row_ok ? nl : 0.f) * LOG2E);
      const float sf = p.scale * f, Dsf = (row_ok ? nd : 0.f) * p.scale * (HF ? p.gscale : 1.f) * f;
      const bool first = (jt == 0), last = (jt == it);
      if (tr) TRACEQ(0, n, 1);
      tc::mbar_wait(&dp_full[grp], m & 1);
      // the products of step n - 2 (and with them those of step n - 3: one issuing thread, in order) are done: the
      // TMEM dS slot and the ring slots this step writes are free
      if (n >= 2) {
        const int q3 = r3 == 0 ? 1 : (r3 == 1 ? 2 : 0);       // (n - 2) % 3
        tc::mbar_wait(&dg_free[q3], ((n - 2) / 3) & 1);
      }
      tc::tc_fence_after();
      if (tr) TRACEQ(0, n, 2);
      // ring slots: the step's lo block in slot n % 3, its hi block in slot (n + 1) % 3
      const uint32_t lo_row = dg_row + r3 * PT_BYTES;
      const uint32_t hi_row = dg_row + (r3 == 2 ? 0 : r3 + 1) * PT_BYTES;
      if (first) {
        // a head's first lo block gets no contribution from a previous step: zero the row (this thread: sub-tile hq),
        // then the row's two threads meet before either stores into it
#pragma unroll
        for (int c = 0; c < 8; ++c)
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(lo_row + hq * TILE + (c << 4)), "r"(0u) : "memory");
        tc::named_bar_sync(1 + grp * 4 + w4, 64);
      }
      // address of band chunk c (16 bytes = 8 columns; 32 chunks: 0-15 the lo block, 16-31 the hi block)
      auto chunk_addr = [&](int c) -> uint32_t {
        return ((c & 16) ? hi_row : lo_row) + ((c >> 3) & 1) * TILE + (((c & 7) ^ a7) << 4);
      };
      uint32_t prev = 0;                            // odd shift: the value waiting for its partner
      // (the hi block of a head's last step is never used -- the diagonal tile's j > i positions, all zero -- and its
      // slot may already belong to the next head: words of chunks >= 16 are not stored there)
      const int wlim = last ? 64 : 1024;
#pragma unroll
      for (int q = 0; q < 4; ++q) {                 // 16 key columns = 8 packed words at a time
        uint32_t dp[16];
        tc::tmem_ld_32x16(dp_addr + 16 * q, dp);
        tc::tmem_ld_wait();
        if (q == 3) {
          tc::tc_fence_before();
          tc::mbar_arrive_warp(&dp_empty[grp]);
        }
        uint32_t D8[8];
#pragma unroll
        for (int y = 0; y < 8; ++y) {
          const uint32_t rv = R[8 * q + y];
          float p0, p1;
          if (HF) { const float2 pf = __half22float2(*reinterpret_cast<const __half2*>(&rv)); p0 = pf.x; p1 = pf.y; }
          else { p0 = __uint_as_float(rv << 16); p1 = __uint_as_float(rv & 0xffff0000u); }
          const float d0 = fmaf(__uint_as_float(dp[2 * y]), sf, -Dsf) * p0;
          const float d1 = fmaf(__uint_as_float(dp[2 * y + 1]), sf, -Dsf) * p1;
          D8[y] = HF ? pack_f16x2(d0, d1) : pack_bf16x2(d0, d1);
        }
        tc::tmem_st_32x8(ds_addr + 8 * q, D8);      // A operand of dS . K_j: row = lane, 16 keys = 8 columns
        // band stores: words base_w + 8 q + y.  Three chunk addresses cover the eight (nine) words.  Branch-free: the
        // shift parity alternates from lane to lane, so the odd case is a byte permute with a per-lane selector (even
        // lanes select the word itself), and only the two ends of the run are 16-bit stores.
        const uint32_t ca0 = chunk_addr(cb + 2 * q), ca1 = chunk_addr(cb + 2 * q + 1), ca2 = chunk_addr(cb + 2 * q + 2);
        const int w0 = base_w + 8 * q;
#pragma unroll
        for (int y = 0; y < 8; ++y) {
          const uint32_t wp = (y < 4 ? (wcarry[y] ? ca1 : ca0) : (wcarry[y - 4] ? ca2 : ca1)) + woff[y & 3];
          const uint32_t ow = __byte_perm(y == 0 ? prev : D8[y - 1], D8[y], sel);
          const bool ok = (w0 + y < wlim);
          if (q == 0 && y == 0) {                   // the run's first word: an odd shift owns its high half only
            sts16_if(wp + 2, ow >> 16, ok);
            sts16_if(wp, ow, ok && !odd);
          } else {
            sts32_if(wp, ow, ok);
          }
        }
        prev = D8[7];
        if (q == 3)                                 // odd shift: the run's last value is the low half of word base_w + 32
          sts16_if(ca2 + woff[0], prev >> 16, odd && (w0 + 8 < wlim));
      }
      if (tr) TRACEQ(0, n, 3);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      tc::fence_proxy_async();
      tc::mbar_arrive_warp(&ready[grp]);
      if (tr) TRACEQ(0, n, 4);
      fetch(R, nl, nd, nm);                         // the group's next tile and statistics
      jt += 2;
      while (jt >= per) { jt -= per; ++item; }
      r3 += 2;
      if (r3 >= 3) r3 -= 3;
      if (tr) TRACEQ(0, n, 5);
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == WQ_MMA_A) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

template <bool HF>
int launch4q(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmDO,
             const CUtensorMap& tmE, const CUtensorMap& tmDE, const BwdQParams& p, dim3 grid, cudaStream_t st) {
  auto kern = rga_bwd4_dqe_kernel<HF>;
  static unsigned long long attr_done = 0; const unsigned long long attr_bit = attr_dev_bit();
  if (!(attr_done & attr_bit)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEMQ);
    if (e != cudaSuccess) { set_error("rga_bwd4_dqe: smem attribute (%d B): %s", SMEMQ, cudaGetErrorString(e)); return (int)e; }
    attr_done |= attr_bit;
  }
  BwdQParams q = p;
  static const bool want_trace = getenv("MT_RGA_TRACE") != nullptr;
  static long long* trace_dev = nullptr;
  const size_t trace_n = 4 * 32 * 8;
  if (want_trace) {
    if (!trace_dev) cudaMalloc(&trace_dev, trace_n * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, trace_n * sizeof(long long), st);
    q.trace = trace_dev;
    q.trace_z = atoi(getenv("MT_RGA_TRACE"));
  }
  kern<<<grid, Q4_THREADS, SMEMQ, st>>>(tmQ, tmK, tmV, tmDO, tmE, tmDE, q);
  if (want_trace) {
    static long long host[4 * 32 * 8];
    cudaMemcpyAsync(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    long long t0 = 0;
    for (size_t x = 0; x < trace_n; ++x) if (host[x] && (!t0 || host[x] < t0)) t0 = host[x];
    static const char* agent[4] = {"MATH", "DP", "DQE", "-"};
    for (int ag = 0; ag < 4; ++ag)
      for (int n = 0; n < 32; ++n) {
        bool any = false;
        for (int e = 0; e < 8; ++e) any |= host[(ag * 32 + n) * 8 + e] != 0;
        if (!any) continue;
        fprintf(stderr, "trace4q %-4s step %2d:", agent[ag], n);
        for (int e = 0; e < 8; ++e) fprintf(stderr, " %8lld", host[(ag * 32 + n) * 8 + e] ? host[(ag * 32 + n) * 8 + e] - t0 : -1LL);
        fprintf(stderr, "\n");
      }
  }
  return check_launch("rga_bwd4_dqe");
}

}  // namespace

// dQ and dE from the forward's P stash (query-tile owner walks the key tiles at or left of it)
int rga_bwd4_dqe(const RgaArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                 const CUtensorMap& tmDO, const CUtensorMap& tmE, int qk_fmt, float gscale, cudaStream_t st) {
  BwdQParams p;
  p.dq = a.dq; p.sb = a.sb; p.sl = a.sl; p.sh = a.sh; p.dE = a.dE;
  p.lse = a.lse; p.delta = a.delta;
  p.B = a.B; p.h = a.h; p.L = a.L; p.max_seq = a.max_seq;
  p.nT = (a.L + TT - 1) / TT;
  p.nTri = p.nT * (p.nT + 1) / 2;
  p.stash = static_cast<const uint8_t*>(a.pstash);
  p.mrow = reinterpret_cast<const float*>(p.stash + (int64_t)a.B * a.h * p.nTri * (int64_t)PT_BYTES);      // then [tile][key half][128 rows] fp32
  p.qk_fmt = qk_fmt;
  p.gscale = gscale;
  p.out_scale = 1.f / gscale;
  p.scale = 1.f / a.inv_scale_div;
  p.trace = nullptr; p.trace_z = 0;
  CUtensorMap tmDE;
  int rc;
  if ((rc = tc::make_tmap_2d_f32(&tmDE, a.dE, a.max_seq, DHC, DHC, 32, TT))) return rc;
  static const int hpc_env = getenv("MT_DQ_HPC") ? atoi(getenv("MT_DQ_HPC")) : 0;
  // (measured at config B, 16 x 8 heads x 16 tiles: 8 heads per CTA 0.546 ms for the two kernels, 4 heads 0.560 ms, 2 heads
  // 0.587 ms -- the per-CTA fixed cost, ~17 k cycles, outweighs the coarser balance down to ~1.5 CTAs per SM)
  // ... but the longest CTA (c heads x nT steps) must stay near the per-SM average of the launch, or it alone sets the
  // makespan (config C, 12 heads x 32 tiles: 8 + 4 heads per CTA made the longest CTA 256 steps against an average of 171)
  int hpc = 1;
  const int64_t avg_steps = (int64_t)a.B * a.h * p.nTri / sm_count();
  for (int c = 8; c > 1; c >>= 1)
    if (2 * (int64_t)((a.h + c - 1) / c) * a.B * p.nT >= 3 * (int64_t)sm_count() && 10 * (int64_t)c * p.nT <= 11 * avg_steps) { hpc = c; break; }
  if (hpc_env > 0) hpc = hpc_env;
  p.heads_per_cta = hpc > a.h ? a.h : hpc;
  const dim3 grid((a.h + p.heads_per_cta - 1) / p.heads_per_cta, a.B, p.nT);
  if (qk_fmt == 0) return launch4q<true>(tmQ, tmK, tmV, tmDO, tmE, tmDE, p, grid, st);
  return launch4q<false>(tmQ, tmK, tmV, tmDO, tmE, tmDE, p, grid, st);
}

}  // namespace mt
