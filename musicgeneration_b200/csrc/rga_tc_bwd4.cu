// Backward of the fused relative global attention, fourth generation: the backward of a TRAINING step
// reads the attention probabilities the forward computed instead of rebuilding them (K2).
//
// Rebuilding P costs the backward S = Q K^T, G = Q E_band^T (two blocks), the skew (a per-row shift of a
// 64-column window: ~200 instructions per thread and step) and the exponentials -- 3 of the 6 tile
// products of the dK/dV role and two thirds of its math-warp instructions (rga_tc_bwd2.cu: a 4.3 k-cycle
// step of which ~3 k are those instructions).  The forward has every P tile in registers, packed as the
// 16-bit UMMA operand, at the moment it hands it to the P.V product; with 180 GB of HBM the cheaper trade
// is to keep them: rga_tc.cu stores each thread's 64 bytes at their place in the 128B-swizzled
// [128 x 128] operand image (B h nT(nT+1)/2 images of 32 KB: 557 MB per layer at config B, 3.3 GB for the
// model) together with the row reference m the tile was exponentiated against (online softmax:
// P_stored = exp2(s - m), the final statistics are not known yet), so that
//     P = P_stored * f,   f[a] = exp2(m[a] - lse2[a])          (one MUFU per row and step)
// and the math of a step shrinks to  dS = P o (dP - D) / sqrt(dh)  (~6 instructions per element pair).
//
// This file: the dK/dV role.  One CTA owns a key tile j of up to four consecutive heads and walks the
// query tiles i >= j:
//     dP = dO_i V_j^T                        (N = 128, K = 64; TMEM, double-buffered)
//     dV += P^T dO_i ;  dK += dS^T Q_i        (N = 64, K = 128; A = the P / dS operand images, MN-major)
//   warps 0-15 : math, two groups of 8 on alternate steps -- row a = 32 (w & 3) + lane, 64 key columns.  The P tile
//                streams from global memory through registers (coalesced 16-byte loads, fetched when the group's
//                previous step is done), P (rescaled) and dS go to shared memory as operand images
//   warp 16    : TMA loader of V_j (per head), Q_i, dO_i; L2 prefetch of the P tiles three steps ahead
//   warp 17    : issues dP (runs a step ahead of the math: two TMEM buffers)
//   warp 18    : issues dV / dK, and -- while the consumers of rga_tc_bwd3.cu still read spilled dS -- the
//                bulk store of the dS image
// The accumulators alternate between two TMEM sets from head to head, so a head's dK / dV leave TMEM one
// step into the next head (no drain bubble).
#include "ops.cuh"
#include "rga_tc_common.cuh"

#include <stdlib.h>

namespace mt {

using namespace rga;

namespace {

constexpr int K4_MATH = 512;
constexpr int W4_LOAD = K4_MATH / 32, W4_MMA_A = W4_LOAD + 1, W4_MMA_B = W4_LOAD + 2;
constexpr int K4_THREADS = K4_MATH + 96;
constexpr int PT_BYTES = 2 * TILE;            // one P (or dS) tile image: two [128 x 64] swizzled sub-tiles

// shared memory (TILE = 16 KB units): V (resident per head), Q x 2, dO x 3, P (rescaled) x 2, dS x 2
struct Lay4 {
  static constexpr int V0 = 0, Q0 = TILE, DO0 = 3 * TILE, PP = 6 * TILE, DS0 = 10 * TILE, BAR = 14 * TILE;
};
constexpr int SMEM4 = Lay4::BAR + 512;
static_assert(SMEM4 <= 232448, "shared memory budget");
// TMEM columns: dP x 2 | {dK, dV} of even heads | {dK, dV} of odd heads
constexpr uint32_t TM4_DP = 0, TM4_ACC = 256;

enum { B4_VF = 0, B4_VE = 2, B4_QF = 4, B4_QE = 6, B4_DOF = 8, B4_DOE = 11, B4_DPF = 14, B4_DPE = 16,
       B4_PPE = 18, B4_DSE = 20, B4_RDY = 22, B4_ACC = 24, B4_ACF = 26, B4_TMEM = 28 };

struct Bwd4Params {
  void* dk; void* dv;
  int64_t sb, sl, sh;
  const float* lse; const float* delta;
  const uint8_t* stash; const float* mrow;     // the forward's P tiles and row references
  uint8_t* ds_ws;                              // != NULL: spill every dS tile image (consumed by rga_tc_bwd3.cu)
  int B, h, L, nT, nTri;
  int heads_per_cta;
  int qk_fmt;                                  // 16-bit format of every MMA operand (1 = bf16; 0 = f16, the first encoder layer:
                                               // dO arrives as f16(g dO), P / g dS are f16, dK / dV leave multiplied by 1/g)
  float gscale, inv_gscale;
  float scale;                                 // 1 / sqrt(dh)
  long long* trace;                            // MT_RGA_TRACE=z: clock64 stamps of CTA (0,0,z), [4 agents][32 steps][8 events]
  int trace_z;
};

#define TRACE4(agent, n, ev)                                                                        \
  do {                                                                                              \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (int)blockIdx.z == p.trace_z && (n) < 32)  \
      p.trace[((agent) * 32 + (n)) * 8 + (ev)] = clock64();                                         \
  } while (0)

__device__ __forceinline__ int64_t tile_index(const Bwd4Params& p, int b, int hh, int it, int jt) {
  return ((int64_t)b * p.h + hh) * p.nTri + (it * (it + 1) / 2 + jt);
}

template <bool HF, bool SPILL>
__global__ void __launch_bounds__(K4_THREADS, 1)
rga_bwd4_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmDO, const Bwd4Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Lay4::BAR);
  uint64_t* v_full = bars + B4_VF;        // loader -> dP issuer: V of head `item`
  uint64_t* v_empty = bars + B4_VE;       // dP issuer -> loader: the head's last dP product is done
  uint64_t* q_full = bars + B4_QF;        // [2]
  uint64_t* q_empty = bars + B4_QE;       // [2] dV/dK issuer -> loader
  uint64_t* do_full = bars + B4_DOF;      // [3]
  uint64_t* do_empty = bars + B4_DOE;     // [3]
  uint64_t* dp_full = bars + B4_DPF;      // [2] dP issuer -> math
  uint64_t* dp_empty = bars + B4_DPE;     // [2] math -> dP issuer (one arrival per warp)
  uint64_t* pp_empty = bars + B4_PPE;     // [2] dV product of a step done: its P slot may be rewritten
  uint64_t* ds_empty = bars + B4_DSE;     // [2] dK product (and the spill's read) of a step done
  uint64_t* ready = bars + B4_RDY;        // [2] math -> dV/dK issuer: P and dS of the step are in shared memory
  uint64_t* acc_done = bars + B4_ACC;     // [2] dV/dK issuer -> math: the head's accumulators are final
  uint64_t* acc_free = bars + B4_ACF;     // [2] math -> dV/dK issuer: the set has been read out (one group: 8 warps).  Without this
                                          // back-pressure the set's reuse two heads later was only ordered by timing: with one
                                          // step per head (the last key tile) a drain delayed by a step -- its code is cold, an
                                          // instruction-cache miss is enough -- let the issuer overwrite the set and complete
                                          // acc_done a second time, after which the drain's parity wait could never succeed
                                          // (one hang in ~10^4 launches)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B4_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.z, b = blockIdx.y, hh0 = (int)blockIdx.x * p.heads_per_cta;
  const int per = p.nT - jt;                                   // steps (query tiles) per head
  const int n_items = min(p.heads_per_cta, p.h - hh0);
  const int nsteps = n_items * per;

  if (warp == W4_LOAD && lane == 0) {
    tc::tma_prefetch_desc(&tmQ); tc::tma_prefetch_desc(&tmV); tc::tma_prefetch_desc(&tmDO);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&pp_empty[s], 1);
      tc::mbar_init(&q_full[s], 1); tc::mbar_init(&q_empty[s], 1);
      tc::mbar_init(&dp_full[s], 1); tc::mbar_init(&dp_empty[s], K4_MATH / 64);      // (one math group: 8 warps)
      tc::mbar_init(&ds_empty[s], 1); tc::mbar_init(&ready[s], K4_MATH / 64);
      tc::mbar_init(&acc_done[s], 1); tc::mbar_init(&acc_free[s], K4_MATH / 64);
    }
    for (int s = 0; s < 3; ++s) { tc::mbar_init(&do_full[s], 1); tc::mbar_init(&do_empty[s], 1); }
    tc::mbar_init(v_full, 1); tc::mbar_init(v_empty, 1);
    tc::fence_barrier_init();
  }
  if (warp == W4_MMA_A) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint64_t TS16 = TILE >> 4;

  if (warp == W4_LOAD) {
    // ================================ loader ===============================================
    if (lane == 0) {
      int k = 0, item = 0;
      for (int n = 0; n < nsteps; ++n) {
        const int hh = hh0 + item, it = jt + k;
        if (k == 0) {             // the head's V tile: the dP products of the previous head must be done with the old one
          if (item >= 1) tc::mbar_wait(v_empty, (item - 1) & 1);
          tc::mbar_arrive_expect_tx(v_full, TILE);
          tc::tma_load_4d(smem + Lay4::V0, &tmV, v_full, 0, hh, jt * TT, b);
        }
        const int d3 = n % 3;
        TRACE4(3, n, 0);
        tc::mbar_wait(&do_empty[d3], ((n / 3) & 1) ^ 1);
        TRACE4(3, n, 1);
        tc::mbar_arrive_expect_tx(&do_full[d3], TILE);
        tc::tma_load_4d(smem + Lay4::DO0 + d3 * TILE, &tmDO, &do_full[d3], 0, hh, it * TT, b);
        {                         // the P tile and the row statistics three steps ahead -> L2 (the math warps fetch them one
                                  // step ahead: a step is shorter than a DRAM round trip)
          int k3 = k + 3, i3 = item;
          while (k3 >= per) { k3 -= per; ++i3; }
          if (i3 < n_items) {
            const int64_t tix = tile_index(p, b, hh0 + i3, jt + k3, jt);
            tc::bulk_prefetch_l2(p.stash + tix * (int64_t)PT_BYTES, PT_BYTES);
            tc::bulk_prefetch_l2(p.mrow + tix * 2 * TT, 2 * TT * 4);
            if ((p.L & 3) == 0) {
              const int i3row = (jt + k3) * TT;
              const int64_t ro = ((int64_t)b * p.h + hh0 + i3) * p.L + i3row;
              const uint32_t nb = (uint32_t)min(TT, p.L - i3row) * 4u;
              tc::bulk_prefetch_l2(p.lse + ro, nb);
              tc::bulk_prefetch_l2(p.delta + ro, nb);
            }
          }
        }
        TRACE4(3, n, 2);
        tc::mbar_wait(&q_empty[n & 1], ((n >> 1) & 1) ^ 1);
        TRACE4(3, n, 3);
        tc::mbar_arrive_expect_tx(&q_full[n & 1], TILE);
        tc::tma_load_4d(smem + Lay4::Q0 + (n & 1) * TILE, &tmQ, &q_full[n & 1], 0, hh, it * TT, b);
        if (++k == per) { k = 0; ++item; }
      }
    }
  } else if (warp == W4_MMA_A) {
    // ================================ dP issuer =============================================
    if (lane == 0) {
      const uint32_t id_kk = tc::make_idesc(TT, TT, p.qk_fmt, p.qk_fmt, 0, 0);       // dO (K-major) x V (K-major), N = 128
      const uint64_t vd0 = tc::make_sdesc(tc::smem_u32(smem + Lay4::V0), 16, 1024);
      const uint64_t dod0 = tc::make_sdesc(tc::smem_u32(smem + Lay4::DO0), 16, 1024);
      int k = 0, item = 0;
      for (int n = 0; n < nsteps; ++n) {
        TRACE4(1, n, 0);
        if (k == 0) tc::mbar_wait(v_full, item & 1);
        tc::mbar_wait(&do_full[n % 3], (n / 3) & 1);
        TRACE4(1, n, 1);
        tc::mbar_wait(&dp_empty[n & 1], ((n >> 1) & 1) ^ 1);          // the math warps have read dP of step n - 2
        tc::tc_fence_after();
        TRACE4(1, n, 2);
        const uint64_t dod = dod0 + (uint64_t)(n % 3) * TS16, vd = vd0;
#pragma unroll
        for (int k4 = 0; k4 < DHC / 16; ++k4)
          tc::umma_f16(tmem + TM4_DP + 128 * (uint32_t)(n & 1), dod + 2 * k4, vd + 2 * k4, id_kk, k4 != 0);
        tc::umma_commit(&dp_full[n & 1]);
        TRACE4(1, n, 3);
        if (++k == per) { tc::umma_commit(v_empty); k = 0; ++item; }
      }
    }
  } else if (warp == W4_MMA_B) {
    // ================================ dV / dK issuer ========================================
    if (lane == 0) {
      const uint32_t id_mnmn = tc::make_idesc(TT, DHC, p.qk_fmt, p.qk_fmt, 1, 1);    // A MN-major (P / dS image), B MN-major (dO / Q), N = 64
      const uint64_t qd_mn0 = tc::make_sdesc(tc::smem_u32(smem + Lay4::Q0), 1024, 1024);
      const uint64_t dod_mn0 = tc::make_sdesc(tc::smem_u32(smem + Lay4::DO0), 1024, 1024);
      const uint64_t ppd0 = tc::make_sdesc(tc::smem_u32(smem + Lay4::PP), TILE, 1024);
      const uint64_t dsd0 = tc::make_sdesc(tc::smem_u32(smem + Lay4::DS0), TILE, 1024);
      int k = 0, item = 0;
      for (int n = 0; n < nsteps; ++n) {
        const uint32_t par = (n >> 1) & 1;
        TRACE4(2, n, 0);
        // Q first (it landed long ago: the wait is off the critical path), the math group's hand-off last.  dO of the
        // step needs no wait here: the dP issuer observed do_full, the math group waited for dP, this thread for the group
        tc::mbar_wait(&q_full[n & 1], par);
        TRACE4(2, n, 1);
        tc::mbar_wait(&ready[n & 1], par);
        tc::tc_fence_after();
        TRACE4(2, n, 2);
        if (SPILL) {          // the dS operand image (32 KB, swizzled) goes to the workspace as it is
          tc::bulk_store_1d(p.ds_ws + tile_index(p, b, hh0 + item, jt + k, jt) * (int64_t)PT_BYTES,
                            smem + Lay4::DS0 + (n & 1) * PT_BYTES, PT_BYTES);
          tc::bulk_commit();
        }
        if (k == 0 && item >= 2) {      // the set held head item - 2: its drain must be over
          tc::mbar_wait(&acc_free[item & 1], ((item >> 1) - 1) & 1);
          tc::tc_fence_after();
        }
        const uint32_t acc = tmem + TM4_ACC + 128 * (uint32_t)(item & 1);
        const uint64_t dod_mn = dod_mn0 + (uint64_t)(n % 3) * TS16, qd_mn = qd_mn0 + (uint64_t)(n & 1) * TS16;
        const uint64_t dsd = dsd0 + (uint64_t)(n & 1) * 2 * TS16, ppd = ppd0 + (uint64_t)(n & 1) * 2 * TS16;
#pragma unroll
        for (int k16 = 0; k16 < TT / 16; ++k16)        // dV += P^T dO (contraction over the 128 query rows)
          tc::umma_f16(acc + 64, ppd + 128 * k16, dod_mn + 128 * k16, id_mnmn, (k | k16) != 0);
        tc::umma_commit(&pp_empty[n & 1]);
        tc::umma_commit(&do_empty[n % 3]);
#pragma unroll
        for (int k16 = 0; k16 < TT / 16; ++k16)        // dK += dS^T Q
          tc::umma_f16(acc, dsd + 128 * k16, qd_mn + 128 * k16, id_mnmn, (k | k16) != 0);
        if (SPILL) tc::bulk_wait_read0();              // (the math warps rewrite the slot once ds_empty is signalled)
        tc::umma_commit(&q_empty[n & 1]);
        tc::umma_commit(&ds_empty[n & 1]);
        TRACE4(2, n, 3);
        if (++k == per) { tc::umma_commit(&acc_done[item & 1]); k = 0; ++item; }
      }
      if (SPILL) tc::bulk_wait0();
    }
  } else {
    // ================================ math warps ============================================
    // Two groups of 8 warps take ALTERNATE steps (group g = warp >> 3 the steps n = g, g + 2, ...): with all 16 warps on
    // one step every phase of the step (TMEM load latency, the unpack / FMA burst, the shared-memory stores, the proxy
    // fence) was exposed in turn -- the traced step was 3.9 k cycles at an issue rate of 0.36 -- whereas now each
    // scheduler holds two warps of either group in different phases, and a group has two step times for its step.  The
    // dP buffer, the P / dS slots and the barriers of parity n & 1 belong to one group.  A thread owns row a and the 64 key
    // columns of sub-tile hq.
    const int w4 = warp & 3, hq = (warp >> 2) & 1, grp = warp >> 3;
    const int a = w4 * 32 + lane, a7 = a & 7;
    const uint32_t lane_base = (uint32_t)(w4 * 32) << 16;
    // stashed tile: [forward warp 4 qt + w4][chunk 0..3][lane][16 B] (rga_tc.cu), qt = 2 hq, 2 hq + 1: eight coalesced loads
    const int64_t my_off = ((int64_t)(8 * hq + w4) * 128 + lane) * 16;
    uint8_t* const pp_row = smem + Lay4::PP + grp * PT_BYTES + hq * TILE + a * 128;
    uint8_t* const ds_row = smem + Lay4::DS0 + grp * PT_BYTES + hq * TILE + a * 128;
    const uint32_t dp_addr = tmem + TM4_DP + 128 * (uint32_t)grp + 64 * (uint32_t)hq + lane_base;
    uint64_t* const my_dp_full = &dp_full[grp];
    uint64_t* const my_dp_empty = &dp_empty[grp];
    uint64_t* const my_pp_empty = &pp_empty[grp];
    uint64_t* const my_ds_empty = &ds_empty[grp];
    uint64_t* const my_ready = &ready[grp];

    // position of step n = grp + 2 m inside the CTA's walk
    int k = grp, item = 0;
    while (k >= per) { k -= per; ++item; }
    int fk = k, fitem = item;             // fetch cursor: the group's next step
    auto fetch = [&](uint32_t (&R)[32], float& lse_raw, float& d_raw, float& mref) {
      if (fitem < n_items) {
        const int hh = hh0 + fitem, it = jt + fk;
        const int64_t tix = tile_index(p, b, hh, it, jt);
        const uint8_t* src = p.stash + tix * (int64_t)PT_BYTES + my_off;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          tc::ldg128_stream(src + 8192 * (c >> 2) + 512 * (c & 3), R[4 * c], R[4 * c + 1], R[4 * c + 2], R[4 * c + 3]);
        const int i = it * TT + a;
        const int64_t ro = ((int64_t)b * p.h + hh) * p.L + i;
        // (raw values only: any arithmetic here would wait for the loads inside the fetch)
        lse_raw = 0.f; d_raw = 0.f;
        if (i < p.L) { lse_raw = __ldg(p.lse + ro); d_raw = __ldg(p.delta + ro); }
        mref = __ldg(p.mrow + (tix * 2 + hq) * TT + a);       // the row reference of this thread's key half
        fk += 2;
        while (fk >= per) { fk -= per; ++fitem; }
      }
    };
    // dK (hq = 0) / dV (hq = 1) of head `it_` -> global: the thread's key row a, all 64 columns
    auto store_acc = [&](int it_) {
      tc::mbar_wait(&acc_done[it_ & 1], (it_ >> 1) & 1);
      tc::tc_fence_after();
      const float osc = HF ? p.inv_gscale : 1.f;
      const int row = jt * TT + a;
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(hq ? p.dv : p.dk) + (int64_t)b * p.sb +
                                            (int64_t)row * p.sl + (int64_t)(hh0 + it_) * p.sh);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem + TM4_ACC + 128 * (uint32_t)(it_ & 1) + 64 * (uint32_t)hq + 32 * hf + lane_base, r);
        tc::tmem_ld_wait();
        if (hf == 1) {                  // this warp's part of the set is in registers
          tc::tc_fence_before();
          tc::mbar_arrive_warp(&acc_free[it_ & 1]);
        }
        if (row < p.L) {
#pragma unroll
          for (int x = 0; x < 4; ++x)
            dst[4 * hf + x] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * x]) * osc, __uint_as_float(r[8 * x + 1]) * osc),
                                         pack_bf16x2(__uint_as_float(r[8 * x + 2]) * osc, __uint_as_float(r[8 * x + 3]) * osc),
                                         pack_bf16x2(__uint_as_float(r[8 * x + 4]) * osc, __uint_as_float(r[8 * x + 5]) * osc),
                                         pack_bf16x2(__uint_as_float(r[8 * x + 6]) * osc, __uint_as_float(r[8 * x + 7]) * osc));
        }
      }
      tc::tc_fence_before();
    };
    uint32_t R[32];
    float nl = 0.f, nd = 0.f, nm = 0.f;
    fetch(R, nl, nd, nm);
    for (int n = grp, m = 0; n < nsteps; n += 2, ++m) {
      const bool tr = (lane == 0 && w4 == 0 && hq == 0);
      if (tr) TRACE4(0, n, 0);
      // (f16 mode: dP arrives scaled by g, so D is scaled to match and dS = g * the true dS)
      const float lse2 = nl * LOG2E, Ds = nd * p.scale * (HF ? p.gscale : 1.f);
      const float f = tc::fast_exp2(nm - lse2);
      if (tr) TRACE4(0, n, 1);        // (the row statistics have arrived)
      tc::mbar_wait(my_dp_full, m & 1);
      if (m >= 1) {                   // the group's slots: dV / dK (and the spill) of step n - 2 have read them
        tc::mbar_wait(my_pp_empty, (m - 1) & 1);
        tc::mbar_wait(my_ds_empty, (m - 1) & 1);
      }
      tc::tc_fence_after();
      if (tr) TRACE4(0, n, 2);
#pragma unroll
      for (int q = 0; q < 4; ++q) {             // 16 key columns at a time: two 16-byte chunks of the row in either image
        uint32_t dp[16];
        tc::tmem_ld_32x16(dp_addr + 16 * q, dp);
        tc::tmem_ld_wait();
        if (q == 3) {
          tc::tc_fence_before();
          tc::mbar_arrive_warp(my_dp_empty);
        }
        uint32_t A[8], D8[8];
#pragma unroll
        for (int y = 0; y < 8; ++y) {
          const uint32_t rv = R[8 * q + y];
          float p0, p1;
          if (HF) { const float2 pf = __half22float2(*reinterpret_cast<const __half2*>(&rv)); p0 = pf.x; p1 = pf.y; }
          else { p0 = __uint_as_float(rv << 16); p1 = __uint_as_float(rv & 0xffff0000u); }
          p0 *= f; p1 *= f;
          const float d0 = fmaf(__uint_as_float(dp[2 * y]), p.scale, -Ds) * p0;
          const float d1 = fmaf(__uint_as_float(dp[2 * y + 1]), p.scale, -Ds) * p1;
          A[y] = HF ? pack_f16x2(p0, p1) : pack_bf16x2(p0, p1);
          D8[y] = HF ? pack_f16x2(d0, d1) : pack_bf16x2(d0, d1);
        }
        const int o0 = ((2 * q) ^ a7) << 4, o1 = ((2 * q + 1) ^ a7) << 4;
        *reinterpret_cast<uint4*>(pp_row + o0) = make_uint4(A[0], A[1], A[2], A[3]);
        *reinterpret_cast<uint4*>(pp_row + o1) = make_uint4(A[4], A[5], A[6], A[7]);
        *reinterpret_cast<uint4*>(ds_row + o0) = make_uint4(D8[0], D8[1], D8[2], D8[3]);
        *reinterpret_cast<uint4*>(ds_row + o1) = make_uint4(D8[4], D8[5], D8[6], D8[7]);
      }
      if (tr) TRACE4(0, n, 3);
      tc::fence_proxy_async();
      tc::mbar_arrive_warp(my_ready);
      if (tr) TRACE4(0, n, 4);
      fetch(R, nl, nd, nm);                   // the group's next tile and statistics (two step times ahead of their use)
      // the step after a head's last one: that head's accumulators (the other TMEM set) are final
      if (k == 0 && item > 0) store_acc(item - 1);
      k += 2;
      while (k >= per) { k -= per; ++item; }
      if (tr) TRACE4(0, n, 5);
    }
    if ((nsteps & 1) == grp) store_acc(n_items - 1);      // (the group that would take step `nsteps`)
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == W4_MMA_A) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

template <bool HF, bool SPILL>
int launch4(const CUtensorMap& tmQ, const CUtensorMap& tmV, const CUtensorMap& tmDO, const Bwd4Params& p, dim3 grid,
            cudaStream_t st) {
  auto kern = rga_bwd4_dkv_kernel<HF, SPILL>;
  static unsigned long long attr_done = 0; const unsigned long long attr_bit = attr_dev_bit();
  if (!(attr_done & attr_bit)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM4);
    if (e != cudaSuccess) { set_error("rga_bwd4: smem attribute (%d B): %s", SMEM4, cudaGetErrorString(e)); return (int)e; }
    attr_done |= attr_bit;
  }
  Bwd4Params q = p;
  static const bool want_trace = getenv("MT_RGA_TRACE") != nullptr;
  static long long* trace_dev = nullptr;
  const size_t trace_n = 4 * 32 * 8;
  if (want_trace) {
    if (!trace_dev) cudaMalloc(&trace_dev, trace_n * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, trace_n * sizeof(long long), st);
    q.trace = trace_dev;
    q.trace_z = atoi(getenv("MT_RGA_TRACE"));
  }
  kern<<<grid, K4_THREADS, SMEM4, st>>>(tmQ, tmV, tmDO, q);
  if (want_trace) {
    static long long host[4 * 32 * 8];
    cudaMemcpyAsync(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    long long t0 = 0;
    for (size_t x = 0; x < trace_n; ++x) if (host[x] && (!t0 || host[x] < t0)) t0 = host[x];
    static const char* agent[4] = {"MATH", "DP", "DVK", "LD"};
    for (int ag = 0; ag < 4; ++ag)
      for (int n = 0; n < 32; ++n) {
        bool any = false;
        for (int e = 0; e < 8; ++e) any |= host[(ag * 32 + n) * 8 + e] != 0;
        if (!any) continue;
        fprintf(stderr, "trace4 %-4s step %2d:", agent[ag], n);
        for (int e = 0; e < 8; ++e) fprintf(stderr, " %8lld", host[(ag * 32 + n) * 8 + e] ? host[(ag * 32 + n) * 8 + e] - t0 : -1LL);
        fprintf(stderr, "\n");
      }
  }
  return check_launch("rga_bwd4_dkv");
}

}  // namespace

// dK, dV from the forward's P stash (key-tile owner walks the query tiles at or below it); ds_ws != NULL: also
// spill the dS tiles for the consumers of rga_tc_bwd3.cu
int rga_bwd4_dkv(const RgaArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmV, const CUtensorMap& tmDO,
                 void* ds_ws, int qk_fmt, float gscale, cudaStream_t st) {
  Bwd4Params p;
  p.dk = a.dk; p.dv = a.dv; p.sb = a.sb; p.sl = a.sl; p.sh = a.sh;
  p.lse = a.lse; p.delta = a.delta;
  p.B = a.B; p.h = a.h; p.L = a.L;
  p.nT = (a.L + TT - 1) / TT;
  p.nTri = p.nT * (p.nT + 1) / 2;
  p.stash = static_cast<const uint8_t*>(a.pstash);
  p.mrow = reinterpret_cast<const float*>(p.stash + (int64_t)a.B * a.h * p.nTri * (int64_t)PT_BYTES);      // then [tile][key half][128 rows] fp32
  p.ds_ws = static_cast<uint8_t*>(ds_ws);
  p.qk_fmt = qk_fmt;
  p.gscale = gscale;
  p.inv_gscale = 1.f / gscale;
  p.scale = 1.f / a.inv_scale_div;
  p.trace = nullptr; p.trace_z = 0;
  // consecutive heads of one (batch row, key tile) share a CTA: as many as leave at least three CTAs per SM
  static const int hpc_env = getenv("MT_DKV_HPC") ? atoi(getenv("MT_DKV_HPC")) : 0;
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
  if (qk_fmt == 0) return ds_ws ? launch4<true, true>(tmQ, tmV, tmDO, p, grid, st) : launch4<true, false>(tmQ, tmV, tmDO, p, grid, st);
  return ds_ws ? launch4<false, true>(tmQ, tmV, tmDO, p, grid, st) : launch4<false, false>(tmQ, tmV, tmDO, p, grid, st);
}

}  // namespace mt
