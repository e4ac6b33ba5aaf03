// Backward of the fused relative global attention on tcgen05 / TMEM / TMA (K2), head dim 64, bf16,
// causal (+ key padding).  Math (per (b,h)):
//     P = exp(S - lse),  S = (Q K^T + skew(Q E_band^T)) / sqrt(dh)
//     dP = dO V^T ;  dS = P o (dP - D) / sqrt(dh),  D = rowsum(dO o O)
//     dV = P^T dO ;  dK = dS^T Q ;  dQ = dS K + dG E_band ;  dE_band = dG^T Q,
// where dG is dS written back into band coordinates (dG[a][127-a+b] = dS[a][b]) -- the inverse of
// the forward skew.  One kernel template, three roles, all sharing the "recompute the P / dS tile"
// core (TMA loads -> S, G_lo, G_hi, dP on the tensor cores -> skew through a private shared
// scratch -> P, dS, dG written as 128B-swizzled UMMA operands):
//   * DKV : CTA owns a key tile, walks the query tiles at or below it; dK, dV accumulate in TMEM;
//   * DQ  : CTA owns a query tile, walks its key tiles; dQ (both the K and the relative-embedding
//           part) accumulates in TMEM;
//   * DE  : CTA owns one tile-diagonal (i0 - j0 fixed => the same two 128-row blocks of E for every
//           step) and walks down it over a slice of (batch, head); dE accumulates in TMEM over the
//           whole walk and is added to global memory once per CTA.
// No output needs per-step global atomics; the price is that S/P are recomputed per role.
// The same shared-memory tile serves as K-major and as MN-major UMMA operand (rows of 128 bytes,
// 8-row swizzle atoms), so no transposed copies exist anywhere.
// Warps 0-7: P / dS math (row a = 32*(w&3)+lane, key columns 64*(w>>2)..+63), 8: TMA, 9: MMA.
#include "ops.cuh"
#include "rga_tc_common.cuh"

#include <stdlib.h>

namespace mt {

using namespace rga;

namespace {

enum { MODE_DKV = 0, MODE_DQ = 1, MODE_DE = 2 };

// TMEM columns
constexpr uint32_t TM_S = 0, TM_GLO = 128, TM_GHI = 256, TM_ACC0 = 384, TM_ACC1 = 448;
constexpr uint32_t TM_DP = TM_GLO;      // dP reuses the G_lo columns once the skew has been read

template <int MODE> struct Lay;
template <> struct Lay<MODE_DKV> {       // K,V resident; {Q,dO,E_lo,E_hi} double buffered
  static constexpr int NST = 2;
  static constexpr int K = 0, V = TILE, STAGE0 = 2 * TILE, STAGE_BYTES = 4 * TILE;
  static constexpr int sQ = 0, sDO = TILE, sELO = 2 * TILE, sEHI = 3 * TILE;
  static constexpr int P = STAGE0 + 2 * STAGE_BYTES, DS = P + 2 * TILE, BAR = DS + 2 * TILE;
  static constexpr int RES_TILES = 2, STAGE_TILES = 4;
};
template <> struct Lay<MODE_DQ> {        // Q,dO resident; {K,V,E_lo,E_hi} per step
  static constexpr int NST = 1;
  static constexpr int Q = 0, DO = TILE, STAGE0 = 2 * TILE, STAGE_BYTES = 4 * TILE;
  static constexpr int sK = 0, sV = TILE, sELO = 2 * TILE, sEHI = 3 * TILE;
  static constexpr int DS = STAGE0 + STAGE_BYTES, DG = DS + 2 * TILE, SCR = DG + 4 * TILE, BAR = SCR + SCR_BYTES;
  static constexpr int RES_TILES = 2, STAGE_TILES = 4;
};
template <> struct Lay<MODE_DE> {        // E_lo,E_hi resident; Q double buffered; {dO,K,V} per step, released early
  static constexpr int NST = 1;
  static constexpr int ELO = 0, EHI = TILE, Q0 = 2 * TILE, STAGE0 = 4 * TILE, STAGE_BYTES = 3 * TILE;
  static constexpr int sDO = 0, sK = TILE, sV = 2 * TILE;
  static constexpr int DG = STAGE0 + STAGE_BYTES, SCR = DG + 4 * TILE, BAR = SCR + SCR_BYTES;
  static constexpr int RES_TILES = 2, STAGE_TILES = 3;
};
template <int MODE> constexpr int smem_bytes() { return Lay<MODE>::BAR + 256 + 1024; }
static_assert(SCR_BYTES <= 2 * TILE, "DKV role parks the skew scratch in the stage's E_lo/E_hi buffers");

struct BwdParams {
  void* dq; void* dk; void* dv;          // 16-bit, q/k/v addressing
  int64_t sb, sl, sh;
  float* dE;
  const float* lse; const float* delta;
  const uint8_t* pad;
  int B, h, L, max_seq, nT;
  int bh_per_cta;                        // DE role
  float scale, scale_log2;
};

struct StepInfo { int it, jt, b, hh; };

template <int MODE>
__device__ __forceinline__ int num_steps(const BwdParams& p, int& bh0) {
  // grid = (h, B, nT) [DKV, DQ] or (slices, 1, nT) [DE]: the tile / diagonal index is the SLOWEST
  // grid dimension, so CTAs are dispatched longest-first over the whole launch
  bh0 = 0;
  if (MODE == MODE_DKV) return p.nT - (int)blockIdx.z;
  if (MODE == MODE_DQ) return p.nT - (int)blockIdx.z;          // it = nT-1-blockIdx.z  -> it+1 steps
  bh0 = (int)blockIdx.x * p.bh_per_cta;
  int nbh = min(p.bh_per_cta, p.B * p.h - bh0);
  return nbh > 0 ? nbh * (p.nT - (int)blockIdx.z) : 0;
}
template <int MODE>
__device__ __forceinline__ StepInfo step_info(const BwdParams& p, int n, int bh0) {
  StepInfo s;
  if (MODE == MODE_DKV) { s.jt = blockIdx.z; s.it = s.jt + n; s.hh = blockIdx.x; s.b = blockIdx.y; }
  else if (MODE == MODE_DQ) { s.it = p.nT - 1 - (int)blockIdx.z; s.jt = n; s.hh = blockIdx.x; s.b = blockIdx.y; }
  else {
    const int per = p.nT - (int)blockIdx.z;
    const int bh = bh0 + n / per, k = n % per;
    s.it = (int)blockIdx.z + k; s.jt = k; s.b = bh / p.h; s.hh = bh % p.h;
  }
  return s;
}

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
rga_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                  const __grid_constant__ CUtensorMap tmE, const BwdParams p) {
  using LY = Lay<MODE>;
  extern __shared__ __align__(1024) uint8_t smem[];      // shared address space kept: LDS/STS, not generic
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LY::BAR);
  uint64_t* bar_res = bars + 0;
  uint64_t* ld_full = bars + 1;       // [2]
  uint64_t* ld_empty = bars + 3;      // [2]
  uint64_t* sg_full = bars + 5;
  uint64_t* sg_consumed = bars + 6;
  uint64_t* dp_full = bars + 7;
  uint64_t* ds_ready = bars + 8;
  uint64_t* step_done = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  uint64_t* q_full = bars + 11;       // [2]  (DE role: double-buffered Q)
  uint64_t* q_empty = bars + 13;      // [2]
  uint8_t* spad = reinterpret_cast<uint8_t*>(bars + 16);      // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bh0;
  const int nsteps = num_steps<MODE>(p, bh0);

  if (warp == 8 && lane == 0) {
    tc::tma_prefetch_desc(&tmQ); tc::tma_prefetch_desc(&tmK); tc::tma_prefetch_desc(&tmV);
    tc::tma_prefetch_desc(&tmDO); tc::tma_prefetch_desc(&tmE);
    tc::mbar_init(bar_res, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&ld_full[s], 1); tc::mbar_init(&ld_empty[s], 1);
      tc::mbar_init(&q_full[s], 1); tc::mbar_init(&q_empty[s], 1);
    }
    tc::mbar_init(sg_full, 1);
    tc::mbar_init(sg_consumed, SM_THREADS);
    tc::mbar_init(dp_full, 1);
    tc::mbar_init(ds_ready, SM_THREADS);
    tc::mbar_init(step_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 9) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (nsteps <= 0) {            // (DE role: empty slice) -- uniform for the whole CTA
    __syncthreads();
    if (warp == 9) tc::tmem_dealloc(tmem, 512);
    return;
  }

  // per-mode buffer lookup
  auto buf_q = [&](int st) -> uint8_t* {
    if (MODE == MODE_DKV) return smem + Lay<MODE_DKV>::STAGE0 + st * Lay<MODE_DKV>::STAGE_BYTES + Lay<MODE_DKV>::sQ;
    if (MODE == MODE_DQ) return smem + Lay<MODE_DQ>::Q;
    return smem + Lay<MODE_DE>::Q0 + st * TILE;          // DE: st = step parity (Q is double buffered)
  };
  auto buf_do = [&](int st) -> uint8_t* {
    if (MODE == MODE_DKV) return smem + Lay<MODE_DKV>::STAGE0 + st * Lay<MODE_DKV>::STAGE_BYTES + Lay<MODE_DKV>::sDO;
    if (MODE == MODE_DQ) return smem + Lay<MODE_DQ>::DO;
    return smem + Lay<MODE_DE>::STAGE0 + Lay<MODE_DE>::sDO;
  };
  auto buf_k = [&](int st) -> uint8_t* {
    if (MODE == MODE_DKV) return smem + Lay<MODE_DKV>::K;
    if (MODE == MODE_DQ) return smem + Lay<MODE_DQ>::STAGE0 + Lay<MODE_DQ>::sK;
    return smem + Lay<MODE_DE>::STAGE0 + Lay<MODE_DE>::sK;
  };
  auto buf_v = [&](int st) -> uint8_t* {
    if (MODE == MODE_DKV) return smem + Lay<MODE_DKV>::V;
    if (MODE == MODE_DQ) return smem + Lay<MODE_DQ>::STAGE0 + Lay<MODE_DQ>::sV;
    return smem + Lay<MODE_DE>::STAGE0 + Lay<MODE_DE>::sV;
  };
  auto buf_elo = [&](int st) -> uint8_t* {
    if (MODE == MODE_DKV) return smem + Lay<MODE_DKV>::STAGE0 + st * Lay<MODE_DKV>::STAGE_BYTES + Lay<MODE_DKV>::sELO;
    if (MODE == MODE_DQ) return smem + Lay<MODE_DQ>::STAGE0 + Lay<MODE_DQ>::sELO;
    return smem + Lay<MODE_DE>::ELO;
  };
  auto buf_ehi = [&](int st) -> uint8_t* {
    if (MODE == MODE_DKV) return smem + Lay<MODE_DKV>::STAGE0 + st * Lay<MODE_DKV>::STAGE_BYTES + Lay<MODE_DKV>::sEHI;
    if (MODE == MODE_DQ) return smem + Lay<MODE_DQ>::STAGE0 + Lay<MODE_DQ>::sEHI;
    return smem + Lay<MODE_DE>::EHI;
  };

  if (warp == 8) {
    // ================================ TMA producer ==========================================
    if (lane == 0) {
      const StepInfo s0 = step_info<MODE>(p, 0, bh0);
      tc::mbar_arrive_expect_tx(bar_res, LY::RES_TILES * TILE);
      if (MODE == MODE_DKV) {
        tc::tma_load_4d(buf_k(0), &tmK, bar_res, 0, s0.hh, s0.jt * TT, s0.b);
        tc::tma_load_4d(buf_v(0), &tmV, bar_res, 0, s0.hh, s0.jt * TT, s0.b);
      } else if (MODE == MODE_DQ) {
        tc::tma_load_4d(buf_q(0), &tmQ, bar_res, 0, s0.hh, s0.it * TT, s0.b);
        tc::tma_load_4d(buf_do(0), &tmDO, bar_res, 0, s0.hh, s0.it * TT, s0.b);
      } else {
        const int c0 = p.max_seq - 1 - (int)blockIdx.z * TT;
        tc::tma_load_2d(buf_elo(0), &tmE, bar_res, 0, c0 - (TT - 1));
        tc::tma_load_2d(buf_ehi(0), &tmE, bar_res, 0, c0 + 1);
      }
      for (int n = 0; n < nsteps; ++n) {
        const StepInfo s = step_info<MODE>(p, n, bh0);
        const int st = (LY::NST == 2) ? (n & 1) : 0;
        const uint32_t ph = (LY::NST == 2) ? ((n >> 1) & 1) : (n & 1);
        if (MODE == MODE_DE && n == 0) {
          // Q lives until the end of the step (dE = dG^T Q), so it gets its own double buffer and
          // is fetched one step ahead; {dO, K, V} are released as soon as S, G and dP are computed
          tc::mbar_arrive_expect_tx(&q_full[0], TILE);
          tc::tma_load_4d(buf_q(0), &tmQ, &q_full[0], 0, s.hh, s.it * TT, s.b);
        }
        tc::mbar_wait(&ld_empty[st], ph ^ 1);
        tc::mbar_arrive_expect_tx(&ld_full[st], LY::STAGE_TILES * TILE);
        const int c0 = p.max_seq - 1 - (s.it - s.jt) * TT;
        if (MODE == MODE_DKV) tc::tma_load_4d(buf_q(st), &tmQ, &ld_full[st], 0, s.hh, s.it * TT, s.b);
        if (MODE != MODE_DQ) tc::tma_load_4d(buf_do(st), &tmDO, &ld_full[st], 0, s.hh, s.it * TT, s.b);
        if (MODE != MODE_DKV) {
          tc::tma_load_4d(buf_k(st), &tmK, &ld_full[st], 0, s.hh, s.jt * TT, s.b);
          tc::tma_load_4d(buf_v(st), &tmV, &ld_full[st], 0, s.hh, s.jt * TT, s.b);
        }
        if (MODE != MODE_DE) {
          tc::tma_load_2d(buf_elo(st), &tmE, &ld_full[st], 0, c0 - (TT - 1));
          tc::tma_load_2d(buf_ehi(st), &tmE, &ld_full[st], 0, c0 + 1);
        }
        if (LY::NST == 1 && n + 1 < nsteps) {
          // single-buffered roles: pull the next step's tiles into L2 while this step computes
          const StepInfo t = step_info<MODE>(p, n + 1, bh0);
          if (MODE == MODE_DE) {
            const int nb = (n + 1) & 1;
            tc::mbar_wait(&q_empty[nb], (((n + 1) >> 1) & 1) ^ 1);
            tc::mbar_arrive_expect_tx(&q_full[nb], TILE);
            tc::tma_load_4d(buf_q(nb), &tmQ, &q_full[nb], 0, t.hh, t.it * TT, t.b);
            tc::tma_prefetch_4d(&tmDO, 0, t.hh, t.it * TT, t.b);
          }
          tc::tma_prefetch_4d(&tmK, 0, t.hh, t.jt * TT, t.b);
          tc::tma_prefetch_4d(&tmV, 0, t.hh, t.jt * TT, t.b);
        }
      }
    }
  } else if (warp == 9) {
    // ================================ MMA issuer ============================================
    if (lane == 0) {
      const uint32_t id_kk = tc::make_idesc(TT, TT, 1, 1, 0, 0);      // S, G, dP : K-major x K-major, N = 128
      const uint32_t id_kmn = tc::make_idesc(TT, DHC, 1, 1, 0, 1);    // dQ      : A K-major, B MN-major, N = 64
      const uint32_t id_mnmn = tc::make_idesc(TT, DHC, 1, 1, 1, 1);   // dK/dV/dE: A MN-major, B MN-major, N = 64
      tc::mbar_wait(bar_res, 0);
      for (int n = 0; n < nsteps; ++n) {
        const int st = (LY::NST == 2) ? (n & 1) : 0;
        const uint32_t ph = (LY::NST == 2) ? ((n >> 1) & 1) : (n & 1);
        const uint32_t par = n & 1;
        tc::mbar_wait(&ld_full[st], ph);
        if (MODE == MODE_DE) tc::mbar_wait(&q_full[n & 1], (n >> 1) & 1);
        tc::tc_fence_after();
        // descriptors: built per step from the buffer addresses, k-steps are adds on the address field
        const uint64_t qk = tc::make_sdesc(tc::smem_u32(buf_q(MODE == MODE_DE ? (n & 1) : st)), 16, 1024);
        const uint64_t qmn = tc::make_sdesc(tc::smem_u32(buf_q(MODE == MODE_DE ? (n & 1) : st)), 1024, 1024);
        const uint64_t dok = tc::make_sdesc(tc::smem_u32(buf_do(st)), 16, 1024);
        const uint64_t domn = tc::make_sdesc(tc::smem_u32(buf_do(st)), 1024, 1024);
        const uint64_t kk = tc::make_sdesc(tc::smem_u32(buf_k(st)), 16, 1024);
        const uint64_t kmn = tc::make_sdesc(tc::smem_u32(buf_k(st)), 1024, 1024);
        const uint64_t vk = tc::make_sdesc(tc::smem_u32(buf_v(st)), 16, 1024);
        const uint64_t elok = tc::make_sdesc(tc::smem_u32(buf_elo(st)), 16, 1024);
        const uint64_t ehik = tc::make_sdesc(tc::smem_u32(buf_ehi(st)), 16, 1024);
        const uint64_t elomn = tc::make_sdesc(tc::smem_u32(buf_elo(st)), 1024, 1024);
        const uint64_t ehimn = tc::make_sdesc(tc::smem_u32(buf_ehi(st)), 1024, 1024);
        // ---- phase A: S = Q K^T, G_lo = Q E_lo^T, G_hi = Q E_hi^T
#pragma unroll
        for (int k4 = 0; k4 < DHC / 16; ++k4) {
          tc::umma_f16(tmem + TM_S, qk + 2 * k4, kk + 2 * k4, id_kk, k4 != 0);
          tc::umma_f16(tmem + TM_GLO, qk + 2 * k4, elok + 2 * k4, id_kk, k4 != 0);
          tc::umma_f16(tmem + TM_GHI, qk + 2 * k4, ehik + 2 * k4, id_kk, k4 != 0);
        }
        tc::umma_commit(sg_full);
        // ---- phase C: dP = dO V^T into the G_lo columns (after the math warps read S / G)
        tc::mbar_wait(sg_consumed, par);
        tc::tc_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < DHC / 16; ++k4)
          tc::umma_f16(tmem + TM_DP, dok + 2 * k4, vk + 2 * k4, id_kk, k4 != 0);
        tc::umma_commit(dp_full);
        if (MODE == MODE_DE) tc::umma_commit(&ld_empty[st]);     // dO, K, V are dead from here on
        // ---- phase E: role MMAs on the P / dS / dG operands written by the math warps
        tc::mbar_wait(ds_ready, par);
        tc::tc_fence_after();
        if (MODE == MODE_DKV) {
          const uint64_t pd = tc::make_sdesc(tc::smem_u32(smem + Lay<MODE_DKV>::P), TILE, 1024);
          const uint64_t dsd = tc::make_sdesc(tc::smem_u32(smem + Lay<MODE_DKV>::DS), TILE, 1024);
#pragma unroll
          for (int k16 = 0; k16 < TT / 16; ++k16) {      // contraction over the 128 query rows
            tc::umma_f16(tmem + TM_ACC1, pd + 128 * k16, domn + 128 * k16, id_mnmn, (n | k16) != 0);
            tc::umma_f16(tmem + TM_ACC0, dsd + 128 * k16, qmn + 128 * k16, id_mnmn, (n | k16) != 0);
          }
        } else if (MODE == MODE_DQ) {
          const uint64_t dsd = tc::make_sdesc(tc::smem_u32(smem + Lay<MODE_DQ>::DS), 16, 1024);
          const uint64_t dgd = tc::make_sdesc(tc::smem_u32(smem + Lay<MODE_DQ>::DG), 16, 1024);
#pragma unroll
          for (int k16 = 0; k16 < TT / 16; ++k16)         // dS . K_j (contraction over the 128 keys)
            tc::umma_f16(tmem + TM_ACC0, dsd + (k16 >> 2) * (TILE >> 4) + 2 * (k16 & 3), kmn + 128 * k16, id_kmn, (n | k16) != 0);
#pragma unroll
          for (int k16 = 0; k16 < 2 * TT / 16; ++k16)     // dG . [E_lo; E_hi] (contraction over the band)
            tc::umma_f16(tmem + TM_ACC0, dgd + (k16 >> 2) * (TILE >> 4) + 2 * (k16 & 3),
                         (k16 < 8 ? elomn + 128 * k16 : ehimn + 128 * (k16 - 8)), id_kmn, 1);
        } else {
          const uint64_t dgd = tc::make_sdesc(tc::smem_u32(smem + Lay<MODE_DE>::DG), TILE, 1024);
#pragma unroll
          for (int k16 = 0; k16 < TT / 16; ++k16) {       // dG_blk^T . Q (contraction over the query rows)
            tc::umma_f16(tmem + TM_ACC0, dgd + 128 * k16, qmn + 128 * k16, id_mnmn, (n | k16) != 0);
            tc::umma_f16(tmem + TM_ACC1, dgd + 2 * (TILE >> 4) + 128 * k16, qmn + 128 * k16, id_mnmn, (n | k16) != 0);
          }
        }
        if (MODE == MODE_DE) tc::umma_commit(&q_empty[n & 1]);
        else tc::umma_commit(&ld_empty[st]);
        tc::umma_commit(step_done);
      }
    }
  } else {
    // ================================ P / dS math warps ======================================
    const int w4 = warp & 3, wg = warp >> 2;
    const int a = w4 * 32 + lane;
    const uint32_t lane_base = (uint32_t)(w4 * 32) << 16;
    uint32_t* scr = nullptr;
    if (MODE == MODE_DQ) scr = reinterpret_cast<uint32_t*>(smem + Lay<MODE_DQ>::SCR) + threadIdx.x * SCR_WORDS;
    if (MODE == MODE_DE) scr = reinterpret_cast<uint32_t*>(smem + Lay<MODE_DE>::SCR) + threadIdx.x * SCR_WORDS;
    uint8_t* dg_base = nullptr;
    if (MODE == MODE_DQ) dg_base = smem + Lay<MODE_DQ>::DG;
    if (MODE == MODE_DE) dg_base = smem + Lay<MODE_DE>::DG;
    if (MODE != MODE_DKV) {
      // dG is zero outside the 128 band columns each row owns; those positions never change
      uint4* z = reinterpret_cast<uint4*>(dg_base);
      for (int x = threadIdx.x; x < 4 * TILE / 16; x += SM_THREADS) z[x] = make_uint4(0, 0, 0, 0);
      tc::fence_proxy_async();
      tc::named_bar_sync(1, SM_THREADS);
    }
    // training batches carry no pad tokens: decide once per CTA whether any key of the sequences
    // this CTA touches is padded; if none is, the unmasked fast path is used for every step
    const uint8_t* pad = p.pad;
    if (pad) {
      const StepInfo sf = step_info<MODE>(p, 0, bh0), sl = step_info<MODE>(p, nsteps - 1, bh0);
      bool mine = false;
      for (int64_t x = (int64_t)sf.b * p.L + threadIdx.x; x < (int64_t)(sl.b + 1) * p.L; x += SM_THREADS)
        mine |= (pad[x] != 0);
      if (!tc::named_bar_red_or(1, SM_THREADS, mine)) pad = nullptr;
    }
    const int base_w = ((127 - a) >> 1) + 32 * wg;   // first 32-bit word of this thread's band run in dG

    for (int n = 0; n < nsteps; ++n) {
      const StepInfo s = step_info<MODE>(p, n, bh0);
      const uint32_t par = n & 1;
      const int st = (LY::NST == 2) ? (n & 1) : 0;
      const int i0 = s.it * TT, j0 = s.jt * TT;
      const int i = i0 + a;
      const bool row_ok = i < p.L;
      const int64_t rowidx = ((int64_t)s.b * p.h + s.hh) * p.L + i;
      const float lse2 = row_ok ? p.lse[rowidx] * LOG2E : 0.f;
      const float Dv = row_ok ? p.delta[rowidx] : 0.f;
      if (pad) {
        tc::named_bar_sync(1, SM_THREADS);
        if (wg == 0) spad[a] = (j0 + a < p.L) ? pad[(int64_t)s.b * p.L + j0 + a] : 1;
        tc::named_bar_sync(1, SM_THREADS);
      }
      if (MODE == MODE_DKV)      // scratch = the E_lo/E_hi buffers of this stage (dead once G is computed)
        scr = reinterpret_cast<uint32_t*>(buf_elo(st)) + threadIdx.x * SCR_WORDS;

      tc::mbar_wait(sg_full, par);
      tc::tc_fence_after();
      float pv[64];
      {
        uint32_t r0[32], r1[32];
        tc::tmem_ld_32x32(tmem + TM_S + lane_base + wg * 64, r0);
        tc::tmem_ld_32x32(tmem + TM_S + lane_base + wg * 64 + 32, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; ++x) { pv[x] = __uint_as_float(r0[x]); pv[32 + x] = __uint_as_float(r1[x]); }
      }
      skew_add_64(pv, tmem + TM_GLO, tmem + TM_GHI, lane_base, w4, wg, lane, scr);
      tc::tc_fence_before();
      tc::mbar_arrive(sg_consumed);

      // P = exp(S - lse) with the reference's mask (causal on the diagonal tile, key padding, tails)
      const bool diag = (i0 == j0);
      const bool tail = (j0 + TT > p.L) || (i0 + TT > p.L);
#pragma unroll
      for (int x = 0; x < 64; ++x) pv[x] = tc::fast_exp2(fmaf(pv[x], p.scale_log2, -lse2));
      if (diag || tail || pad != nullptr) {
        // branch-free: columns x > lim are masked (causal limit on the diagonal tile, ragged tail,
        // rows beyond L), then the key-padding bytes, four per shared-memory word
        int lim = 63;
        if (diag) lim = min(lim, a - wg * 64);
        lim = min(lim, p.L - 1 - j0 - wg * 64);
        if (!row_ok) lim = -1;
#pragma unroll
        for (int x = 0; x < 64; ++x) pv[x] = (x > lim) ? 0.f : pv[x];
        if (pad) {
          const uint32_t* sp = reinterpret_cast<const uint32_t*>(spad + wg * 64);
#pragma unroll
          for (int x4 = 0; x4 < 16; ++x4) {
            const uint32_t w = sp[x4];
#pragma unroll
            for (int e = 0; e < 4; ++e) pv[4 * x4 + e] = ((w >> (8 * e)) & 0xffu) ? 0.f : pv[4 * x4 + e];
          }
        }
      }
      // the previous step's role MMAs read P / dS / dG: they must be done before we overwrite
      if (n > 0) tc::mbar_wait(step_done, (n - 1) & 1);
      if (MODE == MODE_DKV) {
        uint8_t* ptile = smem + Lay<MODE_DKV>::P + wg * TILE;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float* v = pv + c * 8;
          *reinterpret_cast<uint4*>(ptile + swz_chunk(a, c)) =
              make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                         pack_bf16x2(v[6], v[7]));
        }
      }
      // dS = P o (dP - D) / sqrt(dh)
      tc::mbar_wait(dp_full, par);
      tc::tc_fence_after();
      uint32_t A[32];
      {
        uint32_t r0[32], r1[32];
        tc::tmem_ld_32x32(tmem + TM_DP + lane_base + wg * 64, r0);
        tc::tmem_ld_32x32(tmem + TM_DP + lane_base + wg * 64 + 32, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          A[x] = pack_bf16x2(pv[2 * x] * (__uint_as_float(r0[2 * x]) - Dv) * p.scale,
                             pv[2 * x + 1] * (__uint_as_float(r0[2 * x + 1]) - Dv) * p.scale);
          A[16 + x] = pack_bf16x2(pv[32 + 2 * x] * (__uint_as_float(r1[2 * x]) - Dv) * p.scale,
                                  pv[32 + 2 * x + 1] * (__uint_as_float(r1[2 * x + 1]) - Dv) * p.scale);
        }
      }
      if (MODE != MODE_DE) {            // rectangular dS: sub-tile wg of row a
        uint8_t* dstile = smem + (MODE == MODE_DKV ? Lay<MODE_DKV>::DS : Lay<MODE_DQ>::DS) + wg * TILE;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(dstile + swz_chunk(a, c)) = make_uint4(A[4 * c], A[4 * c + 1], A[4 * c + 2], A[4 * c + 3]);
      }
      if (MODE != MODE_DKV) {           // band dG: this thread's 64 values start at band column 127-a+64*wg
        band_store(dg_base, a, base_w, A);
      }
      tc::fence_proxy_async();
      tc::tc_fence_before();
      tc::mbar_arrive(ds_ready);
    }

    // ---- epilogue: accumulators out of TMEM (each thread: 32 of the 64 columns of its row)
    tc::mbar_wait(step_done, (nsteps - 1) & 1);
    tc::tc_fence_after();
    if (MODE == MODE_DE) {
      const int c0 = p.max_seq - 1 - (int)blockIdx.z * TT;
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        const int erow = (blk == 0 ? c0 - (TT - 1) : c0 + 1) + a;
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem + (blk == 0 ? TM_ACC0 : TM_ACC1) + lane_base + wg * 32, r);
        tc::tmem_ld_wait();
        if (erow >= 0 && erow < p.max_seq) {
#pragma unroll
          for (int x = 0; x < 32; ++x) atomicAdd(p.dE + (int64_t)erow * DHC + wg * 32 + x, __uint_as_float(r[x]));
        }
      }
    } else {
      const StepInfo s = step_info<MODE>(p, 0, bh0);
      const int row = (MODE == MODE_DKV ? s.jt : s.it) * TT + a;
#pragma unroll
      for (int which = 0; which < (MODE == MODE_DKV ? 2 : 1); ++which) {
        uint32_t r[32], packed[16];
        tc::tmem_ld_32x32(tmem + (which == 0 ? TM_ACC0 : TM_ACC1) + lane_base + wg * 32, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; x += 2) packed[x / 2] = pack_bf16x2(__uint_as_float(r[x]), __uint_as_float(r[x + 1]));
        if (row < p.L) {
          void* base = (MODE == MODE_DQ) ? p.dq : (which == 0 ? p.dk : p.dv);
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(base) + (int64_t)s.b * p.sb +
                                                (int64_t)row * p.sl + (int64_t)s.hh * p.sh + wg * 32);
#pragma unroll
          for (int x = 0; x < 4; ++x)
            dst[x] = make_uint4(packed[4 * x], packed[4 * x + 1], packed[4 * x + 2], packed[4 * x + 3]);
        }
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 9) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

template <int MODE>
int launch_mode(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmDO,
                const CUtensorMap& tmE, const BwdParams& p, dim3 grid, cudaStream_t st) {
  auto kern = rga_bwd_tc_kernel<MODE>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<MODE>());
    if (e != cudaSuccess) { set_error("rga_bwd_tc: smem attribute (%d B): %s", smem_bytes<MODE>(), cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
  }
  kern<<<grid, NTHREADS, smem_bytes<MODE>(), st>>>(tmQ, tmK, tmV, tmDO, tmE, p);
  return check_launch("rga_bwd_tc");
}

}  // namespace

int rga_bwd2_dkv(const RgaArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                 const CUtensorMap& tmDO, const CUtensorMap& tmE, void* ds_ws, cudaStream_t st);      // rga_tc_bwd2.cu
int rga_bwd3_dq(const RgaArgs& a, const void* ws, const CUtensorMap& tmK, const CUtensorMap& tmE, cudaStream_t st);   // rga_tc_bwd3.cu
int rga_bwd3_de(const RgaArgs& a, const void* ws, const CUtensorMap& tmQ, const CUtensorMap& tmE, cudaStream_t st);
int rga_bwd2_de(const RgaArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                const CUtensorMap& tmDO, const CUtensorMap& tmE, cudaStream_t st);
int rga_bwd2_dq(const RgaArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                const CUtensorMap& tmDO, const CUtensorMap& tmE, cudaStream_t st);

bool rga_bwd_tc_supported(const RgaArgs& a, int dh, int dtype) {
  if (dh != DHC || dtype != MT_BF16 || !a.causal) return false;
  if (a.sl % 8 || a.sh % 8 || a.sb % 8 || a.ol % 8 || a.oh % 8 || a.ob % 8) return false;
  if (!aligned(a.q, 16) || !aligned(a.k, 16) || !aligned(a.v, 16) || !aligned(a.E, 16) || !aligned(a.dO, 16) ||
      !aligned(a.dq, 16) || !aligned(a.dk, 16) || !aligned(a.dv, 16) || !aligned(a.dE, 16))
    return false;
  return mt_device_ok() != 0;
}

int rga_bwd_tc(const RgaArgs& a, int dh, int dtype, void* ws, size_t ws_bytes, cudaStream_t st) {
  int rc;
  // a workspace of rga_bwd3_workspace_bytes() selects the dS-spill variant (S/P/dS computed once, in the
  // dK/dV role); without it every role recomputes them
  const bool spill = ws != nullptr && ws_bytes >= rga_bwd3_workspace_bytes(a.B, a.h, a.L) && aligned(ws, 128);
  if ((rc = rga_delta_launch(a, dh, dtype, st))) return rc;
  CUtensorMap tmQ, tmK, tmV, tmDO, tmE;
  if ((rc = tc::make_tmap_blhd(&tmQ, a.q, dh, a.L, a.h, a.B, a.sl, a.sh, a.sb, DHC, TT))) return rc;
  if ((rc = tc::make_tmap_blhd(&tmK, a.k, dh, a.L, a.h, a.B, a.sl, a.sh, a.sb, DHC, TT))) return rc;
  if ((rc = tc::make_tmap_blhd(&tmV, a.v, dh, a.L, a.h, a.B, a.sl, a.sh, a.sb, DHC, TT))) return rc;
  if ((rc = tc::make_tmap_blhd(&tmDO, a.dO, dh, a.L, a.h, a.B, a.ol, a.oh, a.ob, DHC, TT))) return rc;
  if ((rc = tc::make_tmap_2d(&tmE, a.E, a.max_seq, dh, dh, DHC, TT))) return rc;
  BwdParams p;
  p.dq = a.dq; p.dk = a.dk; p.dv = a.dv; p.sb = a.sb; p.sl = a.sl; p.sh = a.sh;
  p.dE = a.dE; p.lse = a.lse; p.delta = a.delta; p.pad = a.pad;
  p.B = a.B; p.h = a.h; p.L = a.L; p.max_seq = a.max_seq;
  p.nT = (a.L + TT - 1) / TT;
  p.scale = 1.f / a.inv_scale_div;
  p.scale_log2 = LOG2E / a.inv_scale_div;
  p.bh_per_cta = 1;
  dim3 grid(a.h, a.B, p.nT);
  // dK/dV and dE: second-generation two-group pipeline (rga_tc_bwd2.cu); dQ: the role kernel above
  if ((rc = rga_bwd2_dkv(a, tmQ, tmK, tmV, tmDO, tmE, spill ? ws : nullptr, st))) return rc;
  if (spill) {
    if ((rc = rga_bwd3_dq(a, ws, tmK, tmE, st))) return rc;
    return rga_bwd3_de(a, ws, tmQ, tmE, st);
  }
  // MT_RGA_DQ=1: the first-generation dQ role (kept for A/B timing)
  static const bool old_dq = getenv("MT_RGA_DQ") != nullptr && getenv("MT_RGA_DQ")[0] == '1';
  if (old_dq) { if ((rc = launch_mode<MODE_DQ>(tmQ, tmK, tmV, tmDO, tmE, p, grid, st))) return rc; }
  else if ((rc = rga_bwd2_dq(a, tmQ, tmK, tmV, tmDO, tmE, st))) return rc;
  return rga_bwd2_de(a, tmQ, tmK, tmV, tmDO, tmE, st);
}

}  // namespace mt
