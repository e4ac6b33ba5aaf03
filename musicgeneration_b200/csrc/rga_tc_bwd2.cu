// Backward of the fused relative global attention, second generation: the dK/dV role and the dE
// role as a two-group software pipeline (K2).  Math as in rga_tc_bwd.cu:
//     P = exp(S - lse),  S = (Q K^T + skew(Q E_band^T)) / sqrt(dh)
//     dP = dO V^T ;  dS = P o (dP - D) / sqrt(dh) ;  dV = P^T dO ;  dK = dS^T Q ;  dE_band = dG^T Q
// with dG = dS in band coordinates (dG[a][127-a+b] = dS[a][b]).
//
// What changed against the first generation: the 8 math warps did everything of a step in lock
// step (skew -> exp -> dS), so the LSU phase, the MUFU phase and the FMA phase of all warps
// coincided and every tensor-core phase was exposed.  Here the work of a step is split by FUNCTION:
//   * group A (warps 0-7)  : S + skew -> P = exp(S - lse) -> P (bf16, swizzled) to shared memory;
//   * group B (warps 8-15) : dP out of TMEM, P out of shared memory -> dS (and dG) operands;
//   * warp 16 TMA producer, warp 17 MMA issuer.
// A works on step n+1 while B works on step n, so the shared-memory/MUFU-heavy half and the
// FMA/TMEM-heavy half of the math overlap, and the MMAs of step n+1 (S, G_lo, G_hi) are issued as
// soon as A has read S/G of step n and B has read dP of step n (dP reuses the S columns).
// Row a = 32*(w&3)+lane, key columns 64*((w>>2)&1)..+63 for both groups.
#include "ops.cuh"
#include "rga_tc_common.cuh"

#include <stdlib.h>

namespace mt {

using namespace rga;

namespace {

enum { R_DKV = 0, R_DQ = 1, R_DE = 2 };

// 1: the skew of group A runs through a register barrel shifter (no shared-memory scratch traffic);
// 0: through the per-thread shared-memory scratch.  Measured on B200 (config B layer): dK/dV kernel
// 488 us (registers) vs 494 us (scratch), the recompute variant of the whole backward 1.40 vs 1.32 ms --
// removing ~100 KB of LSU traffic per step does not move the kernel, so the scratch stays the default.
#ifndef MT_SKEW_REGS
#define MT_SKEW_REGS 0
#endif

// dK/dV role, third generation (default): the 16 math warps split the tile by COLUMNS instead of by
// function -- a thread owns row a and 32 key columns and does everything for them (S + skew -> P -> dS).
// The trace of the two-group pipeline showed why it stalls: group A reads its second 32-column pass of
// S / G only after the math of the first, so `sg_free` (and with it dP, then group B, then the next S)
// waits ~1.3 k cycles per step, group B idles 70 % of the time.  Here all S / G columns are read at
// once (sg_free after ~350 cycles), dP is computed under the exponentials, P stays in registers for dS
// (no shared-memory hand-off), the skew runs through the register barrel shifter (no scratch for 512
// threads), and the next step's S / G products run under the dS math.
#ifndef MT_DKV_FUSED16
#define MT_DKV_FUSED16 1
#endif
static_assert(MT_DKV_FUSED16 == 1, "the two-group dK/dV variant needs the skew scratch, which the three-slot {Q, dO} ring has replaced");

constexpr int B2_GROUP = 256;                       // threads of one math group
constexpr int B2_THREADS = 2 * B2_GROUP + 96;       // A, B, TMA warp, MMA warp, second MMA warp (dK/dV role)
constexpr int SCRB_WORDS = 34;                      // skew scratch pitch (8-byte stores conflict-free)
constexpr int B2_SCR_BYTES = B2_GROUP * SCRB_WORDS * 4;

// TMEM columns: dP reuses the S columns once group A has read S
constexpr uint32_t TM_S = 0, TM_GLO = 128, TM_GHI = 256, TM_ACC0 = 384, TM_ACC1 = 448;

template <int ROLE> struct Lay2;
template <> struct Lay2<R_DKV> {     // K,V resident; Q x 3, dO x 3 (held until the dV / dK products of their step are done);
                                     // E block x 2 (E_hi of a step = E_lo of the previous one; free as soon as G is computed); P; dS
  static constexpr int K = 0, V = TILE, Q0 = 2 * TILE, DO0 = 5 * TILE, E0 = 8 * TILE;
  static constexpr int P = 10 * TILE, DS = 12 * TILE, BAR = 14 * TILE, SCR = 0;       // (no skew scratch: register barrel shifter)
  static constexpr int ST0 = Q0, ST_BYTES = TILE, sQ = 0, sDO = 0, sE = 0;            // (names the two-stage code paths of the other roles mention)
};
template <> struct Lay2<R_DE> {      // E_lo,E_hi resident; Q x 2; K; {dO, V} (doubles as the P hand-off); dG
  static constexpr int ELO = 0, EHI = TILE, Q0 = 2 * TILE, K = 4 * TILE, DOV = 5 * TILE, DG = 7 * TILE;
  static constexpr int SCR = 11 * TILE, BAR = SCR + B2_SCR_BYTES;
};
template <> struct Lay2<R_DQ> {      // Q,dO resident; K x 2; V; E ring x 3 (one new block per step, as in the forward); dG
  static constexpr int Q = 0, DO = TILE, K0 = 2 * TILE, V = 4 * TILE, E0 = 5 * TILE, DG = 8 * TILE;
  static constexpr int SCR = 12 * TILE, BAR = SCR + B2_SCR_BYTES;
};
template <int ROLE> constexpr int smem2_bytes() { return Lay2<ROLE>::BAR + 512; }
static_assert(smem2_bytes<R_DKV>() <= 232448 && smem2_bytes<R_DE>() <= 232448 && smem2_bytes<R_DQ>() <= 232448,
              "shared memory budget");
// dQ role: the G blocks form a ring (G_hi of step n is G_lo of step n+1), dQ accumulates in 64 columns
// and P / dS (bf16 pairs, 64 columns) never leave TMEM: group A stores P there, group B turns it into
// dS in place, and the dS.K MMA reads it as its A operand
constexpr uint32_t TM_DQ = 384, TM_PS = 448;

// barrier slots (uint64 each)
enum { BR_RES = 0, BR_QF = 1, BR_QE = 3, BR_KF = 5, BR_KE = 6, BR_VF = 7, BR_SFULL = 8, BR_SGFREE = 9,
       BR_DPFULL = 10, BR_DPFREE = 11, BR_PREADY = 12, BR_DSREADY = 13, BR_STEPDONE = 14, BR_TMEM = 15, BR_SPAD = 16,
       BR_QD_FULL = 32, BR_QD_EMPTY = 35, BR_E_FULL = 38 };      // dK/dV role: {Q, dO} ring of 3, E ring of 2

struct Bwd2Params {
  void* dk; void* dv; void* dq;          // 16-bit, q/k/v addressing
  int64_t sb, sl, sh;
  float* dE;
  const float* lse; const float* delta;
  const uint8_t* pad;
  int B, h, L, max_seq, nT;
  int bh_per_cta;                        // DE role
  uint8_t* ds_ws; int nTri;              // DKV role: spill every dS tile image here (rga_tc_bwd3.cu consumes them)
  int heads_per_cta;                     // DKV role: consecutive heads of one (batch row, key tile) walked by one CTA
  int qk_fmt;                            // DKV role: 16-bit format of EVERY MMA operand of the launch (1 = bf16; 0 = f16: the first
                                         // encoder layer -- tcgen05 kind::f16 traps on A / B of different formats, so there dO comes
                                         // in as f16(gscale * dO) and P / dS are packed as f16; dK / dV leave as bf16 either way)
  float gscale, inv_gscale;              // f16 mode: static loss scale of the gradient operands (dO, dS) and its inverse (outputs)
  float scale, scale_log2;
  long long* trace;                      // MT_RGA_TRACE=z: clock64 stamps of CTA (0,0,z), [4 agents][32 steps][8 events]
  int trace_z;
};

// pipeline timeline of one CTA (debug aid, off unless the launcher passes a buffer)
#define TRACE(agent, n, ev)                                                                         \
  do {                                                                                              \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (int)blockIdx.z == p.trace_z && (n) < 32)               \
      p.trace[((agent) * 32 + (n)) * 8 + (ev)] = clock64();                                         \
  } while (0)

struct Step2 { int it, jt, b, hh; };

template <int ROLE>
__device__ __forceinline__ int num_steps2(const Bwd2Params& p, int& bh0) {
  // grid = (h, B, nT) [DKV] or (slices, 1, nT) [DE]: the tile / diagonal index is the SLOWEST grid
  // dimension, so CTAs are dispatched longest-first over the whole launch
  bh0 = 0;
  if (ROLE == R_DKV) return min(p.heads_per_cta, p.h - (int)blockIdx.x * p.heads_per_cta) * (p.nT - (int)blockIdx.z);
  if (ROLE == R_DQ) return p.nT - (int)blockIdx.z;             // it = nT-1-blockIdx.z  ->  it+1 key tiles
  bh0 = (int)blockIdx.x * p.bh_per_cta;
  int nbh = min(p.bh_per_cta, p.B * p.h - bh0);
  return nbh > 0 ? nbh * (p.nT - (int)blockIdx.z) : 0;
}
template <int ROLE>
__device__ __forceinline__ Step2 step2(const Bwd2Params& p, int n, int bh0) {
  Step2 s;
  if (ROLE == R_DKV) {
    const int per = p.nT - (int)blockIdx.z;       // query tiles (steps) per head
    s.jt = blockIdx.z; s.it = s.jt + n % per; s.hh = (int)blockIdx.x * p.heads_per_cta + n / per; s.b = blockIdx.y;
  }
  else if (ROLE == R_DQ) { s.it = p.nT - 1 - (int)blockIdx.z; s.jt = n; s.hh = blockIdx.x; s.b = blockIdx.y; }
  else {
    const int per = p.nT - (int)blockIdx.z;
    const int bh = bh0 + n / per, k = n % per;
    s.it = (int)blockIdx.z + k; s.jt = k; s.b = bh / p.h; s.hh = bh % p.h;
  }
  return s;
}

// step n+1 from step n without the integer divisions of step2 (they sat in every agent's per-step chain)
template <int ROLE>
__device__ __forceinline__ void step_advance(const Bwd2Params& p, Step2& s) {
  if (ROLE == R_DKV) {                                   // next query tile, or the first one of the next head
    if (s.it + 1 < p.nT) ++s.it; else { s.it = (int)blockIdx.z; ++s.hh; }
    return;
  }
  if (ROLE == R_DQ) { ++s.jt; return; }
  if (s.it + 1 < p.nT) { ++s.it; ++s.jt; return; }       // next tile down the diagonal
  s.it = (int)blockIdx.z; s.jt = 0;                      // next (batch, head) of the slice
  if (++s.hh == p.h) { s.hh = 0; ++s.b; }
}

// 64-column band window [w0, w0+64) of this warp's 32 rows -> private scratch as 32 f16 pairs
// (pitch SCRB_WORDS: 8-byte stores)
__device__ __forceinline__ void park64_st64(uint32_t g_lo, uint32_t g_hi, uint32_t lane_base, int w0, uint32_t* scr) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    const int cc = w0 + 32 * c;
    tc::tmem_ld_32x32((cc < 128 ? g_lo + cc : g_hi + (cc - 128)) + lane_base, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int x = 0; x < 32; x += 4)
      *reinterpret_cast<uint2*>(scr + 16 * c + x / 2) =
          make_uint2(pack_f16x2(__uint_as_float(r[x]), __uint_as_float(r[x + 1])),
                     pack_f16x2(__uint_as_float(r[x + 2]), __uint_as_float(r[x + 3])));
  }
}

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// HF (dK/dV role only): the f16 mode described at Bwd2Params::qk_fmt
template <int ROLE, bool HF = false>
__global__ void __launch_bounds__(B2_THREADS, 1)
rga_bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmE, const Bwd2Params p) {
  using LY = Lay2<ROLE>;
  extern __shared__ __align__(1024) uint8_t smem[];      // shared address space kept: LDS/STS, not generic
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LY::BAR);
  uint64_t* bar_res = bars + BR_RES;
  uint64_t* q_full = bars + BR_QF;        // [2]  DKV: stage {Q,dO,E_lo};  DE: Q ring
  uint64_t* q_empty = bars + BR_QE;       // [2]
  uint64_t* k_full = bars + BR_KF;        // DE: K
  uint64_t* k_empty = bars + BR_KE;
  uint64_t* v_full = bars + BR_VF;        // DE: {dO, V}
  uint64_t* s_full = bars + BR_SFULL;     // MMA -> A : S, G_lo, G_hi of the step
  uint64_t* sg_free = bars + BR_SGFREE;   // A -> MMA : S and G read
  uint64_t* dp_full = bars + BR_DPFULL;   // MMA -> B (and A in the DE role)
  uint64_t* dp_free = bars + BR_DPFREE;   // B -> MMA : dP read (the S columns may be overwritten)
  uint64_t* p_ready = bars + BR_PREADY;   // A -> B (and MMA) : P stored
  uint64_t* ds_ready = bars + BR_DSREADY; // B -> MMA (and TMA in the DE role)
  uint64_t* step_done = bars + BR_STEPDONE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BR_TMEM);
  uint8_t* spad = reinterpret_cast<uint8_t*>(bars + BR_SPAD);      // [128]
  uint64_t* qd_full = bars + BR_QD_FULL;      // [3] dK/dV role: Q and dO of a step
  uint64_t* qd_empty = bars + BR_QD_EMPTY;    // [3]
  uint64_t* e_full = bars + BR_E_FULL;        // [2] dK/dV role: the step's new E block

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bh0;
  const int nsteps = num_steps2<ROLE>(p, bh0);

  if (warp == 16 && lane == 0) {
    tc::tma_prefetch_desc(&tmQ); tc::tma_prefetch_desc(&tmK); tc::tma_prefetch_desc(&tmV);
    tc::tma_prefetch_desc(&tmDO); tc::tma_prefetch_desc(&tmE);
    tc::mbar_init(bar_res, 1);
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&q_full[s], 1); tc::mbar_init(&q_empty[s], 1); }
    tc::mbar_init(k_full, 1); tc::mbar_init(k_empty, 1); tc::mbar_init(v_full, 1);
    tc::mbar_init(s_full, 1);
    constexpr int NARR = (ROLE == R_DKV && MT_DKV_FUSED16) ? 2 * B2_GROUP : B2_GROUP;   // arrivals per math barrier
    tc::mbar_init(sg_free, NARR);
    tc::mbar_init(dp_full, 1);
    tc::mbar_init(dp_free, NARR);
    tc::mbar_init(p_ready, NARR);
    tc::mbar_init(ds_ready, NARR);
    tc::mbar_init(step_done, 1);
    for (int q = 0; q < 3; ++q) { tc::mbar_init(&qd_full[q], 1); tc::mbar_init(&qd_empty[q], 1); }
    for (int q = 0; q < 2; ++q) tc::mbar_init(&e_full[q], 1);
    tc::fence_barrier_init();
  }
  if (warp == 17) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (nsteps <= 0) {            // (DE role: empty slice) -- uniform for the whole CTA
    __syncthreads();
    if (warp == 17) tc::tmem_dealloc(tmem, 512);
    return;
  }

  // buffers of step n
  auto buf_q = [&](int n) -> uint8_t* {
    if (ROLE == R_DKV) return smem + Lay2<R_DKV>::Q0 + (n % 3) * TILE;
    if (ROLE == R_DQ) return smem + Lay2<R_DQ>::Q;
    return smem + Lay2<R_DE>::Q0 + (n & 1) * TILE;
  };
  auto buf_do = [&](int n) -> uint8_t* {
    if (ROLE == R_DKV) return smem + Lay2<R_DKV>::DO0 + (n % 3) * TILE;
    if (ROLE == R_DQ) return smem + Lay2<R_DQ>::DO;
    return smem + Lay2<R_DE>::DOV;
  };
  auto buf_k = [&]() -> uint8_t* {          // DQ: slot 0 of the K ring (slot n&1 holds the key tile of step n)
    return smem + (ROLE == R_DKV ? Lay2<R_DKV>::K : (ROLE == R_DQ ? Lay2<R_DQ>::K0 : Lay2<R_DE>::K));
  };
  auto buf_v = [&]() -> uint8_t* {
    return smem + (ROLE == R_DKV ? Lay2<R_DKV>::V : (ROLE == R_DQ ? Lay2<R_DQ>::V : Lay2<R_DE>::DOV + TILE));
  };
  // DQ: E block m (the hi block of step m; block -1 = the lo block of step 0) lives in ring slot (m+1) % 3
  auto buf_eblk = [&](int m) -> uint8_t* { return smem + Lay2<R_DQ>::E0 + ((m + 1) % 3) * TILE; };
  auto buf_elo = [&](int n) -> uint8_t* {       // DKV: slot n & 1
    if (ROLE == R_DKV) return smem + Lay2<R_DKV>::E0 + (n & 1) * TILE;
    return smem + Lay2<R_DE>::ELO;
  };
  auto buf_ehi = [&](int n) -> uint8_t* {       // DKV: the E_lo block of the previous step
    if (ROLE == R_DKV) return smem + Lay2<R_DKV>::E0 + ((n + 1) & 1) * TILE;
    return smem + Lay2<R_DE>::EHI;
  };
  // P hand-off A -> B: the P operand buffer (DKV) or the {dO, V} tiles, dead once dP is computed (DE)
  uint8_t* const pbuf = smem + (ROLE == R_DKV ? Lay2<R_DKV>::P : Lay2<R_DE>::DOV);      // (DQ: P stays in TMEM)

  if (warp == 16) {
    // ================================ TMA producer ==========================================
    if (lane == 0) {
      if (ROLE == R_DKV) {
        Step2 s = step2<ROLE>(p, 0, bh0);
        for (int n = 0; n < nsteps; ++n, step_advance<ROLE>(p, s)) {
          const bool head_start = (s.it == (int)blockIdx.z);
          // {Q, dO} slot n % 3: last read by the dV / dK products of step n-3 -- these loads run two steps ahead
          const int q3 = n % 3;
          tc::mbar_wait(&qd_empty[q3], ((n / 3) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(&qd_full[q3], 2 * TILE);
          tc::tma_load_4d(buf_q(n), &tmQ, &qd_full[q3], 0, s.hh, s.it * TT, s.b);
          tc::tma_load_4d(buf_do(n), &tmDO, &qd_full[q3], 0, s.hh, s.it * TT, s.b);
          if (n + 3 < nsteps) {       // pull the tiles of step n+3 into L2
            Step2 t = s;
            step_advance<ROLE>(p, t); step_advance<ROLE>(p, t); step_advance<ROLE>(p, t);
            tc::tma_prefetch_4d(&tmQ, 0, t.hh, t.it * TT, t.b);
            tc::tma_prefetch_4d(&tmDO, 0, t.hh, t.it * TT, t.b);
            if (t.it == (int)blockIdx.z) {      // ... and the K / V tiles of a head that starts there
              tc::tma_prefetch_4d(&tmK, 0, t.hh, t.jt * TT, t.b);
              tc::tma_prefetch_4d(&tmV, 0, t.hh, t.jt * TT, t.b);
            }
          }
          if (head_start) {
            // the head's resident K / V tiles: the S and dP products of the previous head's last step must be done
            // with the old ones (dp_full is committed after both)
            // (this thread runs up to two steps ahead: dp_full may still be in phase n-2, where a parity-only wait
            // for phase n-1 would pass at once -- so phase n-2 first; phase n-3 is known complete from s_full(n-2))
            if (n > 1) tc::mbar_wait(dp_full, (n - 2) & 1);
            if (n > 0) tc::mbar_wait(dp_full, (n - 1) & 1);
            tc::mbar_arrive_expect_tx(bar_res, 2 * TILE);
            tc::tma_load_4d(buf_k(), &tmK, bar_res, 0, s.hh, s.jt * TT, s.b);
            tc::tma_load_4d(buf_v(), &tmV, bar_res, 0, s.hh, s.jt * TT, s.b);
          }
          // E slot n & 1 held the hi block of step n-1: free once G(n-1) is computed, which s_full(n-1) covers
          if (n >= 1) tc::mbar_wait(s_full, (n - 1) & 1);
          const int c0 = p.max_seq - 1 - (s.it - s.jt) * TT;
          tc::mbar_arrive_expect_tx(&e_full[n & 1], (head_start ? 2 : 1) * TILE);
          tc::tma_load_2d(buf_elo(n), &tmE, &e_full[n & 1], 0, c0 - (TT - 1));
          if (head_start) tc::tma_load_2d(buf_ehi(n), &tmE, &e_full[n & 1], 0, c0 + 1);
        }
      } else if (ROLE == R_DQ) {
        Step2 s = step2<ROLE>(p, 0, bh0);
        tc::mbar_arrive_expect_tx(bar_res, 2 * TILE);
        tc::tma_load_4d(buf_q(0), &tmQ, bar_res, 0, s.hh, s.it * TT, s.b);
        tc::tma_load_4d(buf_do(0), &tmDO, bar_res, 0, s.hh, s.it * TT, s.b);
        for (int n = 0; n < nsteps; ++n, step_advance<ROLE>(p, s)) {
          const int st = n & 1;
          const int c0 = p.max_seq - 1 - (s.it - s.jt) * TT;
          // K slot n&1 and E slot (n+1)%3 were last read by the role MMAs of step n-2
          tc::mbar_wait(&q_empty[st], ((n >> 1) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(&q_full[st], (n == 0 ? 3 : 2) * TILE);
          tc::tma_load_4d(buf_k() + st * TILE, &tmK, &q_full[st], 0, s.hh, s.jt * TT, s.b);
          tc::tma_load_2d(buf_eblk(n), &tmE, &q_full[st], 0, c0 + 1);
          if (n == 0) tc::tma_load_2d(buf_eblk(-1), &tmE, &q_full[st], 0, c0 - (TT - 1));
          // V: single slot, free as soon as dP of the previous step has been computed
          tc::mbar_wait(k_empty, (n & 1) ^ 1);
          tc::mbar_arrive_expect_tx(v_full, TILE);
          tc::tma_load_4d(buf_v(), &tmV, v_full, 0, s.hh, s.jt * TT, s.b);
          if (n + 2 < nsteps) {
            tc::tma_prefetch_4d(&tmK, 0, s.hh, (s.jt + 2) * TT, s.b);
            tc::tma_prefetch_4d(&tmV, 0, s.hh, (s.jt + 2) * TT, s.b);
          }
        }
      } else {
        const int c0 = p.max_seq - 1 - (int)blockIdx.z * TT;
        tc::mbar_arrive_expect_tx(bar_res, 2 * TILE);
        tc::tma_load_2d(smem + Lay2<R_DE>::ELO, &tmE, bar_res, 0, c0 - (TT - 1));
        tc::tma_load_2d(smem + Lay2<R_DE>::EHI, &tmE, bar_res, 0, c0 + 1);
        Step2 s = step2<ROLE>(p, 0, bh0);
        for (int n = 0; n < nsteps; ++n, step_advance<ROLE>(p, s)) {
          tc::mbar_wait(k_empty, (n & 1) ^ 1);
          tc::mbar_arrive_expect_tx(k_full, TILE);
          tc::tma_load_4d(buf_k(), &tmK, k_full, 0, s.hh, s.jt * TT, s.b);
          tc::mbar_wait(&q_empty[n & 1], ((n >> 1) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(&q_full[n & 1], TILE);
          tc::tma_load_4d(buf_q(n), &tmQ, &q_full[n & 1], 0, s.hh, s.it * TT, s.b);
          if (n + 1 < nsteps) {       // pull the next step's tiles into L2 while waiting for the hand-off buffer
            Step2 t = s;
            step_advance<ROLE>(p, t);
            tc::tma_prefetch_4d(&tmK, 0, t.hh, t.jt * TT, t.b);
            tc::tma_prefetch_4d(&tmQ, 0, t.hh, t.it * TT, t.b);
            tc::tma_prefetch_4d(&tmDO, 0, t.hh, t.it * TT, t.b);
            tc::tma_prefetch_4d(&tmV, 0, t.hh, t.jt * TT, t.b);
          }
          // {dO, V} region: group B must have finished reading P of the previous step out of it
          if (n > 0) tc::mbar_wait(ds_ready, (n - 1) & 1);
          tc::mbar_arrive_expect_tx(v_full, 2 * TILE);
          tc::tma_load_4d(buf_do(n), &tmDO, v_full, 0, s.hh, s.it * TT, s.b);
          tc::tma_load_4d(buf_v(), &tmV, v_full, 0, s.hh, s.jt * TT, s.b);
        }
      }
    }
  } else if (warp == 17) {
    // ================================ MMA issuer ============================================
    if (ROLE == R_DKV) {
      if (lane == 0) {         // S, G, dP products (dV / dK: second issuer, warp 18)
        const uint32_t id_kk = tc::make_idesc(TT, TT, p.qk_fmt, p.qk_fmt, 0, 0);      // S, G, dP: K-major x K-major, N = 128
        constexpr uint64_t TS16 = TILE >> 4;
        const uint64_t kd = tc::make_sdesc(tc::smem_u32(buf_k()), 16, 1024);
        const uint64_t vd = tc::make_sdesc(tc::smem_u32(buf_v()), 16, 1024);
        const uint64_t qd0 = tc::make_sdesc(tc::smem_u32(smem + Lay2<R_DKV>::Q0), 16, 1024);
        const uint64_t dod0 = tc::make_sdesc(tc::smem_u32(smem + Lay2<R_DKV>::DO0), 16, 1024);
        const uint64_t ed0 = tc::make_sdesc(tc::smem_u32(smem + Lay2<R_DKV>::E0), 16, 1024);
        auto issue_g = [&](int n) {
          const uint64_t qd = qd0 + (uint64_t)(n % 3) * TS16;
          const uint64_t lo = ed0 + (uint64_t)(n & 1) * TS16, hi = ed0 + (uint64_t)((n + 1) & 1) * TS16;
#pragma unroll
          for (int k4 = 0; k4 < DHC / 16; ++k4) {
            tc::umma_f16(tmem + TM_GLO, qd + 2 * k4, lo + 2 * k4, id_kk, k4 != 0);
            tc::umma_f16(tmem + TM_GHI, qd + 2 * k4, hi + 2 * k4, id_kk, k4 != 0);
          }
        };
        auto issue_s = [&](int n) {
          const uint64_t qd = qd0 + (uint64_t)(n % 3) * TS16;
#pragma unroll
          for (int k4 = 0; k4 < DHC / 16; ++k4)
            tc::umma_f16(tmem + TM_S, qd + 2 * k4, kd + 2 * k4, id_kk, k4 != 0);
          tc::umma_commit(s_full);
        };
        tc::mbar_wait(bar_res, 0);
        tc::mbar_wait(&qd_full[0], 0);
        tc::mbar_wait(&e_full[0], 0);
        tc::tc_fence_after();
        issue_g(0);
        issue_s(0);
        const int per = p.nT - (int)blockIdx.z;   // steps per head
        int k = 0, item = 0;
        for (int n = 0; n < nsteps; ++n) {
          const uint32_t par = n & 1;
          tc::mbar_wait(sg_free, par);            // every math warp has S and G of the step in registers
          tc::tc_fence_after();
          TRACE(3, n, 0);
          const uint64_t dod = dod0 + (uint64_t)(n % 3) * TS16;
#pragma unroll
          for (int k4 = 0; k4 < DHC / 16; ++k4)   // dP = dO V^T into the S columns
            tc::umma_f16(tmem + TM_S, dod + 2 * k4, vd + 2 * k4, id_kk, k4 != 0);
          tc::umma_commit(dp_full);
          TRACE(3, n, 1);
          if (++k == per) { k = 0; ++item; }
          if (n + 1 < nsteps) {
            if (k == 0) tc::mbar_wait(bar_res, item & 1);          // K / V of the head that starts with step n+1
            tc::mbar_wait(&qd_full[(n + 1) % 3], ((n + 1) / 3) & 1);
            tc::mbar_wait(&e_full[(n + 1) & 1], ((n + 1) >> 1) & 1);
            tc::tc_fence_after();
            TRACE(3, n, 2);
            issue_g(n + 1);
            TRACE(3, n, 3);
            tc::mbar_wait(dp_free, par);
            tc::tc_fence_after();
            TRACE(3, n, 4);
            issue_s(n + 1);
          }
        }
      }
    } else if (ROLE == R_DQ) {
      if (lane == 0) {
        const uint32_t id_kk = tc::make_idesc(TT, TT, 1, 1, 0, 0);      // S, G, dP : K-major x K-major, N = 128
        const uint32_t id_kmn = tc::make_idesc(TT, DHC, 1, 1, 0, 1);    // dQ: A K-major (TMEM dS / smem dG), B MN-major, N = 64
        constexpr uint64_t TS16 = TILE >> 4;
        const uint64_t qd = tc::make_sdesc(tc::smem_u32(buf_q(0)), 16, 1024);
        const uint64_t dod = tc::make_sdesc(tc::smem_u32(buf_do(0)), 16, 1024);
        const uint64_t vd = tc::make_sdesc(tc::smem_u32(buf_v()), 16, 1024);
        const uint64_t kd_k0 = tc::make_sdesc(tc::smem_u32(buf_k()), 16, 1024);
        const uint64_t kd_mn0 = tc::make_sdesc(tc::smem_u32(buf_k()), 1024, 1024);
        const uint64_t ed_k0 = tc::make_sdesc(tc::smem_u32(smem + Lay2<R_DQ>::E0), 16, 1024);
        const uint64_t ed_mn0 = tc::make_sdesc(tc::smem_u32(smem + Lay2<R_DQ>::E0), 1024, 1024);
        const uint64_t dgd = tc::make_sdesc(tc::smem_u32(smem + Lay2<R_DQ>::DG), 16, 1024);
        auto eslot = [](int m) -> uint64_t { return (uint64_t)((m + 1) % 3) * TS16; };
        auto issue_ghi = [&](int n) {          // Q . blk(n)^T -> the G region that held G_lo of step n-1
          const uint32_t g_hi = tmem + (((n + 1) & 1) ? TM_GHI : TM_GLO);
#pragma unroll
          for (int k4 = 0; k4 < DHC / 16; ++k4)
            tc::umma_f16(g_hi, qd + 2 * k4, ed_k0 + eslot(n) + 2 * k4, id_kk, k4 != 0);
        };
        auto issue_s = [&](int n) {
#pragma unroll
          for (int k4 = 0; k4 < DHC / 16; ++k4)
            tc::umma_f16(tmem + TM_S, qd + 2 * k4, kd_k0 + (uint64_t)(n & 1) * TS16 + 2 * k4, id_kk, k4 != 0);
          tc::umma_commit(s_full);
        };
        tc::mbar_wait(bar_res, 0);
        tc::mbar_wait(&q_full[0], 0);
        tc::tc_fence_after();
#pragma unroll
        for (int k4 = 0; k4 < DHC / 16; ++k4)   // G_lo of step 0 (block -1)
          tc::umma_f16(tmem + TM_GLO, qd + 2 * k4, ed_k0 + eslot(-1) + 2 * k4, id_kk, k4 != 0);
        issue_ghi(0);
        issue_s(0);
        for (int n = 0; n < nsteps; ++n) {
          const uint32_t par = n & 1;
          tc::mbar_wait(sg_free, par);
          tc::mbar_wait(v_full, par);
          tc::tc_fence_after();
#pragma unroll
          for (int k4 = 0; k4 < DHC / 16; ++k4)   // dP = dO V^T into the S columns
            tc::umma_f16(tmem + TM_S, dod + 2 * k4, vd + 2 * k4, id_kk, k4 != 0);
          tc::umma_commit(dp_full);
          tc::umma_commit(k_empty);               // V slot free
          if (n + 1 < nsteps) {
            tc::mbar_wait(&q_full[(n + 1) & 1], ((n + 1) >> 1) & 1);
            tc::tc_fence_after();
            issue_ghi(n + 1);
            tc::mbar_wait(dp_free, par);
            tc::tc_fence_after();
            issue_s(n + 1);
          }
          tc::mbar_wait(ds_ready, par);
          tc::tc_fence_after();
#pragma unroll
          for (int k16 = 0; k16 < TT / 16; ++k16)         // dQ += dS . K_j : dS is the TMEM A operand (8 columns per 16 keys)
            tc::umma_f16_ts(tmem + TM_DQ, tmem + TM_PS + 8 * k16, kd_mn0 + (uint64_t)(n & 1) * TS16 + 128 * k16, id_kmn,
                            (n | k16) != 0);
#pragma unroll
          for (int k16 = 0; k16 < 2 * TT / 16; ++k16)     // dQ += dG . [E_lo; E_hi] (contraction over the band)
            tc::umma_f16(tmem + TM_DQ, dgd + (uint64_t)(k16 >> 2) * TS16 + 2 * (k16 & 3),
                         ed_mn0 + (k16 < 8 ? eslot(n - 1) + 128 * k16 : eslot(n) + 128 * (k16 - 8)), id_kmn, 1);
          tc::umma_commit(&q_empty[n & 1]);
          tc::umma_commit(step_done);
        }
      }
    } else if (lane == 0) {
      const uint32_t id_kk = tc::make_idesc(TT, TT, 1, 1, 0, 0);      // S, G, dP : K-major x K-major, N = 128
      const uint32_t id_mnmn = tc::make_idesc(TT, DHC, 1, 1, 1, 1);   // dK/dV/dE: A MN-major, B MN-major, N = 64
      // Shared-memory descriptors are built once; inside the loops a k-step is an add on the address
      // field (16-byte units): +2 per 16 elements of a K-major operand, +128 per 16 rows of an MN-major one.
      // (Building them per MMA cost the issuing thread ~100 cycles per instruction.)
      const uint64_t kd = tc::make_sdesc(tc::smem_u32(buf_k()), 16, 1024);
      const uint64_t vd = tc::make_sdesc(tc::smem_u32(buf_v()), 16, 1024);
      // stage st of a double-buffered tile = stage 0 + st * stride (no indexed descriptor arrays:
      // they would live in local memory)
      constexpr uint64_t QSTR = (ROLE == R_DKV ? Lay2<R_DKV>::ST_BYTES : TILE) >> 4;          // Q: both roles x2
      constexpr uint64_t SSTR = (ROLE == R_DKV ? Lay2<R_DKV>::ST_BYTES : 0) >> 4;             // dO, E_lo: DKV only
      const uint64_t qd_k0 = tc::make_sdesc(tc::smem_u32(buf_q(0)), 16, 1024);
      const uint64_t qd_mn0 = tc::make_sdesc(tc::smem_u32(buf_q(0)), 1024, 1024);
      const uint64_t dod_k0 = tc::make_sdesc(tc::smem_u32(buf_do(0)), 16, 1024);
      const uint64_t dod_mn0 = tc::make_sdesc(tc::smem_u32(buf_do(0)), 1024, 1024);
      const uint64_t elod0 = tc::make_sdesc(tc::smem_u32(buf_elo(0)), 16, 1024);
      const uint64_t ehid_de = tc::make_sdesc(tc::smem_u32(smem + Lay2<R_DE>::EHI), 16, 1024);
      const uint64_t opd0 = tc::make_sdesc(tc::smem_u32(smem + (ROLE == R_DKV ? Lay2<R_DKV>::P : Lay2<R_DE>::DG)), TILE, 1024);
      const uint64_t opd1 = tc::make_sdesc(tc::smem_u32(smem + (ROLE == R_DKV ? Lay2<R_DKV>::DS : Lay2<R_DE>::DG + 2 * TILE)), TILE, 1024);
      auto issue_g = [&](int n) {
        const uint64_t st = n & 1;
        const uint64_t qd = qd_k0 + st * QSTR, lo = elod0 + st * SSTR;
        const uint64_t hi = (ROLE == R_DKV) ? elod0 + (st ^ 1) * SSTR : ehid_de;      // DKV: E_lo of the previous step
#pragma unroll
        for (int k4 = 0; k4 < DHC / 16; ++k4) {
          tc::umma_f16(tmem + TM_GLO, qd + 2 * k4, lo + 2 * k4, id_kk, k4 != 0);
          tc::umma_f16(tmem + TM_GHI, qd + 2 * k4, hi + 2 * k4, id_kk, k4 != 0);
        }
      };
      auto issue_s = [&](int n) {
        const uint64_t qd = qd_k0 + (uint64_t)(n & 1) * QSTR;
#pragma unroll
        for (int k4 = 0; k4 < DHC / 16; ++k4)
          tc::umma_f16(tmem + TM_S, qd + 2 * k4, kd + 2 * k4, id_kk, k4 != 0);
        tc::umma_commit(s_full);
        if (ROLE == R_DE) tc::umma_commit(k_empty);
      };
      tc::mbar_wait(bar_res, 0);
      tc::mbar_wait(&q_full[0], 0);
      if (ROLE == R_DE) tc::mbar_wait(k_full, 0);
      tc::tc_fence_after();
      issue_g(0);
      issue_s(0);
      for (int n = 0; n < nsteps; ++n) {
        const uint32_t par = n & 1;
        const uint64_t st = n & 1;
        const uint64_t dod_k = dod_k0 + st * SSTR, dod_mn = dod_mn0 + st * SSTR, qd_mn = qd_mn0 + st * QSTR;
        // ---- dP = dO V^T into the S columns (group A has read S and G)
        tc::mbar_wait(sg_free, par);
        TRACE(3, n, 0);
        if (ROLE == R_DE) tc::mbar_wait(v_full, par);
        tc::tc_fence_after();
        TRACE(3, n, 1);
#pragma unroll
        for (int k4 = 0; k4 < DHC / 16; ++k4)
          tc::umma_f16(tmem + TM_S, dod_k + 2 * k4, vd + 2 * k4, id_kk, k4 != 0);
        tc::umma_commit(dp_full);
        // ---- next step's G (needs only the G columns), then its S once group B has read dP
        if (n + 1 < nsteps) {
          tc::mbar_wait(&q_full[(n + 1) & 1], ((n + 1) >> 1) & 1);
          tc::tc_fence_after();
          TRACE(3, n, 2);
          issue_g(n + 1);
          if (ROLE == R_DE) tc::mbar_wait(k_full, (n + 1) & 1);
          TRACE(3, n, 3);
          tc::mbar_wait(dp_free, par);
          tc::tc_fence_after();
          TRACE(3, n, 4);
          issue_s(n + 1);
        }
        if (ROLE == R_DKV && MT_DKV_FUSED16) continue;      // dV / dK products: second issuer (warp 18)
        // ---- role MMAs on the operands written by the math groups
        tc::mbar_wait(p_ready, par);
        TRACE(3, n, 5);
        tc::mbar_wait(ds_ready, par);
        tc::tc_fence_after();
        TRACE(3, n, 6);
        if (ROLE == R_DKV) {
          if (p.ds_ws) {        // the dS operand image (32 KB, swizzled) goes to the workspace as it is
            const int it = (int)blockIdx.z + n, jt = (int)blockIdx.z;
            uint8_t* dst = p.ds_ws + (((int64_t)blockIdx.y * p.h + blockIdx.x) * p.nTri + (it * (it + 1) / 2 + jt)) *
                                         (int64_t)(2 * TILE);
            tc::bulk_store_1d(dst, smem + Lay2<R_DKV>::DS, 2 * TILE);
            tc::bulk_commit();
          }
#pragma unroll
          for (int k16 = 0; k16 < TT / 16; ++k16) {      // contraction over the 128 query rows
            tc::umma_f16(tmem + TM_ACC1, opd0 + 128 * k16, dod_mn + 128 * k16, id_mnmn, (n | k16) != 0);   // dV += P^T dO
            tc::umma_f16(tmem + TM_ACC0, opd1 + 128 * k16, qd_mn + 128 * k16, id_mnmn, (n | k16) != 0);    // dK += dS^T Q
          }
          // group B overwrites the dS buffer once step_done is signalled: the copy must have read it
          if (p.ds_ws) tc::bulk_wait_read0();
        } else {
#pragma unroll
          for (int k16 = 0; k16 < TT / 16; ++k16) {       // dG_blk^T . Q (contraction over the query rows)
            tc::umma_f16(tmem + TM_ACC0, opd0 + 128 * k16, qd_mn + 128 * k16, id_mnmn, (n | k16) != 0);
            tc::umma_f16(tmem + TM_ACC1, opd1 + 128 * k16, qd_mn + 128 * k16, id_mnmn, (n | k16) != 0);
          }
        }
        tc::umma_commit(&q_empty[n & 1]);
        tc::umma_commit(step_done);
      }
      if (ROLE == R_DKV && !MT_DKV_FUSED16 && p.ds_ws) tc::bulk_wait0();
    }
  } else if (warp == 18) {
    // ================================ second MMA issuer (dK/dV role) ========================
    // The first issuer's loop is a chain of waits (sg_free -> dP, q_full -> G, dp_free -> S); with the 16
    // N = 64 products of dV / dK (and the dS spill) in the same thread every one of those waits was added
    // to the products' issue time (trace: 5.7 k cycles per loop iteration = the step).  The two groups of
    // products touch disjoint TMEM columns and shared-memory stages, so they are issued by two threads.
    if (ROLE == R_DKV && MT_DKV_FUSED16 && lane == 0) {
      const uint32_t id_mnmn = tc::make_idesc(TT, DHC, p.qk_fmt, p.qk_fmt, 1, 1);   // A MN-major (P / dS), B MN-major (dO / Q), N = 64
      constexpr uint64_t STR = TILE >> 4;
      const uint64_t qd_mn0 = tc::make_sdesc(tc::smem_u32(smem + Lay2<R_DKV>::Q0), 1024, 1024);
      const uint64_t dod_mn0 = tc::make_sdesc(tc::smem_u32(smem + Lay2<R_DKV>::DO0), 1024, 1024);
      const uint64_t opd0 = tc::make_sdesc(tc::smem_u32(smem + Lay2<R_DKV>::P), TILE, 1024);
      const uint64_t opd1 = tc::make_sdesc(tc::smem_u32(smem + Lay2<R_DKV>::DS), TILE, 1024);
      const int per = p.nT - (int)blockIdx.z;     // steps per head
      int k = 0, hh = (int)blockIdx.x * p.heads_per_cta;
      for (int n = 0; n < nsteps; ++n) {
        const uint32_t par = n & 1;
        const uint64_t dod_mn = dod_mn0 + (uint64_t)(n % 3) * STR, qd_mn = qd_mn0 + (uint64_t)(n % 3) * STR;
        tc::mbar_wait(p_ready, par);
        tc::mbar_wait(ds_ready, par);
        tc::tc_fence_after();
        if (p.ds_ws) {        // the dS operand image (32 KB, swizzled) goes to the workspace as it is
          const int it = (int)blockIdx.z + k, jt = (int)blockIdx.z;
          uint8_t* dst = p.ds_ws + (((int64_t)blockIdx.y * p.h + hh) * p.nTri + (it * (it + 1) / 2 + jt)) *
                                       (int64_t)(2 * TILE);
          tc::bulk_store_1d(dst, smem + Lay2<R_DKV>::DS, 2 * TILE);
          tc::bulk_commit();
        }
#pragma unroll
        for (int k16 = 0; k16 < TT / 16; ++k16) {      // contraction over the 128 query rows
          tc::umma_f16(tmem + TM_ACC1, opd0 + 128 * k16, dod_mn + 128 * k16, id_mnmn, (k | k16) != 0);   // dV += P^T dO
          tc::umma_f16(tmem + TM_ACC0, opd1 + 128 * k16, qd_mn + 128 * k16, id_mnmn, (k | k16) != 0);    // dK += dS^T Q
        }
        if (p.ds_ws) tc::bulk_wait_read0();            // the math warps overwrite dS once step_done is signalled
        tc::umma_commit(&qd_empty[n % 3]);
        tc::umma_commit(step_done);
        if (++k == per) { k = 0; ++hh; }               // (the first products of the next head restart the accumulators)
      }
      if (p.ds_ws) tc::bulk_wait0();
    }
  } else {
    const bool grpA = warp < 8;
    const int w4 = warp & 3, half = (warp >> 2) & 1;
    const int a = w4 * 32 + lane;
    const uint32_t lane_base = (uint32_t)(w4 * 32) << 16;

    if (ROLE == R_DKV && MT_DKV_FUSED16) {
      // ================================ 16 warps, 32 columns each: S + skew -> P -> dS ==========
      const int grp = warp >> 3;
      const int cfirst = 64 * half + 32 * grp;                // this thread's key columns of the tile
      const uint8_t* pad = p.pad;
      if (pad) {            // any padded key in the sequences this CTA touches?  if not, the unmasked fast path
        const Step2 sf = step2<ROLE>(p, 0, bh0), sl = step2<ROLE>(p, nsteps - 1, bh0);
        bool mine = false;
        for (int64_t x = (int64_t)sf.b * p.L + threadIdx.x; x < (int64_t)(sl.b + 1) * p.L; x += 2 * B2_GROUP)
          mine |= (pad[x] != 0);
        if (!tc::named_bar_red_or(1, 2 * B2_GROUP, mine)) pad = nullptr;
      }
      auto row_stat = [&](const float* src, const Step2& t) -> float {
        const int i = t.it * TT + a;
        return i < p.L ? src[((int64_t)t.b * p.h + t.hh) * p.L + i] : 0.f;
      };
      Step2 s = step2<ROLE>(p, 0, bh0), snext = s;
      float lse_next = row_stat(p.lse, s), d_next = row_stat(p.delta, s);
      uint8_t* const ptile = pbuf + half * TILE;
      uint8_t* const dstile = smem + Lay2<R_DKV>::DS + half * TILE;
      // dK (warps 0-7, ACC0) / dV (warps 8-15, ACC1) of head hh -> global; each thread 32 of the 64 columns of key row a
      auto store_acc = [&](int hh) {
        uint32_t r[32];
        // f16 mode: dV = P^T (g dO) and dK = (g dS)^T Q carry the loss scale g
        const float osc = HF ? p.inv_gscale : 1.f;
        tc::tmem_ld_32x32(tmem + (grp ? TM_ACC1 : TM_ACC0) + lane_base + half * 32, r);
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        const int row = (int)blockIdx.z * TT + a;
        if (row < p.L) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(grp ? p.dv : p.dk) + (int64_t)blockIdx.y * p.sb +
                                                (int64_t)row * p.sl + (int64_t)hh * p.sh + half * 32);
#pragma unroll
          for (int x = 0; x < 4; ++x)
            dst[x] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * x]) * osc, __uint_as_float(r[8 * x + 1]) * osc),
                                pack_bf16x2(__uint_as_float(r[8 * x + 2]) * osc, __uint_as_float(r[8 * x + 3]) * osc),
                                pack_bf16x2(__uint_as_float(r[8 * x + 4]) * osc, __uint_as_float(r[8 * x + 5]) * osc),
                                pack_bf16x2(__uint_as_float(r[8 * x + 6]) * osc, __uint_as_float(r[8 * x + 7]) * osc));
        }
      };
      for (int n = 0; n < nsteps; ++n, s = snext) {
        const uint32_t par = n & 1;
        const int i0 = s.it * TT, j0 = s.jt * TT;
        const bool row_ok = i0 + a < p.L;
        // (f16 mode: dP arrives scaled by g, so D is scaled to match and dS = g * the true dS)
        const float lse2 = lse_next * LOG2E, Ds = d_next * p.scale * (HF ? p.gscale : 1.f);
        step_advance<ROLE>(p, snext);
        if (n + 1 < nsteps) { lse_next = row_stat(p.lse, snext); d_next = row_stat(p.delta, snext); }
        if (pad) {
          tc::named_bar_sync(1, 2 * B2_GROUP);
          if (warp < 4) spad[a] = (j0 + a < p.L) ? pad[(int64_t)s.b * p.L + j0 + a] : 1;
          tc::named_bar_sync(1, 2 * B2_GROUP);
        }
        const bool diag = (i0 == j0);
        const bool need_mask = diag || (j0 + TT > p.L) || (i0 + TT > p.L) || pad != nullptr;
        if (threadIdx.x == 0) TRACE(0, n, 0);
        tc::mbar_wait(s_full, par);
        tc::tc_fence_after();
        if (threadIdx.x == 0) TRACE(0, n, 1);
        uint32_t Wn[32];
        skew_window_64(tmem + TM_GLO, tmem + TM_GHI, lane_base, 96 - 32 * w4 + cfirst, Wn);
        float sv[32];
        {
          uint32_t r[32];
          tc::tmem_ld_32x32(tmem + TM_S + lane_base + cfirst, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int x = 0; x < 32; ++x) sv[x] = __uint_as_float(r[x]);
        }
        tc::tc_fence_before();
        tc::mbar_arrive(sg_free);               // S and G are in registers: dP and the next G may overwrite them
        if (threadIdx.x == 0) TRACE(0, n, 2);
        skew_shift_add_32(sv, Wn, lane);
#pragma unroll
        for (int x = 0; x < 32; ++x) sv[x] = tc::fast_exp2(fmaf(sv[x], p.scale_log2, -lse2));
        if (need_mask) {
          int lim = 31;
          if (diag) lim = min(lim, a - cfirst);
          lim = min(lim, p.L - 1 - j0 - cfirst);
          if (!row_ok) lim = -1;
#pragma unroll
          for (int x = 0; x < 32; ++x) sv[x] = (x > lim) ? 0.f : sv[x];
          if (pad) {
            const uint32_t* sp = reinterpret_cast<const uint32_t*>(spad + cfirst);
#pragma unroll
            for (int x4 = 0; x4 < 8; ++x4) {
              const uint32_t w = sp[x4];
#pragma unroll
              for (int e = 0; e < 4; ++e) sv[4 * x4 + e] = ((w >> (8 * e)) & 0xffu) ? 0.f : sv[4 * x4 + e];
            }
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) pk[x] = HF ? pack_f16x2(sv[2 * x], sv[2 * x + 1]) : pack_bf16x2(sv[2 * x], sv[2 * x + 1]);
        if (threadIdx.x == 0) TRACE(0, n, 3);
        // dP was issued right after sg_free and has long arrived: read it BEFORE the stores of P (which wait for
        // the previous step's dV / dK products), so that dp_free -- the go-ahead of the next step's S product --
        // does not wait behind them
        tc::mbar_wait(dp_full, par);
        tc::tc_fence_after();
        if (threadIdx.x == 0) TRACE(0, n, 4);
        uint32_t A16[16];
        {
          uint32_t dp[32];
          tc::tmem_ld_32x32(tmem + TM_S + lane_base + cfirst, dp);
          tc::tmem_ld_wait();
          tc::tc_fence_before();
          tc::mbar_arrive(dp_free);             // the S columns may take the next step's S
          // dS = P o (dP - D) / sqrt(dh), with the bf16-rounded P the dV product sees
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float p0, p1;
            if (HF) { const float2 pf = __half22float2(*reinterpret_cast<const __half2*>(&pk[k])); p0 = pf.x; p1 = pf.y; }
            else { p0 = bf16lo(pk[k]); p1 = bf16hi(pk[k]); }
            const float d0 = fmaf(__uint_as_float(dp[2 * k]), p.scale, -Ds) * p0;
            const float d1 = fmaf(__uint_as_float(dp[2 * k + 1]), p.scale, -Ds) * p1;
            A16[k] = HF ? pack_f16x2(d0, d1) : pack_bf16x2(d0, d1);
          }
        }
        if (threadIdx.x == 0) TRACE(0, n, 5);
        if (n > 0) tc::mbar_wait(step_done, (n - 1) & 1);      // dV / dK of the previous step have read P and dS
        if (threadIdx.x == 0) TRACE(0, n, 6);
        if (n > 0 && i0 == j0) {
          // first step of a new head: the previous head's dK / dV are complete (step_done above) and must leave
          // TMEM before this step's products -- issued once every thread has arrived on p_ready / ds_ready
          // below -- restart the accumulators
          tc::tc_fence_after();
          store_acc(s.hh - 1);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          *reinterpret_cast<uint4*>(ptile + swz_chunk(a, 4 * grp + c)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          *reinterpret_cast<uint4*>(dstile + swz_chunk(a, 4 * grp + c)) = make_uint4(A16[4 * c], A16[4 * c + 1], A16[4 * c + 2], A16[4 * c + 3]);
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(p_ready);
        tc::mbar_arrive(ds_ready);
        if (threadIdx.x == 0) TRACE(0, n, 7);
      }
      // the last head (s is one step past it: s.hh has already advanced)
      tc::mbar_wait(step_done, (nsteps - 1) & 1);
      tc::tc_fence_after();
      store_acc(s.hh - 1);
    } else if (grpA) {
      // ================================ group A: S + skew -> P ================================
      uint32_t* scr = reinterpret_cast<uint32_t*>(smem + LY::SCR) + threadIdx.x * SCRB_WORDS;
      // training batches carry no pad tokens: decide once per CTA whether any key of the sequences
      // this CTA touches is padded; if none is, the unmasked fast path is used for every step
      const uint8_t* pad = p.pad;
      if (pad) {
        const Step2 sf = step2<ROLE>(p, 0, bh0), sl = step2<ROLE>(p, nsteps - 1, bh0);
        bool mine = false;
        for (int64_t x = (int64_t)sf.b * p.L + threadIdx.x; x < (int64_t)(sl.b + 1) * p.L; x += B2_GROUP)
          mine |= (pad[x] != 0);
        if (!tc::named_bar_red_or(1, B2_GROUP, mine)) pad = nullptr;
      }
      // the per-row statistic of step n+1 is fetched during step n (a global load in the step's
      // dependency chain cost ~600 cycles per step)
      auto row_stat = [&](const float* src, const Step2& t) -> float {
        const int i = t.it * TT + a;
        return i < p.L ? src[((int64_t)t.b * p.h + t.hh) * p.L + i] : 0.f;
      };
      Step2 s = step2<ROLE>(p, 0, bh0), snext = s;
      float lse_next = row_stat(p.lse, s);
      for (int n = 0; n < nsteps; ++n, s = snext) {
        const uint32_t par = n & 1;
        const int i0 = s.it * TT, j0 = s.jt * TT;
        const int i = i0 + a;
        const bool row_ok = i < p.L;
        const float lse2 = lse_next * LOG2E;
        step_advance<ROLE>(p, snext);
        if (n + 1 < nsteps) lse_next = row_stat(p.lse, snext);
        if (pad) {
          tc::named_bar_sync(1, B2_GROUP);
          if (half == 0) spad[a] = (j0 + a < p.L) ? pad[(int64_t)s.b * p.L + j0 + a] : 1;
          tc::named_bar_sync(1, B2_GROUP);
        }
        const bool diag = (i0 == j0);
        const bool need_mask = diag || (j0 + TT > p.L) || (i0 + TT > p.L) || pad != nullptr;

        if (threadIdx.x == 0) TRACE(0, n, 0);
        tc::mbar_wait(s_full, par);
        tc::tc_fence_after();
        if (threadIdx.x == 0) TRACE(0, n, 1);
        uint32_t pk[32];
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          const int cfirst = 64 * half + 32 * ps;
          // DQ: the G blocks alternate (G_hi of step n is G_lo of step n+1)
          const uint32_t g_lo = tmem + ((ROLE == R_DQ && (n & 1)) ? TM_GHI : TM_GLO);
          const uint32_t g_hi = tmem + ((ROLE == R_DQ && (n & 1)) ? TM_GLO : TM_GHI);
#if MT_SKEW_REGS
          uint32_t Wn[32];
          skew_window_64(g_lo, g_hi, lane_base, 96 - 32 * w4 + cfirst, Wn);
#else
          park64_st64(g_lo, g_hi, lane_base, 96 - 32 * w4 + cfirst, scr);
#endif
          float sv[32];
          {
            uint32_t r[32];
            tc::tmem_ld_32x32(tmem + TM_S + lane_base + cfirst, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int x = 0; x < 32; ++x) sv[x] = __uint_as_float(r[x]);
          }
          if (ps == 1) {        // S and both G blocks are in registers / scratch: release the TMEM columns
            tc::tc_fence_before();
            tc::mbar_arrive(sg_free);
            if (threadIdx.x == 0) TRACE(0, n, 2);
          }
#if MT_SKEW_REGS
          skew_shift_add_32(sv, Wn, lane);
#else
          skew_fetch_add_32(sv, scr, lane);
#endif
          // P = exp(S - lse) with the reference's mask (causal on the diagonal tile, key padding, tails)
#pragma unroll
          for (int x = 0; x < 32; ++x) sv[x] = tc::fast_exp2(fmaf(sv[x], p.scale_log2, -lse2));
          if (need_mask) {
            // branch-free: columns x > lim are masked (causal limit on the diagonal tile, ragged tail,
            // rows beyond L), then the key-padding bytes, four per shared-memory word
            int lim = 31;
            if (diag) lim = min(lim, a - cfirst);
            lim = min(lim, p.L - 1 - j0 - cfirst);
            if (!row_ok) lim = -1;
#pragma unroll
            for (int x = 0; x < 32; ++x) sv[x] = (x > lim) ? 0.f : sv[x];
            if (pad) {
              const uint32_t* sp = reinterpret_cast<const uint32_t*>(spad + cfirst);
#pragma unroll
              for (int x4 = 0; x4 < 8; ++x4) {
                const uint32_t w = sp[x4];
#pragma unroll
                for (int e = 0; e < 4; ++e) sv[4 * x4 + e] = ((w >> (8 * e)) & 0xffu) ? 0.f : sv[4 * x4 + e];
              }
            }
          }
#pragma unroll
          for (int x = 0; x < 16; ++x) pk[16 * ps + x] = pack_bf16x2(sv[2 * x], sv[2 * x + 1]);
        }
        // the hand-off buffer must be free: DKV -- dV MMA of the previous step has read P;
        // DE -- dP of THIS step has consumed the {dO, V} tiles it aliases
        if (threadIdx.x == 0) TRACE(0, n, 3);
        if (ROLE == R_DE) tc::mbar_wait(dp_full, par);
        else if (n > 0) tc::mbar_wait(step_done, (n - 1) & 1);      // DKV: dV MMA has read P; DQ: dS.K has read dS
        if (threadIdx.x == 0) TRACE(0, n, 4);
        if (ROLE == R_DQ) {           // P (bf16 pairs) into the TMEM columns group B turns into dS in place
          tc::tc_fence_after();
          tc::tmem_st_32x32(tmem + TM_PS + lane_base + 32 * half, pk);
          tc::tmem_st_wait();
          tc::tc_fence_before();
        } else {
          uint8_t* ptile = pbuf + half * TILE;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(ptile + swz_chunk(a, c)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          tc::fence_proxy_async();
        }
        tc::mbar_arrive(p_ready);
        if (threadIdx.x == 0) TRACE(0, n, 5);
      }
    } else {
      // ================================ group B: dP, P -> dS / dG =============================
      uint8_t* const dg_base = smem + (ROLE == R_DQ ? Lay2<R_DQ>::DG : Lay2<R_DE>::DG);
      if (ROLE != R_DKV) {
        // dG is zero outside the 128 band columns each row owns; those positions never change
        uint4* z = reinterpret_cast<uint4*>(dg_base);
        for (int x = threadIdx.x - B2_GROUP; x < 4 * TILE / 16; x += B2_GROUP) z[x] = make_uint4(0, 0, 0, 0);
        tc::fence_proxy_async();
        tc::named_bar_sync(2, B2_GROUP);
      }
      const int base_w = ((127 - a) >> 1) + 32 * half;   // first 32-bit word of this thread's band run in dG
      auto row_stat = [&](const float* src, const Step2& t) -> float {
        const int i = t.it * TT + a;
        return i < p.L ? src[((int64_t)t.b * p.h + t.hh) * p.L + i] : 0.f;
      };
      Step2 snext = step2<ROLE>(p, 0, bh0);
      float d_next = row_stat(p.delta, snext);
      for (int n = 0; n < nsteps; ++n) {
        const uint32_t par = n & 1;
        const float Ds = d_next * p.scale;
        step_advance<ROLE>(p, snext);
        if (n + 1 < nsteps) d_next = row_stat(p.delta, snext);
        if (threadIdx.x == B2_GROUP) TRACE(1, n, 0);
        tc::mbar_wait(dp_full, par);
        tc::tc_fence_after();
        if (threadIdx.x == B2_GROUP) TRACE(1, n, 1);
        uint32_t dp[64];
        {
          uint32_t r0[32], r1[32];
          tc::tmem_ld_32x32(tmem + TM_S + lane_base + half * 64, r0);
          tc::tmem_ld_32x32(tmem + TM_S + lane_base + half * 64 + 32, r1);
          tc::tmem_ld_wait();
#pragma unroll
          for (int x = 0; x < 32; ++x) { dp[x] = r0[x]; dp[32 + x] = r1[x]; }
        }
        tc::tc_fence_before();
        tc::mbar_arrive(dp_free);
        if (threadIdx.x == B2_GROUP) TRACE(1, n, 2);
        tc::mbar_wait(p_ready, par);
        if (threadIdx.x == B2_GROUP) TRACE(1, n, 3);
        // dS = P o (dP - D) / sqrt(dh)
        uint32_t A[32];
        if (ROLE == R_DQ) {             // P sits in TMEM (this thread's 32 words of row a)
          tc::tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t w[16];
            tc::tmem_ld_32x16(tmem + TM_PS + lane_base + 32 * half + 16 * c, w);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float d0 = fmaf(__uint_as_float(dp[32 * c + 2 * e]), p.scale, -Ds) * bf16lo(w[e]);
              const float d1 = fmaf(__uint_as_float(dp[32 * c + 2 * e + 1]), p.scale, -Ds) * bf16hi(w[e]);
              A[16 * c + e] = pack_bf16x2(d0, d1);
            }
          }
        } else {
          const uint8_t* ptile = pbuf + half * TILE;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 pw = *reinterpret_cast<const uint4*>(ptile + swz_chunk(a, c));
            const uint32_t w[4] = {pw.x, pw.y, pw.z, pw.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float d0 = fmaf(__uint_as_float(dp[8 * c + 2 * e]), p.scale, -Ds) * bf16lo(w[e]);
              const float d1 = fmaf(__uint_as_float(dp[8 * c + 2 * e + 1]), p.scale, -Ds) * bf16hi(w[e]);
              A[4 * c + e] = pack_bf16x2(d0, d1);
            }
          }
        }
        if (ROLE == R_DQ) {
          // dS replaces P in TMEM (A operand of dS.K); the same values in band coordinates go to the
          // shared-memory dG operand of dG.E_band.  Both were last read by the role MMAs of step n-1,
          // which group A waited for before it stored P of this step.
          tc::tmem_st_32x32(tmem + TM_PS + lane_base + 32 * half, A);
          band_store(dg_base, a, base_w, A);
          tc::tmem_st_wait();
          tc::tc_fence_before();
        } else if (ROLE == R_DKV) {            // rectangular dS: sub-tile `half` of row a
          uint8_t* dstile = smem + Lay2<R_DKV>::DS + half * TILE;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(dstile + swz_chunk(a, c)) = make_uint4(A[4 * c], A[4 * c + 1], A[4 * c + 2], A[4 * c + 3]);
        } else {                        // band dG: this thread's 64 values start at band column 127-a+64*half
          if (threadIdx.x == B2_GROUP) TRACE(1, n, 4);
          if (n > 0) tc::mbar_wait(step_done, (n - 1) & 1);      // the previous step's dE MMAs have read dG
          if (threadIdx.x == B2_GROUP) TRACE(1, n, 5);
          band_store(dg_base, a, base_w, A);
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(ds_ready);
        if (threadIdx.x == B2_GROUP) TRACE(1, n, 6);
      }
    }

    // ---- epilogue: accumulators out of TMEM; group A takes ACC0, group B ACC1; each thread 32 of
    // the 64 columns of its row
    if (!(ROLE == R_DKV && MT_DKV_FUSED16)) {
    tc::mbar_wait(step_done, (nsteps - 1) & 1);
    tc::tc_fence_after();
    uint32_t r[32];
    tc::tmem_ld_32x32(tmem + (grpA ? TM_ACC0 : TM_ACC1) + lane_base + half * 32, r);
    tc::tmem_ld_wait();
    if (ROLE == R_DQ) {            // one accumulator (TM_DQ == TM_ACC0): group A writes the query rows
      const Step2 s = step2<ROLE>(p, 0, bh0);
      const int row = s.it * TT + a;
      if (grpA && row < p.L) {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dq) + (int64_t)s.b * p.sb +
                                              (int64_t)row * p.sl + (int64_t)s.hh * p.sh + half * 32);
#pragma unroll
        for (int x = 0; x < 4; ++x)
          dst[x] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * x]), __uint_as_float(r[8 * x + 1])),
                              pack_bf16x2(__uint_as_float(r[8 * x + 2]), __uint_as_float(r[8 * x + 3])),
                              pack_bf16x2(__uint_as_float(r[8 * x + 4]), __uint_as_float(r[8 * x + 5])),
                              pack_bf16x2(__uint_as_float(r[8 * x + 6]), __uint_as_float(r[8 * x + 7])));
      }
    } else if (ROLE == R_DE) {
      const int c0 = p.max_seq - 1 - (int)blockIdx.z * TT;
      const int erow = (grpA ? c0 - (TT - 1) : c0 + 1) + a;
      if (erow >= 0 && erow < p.max_seq) {
#pragma unroll
        for (int x = 0; x < 32; x += 4)      // vector reductions: every CTA of a diagonal flushes into the same rows
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.dE + (int64_t)erow * DHC + half * 32 + x),
                       "f"(__uint_as_float(r[x])), "f"(__uint_as_float(r[x + 1])), "f"(__uint_as_float(r[x + 2])),
                       "f"(__uint_as_float(r[x + 3])) : "memory");
      }
    } else {
      const Step2 s = step2<ROLE>(p, 0, bh0);
      const int row = s.jt * TT + a;
      uint32_t packed[16];
#pragma unroll
      for (int x = 0; x < 32; x += 2) packed[x / 2] = pack_bf16x2(__uint_as_float(r[x]), __uint_as_float(r[x + 1]));
      if (row < p.L) {
        void* base = grpA ? p.dk : p.dv;
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(base) + (int64_t)s.b * p.sb +
                                              (int64_t)row * p.sl + (int64_t)s.hh * p.sh + half * 32);
#pragma unroll
        for (int x = 0; x < 4; ++x)
          dst[x] = make_uint4(packed[4 * x], packed[4 * x + 1], packed[4 * x + 2], packed[4 * x + 3]);
      }
    }
    }   // (!DKV)
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 17) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

template <int ROLE, bool HF = false>
int launch_role2(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmDO,
                 const CUtensorMap& tmE, const Bwd2Params& p, dim3 grid, cudaStream_t st) {
  auto kern = rga_bwd2_kernel<ROLE, HF>;
  static unsigned long long attr_done = 0; const unsigned long long attr_bit = attr_dev_bit();
  if (!(attr_done & attr_bit)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2_bytes<ROLE>());
    if (e != cudaSuccess) { set_error("rga_bwd2: smem attribute (%d B): %s", smem2_bytes<ROLE>(), cudaGetErrorString(e)); return (int)e; }
    attr_done |= attr_bit;
  }
  Bwd2Params q = p;
  static const bool want_trace = getenv("MT_RGA_TRACE") != nullptr;
  static long long* trace_dev = nullptr;
  const size_t trace_n = 4 * 32 * 8;
  if (want_trace) {
    if (!trace_dev) cudaMalloc(&trace_dev, trace_n * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, trace_n * sizeof(long long), st);
    q.trace = trace_dev;
    q.trace_z = atoi(getenv("MT_RGA_TRACE"));
  }
  kern<<<grid, B2_THREADS, smem2_bytes<ROLE>(), st>>>(tmQ, tmK, tmV, tmDO, tmE, q);
  if (want_trace) {
    static long long host[4 * 32 * 8];
    cudaMemcpyAsync(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    long long t0 = 0;
    for (size_t x = 0; x < trace_n; ++x) if (host[x] && (!t0 || host[x] < t0)) t0 = host[x];
    static const char* agent[4] = {"A", "B", "-", "MMA"};
    for (int ag = 0; ag < 4; ++ag)
      for (int n = 0; n < 32; ++n) {
        bool any = false;
        for (int e = 0; e < 8; ++e) any |= host[(ag * 32 + n) * 8 + e] != 0;
        if (!any) continue;
        fprintf(stderr, "trace role %d %-3s step %2d:", ROLE, agent[ag], n);
        for (int e = 0; e < 8; ++e) fprintf(stderr, " %8lld", host[(ag * 32 + n) * 8 + e] ? host[(ag * 32 + n) * 8 + e] - t0 : -1LL);
        fprintf(stderr, "\n");
      }
  }
  return check_launch("rga_bwd2");
}

Bwd2Params make_params2(const RgaArgs& a) {
  Bwd2Params p;
  p.dk = a.dk; p.dv = a.dv; p.dq = a.dq; p.sb = a.sb; p.sl = a.sl; p.sh = a.sh;
  p.dE = a.dE; p.lse = a.lse; p.delta = a.delta; p.pad = a.pad;
  p.B = a.B; p.h = a.h; p.L = a.L; p.max_seq = a.max_seq;
  p.nT = (a.L + TT - 1) / TT;
  p.scale = 1.f / a.inv_scale_div;
  p.scale_log2 = LOG2E / a.inv_scale_div;
  p.bh_per_cta = 1;
  p.ds_ws = nullptr;
  p.heads_per_cta = 1;
  p.qk_fmt = 1;
  p.gscale = p.inv_gscale = 1.f;
  p.nTri = p.nT * (p.nT + 1) / 2;
  p.trace = nullptr;
  p.trace_z = 0;
  return p;
}

}  // namespace

// dK, dV (key-tile owner walks the query tiles at or below it); ds_ws != NULL: also spill the dS tiles
int rga_bwd2_dkv(const RgaArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                 const CUtensorMap& tmDO, const CUtensorMap& tmE, void* ds_ws, int qk_fmt, float gscale, cudaStream_t st) {
  Bwd2Params p = make_params2(a);
  p.ds_ws = static_cast<uint8_t*>(ds_ws);
  p.qk_fmt = qk_fmt;
  p.gscale = gscale;
  p.inv_gscale = 1.f / gscale;
  // consecutive heads of one (batch row, key tile) share a CTA (same walk over the query tiles; the resident K / V
  // tiles are reloaded and dK / dV flushed at the head boundary): as many as leave at least three CTAs per SM
  static const int hpc_env = getenv("MT_DKV_HPC") ? atoi(getenv("MT_DKV_HPC")) : 0;
  int hpc = 1;
  for (int c = 4; c > 1; c >>= 1)
    if ((int64_t)((a.h + c - 1) / c) * a.B * p.nT >= 3 * (int64_t)sm_count()) { hpc = c; break; }
  if (hpc_env > 0) hpc = hpc_env;
  p.heads_per_cta = hpc > a.h ? a.h : hpc;
  const dim3 grid((a.h + p.heads_per_cta - 1) / p.heads_per_cta, a.B, p.nT);
  if (qk_fmt == 0) return launch_role2<R_DKV, true>(tmQ, tmK, tmV, tmDO, tmE, p, grid, st);
  return launch_role2<R_DKV>(tmQ, tmK, tmV, tmDO, tmE, p, grid, st);
}

// dQ (query-tile owner walks the key tiles at or left of it; P / dS stay in TMEM)
int rga_bwd2_dq(const RgaArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                const CUtensorMap& tmDO, const CUtensorMap& tmE, cudaStream_t st) {
  Bwd2Params p = make_params2(a);
  return launch_role2<R_DQ>(tmQ, tmK, tmV, tmDO, tmE, p, dim3(a.h, a.B, p.nT), st);
}

// dE (tile-diagonal owner walks down the diagonal over a slice of (batch, head))
int rga_bwd2_de(const RgaArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                const CUtensorMap& tmDO, const CUtensorMap& tmE, cudaStream_t st) {
  Bwd2Params p = make_params2(a);
  const int bh = a.B * a.h;
  int slices = (2 * sm_count() + p.nT - 1) / p.nT;       // about two CTAs per SM's worth of slices
  if (slices > bh) slices = bh;
  if (slices < 1) slices = 1;
  p.bh_per_cta = (bh + slices - 1) / slices;
  slices = (bh + p.bh_per_cta - 1) / p.bh_per_cta;
  return launch_role2<R_DE>(tmQ, tmK, tmV, tmDO, tmE, p, dim3(slices, 1, p.nT), st);
}

}  // namespace mt
