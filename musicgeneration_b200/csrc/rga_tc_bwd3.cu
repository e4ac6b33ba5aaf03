// Backward of the fused relative global attention, "dS-spill" variant of the dQ and dE roles (K2).
//
// The recompute variant (rga_tc_bwd2.cu) rebuilds S, G, the skew, P and dP in each of its three
// roles: 18 tile products per (query tile, key tile) pair against 6 algorithmic ones, and three
// times the exponentials.  With 180 GB of HBM there is room to do that work ONCE: the dK/dV role
// writes every dS tile (bf16, the 32 KB shared-memory image of its UMMA operand, so no tensor map
// and no layout change) to a workspace of B*h*nT*(nT+1)/2 tiles, and the two roles in this file
// only consume it:
//     dQ role : dQ  = sum_j dS_ij K_j + dG_ij [E_lo; E_hi]      (3 tile products per step)
//     dE role : dE_band += dG_ij^T Q_i                          (2 tile products per step)
// where dG is dS in band coordinates (dG[a][127-a+b] = dS[a][b], the transpose of the reference's
// skew, MT/layers.py:116-125).  No S, no G, no exp, no row statistics.  Workspace traffic: one write
// + two reads of 32 KB per tile pair (config B: 557 MB per layer each way) against the 4/6 of the
// tensor work and 2/3 of the MUFU work it removes.
//
// The dS tiles stream from global memory THROUGH REGISTERS: the band shift is per-row variable, so
// 16 converter warps own the rows anyway (row a = 32*(w&3)+lane, key columns 32*(w>>2)..+31);
// each thread fetches its 64 bytes of a tile with two 256-bit loads, two tiles ahead of the one it
// is converting (a shared-memory staging ring deep enough to cover the HBM latency does not fit
// next to a double-buffered dG operand; 64 K registers do).  From the registers the values go
//   * into the dG operand in shared memory (band_store; A of dG.E_band, or A of dG^T.Q), and
//   * dQ role: as they are into TMEM, where dS is the A operand of dS.K (no shared-memory copy).
//
//   warps 0-15 : converters
//   warp 16    : TMA loader of the K tile + one new E block (dQ role) or the Q tile (dE role)
//   warp 17    : tcgen05.mma issuer
#include "ops.cuh"
#include "rga_tc_common.cuh"

#include <stdlib.h>

namespace mt {

using namespace rga;

namespace {

enum { L_DQ = 0, L_DE = 1, L_DQE = 2 };

constexpr int CV_THREADS = 512;
constexpr int W_LOAD = CV_THREADS / 32, W_MMA = W_LOAD + 1;
constexpr int B3_THREADS = CV_THREADS + 64;
// fused role: + two idle warps (so that the flushers are warps 20-23: TMEM lane quarter = warp & 3) + 4 flusher warps
constexpr int W_FLUSH = 20, FL_THREADS = 128, B3F_THREADS = (W_FLUSH + 4) * 32;
template <int ROLE> constexpr int threads3() { return ROLE == 2 ? B3F_THREADS : B3_THREADS; }
constexpr int DS_BYTES = 2 * TILE;       // one dS tile image: two [128 x 64] swizzled sub-tiles

template <int ROLE> struct Lay3;
template <> struct Lay3<L_DQ> {      // K x 2; E ring x 3; dG x 2 (4 sub-tiles each)
  static constexpr int X0 = 0, E0 = 2 * TILE, DG = 5 * TILE, BAR = 13 * TILE;
  static constexpr uint32_t TMEM_COLS = 256;       // dQ accumulator (64) | dS slot 0 (64) | dS slot 1 (64)
};
template <> struct Lay3<L_DE> {      // Q x 2; dG x 2
  static constexpr int X0 = 0, E0 = 0, DG = 2 * TILE, BAR = 10 * TILE;
  static constexpr uint32_t TMEM_COLS = 128;       // dE_lo | dE_hi
};
constexpr uint32_t TM3_DS = 64;          // dQ role: dS slot s at columns 64 + 64*s
constexpr uint32_t TM3_ACC1 = 192;       // dQ role: second dQ accumulator (heads alternate between columns 0 and 192)
// Fused role (default): the dQ role that ALSO accumulates dE, so every dS tile is read from the workspace ONCE
// and converted to band coordinates once.  A query-tile owner walks the diagonals d = it, it-1, ... 0; the band of
// step n is E blocks {lo, hi} with lo(n+1) = hi(n), so block hi(n) keeps accumulating through step n+1 and is then
// final: four 64-column TMEM accumulators rotate (lo of step n = slot n % 4, hi = slot (n+1) % 4) and ONE
// [128 x 64] fp32 block per step leaves through a reduction.  Four FLUSHER warps (20-23) do nothing else: they wait
// for the step's products (de_full), read the block out of TMEM, release the slot (de_free: the MMA thread
// restarts it three steps later) and reduce it into dE.  (First version: the converters flushed between two
// tiles -- the two blocking waits on the staging tile and four CTA barriers per step put 3.5 k cycles into the
// converters' loop, which is the critical path: 0.44 ms against 0.38 ms for the two separate launches.)
// The flush is a TMA reduction (cp.reduce.async.bulk.tensor .add.f32 through a 128B-swizzled staging tile, two
// halves of 32 columns): measured on B200 with all 148 SMs flushing back to back (scripts/ubench/red_flush.cu),
// red.global.add.v4.f32 from registers costs 3.7 k cycles of the SM's LSU per 32 KB block (2.6 TB/s chip-wide),
// the bulk reduction 1.8 k cycles with no LSU work beyond the staging stores (5.3 TB/s) -- a step needs 2.1 TB/s.
template <> struct Lay3<L_DQE> {     // K (one slot, released right after dS.K); E ring x 3; Q (resident per head);
                                     // STG: [128 x 32] fp32 staging tile of the dE flush; dG x 2 (4 sub-tiles each)
  static constexpr int X0 = 0, E0 = TILE, QT = 4 * TILE, STG = 5 * TILE, DG = 6 * TILE, BAR = 14 * TILE;
  static constexpr uint32_t TMEM_COLS = 512;       // dQ acc 0 | dS slot 0 | dS slot 1 | dQ acc 1 | dE slot 0 | 1 | 2 | 3
};
constexpr uint32_t TM3_DE = 256;         // fused role: dE accumulator slot s at columns 256 + 64*s
template <int ROLE> constexpr int smem3_bytes() { return Lay3<ROLE>::BAR + 256; }
static_assert(smem3_bytes<L_DQ>() <= 232448 && smem3_bytes<L_DE>() <= 232448 && smem3_bytes<L_DQE>() <= 232448,
              "shared memory budget");

enum { B3_XF = 0, B3_XE = 2, B3_DGR = 4, B3_DGF = 6, B3_DONE = 8, B3_TMEM = 10, B3_ALL = 11, B3_KF = 12, B3_KE = 13,
       B3_DEFULL = 14, B3_DEFREE = 18 };      // [4] each     // B3_DONE: [2] (dQ role: one per accumulator)

struct Bwd3Params {
  const uint8_t* ws;                     // dS tiles: [(b*h+hh)][it*(it+1)/2 + jt][32 KB]
  void* dq; int64_t sb, sl, sh;
  float* dE;
  int B, h, L, max_seq, nT, nTri;
  int bh_per_cta;                        // dE role
  int heads_per_cta;                     // dQ role: consecutive heads of one (batch row, query tile) walked by one CTA
  int qk_fmt;                            // 16-bit format of every MMA operand (1 = bf16; 0 = f16: the dS tiles then hold f16(g * dS))
  float out_scale;                       // 1 / g: applied to dQ / dE on the way out (1 in the bf16 mode)
  long long* trace;                      // MT_RGA_TRACE=z: clock64 stamps of CTA (0,0,z), [2 agents][32 steps][4 events]
  int trace_z;
};

// pipeline timeline of one CTA (debug aid, off unless the launcher passes a buffer)
#define TRACE3(agent, n, ev)                                                                         \
  do {                                                                                               \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (int)blockIdx.z == p.trace_z && (n) < 32)   \
      p.trace[((agent) * 32 + (n)) * 4 + (ev)] = clock64();                                          \
  } while (0)

struct Step3 { int it, jt, b, hh; };

template <int ROLE>
__device__ __forceinline__ int num_steps3(const Bwd3Params& p, int& bh0) {
  bh0 = 0;
  if (ROLE != L_DE) {                                          // it = nT-1-blockIdx.z: longest first
    const int items = min(p.heads_per_cta, p.h - (int)blockIdx.x * p.heads_per_cta);
    return items * (p.nT - (int)blockIdx.z);
  }
  bh0 = (int)blockIdx.x * p.bh_per_cta;
  const int nbh = min(p.bh_per_cta, p.B * p.h - bh0);
  return nbh > 0 ? nbh * (p.nT - (int)blockIdx.z) : 0;
}
template <int ROLE>
__device__ __forceinline__ Step3 step3_first(const Bwd3Params& p, int bh0) {
  Step3 s;
  if (ROLE != L_DE) { s.it = p.nT - 1 - (int)blockIdx.z; s.jt = 0; s.hh = blockIdx.x * p.heads_per_cta; s.b = blockIdx.y; }
  else { s.it = (int)blockIdx.z; s.jt = 0; s.b = bh0 / p.h; s.hh = bh0 % p.h; }
  return s;
}
template <int ROLE>
__device__ __forceinline__ void step3_advance(const Bwd3Params& p, Step3& s) {
  if (ROLE != L_DE) {                                    // next key tile, or the first one of the next head
    if (s.jt < s.it) ++s.jt; else { s.jt = 0; ++s.hh; }
    return;
  }
  if (s.it + 1 < p.nT) { ++s.it; ++s.jt; return; }       // next tile down the diagonal
  s.it = (int)blockIdx.z; s.jt = 0;                      // next (batch, head) of the slice
  if (++s.hh == p.h) { s.hh = 0; ++s.b; }
}
__device__ __forceinline__ const uint8_t* ds_tile(const Bwd3Params& p, const Step3& s) {
  return p.ws + (((int64_t)s.b * p.h + s.hh) * p.nTri + (s.it * (s.it + 1) / 2 + s.jt)) * (int64_t)DS_BYTES;
}

template <int ROLE>
__global__ void __launch_bounds__(threads3<ROLE>(), 1)
rga_bwd3_kernel(const __grid_constant__ CUtensorMap tmX,      // K (dQ and fused roles) or Q (dE role)
                const __grid_constant__ CUtensorMap tmE,
                const __grid_constant__ CUtensorMap tmQ,      // fused role: Q (the other roles pass tmX again)
                const __grid_constant__ CUtensorMap tmDE,     // fused role: dE as a [max_seq, 64] fp32 tensor, box {32, 128}
                const Bwd3Params p) {
  using LY = Lay3<ROLE>;
  constexpr bool HAS_DQ = (ROLE != L_DE);       // dQ accumulators, K + E ring, dS slots in TMEM
  constexpr bool FUSED = (ROLE == L_DQE);       // ... and the rotating dE accumulators
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LY::BAR);
  uint64_t* x_full = bars + B3_XF;        // [2] K + E block / Q
  uint64_t* x_empty = bars + B3_XE;       // [2]
  uint64_t* dg_ready = bars + B3_DGR;     // [2] converters -> MMA : dG operand (and the TMEM dS slot) written
  uint64_t* dg_free = bars + B3_DGF;      // [2] MMA -> converters
  uint64_t* acc_done = bars + B3_DONE;
  uint64_t* de_full = bars + B3_DEFULL;   // fused role [4]: MMA -> flushers, the lo block of a step is final (slot = step % 4)
  uint64_t* de_free = bars + B3_DEFREE;   // fused role [4]: flushers -> MMA, the block has left TMEM
  uint64_t* k_full = bars + B3_KF;        // fused role: the single K slot
  uint64_t* k_empty = bars + B3_KE;       //   (released by the dS.K products, a quarter of the way into the step)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B3_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bh0;
  const int nsteps = num_steps3<ROLE>(p, bh0);

  if (warp == W_LOAD && lane == 0) {
    tc::tma_prefetch_desc(&tmX);
    if (HAS_DQ) tc::tma_prefetch_desc(&tmE);
    if (FUSED) { tc::tma_prefetch_desc(&tmQ); tc::tma_prefetch_desc(&tmDE); }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&x_full[s], 1);
      tc::mbar_init(&x_empty[s], 1);
      tc::mbar_init(&dg_ready[s], CV_THREADS / 32);
      tc::mbar_init(&dg_free[s], 1);
    }
    tc::mbar_init(&acc_done[0], 1);
    tc::mbar_init(&acc_done[1], 1);
    for (int q = 0; q < 4; ++q) { tc::mbar_init(&de_full[q], 1); tc::mbar_init(&de_free[q], FL_THREADS / 32); }
    tc::mbar_init(k_full, 1);
    tc::mbar_init(k_empty, 1);
    tc::fence_barrier_init();
  }
  if (warp == W_MMA) tc::tmem_alloc(tmem_slot, LY::TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (nsteps <= 0) {            // (dE role: empty slice) -- uniform for the whole CTA
    __syncthreads();
    if (warp == W_MMA) tc::tmem_dealloc(tmem, LY::TMEM_COLS);
    return;
  }
  constexpr uint64_t TS16 = TILE >> 4;
  // dQ role: E block m (hi block of step m; block -1 = lo block of step 0) lives in ring slot (m+1) % 3
  auto eslot = [](int m) -> int { return (m + 1) % 3; };

  if (warp == W_LOAD) {
    // ================================ loader ===============================================
    if (lane == 0) {
      Step3 s = step3_first<ROLE>(p, bh0);
      for (int n = 0; n < nsteps; ++n, step3_advance<ROLE>(p, s)) {
        const int st = n & 1;
        tc::mbar_wait(&x_empty[st], ((n >> 1) & 1) ^ 1);
        if (HAS_DQ) {
          const int c0 = p.max_seq - 1 - (s.it - s.jt) * TT;
          const bool first = (s.jt == 0);             // first key tile of a head: its lo block is loaded too
          // ... into the slot that holds the hi block of the previous head's last step: wait for that step as well
          // (fused role: the same wait frees the resident Q tile, last read by that step's dE products)
          if (first && n > 0) tc::mbar_wait(&x_empty[(n - 1) & 1], ((n - 1) >> 1) & 1);
          tc::mbar_arrive_expect_tx(&x_full[st], ((FUSED ? 1 : 2) + (first ? (FUSED ? 2 : 1) : 0)) * TILE);
          if (!FUSED) tc::tma_load_4d(smem + LY::X0 + st * TILE, &tmX, &x_full[st], 0, s.hh, s.jt * TT, s.b);
          tc::tma_load_2d(smem + LY::E0 + eslot(n) * TILE, &tmE, &x_full[st], 0, c0 + 1);
          if (first) tc::tma_load_2d(smem + LY::E0 + eslot(n - 1) * TILE, &tmE, &x_full[st], 0, c0 - (TT - 1));
          if (FUSED) {
            if (first) tc::tma_load_4d(smem + Lay3<L_DQE>::QT, &tmQ, &x_full[st], 0, s.hh, s.it * TT, s.b);
            if (n > 0) tc::mbar_wait(k_empty, (n - 1) & 1);      // dS.K of the previous step has read the K slot
            tc::mbar_arrive_expect_tx(k_full, TILE);
            tc::tma_load_4d(smem + LY::X0, &tmX, k_full, 0, s.hh, s.jt * TT, s.b);
          }
        } else {
          tc::mbar_arrive_expect_tx(&x_full[st], TILE);
          tc::tma_load_4d(smem + LY::X0 + st * TILE, &tmX, &x_full[st], 0, s.hh, s.it * TT, s.b);
        }
        if (n + 2 < nsteps) {         // pull the tile of step n+2 into L2
          Step3 t = s;
          step3_advance<ROLE>(p, t);
          step3_advance<ROLE>(p, t);
          tc::tma_prefetch_4d(&tmX, 0, t.hh, (HAS_DQ ? t.jt : t.it) * TT, t.b);
          if (FUSED && n + 3 < nsteps) {      // the dS tile of step n+3 (the converters fetch one tile ahead, out of L2)
            step3_advance<ROLE>(p, t);
            tc::bulk_prefetch_l2(ds_tile(p, t), DS_BYTES);
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ================================ MMA issuer ============================================
    if (lane == 0) {
      if (HAS_DQ) {
        const uint32_t id_kmn = tc::make_idesc(TT, DHC, p.qk_fmt, p.qk_fmt, 0, 1);    // A K-major (TMEM dS / smem dG), B MN-major (K / E), N = 64
        const uint32_t id_mnmn = tc::make_idesc(TT, DHC, p.qk_fmt, p.qk_fmt, 1, 1);   // fused: A MN-major (dG block), B MN-major (Q), N = 64
        const uint64_t kd_mn0 = tc::make_sdesc(tc::smem_u32(smem + LY::X0), 1024, 1024);
        const uint64_t ed_mn0 = tc::make_sdesc(tc::smem_u32(smem + LY::E0), 1024, 1024);
        const uint64_t dgd0 = tc::make_sdesc(tc::smem_u32(smem + LY::DG), 16, 1024);
        const uint64_t qd_mn = tc::make_sdesc(tc::smem_u32(smem + (FUSED ? Lay3<L_DQE>::QT : 0)), 1024, 1024);
        const uint64_t dg_lo0 = tc::make_sdesc(tc::smem_u32(smem + LY::DG), TILE, 1024);
        const uint64_t dg_hi0 = tc::make_sdesc(tc::smem_u32(smem + LY::DG + 2 * TILE), TILE, 1024);
        uint32_t slot_lo = 0;                             // n % 4
        const int per = p.nT - (int)blockIdx.z;           // key tiles (steps) per head
        int jt = 0, item = 0;
        for (int n = 0; n < nsteps; ++n) {
          const uint64_t st = n & 1;
          const uint32_t par = (n >> 1) & 1;
          const uint32_t acc = tmem + ((item & 1) ? TM3_ACC1 : 0u);      // heads alternate between two accumulators
          tc::mbar_wait(&x_full[st], par);
          if (FUSED) tc::mbar_wait(k_full, n & 1);
          TRACE3(1, n, 0);
          tc::mbar_wait(&dg_ready[st], par);
          tc::tc_fence_after();
          TRACE3(1, n, 1);
          const uint64_t kd_mn = kd_mn0 + (FUSED ? 0 : st * TS16);
#pragma unroll
          for (int k16 = 0; k16 < TT / 16; ++k16)         // dQ += dS . K_j : dS is the TMEM A operand (8 columns per 16 keys)
            tc::umma_f16_ts(acc, tmem + TM3_DS + 64 * (uint32_t)st + 8 * k16, kd_mn + 128 * k16, id_kmn,
                            (jt | k16) != 0);
          if (FUSED) tc::umma_commit(k_empty);
          const uint64_t elo = (uint64_t)eslot(n - 1) * TS16, ehi = (uint64_t)eslot(n) * TS16;
          const uint64_t dgd = dgd0 + st * 4 * TS16;
#pragma unroll
          for (int k16 = 0; k16 < 2 * TT / 16; ++k16)     // dQ += dG . [E_lo; E_hi] (contraction over the band)
            tc::umma_f16(acc, dgd + (uint64_t)(k16 >> 2) * TS16 + 2 * (k16 & 3),
                         ed_mn0 + (k16 < 8 ? elo + 128 * k16 : ehi + 128 * (k16 - 8)), id_kmn, 1);
          if (FUSED) {
            // dE_lo += dG_lo^T . Q (continues the block that was this head's hi block one step ago; fresh on the
            // head's first step), dE_hi = dG_hi^T . Q (fresh); contraction over the 128 query rows
            const uint32_t slot_hi = (slot_lo + 1) & 3u;
            const uint32_t d_lo = tmem + TM3_DE + 64 * slot_lo, d_hi = tmem + TM3_DE + 64 * slot_hi;
            // the hi slot held the lo block of step n-3: the flushers must have read it out
            if (n >= 3) { tc::mbar_wait(&de_free[slot_hi], ((n - 3) >> 2) & 1); tc::tc_fence_after(); }
#pragma unroll
            for (int k16 = 0; k16 < TT / 16; ++k16) {
              tc::umma_f16(d_lo, dg_lo0 + st * 4 * TS16 + 128 * k16, qd_mn + 128 * k16, id_mnmn, (jt | k16) != 0);
              tc::umma_f16(d_hi, dg_hi0 + st * 4 * TS16 + 128 * k16, qd_mn + 128 * k16, id_mnmn, k16 != 0);
            }
            tc::umma_commit(&de_full[slot_lo]);
            slot_lo = slot_hi;
          }
          tc::umma_commit(&dg_free[st]);
          tc::umma_commit(&x_empty[st]);
          if (++jt == per) { tc::umma_commit(&acc_done[item & 1]); jt = 0; ++item; }
          TRACE3(1, n, 2);
        }
      } else {
        const uint32_t id_mnmn = tc::make_idesc(TT, DHC, p.qk_fmt, p.qk_fmt, 1, 1);   // A MN-major (dG block), B MN-major (Q), N = 64
        const uint64_t qd_mn0 = tc::make_sdesc(tc::smem_u32(smem + LY::X0), 1024, 1024);
        const uint64_t dg_lo0 = tc::make_sdesc(tc::smem_u32(smem + LY::DG), TILE, 1024);
        const uint64_t dg_hi0 = tc::make_sdesc(tc::smem_u32(smem + LY::DG + 2 * TILE), TILE, 1024);
        for (int n = 0; n < nsteps; ++n) {
          const uint64_t st = n & 1;
          const uint32_t par = (n >> 1) & 1;
          tc::mbar_wait(&x_full[st], par);
          TRACE3(1, n, 0);
          tc::mbar_wait(&dg_ready[st], par);
          tc::tc_fence_after();
          TRACE3(1, n, 1);
#pragma unroll
          for (int k16 = 0; k16 < TT / 16; ++k16) {       // dE_blk += dG_blk^T . Q (contraction over the query rows)
            tc::umma_f16(tmem, dg_lo0 + st * 4 * TS16 + 128 * k16, qd_mn0 + st * TS16 + 128 * k16, id_mnmn, (n | k16) != 0);
            tc::umma_f16(tmem + 64, dg_hi0 + st * 4 * TS16 + 128 * k16, qd_mn0 + st * TS16 + 128 * k16, id_mnmn, (n | k16) != 0);
          }
          tc::umma_commit(&dg_free[st]);
          tc::umma_commit(&x_empty[st]);
          TRACE3(1, n, 2);
        }
      }
      if (ROLE == L_DE) tc::umma_commit(&acc_done[0]);
    }
  } else if (FUSED && warp >= W_FLUSH) {
    // ================================ flushers (fused role): final dE blocks -> dE ===================
    // warp f = warp - 20 reads TMEM lanes 32f..32f+31 (block row a); per block two halves of 32 columns go through
    // the 128B-swizzled staging tile [128 rows x 32 fp32] and leave as TMA reductions (rows outside [0, max_seq)
    // are clipped by the tensor map).  The slot is released as soon as both halves are in registers.
    const int a = (warp - W_FLUSH) * 32 + lane, ftid = threadIdx.x - W_FLUSH * 32;
    const uint32_t lane_base = (uint32_t)((warp - W_FLUSH) * 32) << 16;
    uint8_t* const stg = smem + Lay3<L_DQE>::STG;
    const int fper = p.nT - (int)blockIdx.z;              // steps per head = it + 1
    const float osc = p.out_scale;
    int fjt = 0;
    for (int m = 0; m < nsteps; ++m) {
      const int slot = m & 3;
      const int d = (fper - 1) - fjt;                     // diagonal it - jt of step m
      if (++fjt == fper) fjt = 0;
      const int erow0 = p.max_seq - TT * (d + 1);         // first E row of the lo block: c0 - 127, c0 = max_seq - 1 - 128 d
      tc::mbar_wait(&de_full[slot], (m >> 2) & 1);
      tc::tc_fence_after();
      uint32_t r0[32], r1[32];
      tc::tmem_ld_32x32(tmem + TM3_DE + 64 * (uint32_t)slot + lane_base, r0);
      tc::tmem_ld_32x32(tmem + TM3_DE + 64 * (uint32_t)slot + lane_base + 32, r1);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      tc::mbar_arrive_warp(&de_free[slot]);
      if (erow0 >= p.max_seq || erow0 + TT <= 0) continue;      // (uniform) nothing of the block exists
      if (erow0 < 0) {
        // max_seq is not a multiple of the tile edge and this is the block that straddles E row 0 (once per head, on
        // the farthest diagonal): a bulk-tensor reduction at a negative row coordinate faults on sm_100a (illegal
        // instruction, found with L = max_seq = 64), so these rows leave through vector reductions instead
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
        continue;
      }
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (ftid == 0) tc::bulk_wait_read0();             // the previous reduction has read the staging tile
        tc::named_bar_sync(3, FL_THREADS);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t* r = hf ? r1 : r0;
          *reinterpret_cast<float4*>(stg + swz_chunk(a, c)) =
              make_float4(__uint_as_float(r[4 * c]) * osc, __uint_as_float(r[4 * c + 1]) * osc,
                          __uint_as_float(r[4 * c + 2]) * osc, __uint_as_float(r[4 * c + 3]) * osc);
        }
        tc::fence_proxy_async();
        tc::named_bar_sync(3, FL_THREADS);
        if (ftid == 0) {
          tc::tma_reduce_add_2d(&tmDE, stg, 32 * hf, erow0);
          tc::bulk_commit();
        }
      }
    }
    if (ftid == 0) tc::bulk_wait0();
  } else if (warp < CV_THREADS / 32) {
    // ================================ converters: dS tile -> registers -> band dG (+ TMEM dS) =====
    const int w4 = warp & 3, q4 = warp >> 2;             // quarter q4: key columns 32*q4 .. +31
    const int a = w4 * 32 + lane;
    const int a7 = a & 7;
    const uint32_t lane_base = (uint32_t)(w4 * 32) << 16;
    {
      // dG is zero outside the 128 band columns each row owns; those positions never change
      uint4* z = reinterpret_cast<uint4*>(smem + LY::DG);
      for (int x = threadIdx.x; x < 2 * 4 * TILE / 16; x += CV_THREADS) z[x] = make_uint4(0, 0, 0, 0);
      tc::fence_proxy_async();
      tc::named_bar_sync(1, CV_THREADS);
    }
    const int base_w = ((127 - a) >> 1) + 16 * q4;       // first 32-bit word of this thread's band run in dG
    // This thread's 32 values of a tile are logical 16-byte chunks 4*(q4&1)..+3 of row a in sub-tile q4>>1;
    // the 128B swizzle puts them at physical chunks (4*(q4&1)+j) ^ a7: one aligned 64-byte run, fetched
    // with two 256-bit loads and put back into logical order with two conditional swaps.
    const int64_t my_off = (int64_t)(q4 >> 1) * TILE + a * 128 + (((q4 & 1) ^ (a7 >> 2)) << 6);
    Step3 sf = step3_first<ROLE>(p, bh0);                 // fetch cursor (runs two tiles ahead)
    int nf = 0;
    auto fetch = [&](uint32_t (&R)[16]) {
      if (nf < nsteps) {
        const uint8_t* src = ds_tile(p, sf) + my_off;
        tc::ldg256_stream(src, R[0], R[1], R[2], R[3], R[4], R[5], R[6], R[7]);
        tc::ldg256_stream(src + 32, R[8], R[9], R[10], R[11], R[12], R[13], R[14], R[15]);
        step3_advance<ROLE>(p, sf);
      }
      ++nf;
    };
    auto process = [&](const uint32_t (&R)[16], int n) {
      uint32_t A[16];
      if (threadIdx.x == 0) TRACE3(0, n, 0);
      {
        uint32_t T[16];
        const bool s1 = a7 & 1, s2 = a7 & 2;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 4; ++e) T[4 * c + e] = s1 ? R[4 * (c ^ 1) + e] : R[4 * c + e];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 4; ++e) A[4 * c + e] = s2 ? T[4 * (c ^ 2) + e] : T[4 * c + e];
      }
      const int st = n & 1;
      if (threadIdx.x == 0) TRACE3(0, n, 1);                          // (registers of the tile have arrived)
      if (n >= 2) tc::mbar_wait(&dg_free[st], ((n >> 1) - 1) & 1);   // MMAs of step n-2 have read dG / the dS slot
      if (threadIdx.x == 0) TRACE3(0, n, 2);
      if (HAS_DQ) tc::tc_fence_after();
      if (HAS_DQ) tc::tmem_st_32x16(tmem + TM3_DS + 64 * st + lane_base + 16 * q4, A);
      band_store_n<16>(smem + LY::DG + st * 4 * TILE, a, base_w, A);
      if (HAS_DQ) {
        tc::tmem_st_wait();
        tc::tc_fence_before();
      }
      tc::fence_proxy_async();
      tc::mbar_arrive_warp(&dg_ready[st]);
      if (threadIdx.x == 0) TRACE3(0, n, 3);
    };
    // dQ role: the accumulator of head `item` (columns 0 or TM3_ACC1) -> dq.  Called one step into the next head
    // (the products of the head's last step have finished by then; the accumulators alternate, so the next
    // head's products do not touch it) and once after the loop.
    const int per = HAS_DQ ? p.nT - (int)blockIdx.z : 1;
    auto store_dq = [&](int item) {
      tc::mbar_wait(&acc_done[item & 1], (item >> 1) & 1);
      tc::tc_fence_after();
      uint32_t r[16];
      tc::tmem_ld_32x16(tmem + ((item & 1) ? TM3_ACC1 : 0u) + lane_base + q4 * 16, r);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      const Step3 s = step3_first<ROLE>(p, bh0);
      const int row = s.it * TT + a;
      const float osc = p.out_scale;          // (f16 mode: the dS tiles carry the loss scale)
      if (row < p.L) {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dq) + (int64_t)s.b * p.sb +
                                              (int64_t)row * p.sl + (int64_t)(s.hh + item) * p.sh + q4 * 16);
#pragma unroll
        for (int x = 0; x < 2; ++x)
          dst[x] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * x]) * osc, __uint_as_float(r[8 * x + 1]) * osc),
                              pack_bf16x2(__uint_as_float(r[8 * x + 2]) * osc, __uint_as_float(r[8 * x + 3]) * osc),
                              pack_bf16x2(__uint_as_float(r[8 * x + 4]) * osc, __uint_as_float(r[8 * x + 5]) * osc),
                              pack_bf16x2(__uint_as_float(r[8 * x + 6]) * osc, __uint_as_float(r[8 * x + 7]) * osc));
      }
    };
    int cjt = 0, citem = 0;                // converter-side position inside the head
    auto after = [&]() {
      if (!HAS_DQ) return;
      if (cjt == 0 && citem > 0) store_dq(citem - 1);     // first step of a new head is converted: flush the previous one
      if (++cjt == per) { cjt = 0; ++citem; }
    };
    if (FUSED) {
      // one tile ahead (a step is ~2 us, the fetch is issued a whole step before its tile is converted); the third
      // register buffer of the other roles is what the flush block F needs
      uint32_t R0[16], R1[16];
      fetch(R0);
      for (int n = 0; n < nsteps; n += 2) {
        fetch(R1);
        process(R0, n); after();
        if (n + 1 < nsteps) { fetch(R0); process(R1, n + 1); after(); }
      }
    } else {
      uint32_t R0[16], R1[16], R2[16];
      fetch(R0);
      fetch(R1);
      for (int n = 0; n < nsteps; n += 3) {
        fetch(R2);
        process(R0, n); after();
        if (n + 1 < nsteps) { fetch(R0); process(R1, n + 1); after(); }
        if (n + 2 < nsteps) { fetch(R1); process(R2, n + 2); after(); }
      }
    }

    // ---- epilogue
    if (HAS_DQ) {
      store_dq(citem - 1);        // (the loop ends on the last step of a head: citem = number of heads walked)
    } else {
      tc::mbar_wait(&acc_done[0], 0);
      tc::tc_fence_after();
    }
    if (HAS_DQ) {
    } else {                      // two blocks of 128 E rows x 64: quarters 0,1 the lo block, 2,3 the hi block
      const int c0 = p.max_seq - 1 - (int)blockIdx.z * TT;
      const int erow = ((q4 >> 1) == 0 ? c0 - (TT - 1) : c0 + 1) + a;
      uint32_t r[32];
      tc::tmem_ld_32x32(tmem + lane_base + 32 * q4, r);
      tc::tmem_ld_wait();
      if (erow >= 0 && erow < p.max_seq) {
        // one flush per CTA, but every CTA of a diagonal hits the same rows: vector reductions (8 instead of 32
        // L2 operations per thread; dE rows are 256-byte aligned: mt_rga_bwd checks the base)
        float* dst = p.dE + (int64_t)erow * DHC + 32 * (q4 & 1);
        const float osc = p.out_scale;
#pragma unroll
        for (int x = 0; x < 32; x += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + x), "f"(__uint_as_float(r[x]) * osc),
                       "f"(__uint_as_float(r[x + 1]) * osc), "f"(__uint_as_float(r[x + 2]) * osc),
                       "f"(__uint_as_float(r[x + 3]) * osc) : "memory");
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == W_MMA) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, LY::TMEM_COLS);
  }
}

template <int ROLE>
int launch_role3(const CUtensorMap& tmX, const CUtensorMap& tmE, const CUtensorMap& tmQ, const CUtensorMap& tmDE,
                 const Bwd3Params& p, dim3 grid, cudaStream_t st) {
  auto kern = rga_bwd3_kernel<ROLE>;
  static unsigned long long attr_done = 0; const unsigned long long attr_bit = attr_dev_bit();
  if (!(attr_done & attr_bit)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3_bytes<ROLE>());
    if (e != cudaSuccess) { set_error("rga_bwd3: smem attribute (%d B): %s", smem3_bytes<ROLE>(), cudaGetErrorString(e)); return (int)e; }
    attr_done |= attr_bit;
  }
  Bwd3Params q = p;
  static const bool want_trace = getenv("MT_RGA_TRACE") != nullptr;
  static long long* trace_dev = nullptr;
  const size_t trace_n = 2 * 32 * 4;
  if (want_trace) {
    if (!trace_dev) cudaMalloc(&trace_dev, trace_n * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, trace_n * sizeof(long long), st);
    q.trace = trace_dev;
    q.trace_z = atoi(getenv("MT_RGA_TRACE"));
  }
  kern<<<grid, threads3<ROLE>(), smem3_bytes<ROLE>(), st>>>(tmX, tmE, tmQ, tmDE, q);
  if (want_trace) {
    static long long host[2 * 32 * 4];
    cudaMemcpyAsync(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    long long t0 = 0;
    for (size_t x = 0; x < trace_n; ++x) if (host[x] && (!t0 || host[x] < t0)) t0 = host[x];
    static const char* agent[2] = {"CV", "MMA"};
    for (int ag = 0; ag < 2; ++ag)
      for (int n = 0; n < 32; ++n) {
        bool any = false;
        for (int e = 0; e < 4; ++e) any |= host[(ag * 32 + n) * 4 + e] != 0;
        if (!any) continue;
        fprintf(stderr, "trace3 role %d %-3s step %2d:", ROLE, agent[ag], n);
        for (int e = 0; e < 4; ++e) fprintf(stderr, " %8lld", host[(ag * 32 + n) * 4 + e] ? host[(ag * 32 + n) * 4 + e] - t0 : -1LL);
        fprintf(stderr, "\n");
      }
  }
  return check_launch("rga_bwd3");
}

Bwd3Params make_params3(const RgaArgs& a, const void* ws) {
  Bwd3Params p;
  p.ws = static_cast<const uint8_t*>(ws);
  p.dq = a.dq; p.sb = a.sb; p.sl = a.sl; p.sh = a.sh; p.dE = a.dE;
  p.B = a.B; p.h = a.h; p.L = a.L; p.max_seq = a.max_seq;
  p.nT = (a.L + TT - 1) / TT;
  p.nTri = p.nT * (p.nT + 1) / 2;
  p.bh_per_cta = 1;
  p.heads_per_cta = 1;
  p.qk_fmt = 1;
  p.out_scale = 1.f;
  p.trace = nullptr;
  p.trace_z = 0;
  return p;
}

}  // namespace

size_t rga_bwd3_workspace_bytes(int64_t B, int64_t h, int64_t L) {
  const int64_t nT = (L + TT - 1) / TT;
  return (size_t)(B * h * (nT * (nT + 1) / 2)) * DS_BYTES;
}

// dQ from the spilled dS tiles (query-tile owner walks the key tiles at or left of it)
int rga_bwd3_dq(const RgaArgs& a, const void* ws, const CUtensorMap& tmK, const CUtensorMap& tmE, int qk_fmt, float gscale, cudaStream_t st) {
  Bwd3Params p = make_params3(a, ws);
  p.qk_fmt = qk_fmt;
  p.out_scale = 1.f / gscale;
  // consecutive heads of one (batch row, query tile) share a CTA (same number of key tiles, same E blocks, two
  // alternating accumulators): as many as leave at least three CTAs per SM
  static const int hpc_env = getenv("MT_DQ_HPC") ? atoi(getenv("MT_DQ_HPC")) : 0;
  int hpc = 1;
  for (int c = 4; c > 1; c >>= 1)
    if ((int64_t)((a.h + c - 1) / c) * a.B * p.nT >= 3 * (int64_t)sm_count()) { hpc = c; break; }
  if (hpc_env > 0) hpc = hpc_env;
  p.heads_per_cta = hpc > a.h ? a.h : hpc;
  return launch_role3<L_DQ>(tmK, tmE, tmK, tmK, p, dim3((a.h + p.heads_per_cta - 1) / p.heads_per_cta, a.B, p.nT), st);
}

// dQ AND dE from the spilled dS tiles in one pass over the workspace (the fused role at the top of the file)
int rga_bwd3_dqe(const RgaArgs& a, const void* ws, const CUtensorMap& tmK, const CUtensorMap& tmE, const CUtensorMap& tmQ,
                 int qk_fmt, float gscale, cudaStream_t st) {
  Bwd3Params p = make_params3(a, ws);
  p.qk_fmt = qk_fmt;
  p.out_scale = 1.f / gscale;
  CUtensorMap tmDE;
  int rc;
  if ((rc = tc::make_tmap_2d_f32(&tmDE, a.dE, a.max_seq, DHC, DHC, 32, TT))) return rc;
  static const int hpc_env = getenv("MT_DQ_HPC") ? atoi(getenv("MT_DQ_HPC")) : 0;
  int hpc = 1;
  for (int c = 4; c > 1; c >>= 1)
    if ((int64_t)((a.h + c - 1) / c) * a.B * p.nT >= 3 * (int64_t)sm_count()) { hpc = c; break; }
  if (hpc_env > 0) hpc = hpc_env;
  p.heads_per_cta = hpc > a.h ? a.h : hpc;
  return launch_role3<L_DQE>(tmK, tmE, tmQ, tmDE, p, dim3((a.h + p.heads_per_cta - 1) / p.heads_per_cta, a.B, p.nT), st);
}

// dE from the spilled dS tiles (tile-diagonal owner walks down the diagonal over a slice of (batch, head))
int rga_bwd3_de(const RgaArgs& a, const void* ws, const CUtensorMap& tmQ, const CUtensorMap& tmE, int qk_fmt, float gscale, cudaStream_t st) {
  Bwd3Params p = make_params3(a, ws);
  p.qk_fmt = qk_fmt;
  p.out_scale = 1.f / gscale;
  const int bh = a.B * a.h;
  int slices = (2 * sm_count() + p.nT - 1) / p.nT;       // about two CTAs per SM's worth of slices
  if (slices > bh) slices = bh;
  if (slices < 1) slices = 1;
  p.bh_per_cta = (bh + slices - 1) / slices;
  slices = (bh + p.bh_per_cta - 1) / p.bh_per_cta;
  return launch_role3<L_DE>(tmQ, tmE, tmQ, tmQ, p, dim3(slices, 1, p.nT), st);
}

}  // namespace mt
