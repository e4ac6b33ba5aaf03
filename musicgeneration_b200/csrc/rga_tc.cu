// Fused relative global attention forward on tcgen05 / TMEM / TMA (K1), head dim 64, bf16 or f16.
//
// One CTA = one 128-row query tile of one (batch, head); it walks the key tiles j0 = 0, 128, ...
// (up to the diagonal when causal).  Per key tile, on the tensor cores:
//     S      = Q . K_j^T                     (128 x 128, fp32 in TMEM)
//     G_hi   = Q . E[c0+1 .. c0+128]^T       (the ONE new block of relative-embedding rows;
//                                             c0 = max_seq-1-(i0-j0); the block [c0-127 .. c0]
//                                             is still in TMEM from the previous key tile, so the
//                                             executed relative FLOPs equal the Q.K^T FLOPs)
//     O     += P . V_j                       (P written to shared memory in the UMMA K-major
//                                             128B-swizzled layout by the softmax warps)
// The reference's skew (MT/layers.py:116-125) is index arithmetic on the band [G_lo | G_hi]
// (rga_tc_common.cuh).  E rows >= max_seq (exactly the j > i positions) and rows < 0 are
// zero-filled by TMA, which realises the reference's _qe_masking.  Online softmax in fp32 with
// exp2; LSE saved for the backward.
//
// Warps 0-15: softmax / correction / epilogue (row a = 32*(w&3)+lane, key columns 32*(w>>2)..+31),
// warp 16: TMA producer, warp 17: TMEM allocator + MMA issuer.
#include "ops.cuh"
#include "rga_tc_common.cuh"

#include <stdlib.h>

namespace mt {

using namespace rga;

namespace {

// 16 softmax warps: warp w works on TMEM lanes 32*(w&3)..+31 (query row a = 32*(w&3)+lane) and on
// the key-column quarter qt = w>>2, i.e. 32 logits per thread and step.  Four warps per scheduler
// hide each other's TMEM / shared-memory / MUFU latency; the per-thread state stays below the
// 112-register budget of a 576-thread CTA.
// 1: the skew runs through the register barrel shifter (rga_tc_common.cuh) instead of the per-thread scratch
#ifndef MT_FWD_SKEW_REGS
#define MT_FWD_SKEW_REGS 0
#endif
constexpr int FW_MATH_WARPS = 16;
constexpr int FW_MATH_THREADS = FW_MATH_WARPS * 32;
constexpr int FW_THREADS = FW_MATH_THREADS + 64;     // + TMA producer warp + MMA issuer warp
constexpr int FW_SCR_BYTES = FW_MATH_THREADS * SCR32_WORDS * 4;

// shared memory map (offsets from the 1024-aligned base)
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE;            // 2 stages
constexpr int OFF_V = OFF_K + 2 * TILE;        // 2 stages
constexpr int OFF_E = OFF_V + 2 * TILE;        // 2 stages: the new "hi" block of each step
constexpr int OFF_ELO = OFF_E + 2 * TILE;      // "lo" block of the first step only
constexpr int OFF_SCR = OFF_ELO + TILE;
constexpr int OFF_XCH = OFF_SCR + FW_SCR_BYTES;   // [2 step parities + 1 epilogue][4 quarters][128] floats: row max / row sum exchange
constexpr int OFF_BAR = OFF_XCH + 3 * 4 * TT * 4;
constexpr int FWD_SMEM = OFF_BAR + 256 + 1024;
static_assert(FWD_SMEM <= 232448, "forward kernel exceeds the 227 KB shared-memory limit");

// TMEM columns
constexpr uint32_t TM_S = 0, TM_G0 = 128, TM_G1 = 256, TM_O = 384, TM_P = 448;   // P: 128 x 128 16-bit = 64 columns

// O is rescaled only when the running row maximum grows by more than 2^RESCALE_LOG2: P stays
// below 2^8 (exact in the fp32 row sums, same relative precision in the 16-bit operand)
constexpr float RESCALE_LOG2 = 8.f;

struct FwdParams {
  void* O; int64_t ob, ol, oh;
  float* lse;
  const uint8_t* pad;
  int B, h, L, max_seq, causal, fmt;
  int ofmt;             // 16-bit format of O (1 = bf16, 0 = f16); differs from fmt in the mixed mode (f16 q/k/v/E, bf16 O)
  int heads_per_cta;    // consecutive heads walked by one CTA (same batch row and query tile): halves the per-CTA fixed cost
  float scale_log2;     // log2(e) / sqrt(dh)
  uint8_t* stash;       // training: every P tile (16-bit, relative to the row reference of its step) is kept
  float* mrow;          //           for the backward (rga_tc_bwd4.cu; thread-major layout), with the row references [tile][128] (log2 domain)
  int nTri;             //           tiles per (batch, head): nT (nT + 1) / 2, tile (it, jt) at it (it + 1) / 2 + jt
  long long* trace;     // MT_RGA_TRACE=z: clock64 stamps of CTA (0,0,z), [2 agents][32 steps][8 events]
  int trace_z;
};

// pipeline timeline of one CTA (debug aid, off unless the launcher passes a buffer)
#define FTRACE(agent, n, ev)                                                                        \
  do {                                                                                              \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (int)blockIdx.z == p.trace_z && (n) < 32)  \
      p.trace[((agent) * 32 + (n)) * 8 + (ev)] = clock64();                                         \
  } while (0)

__global__ void __launch_bounds__(FW_THREADS, 1)
rga_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmE,
                  const FwdParams p) {
  // declared (and checked) 1024-byte aligned so that the compiler keeps the shared address space
  // (LDS/STS instead of generic LD/ST) and the 128B-swizzle atoms line up
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* bar_q = bars + 0;
  uint64_t* kv_full = bars + 1;      // [2]
  uint64_t* kv_empty = bars + 3;     // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_consumed = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  uint8_t* spad = reinterpret_cast<uint8_t*>(bars + 10);     // [128] pad flags of the key tile
  uint64_t* q_free = bars + 26;      // MMA -> producer: the S / G products of a head are done with Q (and ELO)
  uint64_t* o_free = bars + 27;      // softmax -> MMA: O of the previous head has been read out
  float* xch = reinterpret_cast<float*>(smem + OFF_XCH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid = (ceil(h / heads_per_cta), B, nT): the query-tile index is the slowest dimension, so over the whole
  // launch the CTAs with the most key tiles are dispatched first
  const int b = blockIdx.y, hh0 = blockIdx.x * p.heads_per_cta;
  const int n_items = min(p.heads_per_cta, p.h - hh0);
  const int i0 = (gridDim.z - 1 - blockIdx.z) * TT;
  const int L = p.L;
  const int n_kt = p.causal ? (i0 / TT + 1) : (L + TT - 1) / TT;
  const int n_g = n_items * n_kt;            // global steps: head `item` = g / n_kt, key tile jt = g % n_kt

  if (warp == FW_MATH_WARPS && lane == 0) {
    tc::tma_prefetch_desc(&tmQ);
    tc::tma_prefetch_desc(&tmK);
    tc::tma_prefetch_desc(&tmV);
    tc::tma_prefetch_desc(&tmE);
    tc::tma_prefetch_4d(&tmQ, 0, hh0, i0, b);      // the CTA's first tiles start towards L2 under the prologue
    tc::tma_prefetch_4d(&tmK, 0, hh0, 0, b);
    tc::tma_prefetch_4d(&tmV, 0, hh0, 0, b);
    tc::mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&kv_full[s], 1);
      tc::mbar_init(&kv_empty[s], 1);
    }
    tc::mbar_init(s_full, 1);
    tc::mbar_init(s_consumed, FW_MATH_THREADS);
    tc::mbar_init(p_full, FW_MATH_THREADS);
    tc::mbar_init(o_done, 1);
    tc::mbar_init(q_free, 1);
    tc::mbar_init(o_free, FW_MATH_THREADS);
    tc::fence_barrier_init();
  }
  if (warp == FW_MATH_WARPS + 1) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == FW_MATH_WARPS) {
    // ================================ TMA producer ==========================================
    if (lane == 0) {
      int item = 0, jt = 0;
      for (int g = 0; g < n_g; ++g) {
        const int hh = hh0 + item;
        if (jt == 0) {
          // Q (and the ELO block) of the previous head: its S / G products must be done with them
          if (item > 0) tc::mbar_wait(q_free, (item - 1) & 1);
          tc::mbar_arrive_expect_tx(bar_q, TILE);
          tc::tma_load_4d(smem + OFF_Q, &tmQ, bar_q, 0, hh, i0, b);
        }
        const int s = g & 1;
        const int j0 = jt * TT;
        const int c0 = p.max_seq - 1 - (i0 - j0);
        tc::mbar_wait(&kv_empty[s], ((g >> 1) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&kv_full[s], (jt == 0 ? 4 : 3) * TILE);
        tc::tma_load_4d(smem + OFF_K + s * TILE, &tmK, &kv_full[s], 0, hh, j0, b);
        tc::tma_load_2d(smem + OFF_E + s * TILE, &tmE, &kv_full[s], 0, c0 + 1);
        if (jt == 0) tc::tma_load_2d(smem + OFF_ELO, &tmE, &kv_full[s], 0, c0 - (TT - 1));
        tc::tma_load_4d(smem + OFF_V + s * TILE, &tmV, &kv_full[s], 0, hh, j0, b);
        if (++jt == n_kt) { jt = 0; ++item; }
      }
    }
  } else if (warp == FW_MATH_WARPS + 1) {
    // ================================ MMA issuer ============================================
    if (lane == 0) {
      const uint32_t idesc_s = tc::make_idesc(TT, TT, p.fmt, p.fmt, 0, 0);     // S, G: K-major x K-major
      const uint32_t idesc_o = tc::make_idesc(TT, DHC, p.fmt, p.fmt, 0, 1);    // O: P K-major, V MN-major
      // descriptors are built once; a k-step inside the loops is an add on the 16-byte address field
      const uint64_t qd0 = tc::make_sdesc(tc::smem_u32(smem + OFF_Q), 16, 1024);
      const uint64_t kd0 = tc::make_sdesc(tc::smem_u32(smem + OFF_K), 16, 1024);
      const uint64_t ed0 = tc::make_sdesc(tc::smem_u32(smem + OFF_E), 16, 1024);
      const uint64_t elod = tc::make_sdesc(tc::smem_u32(smem + OFF_ELO), 16, 1024);
      const uint64_t vd0 = tc::make_sdesc(tc::smem_u32(smem + OFF_V), 1024, 1024);
      // S and the new G block of global step g (key tile jt of its head); the first step of a head also
      // computes the lo block from ELO
      auto issue_s = [&](int g, int jt, bool last_of_head) {
        const uint64_t so = (uint64_t)(g & 1) * (TILE >> 4);
        const uint32_t g_hi = tmem + (((jt + 1) & 1) ? TM_G1 : TM_G0);
#pragma unroll
        for (int k4 = 0; k4 < DHC / 16; ++k4) {
          tc::umma_f16(tmem + TM_S, qd0 + 2 * k4, kd0 + so + 2 * k4, idesc_s, k4 != 0);
          tc::umma_f16(g_hi, qd0 + 2 * k4, ed0 + so + 2 * k4, idesc_s, k4 != 0);
        }
        if (jt == 0) {
#pragma unroll
          for (int k4 = 0; k4 < DHC / 16; ++k4)
            tc::umma_f16(tmem + TM_G0, qd0 + 2 * k4, elod + 2 * k4, idesc_s, k4 != 0);
        }
        tc::umma_commit(s_full);
        if (last_of_head) tc::umma_commit(q_free);        // Q / ELO may take the next head's tiles
      };
      tc::mbar_wait(bar_q, 0);
      tc::mbar_wait(&kv_full[0], 0);
      tc::tc_fence_after();
      issue_s(0, 0, n_kt == 1);
      int item = 0, jt = 0;
      for (int g = 0; g < n_g; ++g) {
        if (g + 1 < n_g) {
          const int jn = (jt + 1 == n_kt) ? 0 : jt + 1;
          tc::mbar_wait(s_consumed, g & 1);
          FTRACE(1, g, 0);
          if (jn == 0) tc::mbar_wait(bar_q, (item + 1) & 1);       // the next head's Q
          tc::mbar_wait(&kv_full[(g + 1) & 1], ((g + 1) >> 1) & 1);
          tc::tc_fence_after();
          FTRACE(1, g, 1);
          issue_s(g + 1, jn, jn + 1 == n_kt);
          FTRACE(1, g, 2);
        }
        tc::mbar_wait(p_full, g & 1);
        if (jt == 0 && item > 0) tc::mbar_wait(o_free, (item - 1) & 1);   // O of the previous head has been read out
        tc::tc_fence_after();
        FTRACE(1, g, 3);
        const uint64_t vd = vd0 + (uint64_t)(g & 1) * (TILE >> 4);
#pragma unroll
        for (int k8 = 0; k8 < TT / 16; ++k8)     // P stays in TMEM (A operand): 16 keys = 8 columns; V: 16 key rows = 2048 B
          tc::umma_f16_ts(tmem + TM_O, tmem + TM_P + 8 * k8, vd + 128 * k8, idesc_o, (jt | k8) != 0);
        tc::umma_commit(&kv_empty[g & 1]);
        tc::umma_commit(o_done);
        FTRACE(1, g, 4);
        if (++jt == n_kt) { jt = 0; ++item; }
      }
    }
  } else {
    // ================================ softmax warps =========================================
    const int w4 = warp & 3, qt = warp >> 2;
    const int a = w4 * 32 + lane;                      // query row within the tile == TMEM lane
    const int i = i0 + a;
    const uint32_t lane_base = (uint32_t)(w4 * 32) << 16;
    uint32_t* scr = reinterpret_cast<uint32_t*>(smem + OFF_SCR) + threadIdx.x * SCR32_WORDS;
    const int w0 = 96 - 32 * w4 + 32 * qt;             // first band column of the warp's window
    const int rowbar = 2 + w4;                         // named barrier of the four warps sharing these rows
    const uint8_t* padrow = p.pad ? p.pad + (int64_t)b * L : nullptr;
    if (padrow) {
      // training batches carry no pad tokens: decide once per CTA whether any key this CTA visits
      // is padded, and drop to the unmasked path if none is
      bool mine = false;
      const int jend = min(L, n_kt * TT);
      for (int j = threadIdx.x; j < jend; j += FW_MATH_THREADS) mine |= (padrow[j] != 0);
      if (!tc::named_bar_red_or(1, FW_MATH_THREADS, mine)) padrow = nullptr;
    }
    float m_run = -INFINITY, l_part = 0.f;             // l_part: this thread's quarter of the row sum

    int item = 0, jt = 0;
    for (int g = 0; g < n_g; ++g) {
      const int j0 = jt * TT;
      if (padrow) {          // stage the key tile's pad flags (overlaps the MMA)
        tc::named_bar_sync(1, FW_MATH_THREADS);
        if (qt == 0) spad[a] = (j0 + a < L) ? padrow[j0 + a] : 1;
        tc::named_bar_sync(1, FW_MATH_THREADS);
      }
      if (threadIdx.x == 0) FTRACE(0, g, 0);
      tc::mbar_wait(s_full, g & 1);
      tc::tc_fence_after();
      if (threadIdx.x == 0) FTRACE(0, g, 1);
      const uint32_t g_lo = tmem + ((jt & 1) ? TM_G1 : TM_G0);
      const uint32_t g_hi = tmem + (((jt + 1) & 1) ? TM_G1 : TM_G0);

      // band window -> scratch (or, MT_FWD_SKEW_REGS, registers), then S; the skewed read of the scratch overlaps the S load
#if MT_FWD_SKEW_REGS
      uint32_t Wn[32];
      skew_window_64(g_lo, g_hi, lane_base, w0, Wn);
#else
      skew_park_64(g_lo, g_hi, lane_base, w0, scr);
#endif
      float sv[32];
      {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem + TM_S + lane_base + qt * 32, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; ++x) sv[x] = __uint_as_float(r[x]);
      }
      // S / G fully read: the MMA warp may overwrite them with the next key tile
      tc::tc_fence_before();
      tc::mbar_arrive(s_consumed);
      if (threadIdx.x == 0) FTRACE(0, g, 2);
#if MT_FWD_SKEW_REGS
      skew_shift_add_32(sv, Wn, lane);
#else
      skew_fetch_add_32(sv, scr, lane);
#endif

      // ---- mask (only on the diagonal / ragged / padded tiles) + online softmax (log2 domain)
      const bool diag = p.causal && (j0 == i0);
      const bool tail = (j0 + TT > L);
      if (diag || tail || padrow != nullptr) {
        // branch-free: columns x > lim are masked (causal limit on the diagonal tile, ragged tail),
        // then the key-padding bytes, four per shared-memory word
        int lim = 31;
        if (diag) lim = min(lim, a - qt * 32);
        lim = min(lim, L - 1 - j0 - qt * 32);
#pragma unroll
        for (int x = 0; x < 32; ++x) sv[x] = (x > lim) ? -INFINITY : sv[x];
        if (padrow) {
          const uint32_t* sp = reinterpret_cast<const uint32_t*>(spad + qt * 32);
#pragma unroll
          for (int x4 = 0; x4 < 8; ++x4) {
            const uint32_t w = sp[x4];
#pragma unroll
            for (int e = 0; e < 4; ++e) sv[4 * x4 + e] = ((w >> (8 * e)) & 0xffu) ? -INFINITY : sv[4 * x4 + e];
          }
        }
      }
      float mx0 = fmaxf(sv[0], sv[1]), mx1 = fmaxf(sv[2], sv[3]);
#pragma unroll
      for (int x = 4; x < 32; x += 4) {
        mx0 = fmaxf(mx0, fmaxf(sv[x], sv[x + 1]));
        mx1 = fmaxf(mx1, fmaxf(sv[x + 2], sv[x + 3]));
      }
      float mx = fmaxf(mx0, mx1);
      // exchange the quarter-row maxima with the three partner threads (other quarters, same row);
      // the slots alternate with the step parity, so a slot is rewritten only after its readers
      // passed the following step's barrier
      float* xs = xch + (g & 1) * 4 * TT;
      xs[qt * TT + a] = mx;
      if (threadIdx.x == 0) FTRACE(0, g, 3);
      tc::named_bar_sync(rowbar, 128);
      if (threadIdx.x == 0) FTRACE(0, g, 4);
      mx = fmaxf(fmaxf(xs[a], xs[TT + a]), fmaxf(xs[2 * TT + a], xs[3 * TT + a])) * p.scale_log2;   // scale > 0
      float alpha = 1.f;
      if (mx > m_run + RESCALE_LOG2) {          // also the first tile (m_run = -inf) unless fully masked
        alpha = tc::fast_exp2(m_run - mx);      // m_run = -inf -> 0
        m_run = mx;
        l_part *= alpha;
      }
      const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
      for (int x = 0; x < 32; x += 4) {
        sv[x] = tc::fast_exp2(fmaf(sv[x], p.scale_log2, -m_use));
        sv[x + 1] = tc::fast_exp2(fmaf(sv[x + 1], p.scale_log2, -m_use));
        sv[x + 2] = tc::fast_exp2(fmaf(sv[x + 2], p.scale_log2, -m_use));
        sv[x + 3] = tc::fast_exp2(fmaf(sv[x + 3], p.scale_log2, -m_use));
        sum0 += sv[x]; sum1 += sv[x + 1]; sum2 += sv[x + 2]; sum3 += sv[x + 3];
      }
      l_part += (sum0 + sum1) + (sum2 + sum3);
      uint32_t pk[16];
#pragma unroll
      for (int x = 0; x < 16; ++x) pk[x] = pack16(sv[2 * x], sv[2 * x + 1], p.fmt);

      // ---- previous P.V must be complete before O is rescaled and P overwritten
      if (threadIdx.x == 0) FTRACE(0, g, 5);
      if (jt > 0) {
        tc::mbar_wait(o_done, (g - 1) & 1);
        tc::tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {       // each quarter rescales 16 of O's 64 columns
          uint32_t r[16];
          tc::tmem_ld_32x16(tmem + TM_O + lane_base + qt * 16, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int x = 0; x < 16; ++x) r[x] = __float_as_uint(__uint_as_float(r[x]) * alpha);
          tc::tmem_st_32x16(tmem + TM_O + lane_base + qt * 16, r);
          tc::tmem_st_wait();
        }
      }
      // ---- P (16-bit pairs) into the TMEM A-operand of the P.V MMA: row = lane, this thread's 32
      // key columns are 16 packed columns (no shared-memory round trip for P)
      if (threadIdx.x == 0) FTRACE(0, g, 6);
      tc::tmem_st_32x16(tmem + TM_P + lane_base + qt * 16, pk);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(p_full);
      if (threadIdx.x == 0) FTRACE(0, g, 7);
      if (p.stash) {
        // the backward reads P instead of rebuilding S, the skew and the exponentials.  Stash layout of a tile
        // (32 KB): [warp 0..15][chunk 0..3][lane][16 B] -- the thread that owns (row a, 32 key columns) here owns them
        // in the backward kernels too, and a warp's store (and the backward's load) of one chunk is 512 contiguous bytes
        const int it = i0 / TT;
        const int64_t tix = ((int64_t)b * p.h + hh0 + item) * p.nTri + (it * (it + 1) / 2 + jt);
        uint4* dst = reinterpret_cast<uint4*>(p.stash + tix * (int64_t)(2 * TILE)) + warp * 128 + lane;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          __stcs(dst + 32 * c, make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]));
        if ((qt & 1) == 0) p.mrow[(tix * 2 + (qt >> 1)) * TT + a] = m_use;      // [tile][key half][row]: one reference for both halves here
      }
      if (++jt < n_kt) continue;

      // ---- end of a head: O / l, LSE (the row sums are exchanged through their own slot: the step slots
      // may already be rewritten by a warp that is ahead)
      {
        const int hh = hh0 + item;
        float* xe = xch + 2 * 4 * TT;
        xe[qt * TT + a] = l_part;
        tc::named_bar_sync(rowbar, 128);
        const float l_run = (xe[a] + xe[TT + a]) + (xe[2 * TT + a] + xe[3 * TT + a]);
        tc::mbar_wait(o_done, g & 1);
        tc::tc_fence_after();
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        uint32_t packed[8];
        {
          uint32_t r[16];
          tc::tmem_ld_32x16(tmem + TM_O + lane_base + qt * 16, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int x = 0; x < 16; x += 2)
            packed[x / 2] = pack16(__uint_as_float(r[x]) * inv, __uint_as_float(r[x + 1]) * inv, p.ofmt);
        }
        tc::tc_fence_before();
        tc::mbar_arrive(o_free);               // the next head's first P.V product may overwrite O
        if (i < L) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.O) + (int64_t)b * p.ob + (int64_t)i * p.ol +
                                                (int64_t)hh * p.oh + qt * 16);
          dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          // natural-log LSE of the scaled logits (what the backward and rga_weights consume)
          if (qt == 0)
            p.lse[((int64_t)b * p.h + hh) * L + i] = l_run > 0.f ? (m_run + log2f(l_run)) * 0.6931471805599453f : 0.f;
        }
        // the partner warps must have read the row sums before this warp's next epilogue rewrites them: the
        // row barriers of the next head's steps lie in between
        m_run = -INFINITY;
        l_part = 0.f;
        jt = 0;
        ++item;
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == FW_MATH_WARPS + 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

// Measured and dropped (round 2): a second-generation forward with TWO SOFTMAX GROUPS ON THE TWO KEY HALVES of a tile --
// groups A / B own keys 0..63 / 64..127 of every tile, each with its own S sub-accumulator (N = 64 product), online
// softmax and O accumulator (P overwriting the S sub-accumulator frees the TMEM columns of the second O), merged at the
// end of a head like split-KV decoding.  Parity-green on every case of tests/test_gpu_rga.py, but 0.312 ms against
// 0.286 ms per config-B layer: the G ring still couples the groups every step (a block may only be overwritten when BOTH
// have parked their windows, and the next S product is committed behind the next G), so they never drift apart by more
// than one parking phase, and the N = 64 score products fetch the Q operand twice.  Real alternation (one group per key
// tile, as in the backward kernels) needs S and the G ring twice: 640 TMEM columns at 128-key steps.

}  // namespace

bool rga_bwd_tc_supported(const RgaArgs& a, int dh, int dtype);   // rga_tc_bwd.cu

// P stash of a training forward: B h nT (nT + 1) / 2 tiles of 32 KB, then per tile [2 key halves][128 rows] fp32 row references
size_t rga_stash_bytes(int64_t B, int64_t h, int64_t L) {
  const int64_t nT = (L + TT - 1) / TT;
  return (size_t)(B * h * (nT * (nT + 1) / 2)) * (size_t)(2 * TILE + 2 * TT * 4);
}

bool rga_tc_supported(const RgaArgs& a, int dh, int dtype, bool backward) {
  if (backward) return rga_bwd_tc_supported(a, dh, dtype);
  if (dh != DHC) return false;
  if (dtype != MT_BF16 && dtype != MT_F16 && dtype != MT_F16_BF16) return false;
  if (a.sl % 8 || a.sh % 8 || a.sb % 8) return false;
  if (!aligned(a.q, 16) || !aligned(a.k, 16) || !aligned(a.v, 16) || !aligned(a.E, 16)) return false;
  if (a.O && (a.ol % 8 || a.oh % 8 || a.ob % 8 || !aligned(a.O, 16))) return false;
  if (a.L < 1) return false;
  return mt_device_ok() != 0;
}

int rga_fwd_tc(const RgaArgs& a, int dh, int dtype, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV, tmE;
  int rc;
  if ((rc = tc::make_tmap_blhd(&tmQ, a.q, dh, a.L, a.h, a.B, a.sl, a.sh, a.sb, DHC, TT))) return rc;
  if ((rc = tc::make_tmap_blhd(&tmK, a.k, dh, a.L, a.h, a.B, a.sl, a.sh, a.sb, DHC, TT))) return rc;
  if ((rc = tc::make_tmap_blhd(&tmV, a.v, dh, a.L, a.h, a.B, a.sl, a.sh, a.sb, DHC, TT))) return rc;
  if ((rc = tc::make_tmap_2d(&tmE, a.E, a.max_seq, dh, dh, DHC, TT))) return rc;
  FwdParams p;
  p.O = a.O; p.ob = a.ob; p.ol = a.ol; p.oh = a.oh; p.lse = a.lse; p.pad = a.pad;
  p.B = a.B; p.h = a.h; p.L = a.L; p.max_seq = a.max_seq; p.causal = a.causal;
  p.fmt = (dtype == MT_BF16) ? 1 : 0;
  p.ofmt = (dtype == MT_F16) ? 0 : 1;
  p.scale_log2 = LOG2E / a.inv_scale_div;
  p.stash = nullptr; p.mrow = nullptr; p.nTri = 0;
  if (a.pstash) {
    if (!a.causal) { set_error("rga_fwd: the P stash exists for the causal mask only"); return MT_E_UNSUPPORTED; }
    const int64_t nTs = (a.L + TT - 1) / TT, tiles = (int64_t)a.B * a.h * (nTs * (nTs + 1) / 2);
    if (a.pstash_bytes < rga_stash_bytes(a.B, a.h, a.L) || !aligned(a.pstash, 128)) {
      set_error("rga_fwd: the P stash needs %zu bytes, 128-byte aligned", rga_stash_bytes(a.B, a.h, a.L));
      return MT_E_WORKSPACE;
    }
    p.stash = static_cast<uint8_t*>(a.pstash);
    p.mrow = reinterpret_cast<float*>(p.stash + tiles * (int64_t)(2 * TILE));
    p.nTri = (int)(nTs * (nTs + 1) / 2);
  }
  static unsigned long long attr_done = 0; const unsigned long long attr_bit = attr_dev_bit();
  if (!(attr_done & attr_bit)) {
    cudaError_t e = cudaFuncSetAttribute(rga_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
    if (e != cudaSuccess) { set_error("rga_fwd_tc: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done |= attr_bit;
  }
  // consecutive heads of one (batch row, query tile) share a CTA: same number of key tiles, same E blocks; the
  // per-CTA fixed cost (launch, TMEM allocation, barrier set-up, pipeline fill) is paid once per pair
  // (measured at config B, 2048 tile rows: 1 head per CTA 0.304 ms, 2 heads 0.287 ms, 4 heads 0.282 ms); as many
  // as leave at least three CTAs per SM
  static const int hpc_env = getenv("MT_FWD_HPC") ? atoi(getenv("MT_FWD_HPC")) : 0;
  const int nT = (a.L + TT - 1) / TT;
  int hpc = 1;
  for (int c = 4; c > 1; c >>= 1)
    if ((int64_t)((a.h + c - 1) / c) * a.B * nT >= 3 * (int64_t)sm_count()) { hpc = c; break; }
  if (hpc_env > 0) hpc = hpc_env;
  p.heads_per_cta = hpc > a.h ? a.h : hpc;
  dim3 grid((a.h + p.heads_per_cta - 1) / p.heads_per_cta, a.B, nT);
  p.trace = nullptr;
  p.trace_z = 0;
  static const bool want_trace = getenv("MT_RGA_TRACE") != nullptr;
  static long long* trace_dev = nullptr;
  const size_t trace_n = 2 * 32 * 8;
  if (want_trace) {
    if (!trace_dev) cudaMalloc(&trace_dev, trace_n * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, trace_n * sizeof(long long), st);
    p.trace = trace_dev;
    p.trace_z = atoi(getenv("MT_RGA_TRACE"));
  }
  rga_fwd_tc_kernel<<<grid, FW_THREADS, FWD_SMEM, st>>>(tmQ, tmK, tmV, tmE, p);
  if (want_trace) {
    static long long host[2 * 32 * 8];
    cudaMemcpyAsync(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    long long t0 = 0;
    for (size_t x = 0; x < trace_n; ++x) if (host[x] && (!t0 || host[x] < t0)) t0 = host[x];
    static const char* agent[2] = {"SM", "MMA"};
    for (int ag = 0; ag < 2; ++ag)
      for (int n = 0; n < 32; ++n) {
        bool any = false;
        for (int e = 0; e < 8; ++e) any |= host[(ag * 32 + n) * 8 + e] != 0;
        if (!any) continue;
        fprintf(stderr, "trace fwd %-3s step %2d:", agent[ag], n);
        for (int e = 0; e < 8; ++e) fprintf(stderr, " %8lld", host[(ag * 32 + n) * 8 + e] ? host[(ag * 32 + n) * 8 + e] - t0 : -1LL);
        fprintf(stderr, "\n");
      }
  }
  return check_launch("rga_fwd_tc");
}

}  // namespace mt
