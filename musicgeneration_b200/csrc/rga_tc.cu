// Fused relative global attention on tcgen05 / TMEM / TMA (K1), head dim 64, bf16 or f16 operands.
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
// The reference's skew (MT/layers.py:116-125) becomes index arithmetic: row a of the tile needs
// the 128 band columns starting at 127-a of [G_lo | G_hi].  TMEM column addresses are
// warp-uniform, so each softmax warp loads the 96-column window common to its 32 rows and the
// per-lane residual shift (31 - lane) is done through a row-private shared-memory scratch
// (written with 128-bit stores, read back at the shifted offset, both bank-conflict free).
// E rows >= max_seq (exactly the j > i positions) and rows < 0 are zero-filled by TMA, which
// realises the reference's _qe_masking.  Online softmax in fp32 with exp2; LSE saved for backward.
//
// Warp roles: 0-3 softmax/correction/epilogue (thread = query row = TMEM lane), 4 = TMA producer,
// 5 = TMEM allocator + MMA issuer.
#include "ops.cuh"
#include "tc_common.cuh"

namespace mt {

namespace {

constexpr int QT = 128;              // query rows per CTA
constexpr int KT = 128;              // keys per step
constexpr int DHC = 64;              // head dim
constexpr int TILE = QT * DHC * 2;   // 16 KB: one [128 x 64] 16-bit operand tile
constexpr int SCR_PITCH = 100;       // floats per scratch row (== 4 mod 32: conflict-free v4 stores)
constexpr int FWD_THREADS = 192;

// shared memory map (offsets from the 1024-aligned base)
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE;            // 2 stages
constexpr int OFF_V = OFF_K + 2 * TILE;        // 2 stages
constexpr int OFF_E = OFF_V + 2 * TILE;        // 2 stages: the new "hi" block of each step
constexpr int OFF_ELO = OFF_E + 2 * TILE;      // "lo" block of the first step only
constexpr int OFF_P = OFF_ELO + TILE;          // 2 K-subtiles of [128 x 64] bf16
constexpr int OFF_SCR = OFF_P + 2 * TILE;
constexpr int OFF_BAR = OFF_SCR + QT * SCR_PITCH * 4;
constexpr int FWD_SMEM = OFF_BAR + 256 + 1024;

// TMEM columns
constexpr uint32_t TM_S = 0, TM_G0 = 128, TM_G1 = 256, TM_O = 384;

struct FwdParams {
  void* O; int64_t ob, ol, oh;
  float* lse;
  const uint8_t* pad;
  int B, h, L, max_seq, causal, fmt;
  float scale_log2;     // log2(e) / sqrt(dh)
  float scale;          // 1 / sqrt(dh)
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(FWD_THREADS, 1)
rga_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmE,
                  const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, hh = blockIdx.y;
  const int i0 = (gridDim.x - 1 - blockIdx.x) * QT;          // longest rows first
  const int L = p.L;
  const int n_kt = p.causal ? (i0 / KT + 1) : (L + KT - 1) / KT;

  if (warp == 4 && lane == 0) {
    tc::tma_prefetch_desc(&tmQ);
    tc::tma_prefetch_desc(&tmK);
    tc::tma_prefetch_desc(&tmV);
    tc::tma_prefetch_desc(&tmE);
    tc::mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&kv_full[s], 1);
      tc::mbar_init(&kv_empty[s], 1);
    }
    tc::mbar_init(s_full, 1);
    tc::mbar_init(s_consumed, 128);
    tc::mbar_init(p_full, 128);
    tc::mbar_init(o_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 5) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ================================ TMA producer ==========================================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(bar_q, TILE);
      tc::tma_load_4d(smem + OFF_Q, &tmQ, bar_q, 0, hh, i0, b);
      for (int jt = 0; jt < n_kt; ++jt) {
        const int s = jt & 1;
        const int j0 = jt * KT;
        const int c0 = p.max_seq - 1 - (i0 - j0);
        tc::mbar_wait(&kv_empty[s], ((jt >> 1) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&kv_full[s], (jt == 0 ? 4 : 3) * TILE);
        tc::tma_load_4d(smem + OFF_K + s * TILE, &tmK, &kv_full[s], 0, hh, j0, b);
        tc::tma_load_4d(smem + OFF_V + s * TILE, &tmV, &kv_full[s], 0, hh, j0, b);
        tc::tma_load_2d(smem + OFF_E + s * TILE, &tmE, &kv_full[s], 0, c0 + 1);
        if (jt == 0) tc::tma_load_2d(smem + OFF_ELO, &tmE, &kv_full[s], 0, c0 - (KT - 1));
      }
    }
  } else if (warp == 5) {
    // ================================ MMA issuer ============================================
    if (lane == 0) {
      const uint32_t idesc_s = tc::make_idesc(QT, KT, p.fmt, p.fmt, 0, 0);     // S, G: K-major x K-major
      const uint32_t idesc_o = tc::make_idesc(QT, DHC, p.fmt, p.fmt, 0, 1);    // O: P K-major, V MN-major
      const uint32_t q_base = tc::smem_u32(smem + OFF_Q);
      auto issue_s = [&](int jt) {
        const int s = jt & 1;
        const uint32_t k_base = tc::smem_u32(smem + OFF_K + s * TILE);
        const uint32_t e_base = tc::smem_u32(smem + OFF_E + s * TILE);
        const uint32_t g_hi = tmem + (((jt + 1) & 1) ? TM_G1 : TM_G0);
#pragma unroll
        for (int k4 = 0; k4 < DHC / 16; ++k4) {
          const uint64_t qd = tc::make_sdesc(q_base + k4 * 32, 16, 1024);
          tc::umma_f16(tmem + TM_S, qd, tc::make_sdesc(k_base + k4 * 32, 16, 1024), idesc_s, k4 != 0);
          tc::umma_f16(g_hi, qd, tc::make_sdesc(e_base + k4 * 32, 16, 1024), idesc_s, k4 != 0);
        }
        if (jt == 0) {
          const uint32_t elo = tc::smem_u32(smem + OFF_ELO);
#pragma unroll
          for (int k4 = 0; k4 < DHC / 16; ++k4)
            tc::umma_f16(tmem + TM_G0, tc::make_sdesc(q_base + k4 * 32, 16, 1024),
                         tc::make_sdesc(elo + k4 * 32, 16, 1024), idesc_s, k4 != 0);
        }
        tc::umma_commit(s_full);
      };
      tc::mbar_wait(bar_q, 0);
      tc::mbar_wait(&kv_full[0], 0);
      tc::tc_fence_after();
      issue_s(0);
      for (int jt = 0; jt < n_kt; ++jt) {
        if (jt + 1 < n_kt) {
          tc::mbar_wait(s_consumed, jt & 1);
          tc::mbar_wait(&kv_full[(jt + 1) & 1], ((jt + 1) >> 1) & 1);
          tc::tc_fence_after();
          issue_s(jt + 1);
        }
        tc::mbar_wait(p_full, jt & 1);
        tc::tc_fence_after();
        const uint32_t p_base = tc::smem_u32(smem + OFF_P);
        const uint32_t v_base = tc::smem_u32(smem + OFF_V + (jt & 1) * TILE);
#pragma unroll
        for (int k8 = 0; k8 < KT / 16; ++k8) {
          const uint64_t pd = tc::make_sdesc(p_base + (k8 >> 2) * TILE + (k8 & 3) * 32, 16, 1024);
          const uint64_t vd = tc::make_sdesc(v_base + k8 * 2048, 1024, 1024);
          tc::umma_f16(tmem + TM_O, pd, vd, idesc_o, (jt | k8) != 0);
        }
        tc::umma_commit(&kv_empty[jt & 1]);
        tc::umma_commit(o_done);
      }
    }
  } else {
    // ================================ softmax warps =========================================
    const int a = threadIdx.x;                         // query row within the tile == TMEM lane
    const int i = i0 + a;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float* scr = reinterpret_cast<float*>(smem + OFF_SCR) + a * SCR_PITCH;
    const uint8_t* padrow = p.pad ? p.pad + (int64_t)b * L : nullptr;
    float m_run = -INFINITY, l_run = 0.f;

    for (int jt = 0; jt < n_kt; ++jt) {
      const int j0 = jt * KT;
      if (padrow) {          // stage the key tile's pad flags (overlaps the MMA)
        tc::named_bar_sync(1, 128);
        spad[a] = (j0 + a < L) ? padrow[j0 + a] : 1;
        tc::named_bar_sync(1, 128);
      }
      tc::mbar_wait(s_full, jt & 1);
      tc::tc_fence_after();
      const uint32_t g_lo = tmem + ((jt & 1) ? TM_G1 : TM_G0);
      const uint32_t g_hi = tmem + (((jt + 1) & 1) ? TM_G1 : TM_G0);

      float sv[KT];
      // ---- S row
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem + TM_S + lane_base + c * 32, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; ++x) sv[c * 32 + x] = __uint_as_float(r[x]);
      }
      // ---- skewed relative term: Srel[a][b] = [G_lo | G_hi][a][127 - a + b]
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int w0 = 96 - 32 * warp + 64 * half;      // first band column of this warp's window
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int cc = w0 + 32 * c;                   // multiple of 32: entirely in lo or in hi
          uint32_t r[32];
          tc::tmem_ld_32x32((cc < 128 ? g_lo + cc : g_hi + (cc - 128)) + lane_base, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int x = 0; x < 32; x += 4)
            *reinterpret_cast<uint4*>(scr + c * 32 + x) = make_uint4(r[x], r[x + 1], r[x + 2], r[x + 3]);
        }
        const float* rd = scr + (31 - lane);
#pragma unroll
        for (int x = 0; x < 64; ++x) sv[half * 64 + x] += rd[x];
      }
      // S / G fully read: the MMA warp may overwrite them with the next key tile
      tc::tc_fence_before();
      tc::mbar_arrive(s_consumed);

      // ---- mask + online softmax (log2 domain)
      const bool diag = p.causal && (j0 == i0);
      const bool tail = (j0 + KT > L);
      float mx = -INFINITY;
#pragma unroll
      for (int x = 0; x < KT; ++x) {
        bool ok = true;
        if (diag) ok = (x <= a);
        if (tail) ok = ok && (j0 + x < L);
        if (padrow) ok = ok && (spad[x] == 0);
        sv[x] = ok ? sv[x] * p.scale_log2 : -INFINITY;
        mx = fmaxf(mx, sv[x]);
      }
      const float m_new = fmaxf(m_run, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = tc::fast_exp2(m_run - m_use);      // m_run = -inf -> 0
      float sum = 0.f;
#pragma unroll
      for (int x = 0; x < KT; ++x) {
        sv[x] = tc::fast_exp2(sv[x] - m_use);
        sum += sv[x];
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;

      // ---- previous P.V must be complete before O is rescaled and P overwritten
      if (jt > 0) {
        tc::mbar_wait(o_done, (jt - 1) & 1);
        tc::tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
          for (int c = 0; c < DHC / 32; ++c) {
            uint32_t r[32];
            tc::tmem_ld_32x32(tmem + TM_O + lane_base + c * 32, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int x = 0; x < 32; ++x) r[x] = __float_as_uint(__uint_as_float(r[x]) * alpha);
            tc::tmem_st_32x32(tmem + TM_O + lane_base + c * 32, r);
          }
          tc::tmem_st_wait();
        }
      }
      // ---- P (16-bit) into the K-major 128B-swizzled operand layout: row a, 16-byte chunk c
      //      lands at chunk (c ^ (a & 7)) of the row
      uint8_t* prow = smem + OFF_P + a * 128;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float* v = sv + sub * 64 + c * 8;
          uint4 w;
          if (p.fmt == 1) {
            w = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          } else {
            w = make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
          }
          *reinterpret_cast<uint4*>(prow + sub * TILE + ((c ^ (a & 7)) << 4)) = w;
        }
      }
      tc::fence_proxy_async();
      tc::tc_fence_before();
      tc::mbar_arrive(p_full);
    }
    // ---- epilogue: O / l, LSE
    tc::mbar_wait(o_done, (n_kt - 1) & 1);
    tc::tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    uint32_t packed[DHC / 2];
#pragma unroll
    for (int c = 0; c < DHC / 32; ++c) {
      uint32_t r[32];
      tc::tmem_ld_32x32(tmem + TM_O + lane_base + c * 32, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; x += 2) {
        const float v0 = __uint_as_float(r[x]) * inv, v1 = __uint_as_float(r[x + 1]) * inv;
        packed[c * 16 + x / 2] = (p.fmt == 1) ? pack_bf16(v0, v1) : pack_f16(v0, v1);
      }
    }
    if (i < L) {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.O) + (int64_t)b * p.ob + (int64_t)i * p.ol +
                                            (int64_t)hh * p.oh);
#pragma unroll
      for (int x = 0; x < DHC / 8; ++x)
        dst[x] = make_uint4(packed[4 * x], packed[4 * x + 1], packed[4 * x + 2], packed[4 * x + 3]);
      // natural-log LSE of the scaled logits (what the backward and rga_weights consume)
      p.lse[((int64_t)b * p.h + hh) * L + i] = l_run > 0.f ? (m_run + log2f(l_run)) * 0.6931471805599453f : 0.f;
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

bool rga_bwd_tc_supported(const RgaArgs& a, int dh, int dtype);   // rga_tc_bwd.cu

bool rga_tc_supported(const RgaArgs& a, int dh, int dtype, bool backward) {
  if (backward) return rga_bwd_tc_supported(a, dh, dtype);
  if (dh != DHC) return false;
  if (dtype != MT_BF16 && dtype != MT_F16) return false;
  if (a.sl % 8 || a.sh % 8 || a.sb % 8) return false;
  if (!aligned(a.q, 16) || !aligned(a.k, 16) || !aligned(a.v, 16) || !aligned(a.E, 16)) return false;
  if (a.O && (a.ol % 8 || a.oh % 8 || a.ob % 8 || !aligned(a.O, 16))) return false;
  if (a.L < 1) return false;
  return mt_device_ok() != 0;
}

int rga_fwd_tc(const RgaArgs& a, int dh, int dtype, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV, tmE;
  int rc;
  if ((rc = tc::make_tmap_blhd(&tmQ, a.q, dh, a.L, a.h, a.B, a.sl, a.sh, a.sb, DHC, QT))) return rc;
  if ((rc = tc::make_tmap_blhd(&tmK, a.k, dh, a.L, a.h, a.B, a.sl, a.sh, a.sb, DHC, KT))) return rc;
  if ((rc = tc::make_tmap_blhd(&tmV, a.v, dh, a.L, a.h, a.B, a.sl, a.sh, a.sb, DHC, KT))) return rc;
  if ((rc = tc::make_tmap_2d(&tmE, a.E, a.max_seq, dh, dh, DHC, KT))) return rc;
  FwdParams p;
  p.O = a.O; p.ob = a.ob; p.ol = a.ol; p.oh = a.oh; p.lse = a.lse; p.pad = a.pad;
  p.B = a.B; p.h = a.h; p.L = a.L; p.max_seq = a.max_seq; p.causal = a.causal;
  p.fmt = (dtype == MT_BF16) ? 1 : 0;
  p.scale = 1.f / a.inv_scale_div;
  p.scale_log2 = 1.4426950408889634f / a.inv_scale_div;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(rga_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
    if (e != cudaSuccess) { set_error("rga_fwd_tc: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
  }
  dim3 grid((a.L + QT - 1) / QT, a.h, a.B);
  rga_fwd_tc_kernel<<<grid, FWD_THREADS, FWD_SMEM, st>>>(tmQ, tmK, tmV, tmE, p);
  return check_launch("rga_fwd_tc");
}

}  // namespace mt
