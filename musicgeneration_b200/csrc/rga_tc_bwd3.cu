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
// skew, MT/layers.py:116-125).  The band shift is per-row variable, so it is done by 8 converter
// warps: swizzled dS image -> registers -> band_store into the dG operand.  No S, no G, no exp, no
// row statistics.  Workspace traffic: one write + two reads of 32 KB per tile pair (config B: 557 MB
// per layer each way) against the 4/6 of the tensor work and 2/3 of the MUFU work it removes.
//
//   warps 0-7 : converters (row a = 32*(w&3)+lane, key columns 64*(w>>2)..+63)
//   warp 8    : loader (1-D bulk copies of the dS images, TMA tiles of K/E or Q)
//   warp 9    : tcgen05.mma issuer
#include "ops.cuh"
#include "rga_tc_common.cuh"

namespace mt {

using namespace rga;

namespace {

enum { L_DQ = 0, L_DE = 1 };

constexpr int CV_THREADS = 256;
constexpr int B3_THREADS = CV_THREADS + 64;
constexpr int DS_BYTES = 2 * TILE;       // one dS tile image: two [128 x 64] swizzled sub-tiles

template <int ROLE> struct Lay3;
template <> struct Lay3<L_DQ> {      // dS x 2; K x 2; E ring x 3; dG (4 sub-tiles)
  static constexpr int DS0 = 0, X0 = 4 * TILE, E0 = 6 * TILE, DG = 9 * TILE, BAR = 13 * TILE;
  static constexpr int NDG = 1;
  static constexpr uint32_t TMEM_COLS = 64;
};
template <> struct Lay3<L_DE> {      // dS x 2; Q x 2; dG x 2
  static constexpr int DS0 = 0, X0 = 4 * TILE, DG = 6 * TILE, BAR = 14 * TILE;
  static constexpr int NDG = 2;
  static constexpr uint32_t TMEM_COLS = 128;
};
template <int ROLE> constexpr int smem3_bytes() { return Lay3<ROLE>::BAR + 256; }
static_assert(smem3_bytes<L_DQ>() <= 232448 && smem3_bytes<L_DE>() <= 232448, "shared memory budget");

enum { B3_DSF = 0, B3_DSE = 2, B3_XF = 4, B3_XE = 6, B3_DGR = 8, B3_DGF = 10, B3_DONE = 12, B3_TMEM = 13 };

struct Bwd3Params {
  const uint8_t* ws;                     // dS tiles: [(b*h+hh)][it*(it+1)/2 + jt][32 KB]
  void* dq; int64_t sb, sl, sh;
  float* dE;
  int B, h, L, max_seq, nT, nTri;
  int bh_per_cta;                        // dE role
};

struct Step3 { int it, jt, b, hh; };

template <int ROLE>
__device__ __forceinline__ int num_steps3(const Bwd3Params& p, int& bh0) {
  bh0 = 0;
  if (ROLE == L_DQ) return p.nT - (int)blockIdx.z;             // it = nT-1-blockIdx.z: longest first
  bh0 = (int)blockIdx.x * p.bh_per_cta;
  const int nbh = min(p.bh_per_cta, p.B * p.h - bh0);
  return nbh > 0 ? nbh * (p.nT - (int)blockIdx.z) : 0;
}
template <int ROLE>
__device__ __forceinline__ Step3 step3_first(const Bwd3Params& p, int bh0) {
  Step3 s;
  if (ROLE == L_DQ) { s.it = p.nT - 1 - (int)blockIdx.z; s.jt = 0; s.hh = blockIdx.x; s.b = blockIdx.y; }
  else { s.it = (int)blockIdx.z; s.jt = 0; s.b = bh0 / p.h; s.hh = bh0 % p.h; }
  return s;
}
template <int ROLE>
__device__ __forceinline__ void step3_advance(const Bwd3Params& p, Step3& s) {
  if (ROLE == L_DQ) { ++s.jt; return; }
  if (s.it + 1 < p.nT) { ++s.it; ++s.jt; return; }       // next tile down the diagonal
  s.it = (int)blockIdx.z; s.jt = 0;                      // next (batch, head) of the slice
  if (++s.hh == p.h) { s.hh = 0; ++s.b; }
}
__device__ __forceinline__ const uint8_t* ds_tile(const Bwd3Params& p, const Step3& s) {
  return p.ws + (((int64_t)s.b * p.h + s.hh) * p.nTri + (s.it * (s.it + 1) / 2 + s.jt)) * (int64_t)DS_BYTES;
}

template <int ROLE>
__global__ void __launch_bounds__(B3_THREADS, 1)
rga_bwd3_kernel(const __grid_constant__ CUtensorMap tmX,      // K (dQ role) or Q (dE role)
                const __grid_constant__ CUtensorMap tmE, const Bwd3Params p) {
  using LY = Lay3<ROLE>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LY::BAR);
  uint64_t* ds_full = bars + B3_DSF;      // [2] loader -> converters (and the dS.K MMA)
  uint64_t* ds_empty = bars + B3_DSE;     // [2] converters (+ MMA commit in the dQ role) -> loader
  uint64_t* x_full = bars + B3_XF;        // [2] K + E block / Q
  uint64_t* x_empty = bars + B3_XE;       // [2]
  uint64_t* dg_ready = bars + B3_DGR;     // [NDG] converters -> MMA
  uint64_t* dg_free = bars + B3_DGF;      // [NDG] MMA -> converters
  uint64_t* acc_done = bars + B3_DONE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B3_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bh0;
  const int nsteps = num_steps3<ROLE>(p, bh0);

  if (warp == 8 && lane == 0) {
    tc::tma_prefetch_desc(&tmX);
    if (ROLE == L_DQ) tc::tma_prefetch_desc(&tmE);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&ds_full[s], 1);
      tc::mbar_init(&ds_empty[s], CV_THREADS + (ROLE == L_DQ ? 1 : 0));
      tc::mbar_init(&x_full[s], 1);
      tc::mbar_init(&x_empty[s], 1);
      tc::mbar_init(&dg_ready[s], CV_THREADS);
      tc::mbar_init(&dg_free[s], 1);
    }
    tc::mbar_init(acc_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 9) tc::tmem_alloc(tmem_slot, LY::TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (nsteps <= 0) {            // (dE role: empty slice) -- uniform for the whole CTA
    __syncthreads();
    if (warp == 9) tc::tmem_dealloc(tmem, LY::TMEM_COLS);
    return;
  }
  constexpr uint64_t TS16 = TILE >> 4;
  // dQ role: E block m (hi block of step m; block -1 = lo block of step 0) lives in ring slot (m+1) % 3
  auto eslot = [](int m) -> int { return (m + 1) % 3; };

  if (warp == 8) {
    // ================================ loader ===============================================
    if (lane == 0) {
      Step3 s = step3_first<ROLE>(p, bh0);
      for (int n = 0; n < nsteps; ++n, step3_advance<ROLE>(p, s)) {
        const int st = n & 1;
        const uint32_t par = ((n >> 1) & 1) ^ 1;
        tc::mbar_wait(&x_empty[st], par);
        if (ROLE == L_DQ) {
          const int c0 = p.max_seq - 1 - (s.it - s.jt) * TT;
          tc::mbar_arrive_expect_tx(&x_full[st], (n == 0 ? 3 : 2) * TILE);
          tc::tma_load_4d(smem + LY::X0 + st * TILE, &tmX, &x_full[st], 0, s.hh, s.jt * TT, s.b);
          tc::tma_load_2d(smem + Lay3<L_DQ>::E0 + eslot(n) * TILE, &tmE, &x_full[st], 0, c0 + 1);
          if (n == 0) tc::tma_load_2d(smem + Lay3<L_DQ>::E0 + eslot(-1) * TILE, &tmE, &x_full[st], 0, c0 - (TT - 1));
        } else {
          tc::mbar_arrive_expect_tx(&x_full[st], TILE);
          tc::tma_load_4d(smem + LY::X0 + st * TILE, &tmX, &x_full[st], 0, s.hh, s.it * TT, s.b);
        }
        tc::mbar_wait(&ds_empty[st], par);
        tc::mbar_arrive_expect_tx(&ds_full[st], DS_BYTES);
        tc::bulk_load_1d(smem + LY::DS0 + st * DS_BYTES, ds_tile(p, s), DS_BYTES, &ds_full[st]);
        if (n + 2 < nsteps) {         // pull the tiles of step n+2 into L2
          Step3 t = s;
          step3_advance<ROLE>(p, t);
          step3_advance<ROLE>(p, t);
          tc::bulk_prefetch_l2(ds_tile(p, t), DS_BYTES);
          tc::tma_prefetch_4d(&tmX, 0, t.hh, (ROLE == L_DQ ? t.jt : t.it) * TT, t.b);
        }
      }
    }
  } else if (warp == 9) {
    // ================================ MMA issuer ============================================
    if (lane == 0) {
      if (ROLE == L_DQ) {
        const uint32_t id_kmn = tc::make_idesc(TT, DHC, 1, 1, 0, 1);    // A K-major (dS / dG), B MN-major (K / E), N = 64
        const uint64_t dsd0 = tc::make_sdesc(tc::smem_u32(smem + LY::DS0), 16, 1024);
        const uint64_t kd_mn0 = tc::make_sdesc(tc::smem_u32(smem + LY::X0), 1024, 1024);
        const uint64_t ed_mn0 = tc::make_sdesc(tc::smem_u32(smem + Lay3<L_DQ>::E0), 1024, 1024);
        const uint64_t dgd = tc::make_sdesc(tc::smem_u32(smem + LY::DG), 16, 1024);
        for (int n = 0; n < nsteps; ++n) {
          const uint64_t st = n & 1;
          const uint32_t par = (n >> 1) & 1;
          tc::mbar_wait(&x_full[st], par);
          tc::mbar_wait(&ds_full[st], par);
          tc::tc_fence_after();
#pragma unroll
          for (int k16 = 0; k16 < TT / 16; ++k16)         // dQ += dS . K_j (contraction over the 128 keys)
            tc::umma_f16(tmem, dsd0 + st * 2 * TS16 + (uint64_t)(k16 >> 2) * TS16 + 2 * (k16 & 3),
                         kd_mn0 + st * TS16 + 128 * k16, id_kmn, (n | k16) != 0);
          tc::umma_commit(&ds_empty[st]);
          tc::mbar_wait(&dg_ready[0], n & 1);
          tc::tc_fence_after();
          const uint64_t elo = (uint64_t)eslot(n - 1) * TS16, ehi = (uint64_t)eslot(n) * TS16;
#pragma unroll
          for (int k16 = 0; k16 < 2 * TT / 16; ++k16)     // dQ += dG . [E_lo; E_hi] (contraction over the band)
            tc::umma_f16(tmem, dgd + (uint64_t)(k16 >> 2) * TS16 + 2 * (k16 & 3),
                         ed_mn0 + (k16 < 8 ? elo + 128 * k16 : ehi + 128 * (k16 - 8)), id_kmn, 1);
          tc::umma_commit(&dg_free[0]);
          tc::umma_commit(&x_empty[st]);
        }
      } else {
        const uint32_t id_mnmn = tc::make_idesc(TT, DHC, 1, 1, 1, 1);   // A MN-major (dG block), B MN-major (Q), N = 64
        const uint64_t qd_mn0 = tc::make_sdesc(tc::smem_u32(smem + LY::X0), 1024, 1024);
        const uint64_t dg_lo0 = tc::make_sdesc(tc::smem_u32(smem + LY::DG), TILE, 1024);
        const uint64_t dg_hi0 = tc::make_sdesc(tc::smem_u32(smem + LY::DG + 2 * TILE), TILE, 1024);
        for (int n = 0; n < nsteps; ++n) {
          const uint64_t st = n & 1;
          const uint32_t par = (n >> 1) & 1;
          tc::mbar_wait(&x_full[st], par);
          tc::mbar_wait(&dg_ready[st], par);
          tc::tc_fence_after();
#pragma unroll
          for (int k16 = 0; k16 < TT / 16; ++k16) {       // dE_blk += dG_blk^T . Q (contraction over the query rows)
            tc::umma_f16(tmem, dg_lo0 + st * 4 * TS16 + 128 * k16, qd_mn0 + st * TS16 + 128 * k16, id_mnmn, (n | k16) != 0);
            tc::umma_f16(tmem + 64, dg_hi0 + st * 4 * TS16 + 128 * k16, qd_mn0 + st * TS16 + 128 * k16, id_mnmn, (n | k16) != 0);
          }
          tc::umma_commit(&dg_free[st]);
          tc::umma_commit(&x_empty[st]);
        }
      }
      tc::umma_commit(acc_done);
    }
  } else {
    // ================================ converters: dS image -> band dG ========================
    const int w4 = warp & 3, half = warp >> 2;
    const int a = w4 * 32 + lane;
    const uint32_t lane_base = (uint32_t)(w4 * 32) << 16;
    {
      // dG is zero outside the 128 band columns each row owns; those positions never change
      uint4* z = reinterpret_cast<uint4*>(smem + LY::DG);
      for (int x = threadIdx.x; x < LY::NDG * 4 * TILE / 16; x += CV_THREADS) z[x] = make_uint4(0, 0, 0, 0);
      tc::fence_proxy_async();
      tc::named_bar_sync(1, CV_THREADS);
    }
    const int base_w = ((127 - a) >> 1) + 32 * half;     // first 32-bit word of this thread's band run in dG
    for (int n = 0; n < nsteps; ++n) {
      const int st = n & 1;
      tc::mbar_wait(&ds_full[st], (n >> 1) & 1);
      uint32_t A[32];
      const uint8_t* img = smem + LY::DS0 + st * DS_BYTES + half * TILE;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 w = *reinterpret_cast<const uint4*>(img + swz_chunk(a, c));
        A[4 * c] = w.x; A[4 * c + 1] = w.y; A[4 * c + 2] = w.z; A[4 * c + 3] = w.w;
      }
      tc::mbar_arrive(&ds_empty[st]);
      const int buf = (LY::NDG == 2) ? st : 0;
      if (LY::NDG == 1) { if (n > 0) tc::mbar_wait(&dg_free[0], (n - 1) & 1); }
      else if (n >= 2) tc::mbar_wait(&dg_free[st], ((n >> 1) - 1) & 1);
      band_store(smem + LY::DG + buf * 4 * TILE, a, base_w, A);
      tc::fence_proxy_async();
      tc::mbar_arrive(&dg_ready[buf]);
    }

    // ---- epilogue
    tc::mbar_wait(acc_done, 0);
    tc::tc_fence_after();
    if (ROLE == L_DQ) {           // 64 accumulator columns: this thread takes 32 of row a
      uint32_t r[32];
      tc::tmem_ld_32x32(tmem + lane_base + half * 32, r);
      tc::tmem_ld_wait();
      const Step3 s = step3_first<ROLE>(p, bh0);
      const int row = s.it * TT + a;
      if (row < p.L) {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dq) + (int64_t)s.b * p.sb +
                                              (int64_t)row * p.sl + (int64_t)s.hh * p.sh + half * 32);
#pragma unroll
        for (int x = 0; x < 4; ++x)
          dst[x] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * x]), __uint_as_float(r[8 * x + 1])),
                              pack_bf16x2(__uint_as_float(r[8 * x + 2]), __uint_as_float(r[8 * x + 3])),
                              pack_bf16x2(__uint_as_float(r[8 * x + 4]), __uint_as_float(r[8 * x + 5])),
                              pack_bf16x2(__uint_as_float(r[8 * x + 6]), __uint_as_float(r[8 * x + 7])));
      }
    } else {                      // two blocks of 128 E rows x 64: warps 0-3 the lo block, 4-7 the hi block
      const int c0 = p.max_seq - 1 - (int)blockIdx.z * TT;
      const int erow = (half == 0 ? c0 - (TT - 1) : c0 + 1) + a;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem + lane_base + 64 * half + 32 * q, r);
        tc::tmem_ld_wait();
        if (erow >= 0 && erow < p.max_seq) {
#pragma unroll
          for (int x = 0; x < 32; ++x) atomicAdd(p.dE + (int64_t)erow * DHC + 32 * q + x, __uint_as_float(r[x]));
        }
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 9) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, LY::TMEM_COLS);
  }
}

template <int ROLE>
int launch_role3(const CUtensorMap& tmX, const CUtensorMap& tmE, const Bwd3Params& p, dim3 grid, cudaStream_t st) {
  auto kern = rga_bwd3_kernel<ROLE>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3_bytes<ROLE>());
    if (e != cudaSuccess) { set_error("rga_bwd3: smem attribute (%d B): %s", smem3_bytes<ROLE>(), cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
  }
  kern<<<grid, B3_THREADS, smem3_bytes<ROLE>(), st>>>(tmX, tmE, p);
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
  return p;
}

}  // namespace

size_t rga_bwd3_workspace_bytes(int64_t B, int64_t h, int64_t L) {
  const int64_t nT = (L + TT - 1) / TT;
  return (size_t)(B * h * (nT * (nT + 1) / 2)) * DS_BYTES;
}

// dQ from the spilled dS tiles (query-tile owner walks the key tiles at or left of it)
int rga_bwd3_dq(const RgaArgs& a, const void* ws, const CUtensorMap& tmK, const CUtensorMap& tmE, cudaStream_t st) {
  Bwd3Params p = make_params3(a, ws);
  return launch_role3<L_DQ>(tmK, tmE, p, dim3(a.h, a.B, p.nT), st);
}

// dE from the spilled dS tiles (tile-diagonal owner walks down the diagonal over a slice of (batch, head))
int rga_bwd3_de(const RgaArgs& a, const void* ws, const CUtensorMap& tmQ, const CUtensorMap& tmE, cudaStream_t st) {
  Bwd3Params p = make_params3(a, ws);
  const int bh = a.B * a.h;
  int slices = (2 * sm_count() + p.nT - 1) / p.nT;       // about two CTAs per SM's worth of slices
  if (slices > bh) slices = bh;
  if (slices < 1) slices = 1;
  p.bh_per_cta = (bh + slices - 1) / slices;
  slices = (bh + p.bh_per_cta - 1) / p.bh_per_cta;
  return launch_role3<L_DE>(tmQ, tmE, p, dim3(slices, 1, p.nT), st);
}

}  // namespace mt
