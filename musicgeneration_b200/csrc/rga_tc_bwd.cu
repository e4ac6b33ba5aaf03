// Backward of the fused relative global attention on tcgen05 (K2): dispatch.
//
// Math (MT/layers.py:86-106 through autograd), per (batch, head):
//     P = exp(S - lse),  S = (Q K^T + skew(Q E_band^T)) / sqrt(dh)
//     dP = dO V^T ;  D = rowsum(dO o O) ;  dS = P o (dP - D) / sqrt(dh)
//     dV = P^T dO ;  dK = dS^T Q ;  dQ = dS K + dG E_band ;  dE_band += dG^T Q
// with dG = dS in band coordinates (dG[a][127-a+b] = dS[a][b], the inverse of the forward skew).
//
// dK/dV accumulate along queries, dQ along keys and dE along tile diagonals -- three owners -- so the
// work is three launches:
//   * dS-spill variant (a workspace is given): rga_tc_bwd2.cu role R_DKV computes S, the skew, P, dP and
//     dS ONCE, accumulates dK / dV and spills every dS tile (bf16 operand image); rga_tc_bwd3.cu turns the
//     tiles into dQ and dE.
//   * recompute variant (no workspace, bf16 only): roles R_DKV, R_DQ, R_DE of rga_tc_bwd2.cu each rebuild
//     P / dS.
// Mixed mode (MT_F16_BF16: f16 q / k / v / E, bf16 O / dO / dq / dk / dv -- the first encoder layer, whose logits
// need an 11-bit mantissa) is served by the dS-spill variant only.  tcgen05.mma kind::f16 takes A and B of ONE
// format (a bf16 x f16 product is an illegal instruction on sm_100a), so every product of the call runs on f16
// operands: the delta kernel also writes dO_h = f16(g * dO) (g = 2^12, a static loss scale: activation gradients
// of the mean-reduced loss are ~1e-6, below f16's normal range), P and g * dS are packed as f16, and dQ / dK / dV /
// dE are multiplied by 1/g on the way out (exact: a power of two).
#include "ops.cuh"
#include "rga_tc_common.cuh"

#include <stdlib.h>

namespace mt {

using namespace rga;

int rga_bwd2_dkv(const RgaArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                 const CUtensorMap& tmDO, const CUtensorMap& tmE, void* ds_ws, int qk_fmt, float gscale, cudaStream_t st);      // rga_tc_bwd2.cu
int rga_bwd3_dq(const RgaArgs& a, const void* ws, const CUtensorMap& tmK, const CUtensorMap& tmE, int qk_fmt, float gscale, cudaStream_t st);   // rga_tc_bwd3.cu
int rga_bwd3_de(const RgaArgs& a, const void* ws, const CUtensorMap& tmQ, const CUtensorMap& tmE, int qk_fmt, float gscale, cudaStream_t st);
int rga_bwd3_dqe(const RgaArgs& a, const void* ws, const CUtensorMap& tmK, const CUtensorMap& tmE, const CUtensorMap& tmQ,
                 int qk_fmt, float gscale, cudaStream_t st);

int rga_bwd4_dkv(const RgaArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmV, const CUtensorMap& tmDO,
                 void* ds_ws, int qk_fmt, float gscale, cudaStream_t st);      // rga_tc_bwd4.cu

int rga_bwd4_dqe(const RgaArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                 const CUtensorMap& tmDO, const CUtensorMap& tmE, int qk_fmt, float gscale, cudaStream_t st);      // rga_tc_bwd4q.cu

constexpr float MIXED_GSCALE = 4096.f;
// bytes of the mixed mode's extra workspace region: the scaled f16 copy of a dense [B, L, h, dh] dO
size_t rga_bwd_mixed_extra_bytes(int64_t B, int64_t h, int64_t L, int64_t dh) { return (size_t)(B * L * h * dh) * 2; }
int rga_bwd2_de(const RgaArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                const CUtensorMap& tmDO, const CUtensorMap& tmE, cudaStream_t st);
int rga_bwd2_dq(const RgaArgs& a, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                const CUtensorMap& tmDO, const CUtensorMap& tmE, cudaStream_t st);

bool rga_bwd_tc_supported(const RgaArgs& a, int dh, int dtype) {
  if (dh != DHC || (dtype != MT_BF16 && dtype != MT_F16_BF16) || !a.causal) return false;
  if (a.sl % 8 || a.sh % 8 || a.sb % 8 || a.ol % 8 || a.oh % 8 || a.ob % 8) return false;
  if (!aligned(a.q, 16) || !aligned(a.k, 16) || !aligned(a.v, 16) || !aligned(a.E, 16) || !aligned(a.dO, 16) ||
      !aligned(a.dq, 16) || !aligned(a.dk, 16) || !aligned(a.dv, 16) || !aligned(a.dE, 16))
    return false;
  return mt_device_ok() != 0;
}

int rga_bwd_tc(const RgaArgs& a_in, int dh, int dtype, void* ws, size_t ws_bytes, cudaStream_t st) {
  int rc;
  RgaArgs a = a_in;
  // a workspace of rga_bwd3_workspace_bytes() selects the dS-spill variant (S/P/dS computed once, in the
  // dK/dV role); without it every role recomputes them
  const size_t tiles = rga_bwd3_workspace_bytes(a.B, a.h, a.L);
  const int qk_fmt = (dtype == MT_F16_BF16) ? 0 : 1;
  const size_t need = tiles + (qk_fmt == 0 ? rga_bwd_mixed_extra_bytes(a.B, a.h, a.L, dh) : 0);
  const bool spill = ws != nullptr && ws_bytes >= need && aligned(ws, 128);
  const float gscale = qk_fmt == 0 ? MIXED_GSCALE : 1.f;
  if (qk_fmt == 0) {
    if (!spill) {
      set_error("rga_bwd: the mixed f16/bf16 mode needs the workspace of mt_rga_bwd_workspace_bytes (%zu bytes)", need);
      return MT_E_WORKSPACE;
    }
    if (!(a.oh == dh && a.ol == (int64_t)a.h * dh && a.ob == (int64_t)a.L * a.h * dh)) {
      set_error("rga_bwd: the mixed f16/bf16 mode takes a dense [B, L, h, dh] O / dO");
      return MT_E_UNSUPPORTED;
    }
    a.dO_h = static_cast<uint8_t*>(ws) + tiles;         // (tiles is a multiple of 32 KB: 128-byte aligned)
    a.gscale = gscale;
  }
  if ((rc = rga_delta_launch(a, dh, MT_BF16, st))) return rc;          // O and dO are bf16 in both modes
  if (qk_fmt == 0) a.dO = a.dO_h;                                      // what the MMAs read from here on
  CUtensorMap tmQ, tmK, tmV, tmDO, tmE;
  if ((rc = tc::make_tmap_blhd(&tmQ, a.q, dh, a.L, a.h, a.B, a.sl, a.sh, a.sb, DHC, TT))) return rc;
  if ((rc = tc::make_tmap_blhd(&tmK, a.k, dh, a.L, a.h, a.B, a.sl, a.sh, a.sb, DHC, TT))) return rc;
  if ((rc = tc::make_tmap_blhd(&tmV, a.v, dh, a.L, a.h, a.B, a.sl, a.sh, a.sb, DHC, TT))) return rc;
  if ((rc = tc::make_tmap_blhd(&tmDO, a.dO, dh, a.L, a.h, a.B, a.ol, a.oh, a.ob, DHC, TT))) return rc;
  if ((rc = tc::make_tmap_2d(&tmE, a.E, a.max_seq, dh, dh, DHC, TT))) return rc;
  if (a.pstash) {
    // training step: the forward kept its P tiles -- both roles read them (rga_tc_bwd4.cu, rga_tc_bwd4q.cu), nothing is
    // spilled.  MT_RGA_STASH_SPILL=1 keeps the intermediate variant (dK/dV from the stash spilling dS, consumers of
    // rga_tc_bwd3.cu) for comparison.
    const char* vs = getenv("MT_RGA_STASH_SPILL");      // (read per call: tests/test_gpu_rga.py switches it)
    const bool via_spill = vs && atoi(vs) != 0;
    if (!via_spill) {
      if ((rc = rga_bwd4_dkv(a, tmQ, tmV, tmDO, nullptr, qk_fmt, gscale, st))) return rc;
      return rga_bwd4_dqe(a, tmQ, tmK, tmV, tmDO, tmE, qk_fmt, gscale, st);
    }
    if (!spill) { set_error("rga_bwd: the stash variant needs the workspace of mt_rga_bwd_workspace_bytes (%zu bytes)", need); return MT_E_WORKSPACE; }
    if ((rc = rga_bwd4_dkv(a, tmQ, tmV, tmDO, ws, qk_fmt, gscale, st))) return rc;
  } else if ((rc = rga_bwd2_dkv(a, tmQ, tmK, tmV, tmDO, tmE, spill ? ws : nullptr, qk_fmt, gscale, st))) return rc;
  if (spill) {
    // one consumer pass over the dS tiles (dQ and dE together); MT_RGA_SPLIT_CONSUMERS=1 keeps the two separate launches
    static const bool split = getenv("MT_RGA_SPLIT_CONSUMERS") && atoi(getenv("MT_RGA_SPLIT_CONSUMERS")) != 0;
    if (!split) return rga_bwd3_dqe(a, ws, tmK, tmE, tmQ, qk_fmt, gscale, st);
    if ((rc = rga_bwd3_dq(a, ws, tmK, tmE, qk_fmt, gscale, st))) return rc;
    return rga_bwd3_de(a, ws, tmQ, tmE, qk_fmt, gscale, st);
  }
  if ((rc = rga_bwd2_dq(a, tmQ, tmK, tmV, tmDO, tmE, st))) return rc;
  return rga_bwd2_de(a, tmQ, tmK, tmV, tmDO, tmE, st);
}

}  // namespace mt
