// C-ABI entry points for the GEMM and relative-attention ops: argument validation and path
// selection (SIMT reference-precision kernels vs tcgen05 tensor-core kernels).
#include "ops.cuh"



using namespace mt;

extern "C" {

size_t mt_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int in_dtype, int path) {
  size_t a = gemm_simt_workspace_bytes(M, N, K);
  size_t b = (in_dtype != MT_F32 && path != 1) ? gemm_tc_workspace_bytes(M, N, K) : 0;
  return a > b ? a : b;
}

int mt_gemm(const void* A, const void* B, void* C, const float* bias, const float* addend,
            const void* aux, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
            int64_t ldc, int transA, int transB, int in_dtype, int out_dtype, int epilogue,
            int path, void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(A && B && C, "gemm: null pointer");
  MT_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%ld N=%ld K=%ld", (long)M, (long)N, (long)K);
  MT_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "gemm: leading dimension too small");
  MT_REQUIRE(!(epilogue & MT_EPI_BIAS) || bias, "gemm: BIAS epilogue without bias");
  MT_REQUIRE(!(epilogue & MT_EPI_ADD) || addend, "gemm: ADD epilogue without addend");
  MT_REQUIRE(!(epilogue & MT_EPI_RELU_MASK) || aux, "gemm: RELU_MASK epilogue without aux");
  MT_REQUIRE(path >= 0 && path <= 2, "gemm: bad path %d", path);
  bool tc_ok = gemm_tc_supported(M, N, K, lda, ldb, ldc, transA, transB, in_dtype, out_dtype, epilogue, A, B, C);
  if (path == 2 && !tc_ok) {
    set_error("gemm: tcgen05 path does not take this problem (M=%ld N=%ld K=%ld tA=%d tB=%d in=%d out=%d)", (long)M, (long)N, (long)K, transA, transB, in_dtype, out_dtype);
    return MT_E_UNSUPPORTED;
  }
  // decode-sized problems (a few dozen activation rows): weight-streaming strip kernel
  if (path == 0 && gemm_skinny_supported(M, N, K, lda, ldb, transA, transB, in_dtype, out_dtype, epilogue, A, B))
    return gemm_skinny(A, B, C, bias, M, N, K, lda, ldb, ldc, in_dtype, out_dtype, epilogue, as_stream(stream));
  if (path == 2 || (path == 0 && tc_ok))
    return gemm_tc(A, B, C, bias, addend, aux, M, N, K, lda, ldb, ldc, transA, transB, in_dtype, out_dtype, epilogue, workspace, workspace_bytes, as_stream(stream));
  return gemm_simt(A, B, C, bias, addend, aux, M, N, K, lda, ldb, ldc, transA, transB, in_dtype, out_dtype, epilogue, workspace, workspace_bytes, as_stream(stream));
}

int mt_wgrad_bias(const void* dy, const void* x, float* dW, float* db, int64_t M, int64_t N, int64_t K,
                  int64_t lddy, int64_t ldx, int64_t lddw, int in_dtype, void* workspace, size_t workspace_bytes,
                  void* stream) {
  MT_REQUIRE(dy && x && dW && db, "wgrad_bias: null pointer");
  MT_REQUIRE(M > 0 && N > 0 && K > 0 && lddy >= M && ldx >= N && lddw >= N, "wgrad_bias: bad shape");
  if (N % 128 != 0 || !gemm_tc_supported(M, N, K, lddy, ldx, lddw, 1, 0, in_dtype, MT_F32, 0, dy, x, dW)) {
    set_error("wgrad_bias: tcgen05 weight-gradient form only (16-bit operands, N %% 128 == 0); use mt_gemm + mt_colsum");
    return MT_E_UNSUPPORTED;
  }
  return gemm_tc(dy, x, dW, nullptr, nullptr, nullptr, M, N, K, lddy, ldx, lddw, 1, 0, in_dtype, MT_F32, 0, workspace,
                 workspace_bytes, as_stream(stream), db);
}

static int fill_rga(RgaArgs& a, const void* q, const void* k, const void* v, int64_t sb, int64_t sl,
                    int64_t sh, const void* E, const uint8_t* pad, int64_t B, int64_t h, int64_t L,
                    int64_t dh, int64_t max_seq, int causal, int dtype) {
  MT_REQUIRE(q && k && E, "rga: null pointer");
  MT_REQUIRE(B > 0 && h > 0 && L > 0 && dh > 0 && max_seq >= L, "rga: bad shape B=%ld h=%ld L=%ld dh=%ld max_seq=%ld", (long)B, (long)h, (long)L, (long)dh, (long)max_seq);
  MT_REQUIRE(B <= 65535 && h <= 65535, "rga: B and h must be <= 65535");
  int es = dtype_size(dtype);
  MT_REQUIRE(sl % 4 == 0 && sb % 4 == 0 && sh % 4 == 0, "rga: strides must be multiples of 4 elements");
  MT_REQUIRE(aligned(q, 4 * es) && aligned(k, 4 * es) && (!v || aligned(v, 4 * es)) && aligned(E, 4 * es), "rga: misaligned q/k/v/E");
  a = RgaArgs{};
  a.q = q; a.k = k; a.v = v; a.sb = sb; a.sl = sl; a.sh = sh; a.E = E; a.pad = pad;
  a.B = (int)B; a.h = (int)h; a.L = (int)L; a.max_seq = (int)max_seq; a.causal = causal;
  a.inv_scale_div = sqrtf((float)dh);
  return 0;
}

int mt_rga_fwd(const void* q, const void* k, const void* v, int64_t sb, int64_t sl, int64_t sh,
               const void* E, const uint8_t* pad_keys, void* O, int64_t ob, int64_t ol,
               int64_t oh, float* lse, int64_t B, int64_t h, int64_t L, int64_t dh,
               int64_t max_seq, int causal, int dtype, int path, void* stream) {
  RgaArgs a;
  int rc = fill_rga(a, q, k, v, sb, sl, sh, E, pad_keys, B, h, L, dh, max_seq, causal, dtype);
  if (rc) return rc;
  MT_REQUIRE(v && O && lse, "rga_fwd: null pointer");
  MT_REQUIRE(path >= 0 && path <= 2, "rga_fwd: bad path %d", path);
  a.O = O; a.ob = ob; a.ol = ol; a.oh = oh; a.lse = lse;
  bool tc_ok = rga_tc_supported(a, (int)dh, dtype, false);
  if (path == 2 && !tc_ok) { set_error("rga_fwd: tcgen05 path does not take this problem"); return MT_E_UNSUPPORTED; }
  if (path == 2 || (path == 0 && tc_ok)) return rga_fwd_tc(a, (int)dh, dtype, as_stream(stream));
  if (dtype == MT_F16_BF16) { set_error("rga_fwd: the mixed f16/bf16 mode exists on the tcgen05 path only"); return MT_E_UNSUPPORTED; }
  return rga_fwd_simt(a, (int)dh, dtype, as_stream(stream));
}

size_t mt_rga_stash_bytes(int64_t B, int64_t h, int64_t L, int64_t dh, int dtype) {
  if (dh != 64 || (dtype != MT_BF16 && dtype != MT_F16_BF16) || B <= 0 || h <= 0 || L <= 0) return 0;
  return rga_stash_bytes(B, h, L);
}

int mt_rga_fwd_stash(const void* q, const void* k, const void* v, int64_t sb, int64_t sl, int64_t sh,
                     const void* E, const uint8_t* pad_keys, void* O, int64_t ob, int64_t ol,
                     int64_t oh, float* lse, int64_t B, int64_t h, int64_t L, int64_t dh,
                     int64_t max_seq, int causal, int dtype, void* stash, size_t stash_bytes, void* stream) {
  RgaArgs a;
  int rc = fill_rga(a, q, k, v, sb, sl, sh, E, pad_keys, B, h, L, dh, max_seq, causal, dtype);
  if (rc) return rc;
  MT_REQUIRE(v && O && lse && stash, "rga_fwd_stash: null pointer");
  a.O = O; a.ob = ob; a.ol = ol; a.oh = oh; a.lse = lse;
  a.pstash = stash; a.pstash_bytes = stash_bytes;
  if (!rga_tc_supported(a, (int)dh, dtype, false) || !rga_tc_supported(a, (int)dh, dtype == MT_F16 ? MT_BF16 : dtype, true) || !causal) {
    set_error("rga_fwd_stash: the tcgen05 training path does not take this problem (head dim 64, causal, bf16 / f16+bf16)");
    return MT_E_UNSUPPORTED;
  }
  return rga_fwd_tc(a, (int)dh, dtype, as_stream(stream));
}

int mt_rga_weights(const void* q, const void* k, int64_t sb, int64_t sl, int64_t sh,
                   const void* E, const uint8_t* pad_keys, const float* lse, float* P, int64_t B,
                   int64_t h, int64_t L, int64_t dh, int64_t max_seq, int causal, int dtype,
                   void* stream) {
  RgaArgs a;
  int rc = fill_rga(a, q, k, nullptr, sb, sl, sh, E, pad_keys, B, h, L, dh, max_seq, causal, dtype);
  if (rc) return rc;
  MT_REQUIRE(lse && P, "rga_weights: null pointer");
  MT_REQUIRE(B * h <= 65535, "rga_weights: B*h must be <= 65535");
  a.lse = const_cast<float*>(lse); a.P = P;
  return rga_weights_simt(a, (int)dh, dtype, as_stream(stream));
}

size_t mt_rga_bwd_workspace_bytes(int64_t B, int64_t h, int64_t L, int64_t dh, int dtype) {
  if (dh != 64 || (dtype != MT_BF16 && dtype != MT_F16_BF16) || B <= 0 || h <= 0 || L <= 0) return 0;
  return rga_bwd3_workspace_bytes(B, h, L) + (dtype == MT_F16_BF16 ? rga_bwd_mixed_extra_bytes(B, h, L, dh) : 0);
}

int mt_rga_bwd(const void* q, const void* k, const void* v, int64_t sb, int64_t sl, int64_t sh,
               const void* E, const uint8_t* pad_keys, const void* O, const void* dO, int64_t ob,
               int64_t ol, int64_t oh, const float* lse, float* delta, void* dq, void* dk,
               void* dv, float* dE, int64_t B, int64_t h, int64_t L, int64_t dh,
               int64_t max_seq, int causal, int dtype, int path, void* stream) {
  return mt_rga_bwd_ws(q, k, v, sb, sl, sh, E, pad_keys, O, dO, ob, ol, oh, lse, delta, dq, dk, dv, dE, B, h, L,
                       dh, max_seq, causal, dtype, path, nullptr, 0, stream);
}

int mt_rga_bwd_ws(const void* q, const void* k, const void* v, int64_t sb, int64_t sl, int64_t sh,
                  const void* E, const uint8_t* pad_keys, const void* O, const void* dO, int64_t ob,
                  int64_t ol, int64_t oh, const float* lse, float* delta, void* dq, void* dk,
                  void* dv, float* dE, int64_t B, int64_t h, int64_t L, int64_t dh,
                  int64_t max_seq, int causal, int dtype, int path, void* workspace,
                  size_t workspace_bytes, void* stream) {
  RgaArgs a;
  int rc = fill_rga(a, q, k, v, sb, sl, sh, E, pad_keys, B, h, L, dh, max_seq, causal, dtype);
  if (rc) return rc;
  MT_REQUIRE(v && O && dO && lse && delta && dq && dk && dv && dE, "rga_bwd: null pointer");
  MT_REQUIRE(path >= 0 && path <= 2, "rga_bwd: bad path %d", path);
  a.O = const_cast<void*>(O); a.dO = dO; a.ob = ob; a.ol = ol; a.oh = oh;
  a.lse = const_cast<float*>(lse); a.delta = delta; a.dq = dq; a.dk = dk; a.dv = dv; a.dE = dE;
  bool tc_ok = rga_tc_supported(a, (int)dh, dtype, true);
  if (path == 2 && !tc_ok) { set_error("rga_bwd: tcgen05 path does not take this problem"); return MT_E_UNSUPPORTED; }
  if (path == 2 || (path == 0 && tc_ok)) return rga_bwd_tc(a, (int)dh, dtype, workspace, workspace_bytes, as_stream(stream));
  if (dtype == MT_F16_BF16) { set_error("rga_bwd: the mixed f16/bf16 mode exists on the tcgen05 path only"); return MT_E_UNSUPPORTED; }
  return rga_bwd_simt(a, (int)dh, dtype, as_stream(stream));
}

int mt_rga_bwd_stash(const void* q, const void* k, const void* v, int64_t sb, int64_t sl, int64_t sh,
                     const void* E, const uint8_t* pad_keys, const void* O, const void* dO, int64_t ob,
                     int64_t ol, int64_t oh, const float* lse, float* delta, void* dq, void* dk,
                     void* dv, float* dE, int64_t B, int64_t h, int64_t L, int64_t dh,
                     int64_t max_seq, int causal, int dtype, const void* stash, size_t stash_bytes,
                     void* workspace, size_t workspace_bytes, void* stream) {
  RgaArgs a;
  int rc = fill_rga(a, q, k, v, sb, sl, sh, E, pad_keys, B, h, L, dh, max_seq, causal, dtype);
  if (rc) return rc;
  MT_REQUIRE(v && O && dO && lse && delta && dq && dk && dv && dE && stash, "rga_bwd_stash: null pointer");
  a.O = const_cast<void*>(O); a.dO = dO; a.ob = ob; a.ol = ol; a.oh = oh;
  a.lse = const_cast<float*>(lse); a.delta = delta; a.dq = dq; a.dk = dk; a.dv = dv; a.dE = dE;
  a.pstash = const_cast<void*>(stash); a.pstash_bytes = stash_bytes;
  if (!rga_tc_supported(a, (int)dh, dtype, true)) { set_error("rga_bwd_stash: tcgen05 path does not take this problem"); return MT_E_UNSUPPORTED; }
  MT_REQUIRE(stash_bytes >= rga_stash_bytes(B, h, L) && aligned(stash, 128), "rga_bwd_stash: the P stash needs %zu bytes, 128-byte aligned",
             rga_stash_bytes(B, h, L));
  return rga_bwd_tc(a, (int)dh, dtype, workspace, workspace_bytes, as_stream(stream));
}

}  // extern "C"
