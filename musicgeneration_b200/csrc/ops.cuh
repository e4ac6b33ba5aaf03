// Internal declarations shared by the launchers (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace mt {

struct RgaArgs {
  const void* q; const void* k; const void* v;
  int64_t sb, sl, sh;
  const void* E; const uint8_t* pad;
  void* O; const void* dO; int64_t ob, ol, oh;
  float* lse; float* delta; float* P;
  void* dq; void* dk; void* dv; float* dE;
  int B, h, L, max_seq, causal;
  float inv_scale_div;
  // f16 gradient mode of the tcgen05 backward (MT_F16_BF16): the delta kernel also writes dO_h = f16(gscale * dO)
  // with dO's own addressing; the MMAs then read dO_h
  void* dO_h; float gscale;
  // training forward / backward of the tcgen05 path: the forward keeps its P tiles here and the backward reads them
  // (rga_stash_bytes(); NULL = inference forward / backward that rebuilds P)
  void* pstash; size_t pstash_bytes;
};

// gemm_simt.cu
size_t gemm_simt_workspace_bytes(int64_t M, int64_t N, int64_t K);
int gemm_simt(const void* A, const void* B, void* C, const float* bias, const float* addend,
              const void* aux, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
              int64_t ldc, int transA, int transB, int in_dtype, int out_dtype, int epilogue,
              void* workspace, size_t workspace_bytes, cudaStream_t stream);
// rga_simt.cu
int rga_fwd_simt(const RgaArgs& a, int dh, int dtype, cudaStream_t st);
int rga_weights_simt(const RgaArgs& a, int dh, int dtype, cudaStream_t st);
int rga_bwd_simt(const RgaArgs& a, int dh, int dtype, cudaStream_t st);
int rga_delta_launch(const RgaArgs& a, int dh, int dtype, cudaStream_t st);   // delta = rowsum(dO * O)
// gemm_tc.cu / rga_tc.cu (tcgen05)
bool gemm_tc_supported(int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                       int transA, int transB, int in_dtype, int out_dtype, int epilogue,
                       const void* A, const void* B, const void* C);
size_t gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t K);
int gemm_tc(const void* A, const void* B, void* C, const float* bias, const float* addend,
            const void* aux, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
            int64_t ldc, int transA, int transB, int in_dtype, int out_dtype, int epilogue,
            void* workspace, size_t workspace_bytes, cudaStream_t stream, float* colsum_out = nullptr);
// gemm_skinny.cu (decode-sized M: mma.sync strip kernel)
bool gemm_skinny_supported(int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int transA, int transB,
                           int in_dtype, int out_dtype, int epilogue, const void* A, const void* B);
int gemm_skinny(const void* A, const void* B, void* C, const float* bias, int64_t M, int64_t N, int64_t K,
                int64_t lda, int64_t ldb, int64_t ldc, int in_dtype, int out_dtype, int epilogue, cudaStream_t st);
bool rga_tc_supported(const RgaArgs& a, int dh, int dtype, bool backward);
int rga_fwd_tc(const RgaArgs& a, int dh, int dtype, cudaStream_t st);
int rga_bwd_tc(const RgaArgs& a, int dh, int dtype, void* ws, size_t ws_bytes, cudaStream_t st);
size_t rga_stash_bytes(int64_t B, int64_t h, int64_t L);            // P stash of a training forward (rga_tc.cu)
size_t rga_bwd3_workspace_bytes(int64_t B, int64_t h, int64_t L);   // dS-spill workspace of the tcgen05 backward
size_t rga_bwd_mixed_extra_bytes(int64_t B, int64_t h, int64_t L, int64_t dh);   // + the scaled f16 copy of dO (MT_F16_BF16)

}  // namespace mt
