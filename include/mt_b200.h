/*
 * mt_b200.h -- C ABI of the B200-native MusicTransformer hot path (libmt_b200.so).
 *
 * The reference (SJTMusicTeam/MusicGeneration, mg/model/MusicTransformer -- "MT/" below) has no
 * FFI: its boundary is the Python module surface of MT/layers.py, MT/network.py and
 * MT/criterion.py.  Every entry point here replaces the group of eager PyTorch ops the cited
 * reference lines execute; the Python mirror in musicgeneration_b200/ binds them with ctypes
 * (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless named host_*;
 *   - the caller owns every buffer (inputs, outputs, workspaces) and keeps it alive until the
 *     stream reaches the op; the library allocates no device memory;
 *   - every op only enqueues work on `stream` (a cudaStream_t passed as void*); no host sync,
 *     CUDA-graph capturable;
 *   - return 0 on success, a negative MT_E_* code for argument errors, or a positive
 *     cudaError_t; mt_last_error() returns a thread-local message; nothing throws or exits;
 *   - dtype codes: MT_F32 = 0, MT_BF16 = 1, MT_F16 = 2.  "lp" = the low-precision activation
 *     type of the bf16 mode; in fp32 mode lp pointers are NULL or dtype is MT_F32.
 *     MT_F16_BF16 = 3 is accepted by mt_rga_fwd / mt_rga_bwd_ws (tcgen05 path) only: q, k, v and E are
 *     f16 (11-bit mantissa), every other 16-bit tensor of the call (O, dO, dq, dk, dv) is bf16 -- the
 *     mode of the FIRST encoder layer, whose input is the un-normalised embedding (MT/layers.py:226-229:
 *     |logit| ~ 1e3, a bf16 operand moves the near-one-hot softmax).  The
 *     backward needs the workspace of mt_rga_bwd_workspace_bytes(..., MT_F16_BF16) and a dense
 *     [B, L, h, dh] O / dO (it keeps a loss-scaled f16 copy of dO there: the tensor cores take one
 *     operand format per product).
 *   - tensors are row-major and dense unless strides are passed (strides are in ELEMENTS).
 */
#ifndef MT_B200_H_
#define MT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MT_F32 0
#define MT_BF16 1
#define MT_F16 2
#define MT_F16_BF16 3

#define MT_E_ARG (-1)      /* bad pointer / shape / alignment                */
#define MT_E_UNSUPPORTED (-2) /* combination not built (e.g. dh not in {32,64,128}) */
#define MT_E_WORKSPACE (-3) /* workspace too small                             */

/* GEMM epilogue flags */
#define MT_EPI_BIAS 1       /* + bias[n]                                        */
#define MT_EPI_RELU 2       /* max(.,0)                                         */
#define MT_EPI_ADD 4        /* + addend[m,n] (fp32, ld = ldc)                   */
#define MT_EPI_RELU_MASK 8  /* zero where aux[m,n] <= 0 (aux has the input dtype, ld = ldc) */

int mt_version(void);
const char* mt_last_error(void);
/* 1 when the running device is sm_100 and the tcgen05 paths are usable */
int mt_device_ok(void);

/* ---- K3: embedding * sqrt(d) + sinusoid + dropout  (MT/layers.py:226-229, :22-39) -------- */
int mt_embed_pos_fwd(const int32_t* ids, const float* emb, const float* pe, float* out_f32,
                     void* out_lp, int lp_dtype, int64_t B, int64_t L, int64_t d, int64_t V,
                     int64_t pos0, float scale, float p_drop, uint64_t seed, uint64_t site,
                     void* stream);
/* demb[V,d] += scatter(dout * scale * dropmask); demb must be zeroed/accumulated by caller */
int mt_embed_pos_bwd(const int32_t* ids, const float* dout, float* demb, int64_t B, int64_t L,
                     int64_t d, int64_t V, float scale, float p_drop, uint64_t seed,
                     uint64_t site, void* stream);

/* ---- K4: out = LayerNorm(dropout(a) + resid)  (MT/layers.py:154-155,159-160) ------------- */
int mt_add_ln_fwd(const void* a, int a_dtype, const float* resid, const float* gamma,
                  const float* beta, float* out_f32, void* out_lp, int lp_dtype, float* mean,
                  float* rstd, int64_t T, int64_t d, float eps, float p_drop, uint64_t seed,
                  uint64_t site, void* stream);
/* dz (fp32, may alias dout) = grad wrt (dropout(a)+resid); da (da_dtype) = dropmask * dz;
 * part[3, nparts, d] receives per-block partial sums of dgamma / dbeta / colsum(da) -- the last is
 * the bias gradient of the linear layer that produced `a` (MT/layers.py:153,158: fc, FFN_suf), so
 * no second pass over da is needed -- reduced by mt_ln_param_grad (dbias may be NULL).
 * nparts = mt_add_ln_bwd_parts(T). */
int64_t mt_add_ln_bwd_parts(int64_t T);
int mt_add_ln_bwd(const float* dout, const void* a, int a_dtype, const float* resid,
                  const float* gamma, const float* mean, const float* rstd, float* dz,
                  void* da, int da_dtype, float* part, int64_t T, int64_t d, float p_drop,
                  uint64_t seed, uint64_t site, void* stream);
int mt_ln_param_grad(const float* part, float* dgamma, float* dbeta, float* dbias, int64_t nparts,
                     int64_t d, void* stream);

/* ---- K5: C = epi(op(A)[M,K] . op(B)[K,N])  (nn.Linear fwd/dgrad/wgrad on the path) ------- */
/* transA: A stored [K,M] (lda = row pitch of the stored matrix); transB: B stored [N,K].
 * in_dtype applies to A, B (and aux); out_dtype to C.  path: 0 = auto, 1 = SIMT fp32-accumulate
 * reference-precision kernel, 2 = tcgen05 tensor-core kernel (bf16/f16 inputs only). */
size_t mt_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int in_dtype, int path);
int mt_gemm(const void* A, const void* B, void* C, const float* bias, const float* addend,
            const void* aux, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
            int64_t ldc, int transA, int transB, int in_dtype, int out_dtype, int epilogue,
            int path, void* workspace, size_t workspace_bytes, void* stream);
/* nn.Linear backward in one pass over dy: dW[M,N] = dy[K,M]^T . x[K,N] (fp32) and db[M] = sum_k dy[k,m].
 * The column sums ride along as one extra 128 x 16 x 16 tensor-core product per k-step against a tile of
 * ones in the CTAs that own the first tile column (no second read of dy, no extra launch).  tcgen05 path
 * only (bf16 / f16 operands, N % 128 == 0): MT_E_UNSUPPORTED otherwise -- callers then use mt_gemm +
 * mt_colsum.  Workspace: mt_gemm_workspace_bytes(M, N, K, in_dtype, 0). */
int mt_wgrad_bias(const void* dy, const void* x, float* dW, float* db, int64_t M, int64_t N, int64_t K,
                  int64_t lddy, int64_t ldx, int64_t lddw, int in_dtype, void* workspace, size_t workspace_bytes,
                  void* stream);
/* out[n] = sum_m X[m,n]  (bias gradients) */
size_t mt_colsum_workspace_bytes(int64_t M, int64_t N);
int mt_colsum(const void* X, int dtype, float* out, int64_t M, int64_t N, int64_t ldx,
              void* workspace, size_t workspace_bytes, void* stream);
int mt_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);
/* dst[r, 0:ldd] = {src[r, 0:cols], 0...}: row-padded operand copy (leading dimension -> multiple
 * of 8 elements, as the TMA-fed GEMM needs; e.g. dlogits [T, 390] -> [T, 392]) */
int mt_cast2d(const void* src, int src_dtype, int64_t lds, void* dst, int dst_dtype, int64_t ldd,
              int64_t rows, int64_t cols, void* stream);
/* dst[c,r] = src[r,c] with dtype conversion (weight shadows for dgrad) */
int mt_transpose_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t rows,
                      int64_t cols, void* stream);

/* ---- K1/K2: fused relative global attention  (MT/layers.py:86-106, :111-133) ------------- */
/* q,k,v: element (b,l,hh,dd) at base[b*sb + l*sl + hh*sh + dd]; O, dO: same addressing with
 * (ob, ol, oh).  E [max_seq, dh] (dtype).  pad_keys [B,L] uint8 or NULL.  lse [B,h,L] fp32.
 * S[i,j] = (q_i.k_j + [j<=i] q_i.E[max_seq-1-(i-j)]) / sqrt(dh); causal: keys j>i excluded.
 * path: 0 auto, 1 SIMT, 2 tcgen05. */
int mt_rga_fwd(const void* q, const void* k, const void* v, int64_t sb, int64_t sl, int64_t sh,
               const void* E, const uint8_t* pad_keys, void* O, int64_t ob, int64_t ol,
               int64_t oh, float* lse, int64_t B, int64_t h, int64_t L, int64_t dh,
               int64_t max_seq, int causal, int dtype, int path, void* stream);
/* attention weights P [B,h,L,L] fp32 (eval-mode return of MT/network.py:40) from q,k,E,lse */
int mt_rga_weights(const void* q, const void* k, int64_t sb, int64_t sl, int64_t sh,
                   const void* E, const uint8_t* pad_keys, const float* lse, float* P, int64_t B,
                   int64_t h, int64_t L, int64_t dh, int64_t max_seq, int causal, int dtype,
                   void* stream);
/* delta [B,h,L] fp32 workspace.  dq,dk,dv use the q/k/v strides.  dE [max_seq,dh] fp32 is
 * ACCUMULATED into (caller zeroes it once per layer). */
int mt_rga_bwd(const void* q, const void* k, const void* v, int64_t sb, int64_t sl, int64_t sh,
               const void* E, const uint8_t* pad_keys, const void* O, const void* dO, int64_t ob,
               int64_t ol, int64_t oh, const float* lse, float* delta, void* dq, void* dk,
               void* dv, float* dE, int64_t B, int64_t h, int64_t L, int64_t dh,
               int64_t max_seq, int causal, int dtype, int path, void* stream);
/* Same with a caller-owned scratch buffer.  On the tcgen05 path a workspace of at least
 * mt_rga_bwd_workspace_bytes() (128-byte aligned) selects the "dS-spill" variant: the dK/dV
 * kernel computes S, the skew, P and dS once and writes every dS tile (bf16) to the workspace,
 * the dQ and dE kernels only consume them.  NULL / too small: each kernel recomputes (no
 * scratch, ~1.4x the time).  The contents are dead when the call's work has run; the buffer
 * may be shared by all layers of a model on one stream.  Results are identical to mt_rga_bwd
 * up to fp32 summation order. */
int mt_rga_bwd_ws(const void* q, const void* k, const void* v, int64_t sb, int64_t sl, int64_t sh,
                  const void* E, const uint8_t* pad_keys, const void* O, const void* dO, int64_t ob,
                  int64_t ol, int64_t oh, const float* lse, float* delta, void* dq, void* dk,
                  void* dv, float* dE, int64_t B, int64_t h, int64_t L, int64_t dh,
                  int64_t max_seq, int causal, int dtype, int path, void* workspace,
                  size_t workspace_bytes, void* stream);
/* B*h*nT*(nT+1)/2 tiles of 32 KB, nT = ceil(L/128) (+ B*L*h*dh*2 bytes for MT_F16_BF16); 0 when the
 * tcgen05 backward does not take the problem (dh != 64, or dtype not MT_BF16 / MT_F16_BF16) */
size_t mt_rga_bwd_workspace_bytes(int64_t B, int64_t h, int64_t L, int64_t dh, int dtype);

/* Training pair of the tcgen05 path (head dim 64, causal, MT_BF16 or MT_F16_BF16).  The reference keeps the
 * attention weights of every layer alive between forward and backward through autograd (MT/layers.py:97-99:
 * softmax output saved for the backward, [B,h,L,L] fp32); here the forward keeps its P tiles -- the causal half,
 * 16-bit, in the UMMA operand layout the backward's products read, plus the online-softmax row reference of each
 * tile -- in a caller-owned stash of mt_rga_stash_bytes() bytes (128-byte aligned; one per layer, alive until
 * that layer's backward), and the backward reads them instead of rebuilding S, the skew and the exponentials.
 * mt_rga_bwd_stash also takes the scratch of mt_rga_bwd_workspace_bytes() (shared by all layers). */
size_t mt_rga_stash_bytes(int64_t B, int64_t h, int64_t L, int64_t dh, int dtype);
int mt_rga_fwd_stash(const void* q, const void* k, const void* v, int64_t sb, int64_t sl, int64_t sh,
                     const void* E, const uint8_t* pad_keys, void* O, int64_t ob, int64_t ol,
                     int64_t oh, float* lse, int64_t B, int64_t h, int64_t L, int64_t dh,
                     int64_t max_seq, int causal, int dtype, void* stash, size_t stash_bytes, void* stream);
int mt_rga_bwd_stash(const void* q, const void* k, const void* v, int64_t sb, int64_t sl, int64_t sh,
                     const void* E, const uint8_t* pad_keys, const void* O, const void* dO, int64_t ob,
                     int64_t ol, int64_t oh, const float* lse, float* delta, void* dq, void* dk,
                     void* dv, float* dE, int64_t B, int64_t h, int64_t L, int64_t dh,
                     int64_t max_seq, int causal, int dtype, const void* stash, size_t stash_bytes,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- K6: label-smoothed cross entropy + step metrics  (MT/criterion.py:43-67, ------------
 *          MT/metrics.py:50-60) */
/* row_lse: 3*T floats (lse, then per-row loss and per-row flags used by the reduction);
 * argmax[T]; sums[4] = {loss_sum, n_valid, n_correct(all positions), loss_mean} */
int mt_smooth_ce_fwd(const float* logits, const int32_t* target, float* row_lse,
                     int32_t* argmax, float* sums, int64_t T, int64_t V, float eps,
                     int32_t ignore, void* stream);
/* dlogits = grad_out * (softmax(z) - q') / n_valid on rows with target != ignore, else 0 */
int mt_smooth_ce_bwd(const float* logits, const int32_t* target, const float* row_lse,
                     const float* sums, const float* grad_out, float* dlogits, int64_t T,
                     int64_t V, float eps, int32_t ignore, void* stream);

/* ---- optimizer ("next" row 1: MT/train.py:143, MT/criterion.py:81-88) -------------------- */
/* Adam (torch.optim.Adam semantics, no weight decay / amsgrad) over a flat fp32 buffer;
 * grad_scale multiplies g first (1/world, 1/accum).  p_lp (bf16 shadow) may be NULL. */
int mt_adam_step(float* p, const float* g, float* m, float* v, void* p_lp, int64_t n, float lr,
                 float beta1, float beta2, float eps, int64_t step, float grad_scale,
                 void* stream);

/* ---- K7/K8: KV-cached single-token decode  (MT/network.py:52-77, causal oracle) ---------- */
/* One decode step for the whole stack is orchestrated by the host mirror from these ops. */
/* s_j = q.(k_j + E[max_seq-1-(t-j)]) / sqrt(dh), j = 0..t; softmax; .V   -- one new token per
 * sequence.  q: element (b,hh,dd) at q[b*q_stride_b + hh*dh + dd] (dtype); kcache/vcache
 * [B, h, max_seq, dh] (dtype), already holding position t; pad_keys [B, max_seq] uint8 or NULL
 * (1 = the token at that position is the pad token: key excluded, MT/utils.py:73);
 * out [B,h,dh] dense (dtype). */
int mt_rga_decode(const void* q, int64_t q_stride_b, const void* kcache, const void* vcache,
                  const void* E, const uint8_t* pad_keys, void* out, int64_t B, int64_t h, int64_t dh, int64_t max_seq,
                  int64_t t, int dtype, void* stream);
/* writes k,v [B,h,dh] of the new token into the caches at position t */
int mt_kv_append(const void* qkv, void* kcache, void* vcache, int64_t B, int64_t h, int64_t dh,
                 int64_t max_seq, int64_t t, int dtype, void* stream);
/* ids_out[b] = sample(logits[b,:]/temperature restricted to top_k) using uniforms u[b];
 * top_k <= 0 or >= V: full distribution; greedy != 0: argmax (first max).  */
int mt_sample(const float* logits, const float* u, int32_t* ids_out, int64_t B, int64_t V,
              float temperature, int32_t top_k, int greedy, void* stream);

/* ---- decode with a DEVICE-RESIDENT step index ------------------------------------------------
 * Same arithmetic as the ops above, but the current position t is read from *t_dev, so ONE CUDA
 * graph of a whole decode step (embed -> per layer: QKV GEMM, append, attend, fc, LN, FFN, LN ->
 * vocabulary GEMM -> sample -> advance) can be replayed for every generated event without any
 * host work in between (MT/network.py:52-77 is a Python loop of full-stack recomputes).
 * ids [B, ld_ids] int32 holds prior + generated tokens; pad_bits [B, max_seq]. */
/* pad_bits (optional): pad_bits[b, t] = (ids[b, t] == pad_token), the key mask of position t */
int mt_decode_embed(const int32_t* ids, int64_t ld_ids, const int32_t* t_dev, const float* emb,
                    const float* pe, float* out_f32, void* out_lp, int lp_dtype, int64_t B, int64_t d,
                    int64_t V, float scale, int32_t pad_token, uint8_t* pad_bits, int64_t max_seq,
                    void* stream);
int mt_decode_kv_append(const void* qkv, void* kcache, void* vcache, const int32_t* ids, int64_t ld_ids,
                        int32_t pad_token, uint8_t* pad_bits, const int32_t* t_dev, int64_t B, int64_t h,
                        int64_t dh, int64_t max_seq, int dtype, void* stream);
/* split-context kernel: one CTA per (head, sequence, 256-key chunk); `workspace`
 * (mt_decode_attend_workspace_bytes, 16-byte aligned) holds the per-chunk partials and one arrival
 * counter per (sequence, head); the caller zeroes it ONCE, the kernel leaves the counters at zero. */
size_t mt_decode_attend_workspace_bytes(int64_t B, int64_t h, int64_t dh, int64_t max_seq);
/* append != 0: q is the fused projection row [3, h, dh] of the new token (q_stride_b apart); its
 * K / V rows are read from there and stored into the caches at position t by this launch. */
int mt_decode_attend(const void* q, int64_t q_stride_b, void* kcache, void* vcache, const void* E,
                     const uint8_t* pad_bits, void* out, const int32_t* t_dev, int64_t B, int64_t h, int64_t dh,
                     int64_t max_seq, int dtype, int append, void* workspace, size_t workspace_bytes,
                     void* stream);
/* writes the sampled id to ids[b, t+1] unless t+1 < prior_len (prior tokens are kept); u holds one
 * row of B uniforms per generated event (row t+1-prior_len) */
int mt_decode_sample(const float* logits, const float* u, int32_t* ids, int64_t ld_ids, const int32_t* t_dev,
                     int32_t prior_len, int64_t B, int64_t V, float temperature, int32_t top_k, int greedy,
                     void* stream);
int mt_decode_advance(int32_t* t_dev, void* stream);

/* ---- the whole generation as ONE persistent launch (MT/network.py:52-77) --------------------------------------------
 * One CTA per SM stays resident and walks, for positions t0 .. t0 + n_steps - 1 of `ids` [B, ld_ids]: embedding + PE ->
 * every encoder layer (QKV projection, KV-cached relative attention incl. the append of the new K / V rows, fc,
 * residual + LayerNorm, FFN, residual + LayerNorm) -> vocabulary projection -> sampler (mt_sample's arithmetic; the
 * event drawn at position t becomes the token at t + 1 unless t + 1 < prior_len), separated by grid-wide barriers.
 * 16-bit mode only (B <= 64, head dim 64, d <= 1024): weights bf16 ([N, K] row-major), except that a layer
 * with layer_f16[l] != 0 has f16 Wqkv / Wfc / E / KV cache (the first layer of the bf16 mode).
 * layer_ptrs: [layers][15] device pointers {Wqkv, bqkv, Wfc, bfc, Wpre, bpre, Wsuf, bsuf, g1, b1, g2, b2, E, kcache,
 * vcache} (biases / LayerNorm parameters fp32; caches [B, h, max_seq, 64]).  uniforms [n_steps.., B] as mt_decode_sample.
 * logits_out (optional) [n_steps, B, V] receives every step's logits.  workspace: mt_decode_run_workspace_bytes(),
 * 256-byte aligned.  Cooperative launch: needs the device to itself for the duration of the call. */
size_t mt_decode_run_workspace_bytes(int64_t B, int64_t d, int64_t V);
int mt_decode_run_supported(int64_t B, int64_t d, int64_t h, int64_t V, int64_t layers);
int mt_decode_run(int32_t* ids, int64_t ld_ids, int64_t B, int64_t t0, int64_t n_steps, int64_t prior_len,
                  const float* emb, const float* pe, const void* const* layer_ptrs, const int32_t* layer_f16,
                  int64_t layers, const void* Wv, const float* bv, int64_t d, int64_t h, int64_t V, int64_t max_seq,
                  int32_t pad_token, uint8_t* pad_bits, const float* uniforms, float temperature, int32_t top_k,
                  int greedy, float* logits_out, void* workspace, size_t workspace_bytes, void* stream);
/* Programmatic dependent launch for the kernels of a decode step (strip GEMM, residual+LayerNorm,
 * mt_decode_*): while enabled, each of them may start launching before its predecessor in the
 * stream has drained and waits for it on the device (griddepcontrol), which hides most of the
 * per-launch latency of the ~46 dependent kernels of a step.  Process-wide switch; leave it off
 * outside a decode step. */
int mt_decode_chain(int enable);

/* ---- data feed: HBM-resident token arena ------------------------------------
 * Replaces Data.batch / slide_seq2seq_batch / seq2seq_batch (MT/data.py:41-67) + the numpy int16 ->
 * device int32 conversion of MT/train.py:258-260.  All `.data` files are concatenated once into one
 * uint8 / uint16 arena on the device; file f is arena[file_off[f] .. file_off[f+1]).
 *
 * mt_window_gather: x[b, t] = arena[starts[b] + t] (t < L);  y[b, t] = arena[starts[b] + y_shift + t]
 * (t < y_len; y may be NULL).  slide_seq2seq_batch: y_len = L, y_shift = 1; seq2seq_batch: y_len = L,
 * y_shift = L.  `starts` is a device array of B arena offsets (drawn by the host with the reference's
 * `random` call sequence, or by mt_window_sample).  Bit-exact integer copy, one launch, no sync. */
int mt_window_gather(const void* arena, int token_bytes, const int64_t* starts, int32_t* x, int32_t* y,
                     int64_t B, int64_t L, int64_t y_len, int64_t y_shift, void* stream);
/* mt_window_sample: on-device counterpart of random.sample(files, B) + random.randrange(0, len - need)
 * (MT/data.py:42,100): row b takes file eligible[perm(b)] (perm = keyed bijection of [0, n_eligible),
 * i.e. without replacement) and a uniform window start in [0, len_f - need); every eligible file must
 * be longer than `need`.  Deterministic in (seed, step); writes arena offsets to starts[B] and, if
 * files != NULL, the chosen file indices.  Not the Mersenne-Twister stream of the reference: same
 * distribution, different draws (the host-drawn path is the bit-exact one). */
int mt_window_sample(const int64_t* file_off, const int64_t* eligible, int64_t n_eligible, int64_t need,
                     uint64_t seed, uint64_t step, int64_t* starts, int64_t* files, int64_t B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MT_B200_H_ */
