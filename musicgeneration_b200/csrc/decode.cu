// KV-cached single-token relative attention (K7) and temperature / top-k sampling (K8).
// Replaces the O(len^2) full-stack recompute of MT/network.py:52-62 and the
// OneHotCategorical draw of :73-74.  For the new token at position t:
//     s_j = q_t . (k_j + E[max_seq-1-(t-j)]) / sqrt(dh),  j = 0..t ;  out = softmax(s) . V
// HBM-bound: each (sequence, head) CTA streams its K and V rows once.
#include "common.cuh"

namespace mt {

template <typename T, int DH>
__global__ void __launch_bounds__(128)
rga_decode_kernel(const T* __restrict__ q, const T* __restrict__ kc, const T* __restrict__ vc,
                  const T* __restrict__ E, const uint8_t* __restrict__ pad_keys, T* __restrict__ out,
                  int64_t q_stride_b, int h, int max_seq, int t_host, const int32_t* __restrict__ t_dev,
                  float sqrt_dh) {
  extern __shared__ float sm[];
  float* qs = sm;              // [DH]
  float* red = qs + DH;        // [128]
  float* sc = red + 128;       // [t+1]
  const int t = t_dev ? *t_dev : t_host;     // device-resident step index: CUDA-graph replayable
  const int tid = threadIdx.x, hh = blockIdx.x, b = blockIdx.y;
  const int n = t + 1;
  const int64_t bh = (int64_t)b * h + hh;
  if (tid < DH) qs[tid] = to_f<T>(q[(int64_t)b * q_stride_b + (int64_t)hh * DH + tid]);
  __syncthreads();
  const uint8_t* pad = pad_keys ? pad_keys + (int64_t)b * max_seq : nullptr;
  const T* kb = kc + bh * (int64_t)max_seq * DH;
  const T* vb = vc + bh * (int64_t)max_seq * DH;
  float mx = -INFINITY;
  for (int j = tid; j < n; j += 128) {
    const T* kp = kb + (int64_t)j * DH;
    const T* ep = E + (int64_t)(max_seq - 1 - (t - j)) * DH;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      float4 kv = load4<T>(kp + d), ev = load4<T>(ep + d);
      acc = fmaf(qs[d + 0], kv.x + ev.x, acc);
      acc = fmaf(qs[d + 1], kv.y + ev.y, acc);
      acc = fmaf(qs[d + 2], kv.z + ev.z, acc);
      acc = fmaf(qs[d + 3], kv.w + ev.w, acc);
    }
    acc = (pad && pad[j]) ? -INFINITY : acc / sqrt_dh;
    sc[j] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < n; j += 128) {
    float pv = (sc[j] == -INFINITY) ? 0.f : expf(sc[j] - mx);
    sc[j] = pv;
    sum += pv;
  }
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  sum = (red[0] + red[1]) + (red[2] + red[3]);
  __syncthreads();
  // out[d] = sum_j p_j v_j[d] / sum ; 128/DH groups split the keys
  constexpr int NG = 128 / DH > 0 ? 128 / DH : 1;
  if (DH <= 128) {
    const int d = tid % DH, grp = tid / DH;
    float acc = 0.f;
    if (grp < NG)
      for (int j = grp; j < n; j += NG) acc = fmaf(sc[j], to_f<T>(vb[(int64_t)j * DH + d]), acc);
    red[tid] = acc;
    __syncthreads();
    if (tid < DH) {
      float tot = 0.f;
#pragma unroll
      for (int g = 0; g < NG; ++g) tot += red[g * DH + tid];
      out[bh * DH + tid] = from_f<T>(sum > 0.f ? tot / sum : 0.f);
    }
  }
}

// qkv row layout of the fused projection: [B, 3, h, dh] -> caches [B, h, max_seq, dh] at t
template <typename T>
__global__ void kv_append_kernel(const T* __restrict__ qkv, T* __restrict__ kc, T* __restrict__ vc,
                                 int B, int h, int dh, int max_seq, int t_host,
                                 const int32_t* __restrict__ t_dev, const int32_t* __restrict__ ids,
                                 int64_t ld_ids, int32_t pad_token, uint8_t* __restrict__ pad_bits) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int n = B * h * dh;
  if (idx >= n) return;
  const int t = t_dev ? *t_dev : t_host;
  if (pad_bits && idx < B)     // the key at position t is masked later iff its token is the pad token
    pad_bits[(int64_t)idx * max_seq + t] = (ids[(int64_t)idx * ld_ids + t] == pad_token) ? 1 : 0;
  int d = idx % dh, hh = (idx / dh) % h, b = idx / (dh * h);
  int64_t src = (int64_t)b * 3 * h * dh + (int64_t)hh * dh + d;
  int64_t dst = (((int64_t)b * h + hh) * max_seq + t) * dh + d;
  kc[dst] = qkv[src + (int64_t)h * dh];
  vc[dst] = qkv[src + 2 * (int64_t)h * dh];
}

// embedding * sqrt(d) + PE[t] of the token at position t = *t_dev of every sequence (MT/layers.py:226-228)
template <typename TL>
__global__ void __launch_bounds__(256)
decode_embed_kernel(const int32_t* __restrict__ ids, int64_t ld_ids, const int32_t* __restrict__ t_dev,
                    const float* __restrict__ emb, const float* __restrict__ pe, float* __restrict__ out,
                    TL* __restrict__ out_lp, int B, int d4, int V, float scale) {
  const int e4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (e4 >= B * d4) return;
  const int b = e4 / d4, c4 = e4 - b * d4;
  const int t = *t_dev;
  int32_t id = ids[(int64_t)b * ld_ids + t];
  id = id < 0 ? 0 : (id >= V ? V - 1 : id);
  const float4 w = *reinterpret_cast<const float4*>(emb + ((int64_t)id * d4 + c4) * 4);
  const float4 q = *reinterpret_cast<const float4*>(pe + ((int64_t)t * d4 + c4) * 4);
  const float4 r = make_float4(__fadd_rn(__fmul_rn(w.x, scale), q.x), __fadd_rn(__fmul_rn(w.y, scale), q.y),
                               __fadd_rn(__fmul_rn(w.z, scale), q.z), __fadd_rn(__fmul_rn(w.w, scale), q.w));
  *reinterpret_cast<float4*>(out + (int64_t)e4 * 4) = r;
  if (out_lp) store4<TL>(out_lp + (int64_t)e4 * 4, r);
}

__global__ void decode_advance_kernel(int32_t* t_dev) { *t_dev += 1; }

// one block per sequence.  z = logits/T; greedy: first arg-max.  Otherwise keep the top_k
// values (ties: lower id first), softmax, inverse CDF over ascending ids with uniform u.
__global__ void __launch_bounds__(256)
sample_kernel(const float* __restrict__ logits, const float* __restrict__ u, int32_t* __restrict__ out,
              int V, float temperature, int top_k, int greedy, const int32_t* __restrict__ t_dev,
              int64_t ld_ids, int prior_len, int B) {
  extern __shared__ float sm[];
  float* z = sm;                          // [V]
  unsigned char* keep = reinterpret_cast<unsigned char*>(z + V);   // [V]
  __shared__ float rv[8];
  __shared__ int ri[8];
  __shared__ float s_thr_v;
  __shared__ int s_thr_i;
  const int tid = threadIdx.x, b = blockIdx.x;
  if (t_dev) {
    // graph-replayed decode: `out` is the id matrix [B, ld_ids]; the event drawn from position t's
    // logits becomes the token at t+1 unless that position still belongs to the prior; `u` holds
    // one row of uniforms per generated event
    const int t = *t_dev;
    if (t + 1 < prior_len) return;
    out = out + (int64_t)b * ld_ids + (t + 1) - b;      // "- b": the stores below index out[b]
    if (u) u = u + (int64_t)(t + 1 - prior_len) * B;
  }
  const float* zb = logits + (int64_t)b * V;
  for (int c = tid; c < V; c += 256) {
    z[c] = greedy ? zb[c] : zb[c] / temperature;
    keep[c] = 0;
  }
  __syncthreads();
  const bool full = greedy || top_k <= 0 || top_k >= V;
  const int rounds = greedy ? 1 : (full ? 0 : top_k);
  for (int it = 0; it < rounds; ++it) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = tid; c < V; c += 256) {
      if (keep[c]) continue;
      float v = z[c];
      if (v > bv || (v == bv && c < bi)) { bv = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { rv[tid >> 5] = bv; ri[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (rv[w] > bv || (rv[w] == bv && ri[w] < bi)) { bv = rv[w]; bi = ri[w]; }
      s_thr_v = bv;
      s_thr_i = bi;
      if (bi < V) keep[bi] = 1;
    }
    __syncthreads();
    if (greedy) {
      if (tid == 0) out[b] = s_thr_i;
      return;
    }
  }
  if (tid == 0) {
    float mx = -INFINITY;
    for (int c = 0; c < V; ++c)
      if (full || keep[c]) mx = fmaxf(mx, z[c]);
    float tot = 0.f;
    for (int c = 0; c < V; ++c) {
      float pv = (full || keep[c]) ? expf(z[c] - mx) : 0.f;
      z[c] = pv;
      tot += pv;
    }
    // cdf over normalised probabilities, ascending ids
    float target = u[b];
    float cdf = 0.f, total = 0.f;
    for (int c = 0; c < V; ++c) total += z[c] / tot;
    target *= total;
    int pick = -1, last_alive = 0;
    for (int c = 0; c < V; ++c) {
      float pr = z[c] / tot;
      if (pr > 0.f) last_alive = c;
      cdf += pr;
      if (pick < 0 && !(cdf <= target)) pick = c;
    }
    if (pick < 0 || pick > last_alive) pick = last_alive;
    out[b] = pick;
  }
}

}  // namespace mt

using namespace mt;

extern "C" {

static int rga_decode_impl(const void* q, int64_t q_stride_b, const void* kcache, const void* vcache,
                  const void* E, const uint8_t* pad_keys, void* out, int64_t B, int64_t h, int64_t dh, int64_t max_seq,
                  int64_t t, const int32_t* t_dev, int dtype, void* stream) {
  MT_REQUIRE(q && kcache && vcache && E && out, "rga_decode: null pointer");
  MT_REQUIRE(B > 0 && h > 0 && max_seq > 0 && t >= 0 && t < max_seq, "rga_decode: bad shape (t=%ld max_seq=%ld)", (long)t, (long)max_seq);
  MT_REQUIRE(aligned(kcache, 16) && aligned(vcache, 16) && aligned(E, 16), "rga_decode: misaligned");
  // with a device-resident step index the score buffer is sized for the longest context
  size_t smem = (size_t)(dh + 128 + (t_dev ? max_seq : t + 1)) * sizeof(float);
  dim3 grid((unsigned)h, (unsigned)B);
  cudaError_t ae = cudaSuccess;
#define MT_LAUNCH_DEC(T, DHC)                                                              \
  {                                                                                        \
    auto kern = rga_decode_kernel<T, DHC>;                                                 \
    if (smem > 48 * 1024) ae = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    kern<<<grid, 128, smem, as_stream(stream)>>>((const T*)q, (const T*)kcache, (const T*)vcache, (const T*)E, pad_keys, (T*)out, q_stride_b, (int)h, (int)max_seq, (int)t, t_dev, sqrtf((float)dh)); \
  }
  MT_DISPATCH_F32_BF16(dtype, T, {
    if (dh == 32) MT_LAUNCH_DEC(T, 32)
    else if (dh == 64) MT_LAUNCH_DEC(T, 64)
    else if (dh == 128) MT_LAUNCH_DEC(T, 128)
    else { set_error("rga_decode: head dim %d not in {32,64,128}", (int)dh); return MT_E_UNSUPPORTED; }
  });
#undef MT_LAUNCH_DEC
  if (ae != cudaSuccess) { set_error("rga_decode: smem attribute: %s", cudaGetErrorString(ae)); return (int)ae; }
  return check_launch("rga_decode");
}

int mt_rga_decode(const void* q, int64_t q_stride_b, const void* kcache, const void* vcache,
                  const void* E, const uint8_t* pad_keys, void* out, int64_t B, int64_t h, int64_t dh, int64_t max_seq,
                  int64_t t, int dtype, void* stream) {
  return rga_decode_impl(q, q_stride_b, kcache, vcache, E, pad_keys, out, B, h, dh, max_seq, t, nullptr, dtype, stream);
}

static int kv_append_impl(const void* qkv, void* kcache, void* vcache, int64_t B, int64_t h, int64_t dh,
                          int64_t max_seq, int64_t t, const int32_t* t_dev, const int32_t* ids, int64_t ld_ids,
                          int32_t pad_token, uint8_t* pad_bits, int dtype, void* stream) {
  MT_REQUIRE(qkv && kcache && vcache, "kv_append: null pointer");
  MT_REQUIRE(B > 0 && h > 0 && dh > 0 && t >= 0 && t < max_seq, "kv_append: bad shape");
  MT_REQUIRE(!pad_bits || ids, "kv_append: pad_bits needs the id matrix");
  int64_t n = B * h * dh;
  MT_DISPATCH_F32_BF16(dtype, T,
      (kv_append_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>((const T*)qkv, (T*)kcache, (T*)vcache, (int)B, (int)h, (int)dh, (int)max_seq, (int)t, t_dev, ids, ld_ids, pad_token, pad_bits)));
  return check_launch("kv_append");
}

int mt_kv_append(const void* qkv, void* kcache, void* vcache, int64_t B, int64_t h, int64_t dh,
                 int64_t max_seq, int64_t t, int dtype, void* stream) {
  return kv_append_impl(qkv, kcache, vcache, B, h, dh, max_seq, t, nullptr, nullptr, 0, 0, nullptr, dtype, stream);
}

static int sample_impl(const float* logits, const float* u, int32_t* ids_out, int64_t B, int64_t V,
                       float temperature, int32_t top_k, int greedy, const int32_t* t_dev, int64_t ld_ids,
                       int prior_len, void* stream) {
  MT_REQUIRE(logits && ids_out && B > 0 && V > 0, "sample: bad args");
  MT_REQUIRE(greedy || (u != nullptr && temperature > 0.f), "sample: need uniforms and temperature > 0");
  size_t smem = (size_t)V * (sizeof(float) + 1) + 16;
  MT_REQUIRE(smem <= 200 * 1024, "sample: vocabulary too large (%ld)", (long)V);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("sample: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  sample_kernel<<<(unsigned)B, 256, smem, as_stream(stream)>>>(logits, u, ids_out, (int)V, temperature, top_k, greedy, t_dev, ld_ids, prior_len, (int)B);
  return check_launch("sample");
}

int mt_sample(const float* logits, const float* u, int32_t* ids_out, int64_t B, int64_t V,
              float temperature, int32_t top_k, int greedy, void* stream) {
  return sample_impl(logits, u, ids_out, B, V, temperature, top_k, greedy, nullptr, 0, 0, stream);
}

// ---- device-resident step index: one CUDA graph of a decode step is replayed per event -------
int mt_decode_embed(const int32_t* ids, int64_t ld_ids, const int32_t* t_dev, const float* emb,
                    const float* pe, float* out_f32, void* out_lp, int lp_dtype, int64_t B, int64_t d,
                    int64_t V, float scale, void* stream) {
  MT_REQUIRE(ids && t_dev && emb && pe && out_f32 && B > 0 && d > 0 && d % 4 == 0 && V > 0, "decode_embed: bad args");
  if (!out_lp) lp_dtype = MT_F32;
  int64_t n4 = B * (d / 4);
  MT_DISPATCH_F32_BF16(lp_dtype, TL,
      (decode_embed_kernel<TL><<<(unsigned)((n4 + 255) / 256), 256, 0, as_stream(stream)>>>(ids, ld_ids, t_dev, emb, pe, out_f32, (TL*)out_lp, (int)B, (int)(d / 4), (int)V, scale)));
  return check_launch("decode_embed");
}

int mt_decode_kv_append(const void* qkv, void* kcache, void* vcache, const int32_t* ids, int64_t ld_ids,
                        int32_t pad_token, uint8_t* pad_bits, const int32_t* t_dev, int64_t B, int64_t h,
                        int64_t dh, int64_t max_seq, int dtype, void* stream) {
  MT_REQUIRE(t_dev, "decode_kv_append: null step index");
  return kv_append_impl(qkv, kcache, vcache, B, h, dh, max_seq, 0, t_dev, ids, ld_ids, pad_token, pad_bits, dtype, stream);
}

int mt_decode_attend(const void* q, int64_t q_stride_b, const void* kcache, const void* vcache, const void* E,
                     const uint8_t* pad_bits, void* out, const int32_t* t_dev, int64_t B, int64_t h, int64_t dh,
                     int64_t max_seq, int dtype, void* stream) {
  MT_REQUIRE(t_dev, "decode_attend: null step index");
  return rga_decode_impl(q, q_stride_b, kcache, vcache, E, pad_bits, out, B, h, dh, max_seq, 0, t_dev, dtype, stream);
}

int mt_decode_sample(const float* logits, const float* u, int32_t* ids, int64_t ld_ids, const int32_t* t_dev,
                     int32_t prior_len, int64_t B, int64_t V, float temperature, int32_t top_k, int greedy,
                     void* stream) {
  MT_REQUIRE(t_dev && ld_ids > 0 && prior_len >= 1, "decode_sample: bad args");
  return sample_impl(logits, u, ids, B, V, temperature, top_k, greedy, t_dev, ld_ids, prior_len, stream);
}

int mt_decode_advance(int32_t* t_dev, void* stream) {
  MT_REQUIRE(t_dev, "decode_advance: null step index");
  decode_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(t_dev);
  return check_launch("decode_advance");
}

}  // extern "C"
