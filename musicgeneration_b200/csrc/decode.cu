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


// ---- split-context variant (the decode hot path) ---------------------------------------------
// One CTA per (head, sequence, 256-key chunk of the context): DH/8 lanes share a key row (16-byte
// loads, a warp reads whole contiguous rows), partial (max, sum, o[DH]) per chunk goes to the
// caller's workspace and the LAST chunk-CTA of a (sequence, head) to finish folds the partials
// (arrival counter, reset for the next launch) -- no second launch.  The one-CTA-per-(sequence,
// head) kernel above streams its whole context through 128 threads and reaches ~1.1 TB/s at
// context 1000; splitting gives every SM several independent streams.
constexpr int DEC_CHUNK = 256;

// eight consecutive elements held raw (one or two 16-byte registers quads) until they are needed
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
  uint4 r;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { r = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = r; }
  __device__ __forceinline__ void to_float(float (&f)[8]) const {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
};
template <> struct Raw8<__half> {
  uint4 r;
  __device__ __forceinline__ void load(const __half* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { r = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = r; }
  __device__ __forceinline__ void to_float(float (&f)[8]) const {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = v.x; f[2 * i + 1] = v.y;
    }
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a; *reinterpret_cast<float4*>(p + 4) = b;
  }
  __device__ __forceinline__ void to_float(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};
template <typename T> __device__ __forceinline__ void load8f(const T* p, float (&f)[8]);
template <> __device__ __forceinline__ void load8f<float>(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void load8f<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}

template <> __device__ __forceinline__ void load8f<__half>(const __half* p, float (&f)[8]) {
  Raw8<__half> r;
  r.load(p);
  r.to_float(f);
}

template <typename T, int DH>
__global__ void __launch_bounds__(128)
rga_decode_split_kernel(const T* __restrict__ q, T* kc, T* vc,
                        const T* __restrict__ E, const uint8_t* __restrict__ pad_keys, T* __restrict__ out,
                        int64_t q_stride_b, int h, int max_seq, const int32_t* __restrict__ t_dev,
                        float inv_sqrt_dh, int32_t* __restrict__ counters, float* __restrict__ part, int nsplit,
                        int append) {
  constexpr int LPK = DH / 8;            // lanes per key row
  constexpr int KPI = 128 / LPK;         // keys per block iteration
  constexpr int NIT = DEC_CHUNK / KPI;
  __shared__ float red[KPI][DH + 1];
  __shared__ float wred[8];
  __shared__ uint8_t spad[DEC_CHUNK];
  __shared__ int s_last;
  chain_prologue();
  const int t = *t_dev;
  const int tid = threadIdx.x, hh = blockIdx.x, b = blockIdx.y, sp = blockIdx.z;
  const int j0 = sp * DEC_CHUNK;
  const int n = min(DEC_CHUNK, t + 1 - j0);
  if (n <= 0) return;                                         // chunk beyond the context (uniform per CTA)
  const int active = (t + DEC_CHUNK) / DEC_CHUNK;             // chunks that hold keys 0..t
  const int64_t bh = (int64_t)b * h + hh;
  const int sub = tid % LPK, grp = tid / LPK;
  // shuffles stay inside a key group: its LPK lanes leave the key loop together, other groups of the warp may not
  const uint32_t LMASK = (LPK == 32 ? 0xffffffffu : ((1u << LPK) - 1u)) << ((tid & 31) / LPK * LPK);
  float q8[8];
  load8f<T>(q + (int64_t)b * q_stride_b + (int64_t)hh * DH + sub * 8, q8);
#pragma unroll
  for (int e = 0; e < 8; ++e) q8[e] *= inv_sqrt_dh;
  const uint8_t* pad = pad_keys ? pad_keys + (int64_t)b * max_seq : nullptr;
  if (pad) {                                                  // the chunk's pad flags, once
    for (int x = tid; x < DEC_CHUNK; x += 128) spad[x] = (x < n) ? pad[j0 + x] : 1;
    __syncthreads();
  }
  T* kb = kc + (bh * (int64_t)max_seq + j0) * DH + sub * 8;
  T* vb = vc + (bh * (int64_t)max_seq + j0) * DH + sub * 8;
  // append != 0: q is the fused projection row [3, h, DH] of the new token; its K / V rows (position t)
  // are taken from there and stored into the caches by the chunk that owns key t (no separate append launch)
  const T* knew = q + (int64_t)b * q_stride_b + (int64_t)(h + hh) * DH + sub * 8;
  const T* vnew = q + (int64_t)b * q_stride_b + (int64_t)(2 * h + hh) * DH + sub * 8;
  const T* eb = E + (int64_t)(max_seq - 1 - t + j0) * DH + sub * 8;      // row of key j0; key j0+jj is jj rows further
  // ---- one pass: every key group (LPK lanes) walks its keys with a running (max, sum, o).  Keys are
  // taken four at a time: the 12 row loads (K, E, V of four keys) are issued back to back BEFORE any
  // of them is used (the compiler does not hoist them across the per-key branches by itself), so a
  // thread keeps 12 independent 16/32-byte loads in flight.
  float m_run = -INFINITY, l_run = 0.f, o8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o8[e] = 0.f;
#pragma unroll 1
  for (int it0 = 0; it0 < NIT; it0 += 4) {
    if (grp + it0 * KPI >= n) break;                // uniform per key group; later groups only feed zeros to the shuffles
    Raw8<T> kr[4], er[4], vr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int jj = grp + (it0 + u) * KPI;
      if (jj < n) {
        const bool is_new = append && (j0 + jj == t);
        kr[u].load(is_new ? knew : kb + (int64_t)jj * DH);
        er[u].load(eb + (int64_t)jj * DH);
        vr[u].load(is_new ? vnew : vb + (int64_t)jj * DH);
        if (is_new) { kr[u].store(kb + (int64_t)jj * DH); vr[u].store(vb + (int64_t)jj * DH); }
      } else {
        kr[u].zero(); er[u].zero(); vr[u].zero();
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int jj = grp + (it0 + u) * KPI;
      float kf[8], ef[8], vf[8];
      kr[u].to_float(kf); er[u].to_float(ef); vr[u].to_float(vf);
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(q8[e], kf[e] + ef[e], acc);
#pragma unroll
      for (int o = LPK / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(LMASK, acc, o);
      const bool use = (jj < n) && !(pad && spad[jj]);
      const float a2 = use ? acc : -INFINITY;
      const float m_new = fmaxf(m_run, a2);
      const float corr = (m_new == -INFINITY) ? 1.f : __expf(m_run - m_new);     // nothing seen yet: keep zeros
      const float pj = use ? __expf(acc - m_new) : 0.f;
      l_run = fmaf(l_run, corr, pj);
#pragma unroll
      for (int e = 0; e < 8; ++e) o8[e] = fmaf(o8[e], corr, pj * vf[e]);
      m_run = m_new;
    }
  }
  // ---- fold the key groups: M = max m_g, weights exp(m_g - M)
  float mx = warp_max(m_run);
  if ((tid & 31) == 0) wred[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(wred[0], wred[1]), fmaxf(wred[2], wred[3]));
  const float wg = (m_run == -INFINITY) ? 0.f : __expf(m_run - mx);
#pragma unroll
  for (int e = 0; e < 8; ++e) red[grp][sub * 8 + e] = o8[e] * wg;
  float sum = (sub == 0) ? l_run * wg : 0.f;
  sum = warp_sum(sum);
  if ((tid & 31) == 0) wred[4 + (tid >> 5)] = sum;
  __syncthreads();
  float* my = part + (bh * nsplit + sp) * (DH + 2);
  if (tid < DH) {
    float tot = 0.f;
#pragma unroll 8
    for (int g = 0; g < KPI; ++g) tot += red[g][tid];
    my[2 + tid] = tot;
  }
  if (tid == 0) { my[0] = mx; my[1] = (wred[4] + wred[5]) + (wred[6] + wred[7]); }
  // ---- the last chunk of this (sequence, head) to arrive folds the partials
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(&counters[bh], 1);
    s_last = (prev == active - 1);
    if (s_last) counters[bh] = 0;                  // ready for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < DH) {
    const float* base = part + bh * nsplit * (DH + 2);
    float M = -INFINITY;
    for (int s2 = 0; s2 < active; ++s2) M = fmaxf(M, __ldcg(base + s2 * (DH + 2)));
    float L = 0.f, acc = 0.f;
    for (int s2 = 0; s2 < active; ++s2) {
      const float ms = __ldcg(base + s2 * (DH + 2));
      const float w = (ms == -INFINITY) ? 0.f : __expf(ms - M);
      L = fmaf(w, __ldcg(base + s2 * (DH + 2) + 1), L);
      acc = fmaf(w, __ldcg(base + s2 * (DH + 2) + 2 + tid), acc);
    }
    out[bh * DH + tid] = from_f<T>(L > 0.f ? acc / L : 0.f);
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
                    TL* __restrict__ out_lp, int B, int d4, int V, float scale, int32_t pad_token,
                    uint8_t* __restrict__ pad_bits, int64_t max_seq) {
  chain_prologue();
  const int e4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (e4 >= B * d4) return;
  const int b = e4 / d4, c4 = e4 - b * d4;
  const int t = *t_dev;
  int32_t id = ids[(int64_t)b * ld_ids + t];
  // the key at position t is masked later iff its token is the pad token (MT/utils.py:73)
  if (pad_bits && c4 == 0) pad_bits[(int64_t)b * max_seq + t] = (id == pad_token) ? 1 : 0;
  id = id < 0 ? 0 : (id >= V ? V - 1 : id);
  const float4 w = *reinterpret_cast<const float4*>(emb + ((int64_t)id * d4 + c4) * 4);
  const float4 q = *reinterpret_cast<const float4*>(pe + ((int64_t)t * d4 + c4) * 4);
  const float4 r = make_float4(__fadd_rn(__fmul_rn(w.x, scale), q.x), __fadd_rn(__fmul_rn(w.y, scale), q.y),
                               __fadd_rn(__fmul_rn(w.z, scale), q.z), __fadd_rn(__fmul_rn(w.w, scale), q.w));
  *reinterpret_cast<float4*>(out + (int64_t)e4 * 4) = r;
  if (out_lp) store4<TL>(out_lp + (int64_t)e4 * 4, r);
}

__global__ void decode_advance_kernel(int32_t* t_dev) { chain_prologue(); *t_dev += 1; }

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
  chain_prologue();
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
  MT_DISPATCH_DTYPE(dtype, T, {
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
  MT_DISPATCH_DTYPE(dtype, T,
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
  launch_chain(sample_kernel, dim3((unsigned)B), dim3(256), smem, as_stream(stream), logits, u, ids_out, (int)V, temperature, (int)top_k, greedy, t_dev, ld_ids, (int)prior_len, (int)B);
  return check_launch("sample");
}

int mt_sample(const float* logits, const float* u, int32_t* ids_out, int64_t B, int64_t V,
              float temperature, int32_t top_k, int greedy, void* stream) {
  return sample_impl(logits, u, ids_out, B, V, temperature, top_k, greedy, nullptr, 0, 0, stream);
}

// ---- device-resident step index: one CUDA graph of a decode step is replayed per event -------
int mt_decode_embed(const int32_t* ids, int64_t ld_ids, const int32_t* t_dev, const float* emb,
                    const float* pe, float* out_f32, void* out_lp, int lp_dtype, int64_t B, int64_t d,
                    int64_t V, float scale, int32_t pad_token, uint8_t* pad_bits, int64_t max_seq, void* stream) {
  MT_REQUIRE(ids && t_dev && emb && pe && out_f32 && B > 0 && d > 0 && d % 4 == 0 && V > 0, "decode_embed: bad args");
  if (!out_lp) lp_dtype = MT_F32;
  int64_t n4 = B * (d / 4);
  MT_DISPATCH_DTYPE(lp_dtype, TL,
      (launch_chain(decode_embed_kernel<TL>, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, as_stream(stream), ids, ld_ids, t_dev, emb, pe, out_f32, (TL*)out_lp, (int)B, (int)(d / 4), (int)V, scale, pad_token, pad_bits, max_seq)));
  return check_launch("decode_embed");
}

int mt_decode_kv_append(const void* qkv, void* kcache, void* vcache, const int32_t* ids, int64_t ld_ids,
                        int32_t pad_token, uint8_t* pad_bits, const int32_t* t_dev, int64_t B, int64_t h,
                        int64_t dh, int64_t max_seq, int dtype, void* stream) {
  MT_REQUIRE(t_dev, "decode_kv_append: null step index");
  return kv_append_impl(qkv, kcache, vcache, B, h, dh, max_seq, 0, t_dev, ids, ld_ids, pad_token, pad_bits, dtype, stream);
}

size_t mt_decode_attend_workspace_bytes(int64_t B, int64_t h, int64_t dh, int64_t max_seq) {
  const int64_t nsplit = (max_seq + DEC_CHUNK - 1) / DEC_CHUNK;
  return (size_t)(B * h) * sizeof(int32_t) + (size_t)(B * h * nsplit * (dh + 2)) * sizeof(float);
}

int mt_decode_attend(const void* q, int64_t q_stride_b, void* kcache, void* vcache, const void* E,
                     const uint8_t* pad_bits, void* out, const int32_t* t_dev, int64_t B, int64_t h, int64_t dh,
                     int64_t max_seq, int dtype, int append, void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(t_dev, "decode_attend: null step index");
  MT_REQUIRE(q && kcache && vcache && E && out, "decode_attend: null pointer");
  MT_REQUIRE(B > 0 && h > 0 && max_seq > 0 && B <= 65535, "decode_attend: bad shape");
  MT_REQUIRE(aligned(kcache, 16) && aligned(vcache, 16) && aligned(E, 16) && aligned(q, 16) && q_stride_b % 8 == 0, "decode_attend: misaligned");
  const size_t need = mt_decode_attend_workspace_bytes(B, h, dh, max_seq);
  if (!workspace || workspace_bytes < need || !aligned(workspace, 16)) {
    set_error("decode_attend: workspace too small or misaligned (%zu < %zu); it must be zeroed once before the first call", workspace_bytes, need);
    return MT_E_WORKSPACE;
  }
  const int nsplit = (int)((max_seq + DEC_CHUNK - 1) / DEC_CHUNK);
  int32_t* counters = reinterpret_cast<int32_t*>(workspace);
  // partials start at a 16-byte boundary after the counters
  float* part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + (((size_t)(B * h) * sizeof(int32_t) + 15) / 16) * 16);
  MT_REQUIRE((size_t)(reinterpret_cast<uint8_t*>(part) - reinterpret_cast<uint8_t*>(workspace)) + (size_t)(B * h * nsplit * (dh + 2)) * sizeof(float) <= workspace_bytes + 16, "decode_attend: workspace layout");
  dim3 grid((unsigned)h, (unsigned)B, (unsigned)nsplit);
  const float isd = 1.f / sqrtf((float)dh);
#define MT_LAUNCH_DECS(T, DHC) \
  launch_chain(rga_decode_split_kernel<T, DHC>, grid, dim3(128), 0, as_stream(stream), (const T*)q, (T*)kcache, (T*)vcache, (const T*)E, pad_bits, (T*)out, q_stride_b, (int)h, (int)max_seq, t_dev, isd, counters, part, nsplit, append);
  MT_DISPATCH_DTYPE(dtype, T, {
    if (dh == 32) { MT_LAUNCH_DECS(T, 32) }
    else if (dh == 64) { MT_LAUNCH_DECS(T, 64) }
    else if (dh == 128) { MT_LAUNCH_DECS(T, 128) }
    else { set_error("decode_attend: head dim %ld not built (32, 64, 128)", (long)dh); return MT_E_UNSUPPORTED; }
  });
#undef MT_LAUNCH_DECS
  return check_launch("decode_attend");
}

int mt_decode_sample(const float* logits, const float* u, int32_t* ids, int64_t ld_ids, const int32_t* t_dev,
                     int32_t prior_len, int64_t B, int64_t V, float temperature, int32_t top_k, int greedy,
                     void* stream) {
  MT_REQUIRE(t_dev && ld_ids > 0 && prior_len >= 1, "decode_sample: bad args");
  return sample_impl(logits, u, ids, B, V, temperature, top_k, greedy, t_dev, ld_ids, prior_len, stream);
}

int mt_decode_advance(int32_t* t_dev, void* stream) {
  MT_REQUIRE(t_dev, "decode_advance: null step index");
  launch_chain(decode_advance_kernel, dim3(1), dim3(1), 0, as_stream(stream), t_dev);
  return check_launch("decode_advance");
}

}  // extern "C"
