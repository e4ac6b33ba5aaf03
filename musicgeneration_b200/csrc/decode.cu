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
                  int64_t q_stride_b, int h, int max_seq, int t, float sqrt_dh) {
  extern __shared__ float sm[];
  float* qs = sm;              // [DH]
  float* red = qs + DH;        // [128]
  float* sc = red + 128;       // [t+1]
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
                                 int B, int h, int dh, int max_seq, int t) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int n = B * h * dh;
  if (idx >= n) return;
  int d = idx % dh, hh = (idx / dh) % h, b = idx / (dh * h);
  int64_t src = (int64_t)b * 3 * h * dh + (int64_t)hh * dh + d;
  int64_t dst = (((int64_t)b * h + hh) * max_seq + t) * dh + d;
  kc[dst] = qkv[src + (int64_t)h * dh];
  vc[dst] = qkv[src + 2 * (int64_t)h * dh];
}

// one block per sequence.  z = logits/T; greedy: first arg-max.  Otherwise keep the top_k
// values (ties: lower id first), softmax, inverse CDF over ascending ids with uniform u.
__global__ void __launch_bounds__(256)
sample_kernel(const float* __restrict__ logits, const float* __restrict__ u, int32_t* __restrict__ out,
              int V, float temperature, int top_k, int greedy) {
  extern __shared__ float sm[];
  float* z = sm;                          // [V]
  unsigned char* keep = reinterpret_cast<unsigned char*>(z + V);   // [V]
  __shared__ float rv[8];
  __shared__ int ri[8];
  __shared__ float s_thr_v;
  __shared__ int s_thr_i;
  const int tid = threadIdx.x, b = blockIdx.x;
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

int mt_rga_decode(const void* q, int64_t q_stride_b, const void* kcache, const void* vcache,
                  const void* E, const uint8_t* pad_keys, void* out, int64_t B, int64_t h, int64_t dh, int64_t max_seq,
                  int64_t t, int dtype, void* stream) {
  MT_REQUIRE(q && kcache && vcache && E && out, "rga_decode: null pointer");
  MT_REQUIRE(B > 0 && h > 0 && max_seq > 0 && t >= 0 && t < max_seq, "rga_decode: bad shape (t=%ld max_seq=%ld)", (long)t, (long)max_seq);
  MT_REQUIRE(aligned(kcache, 16) && aligned(vcache, 16) && aligned(E, 16), "rga_decode: misaligned");
  size_t smem = (size_t)(dh + 128 + t + 1) * sizeof(float);
  dim3 grid((unsigned)h, (unsigned)B);
  cudaError_t ae = cudaSuccess;
#define MT_LAUNCH_DEC(T, DHC)                                                              \
  {                                                                                        \
    auto kern = rga_decode_kernel<T, DHC>;                                                 \
    if (smem > 48 * 1024) ae = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    kern<<<grid, 128, smem, as_stream(stream)>>>((const T*)q, (const T*)kcache, (const T*)vcache, (const T*)E, pad_keys, (T*)out, q_stride_b, (int)h, (int)max_seq, (int)t, sqrtf((float)dh)); \
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

int mt_kv_append(const void* qkv, void* kcache, void* vcache, int64_t B, int64_t h, int64_t dh,
                 int64_t max_seq, int64_t t, int dtype, void* stream) {
  MT_REQUIRE(qkv && kcache && vcache, "kv_append: null pointer");
  MT_REQUIRE(B > 0 && h > 0 && dh > 0 && t >= 0 && t < max_seq, "kv_append: bad shape");
  int64_t n = B * h * dh;
  MT_DISPATCH_F32_BF16(dtype, T,
      (kv_append_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>((const T*)qkv, (T*)kcache, (T*)vcache, (int)B, (int)h, (int)dh, (int)max_seq, (int)t)));
  return check_launch("kv_append");
}

int mt_sample(const float* logits, const float* u, int32_t* ids_out, int64_t B, int64_t V,
              float temperature, int32_t top_k, int greedy, void* stream) {
  MT_REQUIRE(logits && ids_out && B > 0 && V > 0, "sample: bad args");
  MT_REQUIRE(greedy || (u != nullptr && temperature > 0.f), "sample: need uniforms and temperature > 0");
  size_t smem = (size_t)V * (sizeof(float) + 1) + 16;
  MT_REQUIRE(smem <= 200 * 1024, "sample: vocabulary too large (%ld)", (long)V);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("sample: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  sample_kernel<<<(unsigned)B, 256, smem, as_stream(stream)>>>(logits, u, ids_out, (int)V, temperature, top_k, greedy);
  return check_launch("sample");
}

}  // extern "C"
