// Relative global attention, reference-precision (FFMA, fp32) flash-style kernels.
// Implements MT/layers.py:86-106 (+ :111-133 skew / QE masking) in the closed form
//     S[i,j] = ( q_i.k_j + [j<=i] q_i.E[max_seq-1-(i-j)] ) / sqrt(dh)
// without materialising any L x L tensor: per 64x64 tile the needed E rows form a band of 127
// consecutive rows (row = max_seq-1-(i0-j0)-63 + g, g = 63 - a + b), staged in shared memory;
// rows >= max_seq are exactly the j>i positions and are loaded as zeros, which realises the
// reference's _qe_masking.  Online softmax (fp32), LSE saved for the backward.
// Backward follows SURVEY Appendix A: dS = P o (dP - D)/sqrt(dh); dV = P^T dO; dK = dS^T Q;
// dQ = dS (K + E_band);  dE[r] += sum dS[i,i-r] q_i  (fp32 atomics across (b,h) and tiles).
#include "ops.cuh"

namespace mt {

constexpr int RT = 64;          // tile edge (queries and keys)
constexpr int RTHREADS = 256;   // 16 x 16 threads, each 4 rows x 4 strided columns


template <typename T, int DH>
__device__ __forceinline__ void load_rows(float* dst, const T* base, int64_t sl, int row0, int L,
                                          int tid) {
  // dst[r][DH+1] <- base[(row0+r)*sl + d], zero beyond L
  constexpr int PD = DH + 1;
  for (int idx = tid; idx < RT * (DH / 4); idx += RTHREADS) {
    int r = idx / (DH / 4), c4 = idx - r * (DH / 4);
    int row = row0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < L) v = load4<T>(base + (int64_t)row * sl + c4 * 4);
    float* d = dst + r * PD + c4 * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
}

template <typename T, int DH>
__device__ __forceinline__ void load_band(float* dst, const T* E, int erow0, int max_seq, int tid) {
  // dst[g][DH+1] <- E[erow0+g], g in [0,127); rows outside [0,max_seq) are zero
  constexpr int PD = DH + 1;
  for (int idx = tid; idx < (2 * RT - 1) * (DH / 4); idx += RTHREADS) {
    int g = idx / (DH / 4), c4 = idx - g * (DH / 4);
    int row = erow0 + g;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row >= 0 && row < max_seq) v = load4<T>(E + (int64_t)row * DH + c4 * 4);
    float* d = dst + g * PD + c4 * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
}

// s[r][c] = sum_d q[ty*4+r][d] * (k[tx+16c][d] + e[63-(ty*4+r)+tx+16c][d])
template <int DH>
__device__ __forceinline__ void tile_scores(float (&s)[4][4], const float* Qs, const float* Ks,
                                            const float* Es, int tx, int ty) {
  constexpr int PD = DH + 1;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) s[r][c] = 0.f;
  const float* qp = Qs + (ty * 4) * PD;
  const float* kp = Ks + tx * PD;
  const float* ep = Es + (63 - ty * 4 + tx) * PD;
#pragma unroll 4
  for (int d = 0; d < DH; ++d) {
    float qv[4], kv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) qv[r] = qp[r * PD + d];
#pragma unroll
    for (int c = 0; c < 4; ++c) kv[c] = kp[c * 16 * PD + d];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        s[r][c] = fmaf(qv[r], kv[c] + ep[(16 * c - r) * PD + d], s[r][c]);
  }
}

__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// =======================================================================================
// forward
// =======================================================================================
template <typename T, int DH>
__global__ void __launch_bounds__(RTHREADS) rga_fwd_simt_kernel(RgaArgs p) {
  extern __shared__ float smem[];
  constexpr int PD = DH + 1;
  float* Qs = smem;
  float* Ks = Qs + RT * PD;
  float* Vs = Ks + RT * PD;
  float* Es = Vs + RT * PD;             // [127][PD]
  float* Ps = Es + (2 * RT - 1) * PD;   // [64][65]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z, hh = blockIdx.y;
  const int i0 = (gridDim.x - 1 - blockIdx.x) * RT;   // long rows first
  const int L = p.L;
  const T* qb = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.sb + (int64_t)hh * p.sh;
  const T* kb = reinterpret_cast<const T*>(p.k) + (int64_t)b * p.sb + (int64_t)hh * p.sh;
  const T* vb = reinterpret_cast<const T*>(p.v) + (int64_t)b * p.sb + (int64_t)hh * p.sh;
  const T* E = reinterpret_cast<const T*>(p.E);
  const uint8_t* pad = p.pad ? p.pad + (int64_t)b * L : nullptr;

  load_rows<T, DH>(Qs, qb, p.sl, i0, L, tid);

  float m[4], l[4], o[4][DH / 16];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
    for (int c = 0; c < DH / 16; ++c) o[r][c] = 0.f;
  }
  const int jend = p.causal ? min(L, i0 + RT) : L;
  for (int j0 = 0; j0 < jend; j0 += RT) {
    __syncthreads();
    load_rows<T, DH>(Ks, kb, p.sl, j0, L, tid);
    load_rows<T, DH>(Vs, vb, p.sl, j0, L, tid);
    load_band<T, DH>(Es, E, p.max_seq - 1 - (i0 - j0) - (RT - 1), p.max_seq, tid);
    __syncthreads();
    float s[4][4];
    tile_scores<DH>(s, Qs, Ks, Es, tx, ty);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty * 4 + r;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + tx + 16 * c;
        bool ok = (j < L) && (!p.causal || j <= i) && !(pad && pad[j]);
        s[r][c] = ok ? s[r][c] / p.inv_scale_div : -INFINITY;
        mx = fmaxf(mx, s[r][c]);
      }
      mx = half_warp_max(mx);
      const float m_new = fmaxf(m[r], mx);
      const float alpha = (m_new == -INFINITY) ? 1.f : expf(m[r] - m_new);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float pv = (s[r][c] == -INFINITY) ? 0.f : expf(s[r][c] - m_new);
        Ps[(ty * 4 + r) * (RT + 1) + tx + 16 * c] = pv;
        sum += pv;
      }
      sum = half_warp_sum(sum);
      l[r] = l[r] * alpha + sum;
      m[r] = m_new;
#pragma unroll
      for (int c = 0; c < DH / 16; ++c) o[r][c] *= alpha;
    }
    __syncthreads();
#pragma unroll 4
    for (int bb = 0; bb < RT; ++bb) {
      float pv[4], vv[DH / 16];
#pragma unroll
      for (int r = 0; r < 4; ++r) pv[r] = Ps[(ty * 4 + r) * (RT + 1) + bb];
#pragma unroll
      for (int c = 0; c < DH / 16; ++c) vv[c] = Vs[bb * PD + tx + 16 * c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < DH / 16; ++c) o[r][c] = fmaf(pv[r], vv[c], o[r][c]);
    }
  }
  T* ob = reinterpret_cast<T*>(p.O) + (int64_t)b * p.ob + (int64_t)hh * p.oh;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= L) continue;
    const float inv = l[r] > 0.f ? 1.f / l[r] : 0.f;
#pragma unroll
    for (int c = 0; c < DH / 16; ++c)
      ob[(int64_t)i * p.ol + tx + 16 * c] = from_f<T>(o[r][c] * inv);
    if (tx == 0) p.lse[((int64_t)b * p.h + hh) * L + i] = l[r] > 0.f ? m[r] + logf(l[r]) : 0.f;
  }
}

// =======================================================================================
// attention weights (eval-mode return value): P[b,h,i,j] = exp(S_ij - lse_i), 0 where masked
// =======================================================================================
template <typename T, int DH>
__global__ void __launch_bounds__(RTHREADS) rga_weights_kernel(RgaArgs p) {
  extern __shared__ float smem[];
  constexpr int PD = DH + 1;
  float* Qs = smem;
  float* Ks = Qs + RT * PD;
  float* Es = Ks + RT * PD;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z / p.h, hh = blockIdx.z % p.h;
  const int i0 = blockIdx.y * RT, j0 = blockIdx.x * RT;
  const int L = p.L;
  const T* qb = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.sb + (int64_t)hh * p.sh;
  const T* kb = reinterpret_cast<const T*>(p.k) + (int64_t)b * p.sb + (int64_t)hh * p.sh;
  const uint8_t* pad = p.pad ? p.pad + (int64_t)b * L : nullptr;
  load_rows<T, DH>(Qs, qb, p.sl, i0, L, tid);
  load_rows<T, DH>(Ks, kb, p.sl, j0, L, tid);
  load_band<T, DH>(Es, reinterpret_cast<const T*>(p.E), p.max_seq - 1 - (i0 - j0) - (RT - 1), p.max_seq, tid);
  __syncthreads();
  float s[4][4];
  tile_scores<DH>(s, Qs, Ks, Es, tx, ty);
  float* Pb = p.P + ((int64_t)b * p.h + hh) * L * L;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= L) continue;
    const float lse = p.lse[((int64_t)b * p.h + hh) * L + i];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx + 16 * c;
      if (j >= L) continue;
      bool ok = (!p.causal || j <= i) && !(pad && pad[j]);
      Pb[(int64_t)i * L + j] = ok ? expf(s[r][c] / p.inv_scale_div - lse) : 0.f;
    }
  }
}

// =======================================================================================
// backward
// =======================================================================================
// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d].  dh/8 lanes per (b,i,h) row, 16-byte loads (a
// warp reads 512 contiguous bytes of O and of dO when the heads of a position are adjacent),
// shuffle reduction inside the lane group; dh in {32, 64, 128} (other head sizes: one warp per row).
__device__ __forceinline__ uint32_t rga_pack_h2(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <typename T, int LPR>         // LPR = lanes per row = dh / 8
__global__ void __launch_bounds__(256) rga_delta_vec_kernel(RgaArgs p) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t row = t / LPR;                    // (b, i, h) flattened, h fastest
  const int sub = (int)(t % LPR);
  const int64_t n = (int64_t)p.B * p.L * p.h;
  float s = 0.f;
  int hh = 0, i = 0, b = 0;
  if (row < n) {
    hh = (int)(row % p.h);
    i = (int)((row / p.h) % p.L);
    b = (int)(row / ((int64_t)p.h * p.L));
    const int64_t off = (int64_t)b * p.ob + (int64_t)hh * p.oh + (int64_t)i * p.ol + sub * 8;
    const uint4 o = *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.O) + off);
    const uint4 g = *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.dO) + off);
    const T* ov = reinterpret_cast<const T*>(&o);
    const T* gv = reinterpret_cast<const T*>(&g);
    float gf[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { gf[e] = to_f<T>(gv[e]); s += to_f<T>(ov[e]) * gf[e]; }
    if (p.dO_h) {          // loss-scaled f16 copy of dO for the f16 gradient mode (same addressing)
      uint4 w;
      w.x = rga_pack_h2(gf[0] * p.gscale, gf[1] * p.gscale); w.y = rga_pack_h2(gf[2] * p.gscale, gf[3] * p.gscale);
      w.z = rga_pack_h2(gf[4] * p.gscale, gf[5] * p.gscale); w.w = rga_pack_h2(gf[6] * p.gscale, gf[7] * p.gscale);
      *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.dO_h) + off) = w;
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (row < n && sub == 0) p.delta[((int64_t)b * p.h + hh) * p.L + i] = s;
}

template <typename T>
__global__ void __launch_bounds__(256) rga_delta_kernel(RgaArgs p, int dh) {
  int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int64_t n = (int64_t)p.B * p.h * p.L;
  if (w >= n) return;
  int i = (int)(w % p.L);
  int hh = (int)((w / p.L) % p.h);
  int b = (int)(w / ((int64_t)p.L * p.h));
  const T* o = reinterpret_cast<const T*>(p.O) + (int64_t)b * p.ob + (int64_t)hh * p.oh + (int64_t)i * p.ol;
  const T* g = reinterpret_cast<const T*>(p.dO) + (int64_t)b * p.ob + (int64_t)hh * p.oh + (int64_t)i * p.ol;
  float s = 0.f;
  for (int d = lane; d < dh; d += 32) s += to_f<T>(o[d]) * to_f<T>(g[d]);
  s = warp_sum(s);
  if (lane == 0) p.delta[w] = s;
}

// recompute P tile and dS tile into shared memory (Ps, dSs as [64][65]); needs Qs,Ks,Es,Vs,dOs
template <int DH>
__device__ __forceinline__ void tile_p_ds(const RgaArgs& p, const float* Qs, const float* Ks,
                                          const float* Vs, const float* Es, const float* dOs,
                                          float* Ps, float* dSs, const float* lse_s,
                                          const float* delta_s, const uint8_t* pad, int i0, int j0,
                                          int tx, int ty) {
  constexpr int PD = DH + 1;
  float s[4][4];
  tile_scores<DH>(s, Qs, Ks, Es, tx, ty);
  // dP = dO . V^T
  float dp[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) dp[r][c] = 0.f;
#pragma unroll 4
  for (int d = 0; d < DH; ++d) {
    float gv[4], vv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) gv[r] = dOs[(ty * 4 + r) * PD + d];
#pragma unroll
    for (int c = 0; c < 4; ++c) vv[c] = Vs[(tx + 16 * c) * PD + d];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) dp[r][c] = fmaf(gv[r], vv[c], dp[r][c]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int a = ty * 4 + r, i = i0 + a;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int bb = tx + 16 * c, j = j0 + bb;
      bool ok = (i < p.L) && (j < p.L) && (!p.causal || j <= i) && !(pad && pad[j]);
      float pv = ok ? expf(s[r][c] / p.inv_scale_div - lse_s[a]) : 0.f;
      Ps[a * (RT + 1) + bb] = pv;
      dSs[a * (RT + 1) + bb] = pv * (dp[r][c] - delta_s[a]) / p.inv_scale_div;
    }
  }
}

// grid over query tiles: dQ and dE
template <typename T, int DH>
__global__ void __launch_bounds__(RTHREADS) rga_bwd_dq_kernel(RgaArgs p) {
  extern __shared__ float smem[];
  constexpr int PD = DH + 1;
  float* Qs = smem;
  float* Ks = Qs + RT * PD;
  float* Vs = Ks + RT * PD;
  float* dOs = Vs + RT * PD;
  float* Es = dOs + RT * PD;              // [127][PD]
  float* Ps = Es + (2 * RT - 1) * PD;     // [64][65]
  float* dSs = Ps + RT * (RT + 1);
  float* lse_s = dSs + RT * (RT + 1);     // [64]
  float* delta_s = lse_s + RT;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z, hh = blockIdx.y;
  const int i0 = (gridDim.x - 1 - blockIdx.x) * RT;
  const int L = p.L;
  const int64_t qoff = (int64_t)b * p.sb + (int64_t)hh * p.sh;
  const int64_t ooff = (int64_t)b * p.ob + (int64_t)hh * p.oh;
  const T* qb = reinterpret_cast<const T*>(p.q) + qoff;
  const T* kb = reinterpret_cast<const T*>(p.k) + qoff;
  const T* vb = reinterpret_cast<const T*>(p.v) + qoff;
  const T* gb = reinterpret_cast<const T*>(p.dO) + ooff;
  const T* E = reinterpret_cast<const T*>(p.E);
  const uint8_t* pad = p.pad ? p.pad + (int64_t)b * L : nullptr;

  load_rows<T, DH>(Qs, qb, p.sl, i0, L, tid);
  load_rows<T, DH>(dOs, gb, p.ol, i0, L, tid);
  if (tid < RT) {
    int i = i0 + tid;
    lse_s[tid] = i < L ? p.lse[((int64_t)b * p.h + hh) * L + i] : 0.f;
    delta_s[tid] = i < L ? p.delta[((int64_t)b * p.h + hh) * L + i] : 0.f;
  }
  float dq[4][DH / 16];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < DH / 16; ++c) dq[r][c] = 0.f;

  const int jend = p.causal ? min(L, i0 + RT) : L;
  for (int j0 = 0; j0 < jend; j0 += RT) {
    __syncthreads();
    load_rows<T, DH>(Ks, kb, p.sl, j0, L, tid);
    load_rows<T, DH>(Vs, vb, p.sl, j0, L, tid);
    const int erow0 = p.max_seq - 1 - (i0 - j0) - (RT - 1);
    load_band<T, DH>(Es, E, erow0, p.max_seq, tid);
    __syncthreads();
    tile_p_ds<DH>(p, Qs, Ks, Vs, Es, dOs, Ps, dSs, lse_s, delta_s, pad, i0, j0, tx, ty);
    __syncthreads();
    // dQ[a][d] += sum_b dS[a][b] * (K[b][d] + E[63-a+b][d]),  d = tx + 16c
#pragma unroll 2
    for (int bb = 0; bb < RT; ++bb) {
      float ds[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) ds[r] = dSs[(ty * 4 + r) * (RT + 1) + bb];
#pragma unroll
      for (int c = 0; c < DH / 16; ++c) {
        const int d = tx + 16 * c;
        const float kv = Ks[bb * PD + d];
#pragma unroll
        for (int r = 0; r < 4; ++r)
          dq[r][c] = fmaf(ds[r], kv + Es[(63 - (ty * 4 + r) + bb) * PD + d], dq[r][c]);
      }
    }
    // dE band: dEb[g][d] = sum_a dS[a][g-63+a] * Q[a][d]   (0 <= g-63+a < 64)
    for (int idx = tid; idx < (2 * RT - 1) * DH; idx += RTHREADS) {
      const int g = idx / DH, d = idx - g * DH;
      const int erow = erow0 + g;
      if (erow < 0 || erow >= p.max_seq) continue;
      const int a_lo = max(0, 63 - g), a_hi = min(RT - 1, 126 - g);
      float acc = 0.f;
      for (int a = a_lo; a <= a_hi; ++a)
        acc = fmaf(dSs[a * (RT + 1) + (g - 63 + a)], Qs[a * PD + d], acc);
      atomicAdd(p.dE + (int64_t)erow * DH + d, acc);
    }
  }
  T* dqb = reinterpret_cast<T*>(p.dq) + qoff;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= L) continue;
#pragma unroll
    for (int c = 0; c < DH / 16; ++c)
      dqb[(int64_t)i * p.sl + tx + 16 * c] = from_f<T>(dq[r][c]);
  }
}

// grid over key tiles: dK and dV
template <typename T, int DH>
__global__ void __launch_bounds__(RTHREADS) rga_bwd_dkv_kernel(RgaArgs p) {
  extern __shared__ float smem[];
  constexpr int PD = DH + 1;
  float* Qs = smem;
  float* Ks = Qs + RT * PD;
  float* Vs = Ks + RT * PD;
  float* dOs = Vs + RT * PD;
  float* Es = dOs + RT * PD;
  float* Ps = Es + (2 * RT - 1) * PD;
  float* dSs = Ps + RT * (RT + 1);
  float* lse_s = dSs + RT * (RT + 1);
  float* delta_s = lse_s + RT;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z, hh = blockIdx.y;
  const int j0 = blockIdx.x * RT;
  const int L = p.L;
  const int64_t qoff = (int64_t)b * p.sb + (int64_t)hh * p.sh;
  const int64_t ooff = (int64_t)b * p.ob + (int64_t)hh * p.oh;
  const T* qb = reinterpret_cast<const T*>(p.q) + qoff;
  const T* kb = reinterpret_cast<const T*>(p.k) + qoff;
  const T* vb = reinterpret_cast<const T*>(p.v) + qoff;
  const T* gb = reinterpret_cast<const T*>(p.dO) + ooff;
  const T* E = reinterpret_cast<const T*>(p.E);
  const uint8_t* pad = p.pad ? p.pad + (int64_t)b * L : nullptr;

  load_rows<T, DH>(Ks, kb, p.sl, j0, L, tid);
  load_rows<T, DH>(Vs, vb, p.sl, j0, L, tid);
  // this thread owns key rows bb = ty*4+r and features d = tx+16c of dK / dV
  float dk[4][DH / 16], dv[4][DH / 16];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < DH / 16; ++c) dk[r][c] = dv[r][c] = 0.f;

  const int ibeg = p.causal ? j0 : 0;
  for (int i0 = ibeg; i0 < L; i0 += RT) {
    __syncthreads();
    load_rows<T, DH>(Qs, qb, p.sl, i0, L, tid);
    load_rows<T, DH>(dOs, gb, p.ol, i0, L, tid);
    load_band<T, DH>(Es, E, p.max_seq - 1 - (i0 - j0) - (RT - 1), p.max_seq, tid);
    if (tid < RT) {
      int i = i0 + tid;
      lse_s[tid] = i < L ? p.lse[((int64_t)b * p.h + hh) * L + i] : 0.f;
      delta_s[tid] = i < L ? p.delta[((int64_t)b * p.h + hh) * L + i] : 0.f;
    }
    __syncthreads();
    tile_p_ds<DH>(p, Qs, Ks, Vs, Es, dOs, Ps, dSs, lse_s, delta_s, pad, i0, j0, tx, ty);
    __syncthreads();
#pragma unroll 2
    for (int a = 0; a < RT; ++a) {
      float pv[4], ds[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        pv[r] = Ps[a * (RT + 1) + ty * 4 + r];
        ds[r] = dSs[a * (RT + 1) + ty * 4 + r];
      }
#pragma unroll
      for (int c = 0; c < DH / 16; ++c) {
        const int d = tx + 16 * c;
        const float gv = dOs[a * PD + d], qv = Qs[a * PD + d];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          dv[r][c] = fmaf(pv[r], gv, dv[r][c]);
          dk[r][c] = fmaf(ds[r], qv, dk[r][c]);
        }
      }
    }
  }
  T* dkb = reinterpret_cast<T*>(p.dk) + qoff;
  T* dvb = reinterpret_cast<T*>(p.dv) + qoff;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = j0 + ty * 4 + r;
    if (j >= L) continue;
#pragma unroll
    for (int c = 0; c < DH / 16; ++c) {
      dkb[(int64_t)j * p.sl + tx + 16 * c] = from_f<T>(dk[r][c]);
      dvb[(int64_t)j * p.sl + tx + 16 * c] = from_f<T>(dv[r][c]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
template <typename K>
static int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e));
      return (int)e;
    }
  }
  return 0;
}

#define MT_DISPATCH_DH(dh, DHC, ...)                                   \
  if ((dh) == 32) { constexpr int DHC = 32; __VA_ARGS__; }             \
  else if ((dh) == 64) { constexpr int DHC = 64; __VA_ARGS__; }        \
  else if ((dh) == 128) { constexpr int DHC = 128; __VA_ARGS__; }      \
  else { set_error("rga: head dim %d not in {32,64,128}", (int)(dh)); return MT_E_UNSUPPORTED; }

int rga_fwd_simt(const RgaArgs& a, int dh, int dtype, cudaStream_t st) {
  int rc = 0;
  MT_DISPATCH_F32_BF16(dtype, T, MT_DISPATCH_DH(dh, DHC, {
    size_t smem = ((3 * RT + 2 * RT - 1) * (DHC + 1) + RT * (RT + 1)) * sizeof(float);
    auto kern = rga_fwd_simt_kernel<T, DHC>;
    if ((rc = set_smem(kern, smem))) return rc;
    dim3 grid((a.L + RT - 1) / RT, a.h, a.B);
    kern<<<grid, RTHREADS, smem, st>>>(a);
  }));
  return check_launch("rga_fwd_simt");
}

int rga_weights_simt(const RgaArgs& a, int dh, int dtype, cudaStream_t st) {
  int rc = 0;
  MT_DISPATCH_DTYPE(dtype, T, MT_DISPATCH_DH(dh, DHC, {
    size_t smem = ((2 * RT + 2 * RT - 1) * (DHC + 1)) * sizeof(float);
    auto kern = rga_weights_kernel<T, DHC>;
    if ((rc = set_smem(kern, smem))) return rc;
    int nt = (a.L + RT - 1) / RT;
    dim3 grid(nt, nt, a.B * a.h);
    kern<<<grid, RTHREADS, smem, st>>>(a);
  }));
  return check_launch("rga_weights");
}

int rga_delta_launch(const RgaArgs& a, int dh, int dtype, cudaStream_t st) {
  const int64_t rows = (int64_t)a.B * a.h * a.L;
  if (a.dO_h && !(dtype == MT_BF16 && (dh == 32 || dh == 64 || dh == 128) && a.ob % 8 == 0 && a.oh % 8 == 0 &&
                  a.ol % 8 == 0 && aligned(a.O, 16) && aligned(a.dO, 16) && aligned(a.dO_h, 16))) {
    set_error("rga_delta: the scaled f16 copy of dO needs the vector kernel (bf16, 16-byte aligned rows)");
    return MT_E_UNSUPPORTED;
  }
  const bool vec = (dtype == MT_BF16 || dtype == MT_F16) && (dh == 32 || dh == 64 || dh == 128) &&
                   a.ob % 8 == 0 && a.oh % 8 == 0 && a.ol % 8 == 0 && aligned(a.O, 16) && aligned(a.dO, 16);
  if (vec) {
    const int lpr = dh / 8;
    const unsigned grid = (unsigned)((rows * lpr + 255) / 256);
    if (dtype == MT_BF16) {
      if (lpr == 4) rga_delta_vec_kernel<__nv_bfloat16, 4><<<grid, 256, 0, st>>>(a);
      else if (lpr == 8) rga_delta_vec_kernel<__nv_bfloat16, 8><<<grid, 256, 0, st>>>(a);
      else rga_delta_vec_kernel<__nv_bfloat16, 16><<<grid, 256, 0, st>>>(a);
    } else {
      if (lpr == 4) rga_delta_vec_kernel<__half, 4><<<grid, 256, 0, st>>>(a);
      else if (lpr == 8) rga_delta_vec_kernel<__half, 8><<<grid, 256, 0, st>>>(a);
      else rga_delta_vec_kernel<__half, 16><<<grid, 256, 0, st>>>(a);
    }
    return check_launch("rga_delta");
  }
  MT_DISPATCH_DTYPE(dtype, T, {
    rga_delta_kernel<T><<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(a, dh);
  });
  return check_launch("rga_delta");
}

int rga_bwd_simt(const RgaArgs& a, int dh, int dtype, cudaStream_t st) {
  int rc = 0;
  MT_DISPATCH_F32_BF16(dtype, T, MT_DISPATCH_DH(dh, DHC, {
    int64_t rows = (int64_t)a.B * a.h * a.L;
    rga_delta_kernel<T><<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(a, dh);
    if ((rc = check_launch("rga_delta"))) return rc;
    size_t smem = ((4 * RT + 2 * RT - 1) * (DHC + 1) + 2 * RT * (RT + 1) + 2 * RT) * sizeof(float);
    auto k1 = rga_bwd_dq_kernel<T, DHC>;
    auto k2 = rga_bwd_dkv_kernel<T, DHC>;
    if ((rc = set_smem(k1, smem))) return rc;
    if ((rc = set_smem(k2, smem))) return rc;
    dim3 grid((a.L + RT - 1) / RT, a.h, a.B);
    k1<<<grid, RTHREADS, smem, st>>>(a);
    if ((rc = check_launch("rga_bwd_dq"))) return rc;
    k2<<<grid, RTHREADS, smem, st>>>(a);
  }));
  return check_launch("rga_bwd_dkv");
}

}  // namespace mt
