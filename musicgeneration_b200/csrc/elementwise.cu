// HBM-bound kernels of the MusicTransformer path: embedding+sinusoid+dropout (K3),
// dropout+residual+LayerNorm fwd/bwd (K4), label-smoothed CE + metrics (K6), bias-gradient
// column sums, casts, Adam.  All are one-pass, 128-bit vectorised, warp-shuffle reductions.
#include <stdarg.h>

#include "common.cuh"

namespace mt {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}
static bool g_chain = false;
bool chain_enabled() { return g_chain; }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// =======================================================================================
// K3  embedding * sqrt(d) + PE + dropout      (MT/layers.py:226-229)
// one thread per 4 features; grid-stride over T*d/4
// =======================================================================================
template <typename TL>
__global__ void __launch_bounds__(256)
embed_pos_fwd_kernel(const int32_t* __restrict__ ids, const float* __restrict__ emb,
                     const float* __restrict__ pe, float* __restrict__ out, TL* __restrict__ out_lp,
                     int64_t T, int64_t L, int d4, int64_t V, int64_t pos0, float scale, float p,
                     float inv_keep, uint64_t seed, uint64_t site) {
  int64_t n4 = T * d4;
  for (int64_t e4 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e4 < n4;
       e4 += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = e4 / d4;
    int c4 = (int)(e4 - t * d4);
    int64_t pos = pos0 + (t % L);
    int32_t id = ids[t];
    id = id < 0 ? 0 : (id >= V ? (int32_t)(V - 1) : id);
    float4 w = *reinterpret_cast<const float4*>(emb + (int64_t)id * d4 * 4 + c4 * 4);
    float4 q = *reinterpret_cast<const float4*>(pe + pos * d4 * 4 + c4 * 4);
    // x *= sqrt(d) then x + PE: two roundings, like the reference (mul_, then add)
    float4 r = make_float4(__fadd_rn(__fmul_rn(w.x, scale), q.x), __fadd_rn(__fmul_rn(w.y, scale), q.y),
                           __fadd_rn(__fmul_rn(w.z, scale), q.z), __fadd_rn(__fmul_rn(w.w, scale), q.w));
    if (p > 0.f) {
      float4 m = dropout_mult4(seed, site, (uint64_t)e4, p, inv_keep);
      r.x *= m.x; r.y *= m.y; r.z *= m.z; r.w *= m.w;
    }
    *reinterpret_cast<float4*>(out + e4 * 4) = r;
    if (out_lp) store4<TL>(out_lp + e4 * 4, r);
  }
}

// backward: demb[ids[t]] += dout[t] * scale * mask   (fp32 RED; V is a few hundred rows)
__global__ void __launch_bounds__(256)
embed_pos_bwd_kernel(const int32_t* __restrict__ ids, const float* __restrict__ dout,
                     float* __restrict__ demb, int64_t T, int d4, int64_t V, float scale, float p,
                     float inv_keep, uint64_t seed, uint64_t site) {
  int64_t n4 = T * d4;
  for (int64_t e4 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e4 < n4;
       e4 += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = e4 / d4;
    int c4 = (int)(e4 - t * d4);
    int32_t id = ids[t];
    if (id < 0 || id >= V) continue;
    float4 g = *reinterpret_cast<const float4*>(dout + e4 * 4);
    if (p > 0.f) {
      float4 m = dropout_mult4(seed, site, (uint64_t)e4, p, inv_keep);
      g.x *= m.x; g.y *= m.y; g.z *= m.z; g.w *= m.w;
    }
    float* dst = demb + (int64_t)id * d4 * 4 + c4 * 4;
    // one vector reduction per quad (demb is 16-byte aligned, checked by the launcher): the V ~ 400 rows
    // are hit by T = 32 k tokens, so the L2 operation count is what the kernel costs
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(g.x * scale), "f"(g.y * scale),
                 "f"(g.z * scale), "f"(g.w * scale) : "memory");
  }
}

// =======================================================================================
// K4  out = LN(dropout(a) + resid) * gamma + beta       (MT/layers.py:154-155,159-160)
// one warp per row, row cached in registers (d <= 1024), two-pass mean / variance
// =======================================================================================

template <typename TA, typename TL, int NV>
__global__ void __launch_bounds__(256)
add_ln_fwd_kernel(const TA* __restrict__ a, const float* __restrict__ resid,
                  const float* __restrict__ gamma, const float* __restrict__ beta,
                  float* __restrict__ out, TL* __restrict__ out_lp, float* __restrict__ mean,
                  float* __restrict__ rstd, int64_t T, int d, float eps, float p, float inv_keep,
                  uint64_t seed, uint64_t site) {
  chain_prologue();
  const int lane = threadIdx.x & 31;
  const int d4 = d >> 2;
  int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; row < T; row += nwarps) {
    float4 z[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c4 = lane + i * 32;
      if (c4 < d4) {
        int64_t e4 = row * d4 + c4;
        float4 av = load4<TA>(a + e4 * 4);
        if (p > 0.f) {
          float4 m = dropout_mult4(seed, site, (uint64_t)e4, p, inv_keep);
          av.x *= m.x; av.y *= m.y; av.z *= m.z; av.w *= m.w;
        }
        float4 rv = *reinterpret_cast<const float4*>(resid + e4 * 4);
        z[i] = make_float4(av.x + rv.x, av.y + rv.y, av.z + rv.z, av.w + rv.w);
        s += (z[i].x + z[i].y) + (z[i].z + z[i].w);
      }
    }
    float mu = warp_sum(s) / (float)d;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c4 = lane + i * 32;
      if (c4 < d4) {
        float dx = z[i].x - mu, dy = z[i].y - mu, dz = z[i].z - mu, dw = z[i].w - mu;
        v += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    float var = warp_sum(v) / (float)d;
    float rs = 1.0f / sqrtf(var + eps);
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c4 = lane + i * 32;
      if (c4 < d4) {
        float4 g = *reinterpret_cast<const float4*>(gamma + c4 * 4);
        float4 b = *reinterpret_cast<const float4*>(beta + c4 * 4);
        float4 o = make_float4((z[i].x - mu) * rs * g.x + b.x, (z[i].y - mu) * rs * g.y + b.y,
                               (z[i].z - mu) * rs * g.z + b.z, (z[i].w - mu) * rs * g.w + b.w);
        int64_t e4 = row * d4 + c4;
        *reinterpret_cast<float4*>(out + e4 * 4) = o;
        if (out_lp) store4<TL>(out_lp + e4 * 4, o);
      }
    }
  }
}

// backward.  xhat = (dropout(a)+resid - mean)*rstd recomputed; g = dout*gamma;
// dz = rstd*(g - mean_d(g) - xhat*mean_d(g*xhat)); da = mask*dz.
// Each warp keeps running dgamma/dbeta partial sums for its columns over the rows it owns and
// the block folds its 8 warps through shared memory into part[0/1][block][d].
template <typename TA, typename TD, int NV>
__global__ void __launch_bounds__(256, (NV <= 4 ? 2 : 1))       // two CTAs per SM (<= 128 registers) up to d = 512
add_ln_bwd_kernel(const float* dout /* may alias dz_out */, const TA* __restrict__ a,
                  const float* __restrict__ resid, const float* __restrict__ gamma,
                  const float* __restrict__ mean, const float* __restrict__ rstd, float* dz_out,
                  TD* __restrict__ da, float* __restrict__ part, int64_t T, int d, float p,
                  float inv_keep, uint64_t seed, uint64_t site) {
  extern __shared__ float sh[];  // [3][8][d]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d4 = d >> 2;
  float4 dg[NV], db[NV], dc[NV];      // dc: column sums of da = bias gradient of the linear that produced `a`
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = dc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t row = blockIdx.x * 8 + warp;
  const int64_t stride = (int64_t)gridDim.x * 8;
  for (; row < T; row += stride) {
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NV], g[NV], msk[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c4 = lane + i * 32;
      if (c4 < d4) {
        int64_t e4 = row * d4 + c4;
        float4 av = load4<TA>(a + e4 * 4);
        msk[i] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p > 0.f) {
          msk[i] = dropout_mult4(seed, site, (uint64_t)e4, p, inv_keep);
          av.x *= msk[i].x; av.y *= msk[i].y; av.z *= msk[i].z; av.w *= msk[i].w;
        }
        float4 rv = *reinterpret_cast<const float4*>(resid + e4 * 4);
        xh[i] = make_float4((av.x + rv.x - mu) * rs, (av.y + rv.y - mu) * rs,
                            (av.z + rv.z - mu) * rs, (av.w + rv.w - mu) * rs);
        float4 dy = *reinterpret_cast<const float4*>(dout + e4 * 4);
        float4 gm = *reinterpret_cast<const float4*>(gamma + c4 * 4);
        g[i] = make_float4(dy.x * gm.x, dy.y * gm.y, dy.z * gm.z, dy.w * gm.w);
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
        dg[i].x += dy.x * xh[i].x; dg[i].y += dy.y * xh[i].y;
        dg[i].z += dy.z * xh[i].z; dg[i].w += dy.w * xh[i].w;
        db[i].x += dy.x; db[i].y += dy.y; db[i].z += dy.z; db[i].w += dy.w;
      }
    }
    float m1 = warp_sum(s1) / (float)d, m2 = warp_sum(s2) / (float)d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c4 = lane + i * 32;
      if (c4 < d4) {
        int64_t e4 = row * d4 + c4;
        float4 o = make_float4(rs * (g[i].x - m1 - xh[i].x * m2), rs * (g[i].y - m1 - xh[i].y * m2),
                               rs * (g[i].z - m1 - xh[i].z * m2), rs * (g[i].w - m1 - xh[i].w * m2));
        *reinterpret_cast<float4*>(dz_out + e4 * 4) = o;
        o.x *= msk[i].x; o.y *= msk[i].y; o.z *= msk[i].z; o.w *= msk[i].w;
        store4<TD>(da + e4 * 4, o);
        // sum what the weight-gradient GEMM will read (the rounded value)
        dc[i].x += to_f<TD>(from_f<TD>(o.x)); dc[i].y += to_f<TD>(from_f<TD>(o.y));
        dc[i].z += to_f<TD>(from_f<TD>(o.z)); dc[i].w += to_f<TD>(from_f<TD>(o.w));
      }
    }
  }
  // fold the 8 warps
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int c4 = lane + i * 32;
    if (c4 < d4) {
      *reinterpret_cast<float4*>(sh + (0 * 8 + warp) * d + c4 * 4) = dg[i];
      *reinterpret_cast<float4*>(sh + (1 * 8 + warp) * d + c4 * 4) = db[i];
      *reinterpret_cast<float4*>(sh + (2 * 8 + warp) * d + c4 * 4) = dc[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * d; c += blockDim.x) {
    int which = c / d, col = c - which * d;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sh[(which * 8 + w) * d + col];
    part[((int64_t)which * gridDim.x + blockIdx.x) * d + col] = s;
  }
}

// one warp per (which, column): the nparts partials are strided by d, lanes take every 32nd
__global__ void __launch_bounds__(256)
ln_param_grad_kernel(const float* __restrict__ part, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ dbias, int64_t nparts, int d) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= 3 * d) return;
  const int which = c / d, col = c - which * d;
  float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dbias);
  if (!dst) return;
  const float* src = part + (int64_t)which * nparts * d + col;
  float s = 0.f;
  for (int64_t i = lane; i < nparts; i += 32) s += src[i * d];
  s = warp_sum(s);
  if (lane == 0) dst[col] = s;
}

// =======================================================================================
// column sums (bias gradients): out[n] = sum_m X[m,n]; deterministic two-stage inside one
// kernel launch per 32-column strip: block = 32 x 8 threads, rows strided by 8, smem fold.
// =======================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ X, float* __restrict__ out, int64_t M, int64_t N, int64_t ldx) {
  __shared__ float sh[8][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int64_t col = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (col < N)
    for (int64_t r = ty; r < M; r += 8) s += to_f<T>(X[r * ldx + col]);
  sh[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][tx];
    out[col] = t;
  }
}
// higher-parallelism variant: grid.y row chunks write partials, then a fold kernel
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ X, float* __restrict__ part, int64_t M, int64_t N,
                      int64_t ldx, int64_t rows_per_chunk) {
  __shared__ float sh[8][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int64_t col = blockIdx.x * 32 + tx;
  int64_t r0 = blockIdx.y * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  float s = 0.f;
  if (col < N)
    for (int64_t r = r0 + ty; r < r1; r += 8) s += to_f<T>(X[r * ldx + col]);
  sh[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][tx];
    part[blockIdx.y * N + col] = t;
  }
}

// vectorised variant (4 columns per thread, 64/128-bit loads): N % 4 == 0, ldx % 4 == 0, aligned X
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_vec_kernel(const T* __restrict__ X, float* __restrict__ part, int64_t M, int64_t N,
                          int64_t ldx, int64_t rows_per_chunk) {
  __shared__ float4 sh[8][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int64_t col = (blockIdx.x * 32 + tx) * 4;
  int64_t r0 = blockIdx.y * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N) {
#pragma unroll 4
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float4 v = load4<T>(X + r * ldx + col);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  sh[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && col < N) {
    float4 t = sh[0][tx];
#pragma unroll
    for (int w = 1; w < 8; ++w) { t.x += sh[w][tx].x; t.y += sh[w][tx].y; t.z += sh[w][tx].z; t.w += sh[w][tx].w; }
    *reinterpret_cast<float4*>(part + blockIdx.y * N + col) = t;
  }
}

// dst[r, c] = src[r, c] for c < cols, 0 for cols <= c < ldd  (row-padded operand copies)
template <typename TS, typename TD>
__global__ void __launch_bounds__(256)
cast2d_kernel(const TS* __restrict__ s, TD* __restrict__ d, int64_t rows, int64_t cols, int64_t lds, int64_t ldd) {
  int64_t n = rows * ldd;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / ldd, c = i - r * ldd;
    d[i] = from_f<TD>(c < cols ? to_f<TS>(s[r * lds + c]) : 0.f);
  }
}

// 256 threads = 8 warps x 32 columns: warp w sums chunks w, w+8, ... (all loads in flight: one thread walking
// the <= 128 chunks took 11 us of serial L2 latency per launch), fixed-order fold through shared memory
__global__ void __launch_bounds__(256) colsum_fold_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                          int64_t chunks, int64_t N) {
  __shared__ float sh[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < N) {
#pragma unroll 8
    for (int64_t i = ty; i < chunks; i += 8) s += part[i * N + c];
  }
  sh[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < N) {
    float t = sh[0][tx];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += sh[w][tx];
    out[c] = t;
  }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, int64_t n) {
  int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x)
    store4<TD>(d + i * 4, load4<TS>(s + i * 4));
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    d[i] = from_f<TD>(to_f<TS>(s[i]));
  }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256)
transpose_cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, int64_t rows, int64_t cols) {
  __shared__ float tile[32][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  int64_t c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = ty; i < 32; i += 8) {
    int64_t r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? to_f<TS>(s[r * cols + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int64_t c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) d[c * rows + r] = from_f<TD>(tile[tx][i]);
  }
}

// =======================================================================================
// K6  label-smoothed CE (+ argmax / accuracy)     (MT/criterion.py:43-67, MT/metrics.py:50-60)
// one warp per row: loss_row = lse - (1-eps) z_t - (eps/V) sum_v z_v
// =======================================================================================
__global__ void __launch_bounds__(256)
smooth_ce_fwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ target,
                     float* __restrict__ row_lse, int32_t* __restrict__ argmax,
                     float* __restrict__ row_loss, float* __restrict__ row_flags, int64_t T, int V,
                     float eps, int32_t ignore) {
  const int lane = threadIdx.x & 31;
  int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= T) return;
  const float* z = logits + row * V;
  float mx = -INFINITY;
  int am = 0x7fffffff;
  float sz = 0.f;
  for (int c = lane; c < V; c += 32) {
    float v = z[c];
    sz += v;
    if (v > mx) { mx = v; am = c; }
  }
  // arg-max with first-index tie-break (torch.argmax on equal values returns the first)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    int oam = __shfl_xor_sync(0xffffffffu, am, o);
    if (omx > mx || (omx == mx && oam < am)) { mx = omx; am = oam; }
  }
  sz = warp_sum(sz);
  float se = 0.f;
  for (int c = lane; c < V; c += 32) se += expf(z[c] - mx);
  se = warp_sum(se);
  float lse = mx + logf(se);
  if (lane == 0) {
    int32_t t = target[row];
    bool valid = (t != ignore);
    float zt = (valid && t >= 0 && t < V) ? z[t] : 0.f;
    row_lse[row] = lse;
    argmax[row] = am;
    row_loss[row] = valid ? (lse - (1.f - eps) * zt - (eps / (float)V) * sz) : 0.f;
    row_flags[row] = (valid ? 1.f : 0.f) + ((am == t) ? 65536.f : 0.f);  // packed: valid + 65536*correct
  }
}

// single block, deterministic tree: sums[0]=loss_sum, [1]=n_valid, [2]=n_correct, [3]=mean loss
__global__ void __launch_bounds__(1024)
smooth_ce_reduce_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_flags,
                        float* __restrict__ sums, int64_t T) {
  __shared__ double sl[32];
  __shared__ unsigned long long sv[32], sc[32];
  double l = 0.0;
  unsigned long long nv = 0, nc = 0;
#pragma unroll 8
  for (int64_t i = threadIdx.x; i < T; i += blockDim.x) {        // (loads of 8 iterations in flight; same add order)
    l += (double)row_loss[i];
    float f = row_flags[i];
    int fi = (int)f;
    nv += (fi & 1);
    nc += (fi >> 16);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l += __shfl_xor_sync(0xffffffffu, l, o);
    nv += __shfl_xor_sync(0xffffffffu, nv, o);
    nc += __shfl_xor_sync(0xffffffffu, nc, o);
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sl[w] = l; sv[w] = nv; sc[w] = nc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tl = 0.0;
    unsigned long long tv = 0, tc = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { tl += sl[i]; tv += sv[i]; tc += sc[i]; }
    sums[0] = (float)tl;
    sums[1] = (float)tv;
    sums[2] = (float)tc;
    sums[3] = (float)(tl / (double)tv);
  }
}

__global__ void __launch_bounds__(256)
smooth_ce_bwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ target,
                     const float* __restrict__ row_lse, const float* __restrict__ sums,
                     const float* __restrict__ grad_out, float* __restrict__ dlogits, int64_t T,
                     int V, float eps, int32_t ignore) {
  const int lane = threadIdx.x & 31;
  int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= T) return;
  const float* z = logits + row * V;
  float* dz = dlogits + row * V;
  int32_t t = target[row];
  if (t == ignore) {
    for (int c = lane; c < V; c += 32) dz[c] = 0.f;
    return;
  }
  float sc = (grad_out ? grad_out[0] : 1.f) / sums[1];
  float lse = row_lse[row];
  float u = eps / (float)V;
  for (int c = lane; c < V; c += 32) {
    float pr = expf(z[c] - lse);
    float q = u + ((c == t) ? (1.f - eps) : 0.f);
    dz[c] = (pr - q) * sc;
  }
}

// =======================================================================================
// Adam over a flat fp32 buffer (torch.optim.Adam: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)), optional bf16 shadow write
// =======================================================================================
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, __nv_bfloat16* __restrict__ p_lp, int64_t n, float lr, float b1,
            float b2, float eps, float bc1, float bc2_sqrt, float gscale) {
  int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 pv = *reinterpret_cast<float4*>(p + i * 4);
    float4 gv = *reinterpret_cast<const float4*>(g + i * 4);
    float4 mv = *reinterpret_cast<float4*>(m + i * 4);
    float4 vv = *reinterpret_cast<float4*>(v + i * 4);
    float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gg = gp[k] * gscale;
      mp[k] = b1 * mp[k] + (1.f - b1) * gg;
      vp[k] = b2 * vp[k] + (1.f - b2) * gg * gg;
      float denom = sqrtf(vp[k]) / bc2_sqrt + eps;
      pp[k] -= (lr / bc1) * (mp[k] / denom);
    }
    *reinterpret_cast<float4*>(p + i * 4) = pv;
    *reinterpret_cast<float4*>(m + i * 4) = mv;
    *reinterpret_cast<float4*>(v + i * 4) = vv;
    if (p_lp) store4<__nv_bfloat16>(p_lp + i * 4, pv);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    float gg = g[i] * gscale;
    float mm = b1 * m[i] + (1.f - b1) * gg;
    float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    m[i] = mm; v[i] = vv;
    float np = p[i] - (lr / bc1) * (mm / (sqrtf(vv) / bc2_sqrt + eps));
    p[i] = np;
    if (p_lp) p_lp[i] = __float2bfloat16_rn(np);
  }
}

static inline int grid_for(int64_t work_items, int threads, int max_waves = 8) {
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace mt

using namespace mt;

// =======================================================================================
// C ABI
// =======================================================================================
extern "C" {

int mt_version(void) { return 100; }
int mt_decode_chain(int enable) { g_chain = enable != 0; return 0; }
const char* mt_last_error(void) { return mt::g_err; }
int mt_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

int mt_embed_pos_fwd(const int32_t* ids, const float* emb, const float* pe, float* out_f32,
                     void* out_lp, int lp_dtype, int64_t B, int64_t L, int64_t d, int64_t V,
                     int64_t pos0, float scale, float p_drop, uint64_t seed, uint64_t site,
                     void* stream) {
  MT_REQUIRE(ids && emb && pe && out_f32, "embed_pos_fwd: null pointer");
  MT_REQUIRE(B > 0 && L > 0 && d > 0 && d % 4 == 0 && V > 0 && pos0 >= 0, "embed_pos_fwd: bad shape B=%ld L=%ld d=%ld", (long)B, (long)L, (long)d);
  MT_REQUIRE(aligned(emb, 16) && aligned(pe, 16) && aligned(out_f32, 16) && aligned(out_lp, 8), "embed_pos_fwd: misaligned buffer");
  MT_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "embed_pos_fwd: p_drop out of range");
  int64_t T = B * L;
  float inv_keep = 1.f / (1.f - p_drop);
  int grid = grid_for(T * (d / 4), 256);
  if (!out_lp) lp_dtype = MT_F32;
  MT_DISPATCH_DTYPE(lp_dtype, TL,
      (embed_pos_fwd_kernel<TL><<<grid, 256, 0, as_stream(stream)>>>(ids, emb, pe, out_f32, (TL*)out_lp, T, L, (int)(d / 4), V, pos0, scale, p_drop, inv_keep, seed, site)));
  return check_launch("embed_pos_fwd");
}

int mt_embed_pos_bwd(const int32_t* ids, const float* dout, float* demb, int64_t B, int64_t L,
                     int64_t d, int64_t V, float scale, float p_drop, uint64_t seed,
                     uint64_t site, void* stream) {
  MT_REQUIRE(ids && dout && demb, "embed_pos_bwd: null pointer");
  MT_REQUIRE(B > 0 && L > 0 && d > 0 && d % 4 == 0 && V > 0, "embed_pos_bwd: bad shape");
  MT_REQUIRE(aligned(dout, 16) && aligned(demb, 16), "embed_pos_bwd: misaligned buffer");
  int64_t T = B * L;
  float inv_keep = 1.f / (1.f - p_drop);
  int grid = grid_for(T * (d / 4), 256);
  embed_pos_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(ids, dout, demb, T, (int)(d / 4), V, scale, p_drop, inv_keep, seed, site);
  return check_launch("embed_pos_bwd");
}

#define MT_LN_NV_DISPATCH(d4, NVC, ...)                 \
  if ((d4) <= 32 * 2) { constexpr int NVC = 2; __VA_ARGS__; }        \
  else if ((d4) <= 32 * 4) { constexpr int NVC = 4; __VA_ARGS__; }   \
  else { constexpr int NVC = 8; __VA_ARGS__; }

int mt_add_ln_fwd(const void* a, int a_dtype, const float* resid, const float* gamma,
                  const float* beta, float* out_f32, void* out_lp, int lp_dtype, float* mean,
                  float* rstd, int64_t T, int64_t d, float eps, float p_drop, uint64_t seed,
                  uint64_t site, void* stream) {
  MT_REQUIRE(a && resid && gamma && beta && out_f32 && mean && rstd, "add_ln_fwd: null pointer");
  MT_REQUIRE(T > 0 && d > 0 && d % 4 == 0 && d <= 1024, "add_ln_fwd: bad shape T=%ld d=%ld", (long)T, (long)d);
  MT_REQUIRE(aligned(a, 8) && aligned(resid, 16) && aligned(gamma, 16) && aligned(beta, 16) && aligned(out_f32, 16) && aligned(out_lp, 8), "add_ln_fwd: misaligned buffer");
  MT_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "add_ln_fwd: p_drop out of range");
  if (a_dtype == MT_F32) MT_REQUIRE(aligned(a, 16), "add_ln_fwd: misaligned a");
  float inv_keep = 1.f / (1.f - p_drop);
  int grid = grid_for(T * 32, 256);
  int d4 = (int)(d / 4);
  if (!out_lp) lp_dtype = MT_F32;
  MT_DISPATCH_F32_BF16(a_dtype, TA, MT_DISPATCH_F32_BF16(lp_dtype, TL, MT_LN_NV_DISPATCH(d4, NVC,
      (launch_chain(add_ln_fwd_kernel<TA, TL, NVC>, dim3(grid), dim3(256), 0, as_stream(stream), (const TA*)a, resid, gamma, beta, out_f32, (TL*)out_lp, mean, rstd, T, (int)d, eps, p_drop, inv_keep, seed, site)))));
  return check_launch("add_ln_fwd");
}

int64_t mt_add_ln_bwd_parts(int64_t T) {
  int64_t blocks = (T + 7) / 8;
  int64_t cap = (int64_t)sm_count() * 2;
  return blocks < cap ? (blocks < 1 ? 1 : blocks) : cap;
}

int mt_add_ln_bwd(const float* dout, const void* a, int a_dtype, const float* resid,
                  const float* gamma, const float* mean, const float* rstd, float* dz,
                  void* da, int da_dtype, float* part, int64_t T, int64_t d, float p_drop,
                  uint64_t seed, uint64_t site, void* stream) {
  MT_REQUIRE(dout && a && resid && gamma && mean && rstd && dz && da && part, "add_ln_bwd: null pointer");
  MT_REQUIRE(T > 0 && d > 0 && d % 4 == 0 && d <= 1024, "add_ln_bwd: bad shape");
  MT_REQUIRE(aligned(dout, 16) && aligned(a, 8) && aligned(resid, 16) && aligned(gamma, 16) && aligned(dz, 16) && aligned(da, 8), "add_ln_bwd: misaligned buffer");
  float inv_keep = 1.f / (1.f - p_drop);
  int grid = (int)mt_add_ln_bwd_parts(T);
  int d4 = (int)(d / 4);
  size_t smem = 3 * 8 * d * sizeof(float);
  MT_REQUIRE(smem <= 48 * 1024 || d <= 1024, "add_ln_bwd: d too large");
  cudaError_t attr_err = cudaSuccess;
  MT_DISPATCH_F32_BF16(a_dtype, TA, MT_DISPATCH_F32_BF16(da_dtype, TD, MT_LN_NV_DISPATCH(d4, NVC, {
      auto kern = add_ln_bwd_kernel<TA, TD, NVC>;
      if (smem > 48 * 1024) attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<grid, 256, smem, as_stream(stream)>>>(dout, (const TA*)a, resid, gamma, mean, rstd, dz, (TD*)da, part, T, (int)d, p_drop, inv_keep, seed, site);
  })));
  if (attr_err != cudaSuccess) { set_error("add_ln_bwd: smem attribute: %s", cudaGetErrorString(attr_err)); return (int)attr_err; }
  return check_launch("add_ln_bwd");
}

int mt_ln_param_grad(const float* part, float* dgamma, float* dbeta, float* dbias, int64_t nparts, int64_t d,
                     void* stream) {
  MT_REQUIRE(part && dgamma && dbeta && nparts > 0 && d > 0, "ln_param_grad: bad args");
  const int64_t warps = 3 * d;
  ln_param_grad_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(part, dgamma, dbeta, dbias, nparts, (int)d);
  return check_launch("ln_param_grad");
}

size_t mt_colsum_workspace_bytes(int64_t M, int64_t N) {
  int64_t chunks = (M + 255) / 256;
  if (chunks > 128) chunks = 128;
  return chunks <= 1 ? 0 : (size_t)(chunks * N * sizeof(float));
}

int mt_colsum(const void* X, int dtype, float* out, int64_t M, int64_t N, int64_t ldx,
              void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(X && out && M > 0 && N > 0 && ldx >= N, "colsum: bad args");
  int64_t chunks = (M + 255) / 256;
  if (chunks > 128) chunks = 128;
  if (chunks <= 1) {
    dim3 grid((unsigned)((N + 31) / 32));
    MT_DISPATCH_DTYPE(dtype, T, (colsum_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)X, out, M, N, ldx)));
    return check_launch("colsum");
  }
  if (!workspace || workspace_bytes < (size_t)(chunks * N * sizeof(float))) {
    set_error("colsum: workspace too small (%zu < %zu)", workspace_bytes, (size_t)(chunks * N * sizeof(float)));
    return MT_E_WORKSPACE;
  }
  int64_t rows_per_chunk = (M + chunks - 1) / chunks;
  const bool vec = (N % 4 == 0) && (ldx % 4 == 0) && aligned(X, 4 * dtype_size(dtype)) && aligned(workspace, 16);
  if (vec) {
    dim3 grid((unsigned)((N / 4 + 31) / 32), (unsigned)chunks);
    MT_DISPATCH_DTYPE(dtype, T, (colsum_partial_vec_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)X, (float*)workspace, M, N, ldx, rows_per_chunk)));
  } else {
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)chunks);
    MT_DISPATCH_DTYPE(dtype, T, (colsum_partial_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)X, (float*)workspace, M, N, ldx, rows_per_chunk)));
  }
  int rc = check_launch("colsum_partial");
  if (rc) return rc;
  colsum_fold_kernel<<<(unsigned)((N + 31) / 32), 256, 0, as_stream(stream)>>>((const float*)workspace, out, chunks, N);
  return check_launch("colsum_fold");
}

int mt_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream) {
  MT_REQUIRE(src && dst && n > 0, "cast: bad args");
  MT_REQUIRE(aligned(src, 8) && aligned(dst, 8), "cast: misaligned");
  if (src_dtype == MT_F32) MT_REQUIRE(aligned(src, 16), "cast: misaligned src");
  if (dst_dtype == MT_F32) MT_REQUIRE(aligned(dst, 16), "cast: misaligned dst");
  int grid = grid_for(n / 4 + 1, 256);
  MT_DISPATCH_DTYPE(src_dtype, TS, MT_DISPATCH_DTYPE(dst_dtype, TD,
      (cast_kernel<TS, TD><<<grid, 256, 0, as_stream(stream)>>>((const TS*)src, (TD*)dst, n))));
  return check_launch("cast");
}

int mt_cast2d(const void* src, int src_dtype, int64_t lds, void* dst, int dst_dtype, int64_t ldd,
              int64_t rows, int64_t cols, void* stream) {
  MT_REQUIRE(src && dst && rows > 0 && cols > 0 && lds >= cols && ldd >= cols, "cast2d: bad args");
  int grid = grid_for(rows * ldd, 256);
  MT_DISPATCH_F32_BF16(src_dtype, TS, MT_DISPATCH_F32_BF16(dst_dtype, TD,
      (cast2d_kernel<TS, TD><<<grid, 256, 0, as_stream(stream)>>>((const TS*)src, (TD*)dst, rows, cols, lds, ldd))));
  return check_launch("cast2d");
}

int mt_transpose_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t rows,
                      int64_t cols, void* stream) {
  MT_REQUIRE(src && dst && rows > 0 && cols > 0, "transpose_cast: bad args");
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  MT_DISPATCH_DTYPE(src_dtype, TS, MT_DISPATCH_DTYPE(dst_dtype, TD,
      (transpose_cast_kernel<TS, TD><<<grid, 256, 0, as_stream(stream)>>>((const TS*)src, (TD*)dst, rows, cols))));
  return check_launch("transpose_cast");
}

// row_loss / row_flags live behind row_lse: caller passes row_lse with room for 3*T floats
int mt_smooth_ce_fwd(const float* logits, const int32_t* target, float* row_lse,
                     int32_t* argmax, float* sums, int64_t T, int64_t V, float eps,
                     int32_t ignore, void* stream) {
  MT_REQUIRE(logits && target && row_lse && argmax && sums, "smooth_ce_fwd: null pointer");
  MT_REQUIRE(T > 0 && V > 0 && V < (1 << 30), "smooth_ce_fwd: bad shape");
  MT_REQUIRE(eps >= 0.f && eps <= 1.f, "smooth_ce_fwd: label_smoothing out of range");
  float* row_loss = row_lse + T;
  float* row_flags = row_lse + 2 * T;
  int grid = (int)((T * 32 + 255) / 256);
  smooth_ce_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(logits, target, row_lse, argmax, row_loss, row_flags, T, (int)V, eps, ignore);
  int rc = check_launch("smooth_ce_fwd");
  if (rc) return rc;
  smooth_ce_reduce_kernel<<<1, 1024, 0, as_stream(stream)>>>(row_loss, row_flags, sums, T);
  return check_launch("smooth_ce_reduce");
}

int mt_smooth_ce_bwd(const float* logits, const int32_t* target, const float* row_lse,
                     const float* sums, const float* grad_out, float* dlogits, int64_t T,
                     int64_t V, float eps, int32_t ignore, void* stream) {
  MT_REQUIRE(logits && target && row_lse && sums && dlogits, "smooth_ce_bwd: null pointer");
  MT_REQUIRE(T > 0 && V > 0, "smooth_ce_bwd: bad shape");
  int grid = (int)((T * 32 + 255) / 256);
  smooth_ce_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(logits, target, row_lse, sums, grad_out, dlogits, T, (int)V, eps, ignore);
  return check_launch("smooth_ce_bwd");
}

int mt_adam_step(float* p, const float* g, float* m, float* v, void* p_lp, int64_t n, float lr,
                 float beta1, float beta2, float eps, int64_t step, float grad_scale,
                 void* stream) {
  MT_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adam_step: bad args");
  MT_REQUIRE(aligned(p, 16) && aligned(g, 16) && aligned(m, 16) && aligned(v, 16) && aligned(p_lp, 8), "adam_step: misaligned");
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2 = 1.f - powf(beta2, (float)step);
  int grid = grid_for(n / 4 + 1, 256);
  adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(p, g, m, v, (__nv_bfloat16*)p_lp, n, lr, beta1, beta2, eps, bc1, sqrtf(bc2), grad_scale);
  return check_launch("adam_step");
}

}  // extern "C"
