// KV-cached autoregressive decode as ONE persistent launch (K5/K7/K8 of the decode path; replaces the loop of
// MT/network.py:52-77 for a whole generation).
//
// A decode step of the stack is 46 dependent kernels of a few dozen activation rows each: even inside one CUDA graph
// with programmatic dependent launch its time is the SUM of their critical paths (launch, fill, drain: 5-10 us
// each, 0.47 ms per event at 32 sequences), not their work.  Here one grid of one CTA per SM stays resident for the
// whole generation and walks the phases of every step itself, separated by grid-wide barriers (a counter in global
// memory: ~2 us):
//     per layer:  QKV projection (prologue: embedding + PE, or the previous layer's second residual+LayerNorm)
//                 attention over the KV cache (+ append of the new K / V rows)
//                 fc
//                 FFN_pre + ReLU (prologue: first residual+LayerNorm)
//                 FFN_suf
//     then:       vocabulary projection (prologue: last residual+LayerNorm), sampler, next position
// 5 barriers per layer + 2.  A projection phase gives each CTA strips of 8 output columns (mma.sync.m16n8k16, the 8
// warps split K); the strip's weight rows are requested BEFORE the barrier that ends the previous phase -- weights
// are constant, so their latency runs under the predecessor's tail -- and the LayerNorm of the few activation rows is
// recomputed by every CTA that owns a strip in its prologue instead of being a phase of its own.  The attention phase
// streams the cached K / V / E rows through a shared-memory ring of bulk copies (requested across the CTA's units and
// before the phase's barrier); the sampler is one warp per row on register-resident logits.
// One rule shaped every phase: the gpu-scope fence of a grid barrier INVALIDATES L1, so nothing is "cached" from phase to
// phase -- every operand a dependency chain needs (pad flags, gamma / beta, biases, uniforms, the new key's rows) is
// requested at the top of the phase, before the chain, or its L2 round trip is exposed once per use.
// Arithmetic is that of the per-kernel path (gemm_skinny.cu, decode.cu, elementwise.cu): 16-bit operands (f16 in the
// first layer's attention block), fp32 accumulation, fp32 residual stream / LayerNorm / softmax.
#include "ops.cuh"
#include "tc_common.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace mt {

namespace {

constexpr int DS_THREADS = 256, DS_WARPS = 8;
constexpr int DS_MAXL = 16;          // layers
constexpr int DS_PAD = 8;            // 16-bit elements of row padding in the shared-memory panels
constexpr int DS_BN = 8;             // output columns of a strip

struct DsLayer {
  const void* Wqkv; const float* bqkv; const void* Wfc; const float* bfc;
  const void* Wpre; const float* bpre; const void* Wsuf; const float* bsuf;
  const float* g1; const float* b1; const float* g2; const float* b2;
  const void* E; void* kc; void* vc;
  int f16;                           // 16-bit type of the layer's attention block: 1 = f16, 0 = bf16
};

struct DsParams {
  DsLayer L[DS_MAXL];
  int layers;
  int32_t* ids; int64_t ld_ids;
  int B, d, h, V, max_seq, mtiles;
  int t0, n_steps, prior_len;
  const float* emb; const float* pe; const void* Wv; const float* bv;
  int32_t pad_token; uint8_t* pad_bits;
  const float* uniforms; float temperature; int top_k, greedy;
  float* logits_out;
  // workspace
  unsigned* barrier;
  float* x; float* a; float* out1; float* f; float* logits;
  void* qkv; void* o; void* hmid;
  long long* prof;                   // MT_DECODE_PROF: [16] accumulated ns of CTA 0 per phase kind (debug aid)
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, bool f16) {
  if (f16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack16(float a, float b, bool f16) {
  if (f16) { __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8], bool f16) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (f16) { const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[i])); f[2 * i] = v.x; f[2 * i + 1] = v.y; }
    else { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
}
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// grid-wide barrier: a monotone counter (zeroed by the launcher), `epoch` counts the barriers this CTA has passed
__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& epoch) {
  ++epoch;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const unsigned target = epoch * gridDim.x;
    unsigned spins = 0;
    while (ld_acquire(ctr) < target) {
      if (++spins > (1u << 26)) { printf("decode_step: grid barrier timeout (block %d epoch %u)\n", blockIdx.x, epoch); __trap(); }
    }
    __threadfence();
  }
  __syncthreads();
}

// shared memory
struct DsSmem {
  uint16_t* A;        // [mtiles * 16][K + PAD] activation panel
  uint16_t* W;        // [2][8][K + PAD] weight strips
  float* red;         // [8 warps][mtiles * 16][8], also the attention / sampler scratch
};

// bytes of the scratch region behind the panels: the split-K fold of the projections / the attention fold
__host__ __device__ inline size_t ds_scratch_bytes(int mtiles, int V) {
  size_t scratch = (size_t)DS_WARPS * mtiles * 16 * 8 * 4;
  const size_t att = (size_t)(32 * 65 + 16) * 4;
  if (att > scratch) scratch = att;
  (void)V;
  return (scratch + 15) / 16 * 16;
}

// request the weight rows of this CTA's strips of a projection (first two strips: the shared-memory slots)
__device__ __forceinline__ void prefetch_strips(const DsSmem& s, const void* W, int N, int K, int nstrips) {
  const int pitch = K + DS_PAD, kv = K / 8;
  const uint16_t* Wg = reinterpret_cast<const uint16_t*>(W);
  int slot = 0;
  for (int st = blockIdx.x; st < nstrips && slot < 2; st += gridDim.x, ++slot) {
    for (int v = threadIdx.x; v < DS_BN * kv; v += DS_THREADS) {
      const int r = v / kv, c = v - r * kv;
      const int n = min(st * DS_BN + r, N - 1);
      cp_async16(s.W + (slot * DS_BN + r) * pitch + c * 8, Wg + (int64_t)n * K + c * 8);
    }
  }
}

// panel <- 16-bit rows [B, K] from global memory (written by another CTA: L2 loads)
__device__ __forceinline__ void stage_plain(const DsSmem& s, const void* src, int B, int K, int mtiles) {
  const int pitch = K + DS_PAD, kv = K / 8;
  const uint16_t* g = reinterpret_cast<const uint16_t*>(src);
  for (int v = threadIdx.x; v < mtiles * 16 * kv; v += DS_THREADS) {
    const int r = v / kv, c = v - r * kv;
    uint16_t* dst = s.A + r * pitch + c * 8;
    if (r < B) cp_async16(dst, g + (int64_t)r * K + c * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
}

// panel <- 16-bit copy of LayerNorm(u + r) gamma + beta (every CTA normalises all rows itself; the CTAs that own a row
// also write the fp32 result for the residual path).  One warp per row, two-pass statistics in registers.
template <int NV>
__device__ __forceinline__ void stage_ln(const DsSmem& s, const float* u, const float* r, const float* gamma, const float* beta,
                                         float* out_f32, int B, int d, int mtiles, bool f16) {
  const int pitch = d + DS_PAD, d4 = d >> 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // gamma / beta of this lane's columns, requested first: every grid barrier invalidates L1 (gpu-scope fence), so they come
  // from L2 in every phase -- loaded after the reductions, that round trip sat at the end of each row's chain
  float4 gam[NV], bet[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + i * 32;
    gam[i] = bet[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < d4) { gam[i] = *reinterpret_cast<const float4*>(gamma + c4 * 4); bet[i] = *reinterpret_cast<const float4*>(beta + c4 * 4); }
  }
  for (int row0 = warp; row0 < mtiles * 16; row0 += 4 * DS_WARPS) {
    // four rows of this warp at a time, every load issued before the first reduction (a row by itself is two L2 round
    // trips plus two warp reductions: ~2 us, four rows in sequence were the most expensive part of a phase)
    float4 z[4][NV];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int row = row0 + q * DS_WARPS;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c4 = lane + i * 32;
        z[q][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < B && c4 < d4) {
          const float4 av = __ldcg(reinterpret_cast<const float4*>(u + (int64_t)row * d) + c4);
          const float4 rv = __ldcg(reinterpret_cast<const float4*>(r + (int64_t)row * d) + c4);
          z[q][i] = make_float4(av.x + rv.x, av.y + rv.y, av.z + rv.z, av.w + rv.w);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int row = row0 + q * DS_WARPS;
      if (row >= mtiles * 16) continue;
      uint16_t* dst = s.A + row * pitch;
      if (row >= B) {
        for (int c = lane; c < d / 8; c += 32) *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(0, 0, 0, 0);
        continue;
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (lane + i * 32 < d4) sum += (z[q][i].x + z[q][i].y) + (z[q][i].z + z[q][i].w);
      const float mu = warp_sum(sum) / (float)d;
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane + i * 32 < d4) {
          const float dx = z[q][i].x - mu, dy = z[q][i].y - mu, dz = z[q][i].z - mu, dw = z[q][i].w - mu;
          v += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
      }
      const float rs = 1.0f / sqrtf(warp_sum(v) / (float)d + 1e-6f);
      const bool mine = (row % (int)gridDim.x) == (int)blockIdx.x;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c4 = lane + i * 32;
        if (c4 < d4) {
          const float4 g = gam[i], bt = bet[i];
          const float4 o = make_float4((z[q][i].x - mu) * rs * g.x + bt.x, (z[q][i].y - mu) * rs * g.y + bt.y,
                                       (z[q][i].z - mu) * rs * g.z + bt.z, (z[q][i].w - mu) * rs * g.w + bt.w);
          *reinterpret_cast<uint2*>(dst + c4 * 4) = make_uint2(pack16(o.x, o.y, f16), pack16(o.z, o.w, f16));
          if (mine) *(reinterpret_cast<float4*>(out_f32 + (int64_t)row * d) + c4) = o;
        }
      }
    }
  }
}

// panel <- embedding * sqrt(d) + PE[t] of the token at position t (MT/layers.py:226-228); pad bit of the position
__device__ __forceinline__ void stage_embed(const DsSmem& s, const DsParams& p, int t, bool f16) {
  const int d = p.d, pitch = d + DS_PAD, d4 = d >> 2;
  const float scale = sqrtf((float)d);
  for (int e = threadIdx.x; e < p.mtiles * 16 * d4; e += DS_THREADS) {
    const int row = e / d4, c4 = e - row * d4;
    uint16_t* dst = s.A + row * pitch + c4 * 4;
    if (row >= p.B) { *reinterpret_cast<uint2*>(dst) = make_uint2(0, 0); continue; }
    int32_t id = __ldcg(p.ids + (int64_t)row * p.ld_ids + t);
    if (blockIdx.x == 0 && c4 == 0) p.pad_bits[(int64_t)row * p.max_seq + t] = (id == p.pad_token) ? 1 : 0;
    id = id < 0 ? 0 : (id >= p.V ? p.V - 1 : id);
    const float4 w = *(reinterpret_cast<const float4*>(p.emb + (int64_t)id * d) + c4);
    const float4 q = *(reinterpret_cast<const float4*>(p.pe + (int64_t)t * d) + c4);
    const float4 o = make_float4(__fadd_rn(__fmul_rn(w.x, scale), q.x), __fadd_rn(__fmul_rn(w.y, scale), q.y),
                                 __fadd_rn(__fmul_rn(w.z, scale), q.z), __fadd_rn(__fmul_rn(w.w, scale), q.w));
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack16(o.x, o.y, f16), pack16(o.z, o.w, f16));
    if ((row % (int)gridDim.x) == (int)blockIdx.x) *(reinterpret_cast<float4*>(p.x + (int64_t)row * d) + c4) = o;
  }
}

enum { OUT_F32 = 0, OUT_16 = 1 };

// C[B, N] = epi(panel . W^T + bias) for this CTA's strips; the panel and the first two strips are in shared memory
__device__ __forceinline__ void run_strips(const DsSmem& s, const DsParams& p, const void* W, const float* bias, int N, int K,
                                           int nstrips, bool f16_in, bool relu, int out_kind, bool out_f16, void* C, float* C2) {
  const int pitch = K + DS_PAD, kv = K / 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = p.mtiles;
  const int kq = ((K / DS_WARPS + 15) / 16) * 16, kbeg = warp * kq, kend = min(K, kbeg + kq);      // (short K: the last warps idle)
  const int lrow = lane & 15, lcol = (lane >> 4) * 8;
  const uint32_t sA_u = static_cast<uint32_t>(__cvta_generic_to_shared(s.A));
  int slot = 0;
  for (int st = blockIdx.x; st < nstrips; st += gridDim.x, ++slot) {
    if (slot >= 2) {          // (more than two strips per CTA: only with very few SMs) -- reload slot 0
      __syncthreads();
      const uint16_t* Wg = reinterpret_cast<const uint16_t*>(W);
      for (int v = tid; v < DS_BN * kv; v += DS_THREADS) {
        const int r = v / kv, c = v - r * kv;
        cp_async16(s.W + r * pitch + c * 8, Wg + (int64_t)min(st * DS_BN + r, N - 1) * K + c * 8);
      }
      cp_async_wait_all();
      __syncthreads();
    }
    const uint16_t* sW = s.W + (slot >= 2 ? 0 : slot) * DS_BN * pitch;
    const int ncol = st * DS_BN + (tid & 7);                     // the column every epilogue element of this thread falls into
    const float bias_c = ncol < N ? __ldg(bias + ncol) : 0.f;    // (requested before the products: L1 is cold after a barrier)
    float acc[4][4];
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
    const uint16_t* wrow = sW + (lane >> 2) * pitch + 2 * (lane & 3);
    for (int k0 = kbeg; k0 < kend; k0 += 16) {
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wrow + k0);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wrow + k0 + 8);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        if (m < mt) {
          uint32_t a[4];
          ldmatrix_x4(a, sA_u + (uint32_t)(((m * 16 + lrow) * pitch + k0 + lcol) * 2));
          mma_16816(acc[m], a, b0, b1, f16_in);
        }
      }
    }
    // fold the eight K slices; C fragment: c0,c1 -> row lane/4, cols 2*(lane%4)+{0,1}; c2,c3 -> row + 8
    __syncthreads();          // (the previous strip's fold has been read)
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (m < mt) {
        float* dst = s.red + ((warp * mt + m) * 16) * 8;
        const int r = lane >> 2, c = 2 * (lane & 3);
        dst[r * 8 + c] = acc[m][0]; dst[r * 8 + c + 1] = acc[m][1];
        dst[(r + 8) * 8 + c] = acc[m][2]; dst[(r + 8) * 8 + c + 1] = acc[m][3];
      }
    }
    __syncthreads();
    const int n0 = st * DS_BN;
    for (int e = tid; e < mt * 16 * 8; e += DS_THREADS) {
      const int r = e >> 3, c = e & 7, n = n0 + c;
      if (r >= p.B || n >= N) continue;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < DS_WARPS; ++w) v += s.red[w * mt * 16 * 8 + e];
      v += bias_c;
      if (relu) v = fmaxf(v, 0.f);
      if (out_kind == OUT_F32) {
        reinterpret_cast<float*>(C)[(int64_t)r * N + n] = v;
        if (C2) C2[(int64_t)r * N + n] = v;
      } else if (out_f16) {
        reinterpret_cast<__half*>(C)[(int64_t)r * N + n] = __float2half_rn(v);
      } else {
        reinterpret_cast<__nv_bfloat16*>(C)[(int64_t)r * N + n] = __float2bfloat16_rn(v);
      }
    }
  }
}

// attention of the new token (position t) of (sequence b, head hh) over keys 0 .. t.  The cached rows 0 .. t-1 of K and
// V and the matching E rows stream through a shared-memory ring of 128-key blocks as 1-D bulk copies (three 16 KB copies
// per block, issued by thread 0 up to DS_STAGES blocks ahead: ~100 KB in flight per SM with no registers spent on it --
// with the loads in registers, one 256-thread CTA per SM kept ~1 TB/s of the 6.5 moving); the new token's K / V rows come
// from the fused projection row and are appended to the caches.  8 lanes share a key row, 32 keys per pass, online
// softmax per key group, folded across the groups at the end (decode.cu: rga_decode_split_kernel's arithmetic).
constexpr int DS_STAGES = 3, DS_BLK = 128, DS_STAGE_BYTES = 3 * DS_BLK * 128;

struct DsRing {
  uint8_t* buf;         // [DS_STAGES][3][DS_BLK * 128 B]
  uint64_t* full;       // [DS_STAGES]
  uint64_t* empty;      // [DS_STAGES] (one arrival per warp)
  uint32_t issued, consumed;      // block counters of the whole launch (slot = counter % DS_STAGES, phase = counter / DS_STAGES)
  int cu, cblk;         // producer cursor (thread 0): next (unit, block) of the current attention phase
};

// thread 0: request blocks of the CTA's units of this attention phase, in consumption order, while ring slots are free.
// The cached rows do not depend on the step's projection, so the first blocks are requested BEFORE the grid barrier
// that precedes the phase, and a unit's first blocks while the previous unit is still being consumed.
__device__ __forceinline__ void ring_fill(DsRing& ring, const DsParams& p, const DsLayer& Ly, int t) {
  const int nblk = (t + DS_BLK - 1) / DS_BLK, units = p.B * p.h;
  while (ring.cu < units && nblk > 0 && ring.issued - ring.consumed < (uint32_t)DS_STAGES) {
    const uint32_t c = ring.issued, slot = c % DS_STAGES;
    if (c >= DS_STAGES) tc::mbar_wait(&ring.empty[slot], ((c / DS_STAGES) - 1) & 1);
    const int blk = ring.cblk;
    const int64_t bh = ring.cu;               // unit index = b * h + hh
    const uint8_t* kb = reinterpret_cast<const uint8_t*>(Ly.kc) + (bh * (int64_t)p.max_seq + (int64_t)blk * DS_BLK) * 128;
    const uint8_t* vb = reinterpret_cast<const uint8_t*>(Ly.vc) + (bh * (int64_t)p.max_seq + (int64_t)blk * DS_BLK) * 128;
    const uint8_t* eb = reinterpret_cast<const uint8_t*>(Ly.E) + ((int64_t)(p.max_seq - 1 - t) + (int64_t)blk * DS_BLK) * 128;
    const uint32_t bytes = (uint32_t)min(DS_BLK, t - blk * DS_BLK) * 128u;
    uint8_t* dst = ring.buf + slot * DS_STAGE_BYTES;
    tc::mbar_arrive_expect_tx(&ring.full[slot], 3 * bytes);
    tc::bulk_load_1d(dst, kb, bytes, &ring.full[slot]);
    tc::bulk_load_1d(dst + DS_BLK * 128, eb, bytes, &ring.full[slot]);
    tc::bulk_load_1d(dst + 2 * DS_BLK * 128, vb, bytes, &ring.full[slot]);
    ring.issued = c + 1;
    if (++ring.cblk == nblk) { ring.cblk = 0; ring.cu += (int)gridDim.x; }
  }
}

__device__ __forceinline__ void attend_unit(const DsSmem& s, DsRing& ring, const DsParams& p, const DsLayer& Ly, int b, int hh,
                                            int t) {
  constexpr int DH = 64, LPK = 8, KPI = DS_THREADS / LPK;     // 32 keys per pass
  const bool f16 = Ly.f16 != 0;
  const int tid = threadIdx.x, sub = tid % LPK, grp = tid / LPK;
  const int h = p.h, d = p.d, max_seq = p.max_seq;
  const int64_t bh = (int64_t)b * h + hh;
  const uint32_t LMASK = ((1u << LPK) - 1u) << ((tid & 31) / LPK * LPK);
  const uint16_t* qkv = reinterpret_cast<const uint16_t*>(p.qkv) + (int64_t)b * 3 * d;
  float q8[8];
  {
    const uint4 qr = __ldcg(reinterpret_cast<const uint4*>(qkv + hh * DH + sub * 8));
    unpack8(qr, q8, f16);
#pragma unroll
    for (int e = 0; e < 8; ++e) q8[e] *= 0.125f;              // 1 / sqrt(64)
  }
  const uint8_t* pad = p.pad_bits + (int64_t)b * max_seq;
  const uint8_t* kb = reinterpret_cast<const uint8_t*>(Ly.kc) + bh * (int64_t)max_seq * 128;
  const uint8_t* vb = reinterpret_cast<const uint8_t*>(Ly.vc) + bh * (int64_t)max_seq * 128;
  const uint8_t* eb = reinterpret_cast<const uint8_t*>(Ly.E) + (int64_t)(max_seq - 1 - t) * 128;     // E row of key 0
  const int nblk = (t + DS_BLK - 1) / DS_BLK;                 // blocks of cached keys 0 .. t-1
  const uint4 enew = *reinterpret_cast<const uint4*>(eb + (int64_t)t * 128 + sub * 16);      // (requested with q: cold L1)
  const bool pad_new = pad[t] != 0;
  float m_run = -INFINITY, l_run = 0.f, o8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o8[e] = 0.f;
  auto one_key = [&](const uint4& kr, const uint4& er, const uint4& vr, bool use) {
    float kf[8], ef[8], vf[8];
    unpack8(kr, kf, f16); unpack8(er, ef, f16); unpack8(vr, vf, f16);
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc = fmaf(q8[e], kf[e] + ef[e], acc);
#pragma unroll
    for (int o = LPK / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(LMASK, acc, o);
    const float a2 = use ? acc : -INFINITY;
    const float m_new = fmaxf(m_run, a2);
    const float corr = (m_new == -INFINITY) ? 1.f : __expf(m_run - m_new);
    const float pj = use ? __expf(acc - m_new) : 0.f;
    l_run = fmaf(l_run, corr, pj);
#pragma unroll
    for (int e = 0; e < 8; ++e) o8[e] = fmaf(o8[e], corr, pj * vf[e]);
    m_run = m_new;
  };
  // ---- the new token (key t): rows from the projection, stored into the caches; handled by key group 0
  {
    const uint4 knew = __ldcg(reinterpret_cast<const uint4*>(qkv + (h + hh) * DH + sub * 8));
    const uint4 vnew = __ldcg(reinterpret_cast<const uint4*>(qkv + (2 * h + hh) * DH + sub * 8));
    if (grp == 0) {
      *reinterpret_cast<uint4*>(const_cast<uint8_t*>(kb) + (int64_t)t * 128 + sub * 16) = knew;
      *reinterpret_cast<uint4*>(const_cast<uint8_t*>(vb) + (int64_t)t * 128 + sub * 16) = vnew;
    }
    if (grp == 0) one_key(knew, enew, vnew, !pad_new);       // (whole key groups take the branch: the shuffles stay inside a group)
  }
  // ---- the cached keys, block by block out of the ring
  for (int blk = 0; blk < nblk; ++blk) {
    const uint32_t c = ring.consumed, slot = c % DS_STAGES;
    const int j0 = blk * DS_BLK, nk = min(DS_BLK, t - j0);
    // the block's pad flags BEFORE the wait for its rows: every grid barrier invalidates L1 (gpu-scope fence), so each
    // phase re-reads the flags from L2 -- inside the softmax chain that round trip was exposed once per block
    bool padk[DS_BLK / KPI];
#pragma unroll
    for (int ps = 0; ps < DS_BLK / KPI; ++ps) {
      const int jj = ps * KPI + grp;
      padk[ps] = (jj < nk) ? (pad[j0 + jj] != 0) : true;
    }
    tc::mbar_wait(&ring.full[slot], (c / DS_STAGES) & 1);
    const uint8_t* src = ring.buf + slot * DS_STAGE_BYTES + sub * 16;
    // the block's four keys of this key group at once: four independent dot products and group reductions, then ONE
    // update of the running (max, sum, o) -- key by key the update chain (reduction -> max -> exp -> fma, ~250 cycles)
    // ran four times per block with only two warps per scheduler to hide it
    float acc4[4];
    uint4 vr4[4];
#pragma unroll
    for (int ps = 0; ps < DS_BLK / KPI; ++ps) {
      const int jj = ps * KPI + grp;
      acc4[ps] = -INFINITY;
      vr4[ps] = make_uint4(0, 0, 0, 0);
      if (jj < nk) {       // (uniform per key group)
        const uint4 kr = *reinterpret_cast<const uint4*>(src + jj * 128);
        const uint4 er = *reinterpret_cast<const uint4*>(src + DS_BLK * 128 + jj * 128);
        vr4[ps] = *reinterpret_cast<const uint4*>(src + 2 * DS_BLK * 128 + jj * 128);
        float kf[8], ef[8];
        unpack8(kr, kf, f16); unpack8(er, ef, f16);
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = fmaf(q8[e], kf[e] + ef[e], acc);
#pragma unroll
        for (int o = LPK / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(LMASK, acc, o);
        acc4[ps] = padk[ps] ? -INFINITY : acc;
      }
    }
    {
      const float m_new = fmaxf(fmaxf(m_run, fmaxf(acc4[0], acc4[1])), fmaxf(acc4[2], acc4[3]));
      const float corr = (m_new == -INFINITY) ? 1.f : __expf(m_run - m_new);
      float pj[4];
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) pj[ps] = (acc4[ps] == -INFINITY) ? 0.f : __expf(acc4[ps] - m_new);
      l_run = fmaf(l_run, corr, (pj[0] + pj[1]) + (pj[2] + pj[3]));
#pragma unroll
      for (int e = 0; e < 8; ++e) o8[e] *= corr;
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        float vf[8];
        unpack8(vr4[ps], vf, f16);
#pragma unroll
        for (int e = 0; e < 8; ++e) o8[e] = fmaf(pj[ps], vf[e], o8[e]);
      }
      m_run = m_new;
    }
    ring.consumed = c + 1;
    tc::mbar_arrive_warp(&ring.empty[slot]);
    if (tid == 0) ring_fill(ring, p, Ly, t);
  }
  // fold the 32 key groups: M = max m_g, weights exp(m_g - M)
  float* red = s.red;                       // [KPI][DH + 1]
  float* wred = s.red + KPI * (DH + 1);     // [16]
  float mx = warp_max(m_run);
  if ((tid & 31) == 0) wred[tid >> 5] = mx;
  __syncthreads();
  mx = wred[0];
#pragma unroll
  for (int w = 1; w < DS_WARPS; ++w) mx = fmaxf(mx, wred[w]);
  const float wg = (m_run == -INFINITY) ? 0.f : __expf(m_run - mx);
#pragma unroll
  for (int e = 0; e < 8; ++e) red[grp * (DH + 1) + sub * 8 + e] = o8[e] * wg;
  float sum = (sub == 0) ? l_run * wg : 0.f;
  sum = warp_sum(sum);
  if ((tid & 31) == 0) wred[8 + (tid >> 5)] = sum;
  __syncthreads();
  if (tid < DH) {
    float tot = 0.f;
#pragma unroll 8
    for (int g = 0; g < KPI; ++g) tot += red[g * (DH + 1) + tid];
    float L = 0.f;
#pragma unroll
    for (int w = 0; w < DS_WARPS; ++w) L += wred[8 + w];
    const float r = L > 0.f ? tot / L : 0.f;
    uint16_t* out = reinterpret_cast<uint16_t*>(p.o) + (int64_t)b * d + hh * DH + tid;
    if (f16) *reinterpret_cast<__half*>(out) = __float2half_rn(r);
    else *reinterpret_cast<__nv_bfloat16*>(out) = __float2bfloat16_rn(r);
  }
  __syncthreads();                          // (the scratch is reused by the CTA's next unit)
}

// temperature / top-k / inverse-CDF sampler of one row by ONE warp (decode.cu: sample_kernel's rule: keep the top_k
// values, ties to the lower id; softmax over the kept ones; inverse CDF over ascending ids with the given uniform;
// greedy = first arg-max).  Lane l holds ids l*C .. l*C+C-1 (C = ceil(V / 32)): ascending ids = lane order, then the
// order inside a lane, so the CDF is a lane-local prefix plus a warp scan of the lane sums.
template <int C>
__device__ __forceinline__ void sample_row_warp(const DsParams& p, int b, int t) {
  const int V = p.V, lane = threadIdx.x & 31;
  const float* zb = p.logits + (int64_t)b * V;
  float z[C];
  bool keep[C];
#pragma unroll
  for (int i = 0; i < C; ++i) {
    const int c = lane * C + i;
    const float v = c < V ? __ldcg(zb + c) : -INFINITY;
    z[i] = (p.greedy || c >= V) ? v : v / p.temperature;
    keep[i] = false;
  }
  int32_t* out = p.ids + (int64_t)b * p.ld_ids + (t + 1);
  const float u_row = p.greedy ? 0.f : __ldg(p.uniforms + (int64_t)(t + 1 - p.prior_len) * p.B + b);
  const bool full = p.greedy || p.top_k <= 0 || p.top_k >= V;
  const int rounds = p.greedy ? 1 : (full ? 0 : p.top_k);
  for (int it = 0; it < rounds; ++it) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < C; ++i) {
      const int c = lane * C + i;
      if (c < V && !keep[i] && (z[i] > bv || (z[i] == bv && c < bi))) { bv = z[i]; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (p.greedy) {
      if (lane == 0) *out = bi;
      return;
    }
#pragma unroll
    for (int i = 0; i < C; ++i)
      if (lane * C + i == bi) keep[i] = true;
  }
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < C; ++i)
    if (lane * C + i < V && (full || keep[i])) mx = fmaxf(mx, z[i]);
  mx = warp_max(mx);
  float pr[C], lsum = 0.f;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    pr[i] = (lane * C + i < V && (full || keep[i])) ? expf(z[i] - mx) : 0.f;
    lsum += pr[i];
  }
  const float tot = warp_sum(lsum);
  float lnorm = 0.f;
  int last_alive = -1;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    pr[i] = pr[i] / tot;
    lnorm += pr[i];
    if (pr[i] > 0.f) last_alive = lane * C + i;
  }
  // exclusive scan of the lane sums, total
  float incl = lnorm;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const float total = __shfl_sync(0xffffffffu, incl, 31);
  float cdf = incl - lnorm;
  const float target = u_row * total;
  int pick = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    cdf += pr[i];
    if (pick == 0x7fffffff && lane * C + i < V && !(cdf <= target)) pick = lane * C + i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pick = min(pick, __shfl_xor_sync(0xffffffffu, pick, o));
    last_alive = max(last_alive, __shfl_xor_sync(0xffffffffu, last_alive, o));
  }
  if (last_alive < 0) last_alive = 0;
  if (pick == 0x7fffffff || pick > last_alive) pick = last_alive;
  if (lane == 0) *out = pick;
}

__device__ __forceinline__ void sample_row(const DsParams& p, int b, int t) {
  const int c = (p.V + 31) / 32;
  if (c <= 8) sample_row_warp<8>(p, b, t);
  else sample_row_warp<16>(p, b, t);        // (V <= 512: mt_decode_run_supported)
}

template <int NV>
__global__ void __launch_bounds__(DS_THREADS, 1)
decode_run_kernel(const __grid_constant__ DsParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int d = p.d, pitchA = d + DS_PAD;
  DsSmem s;
  s.A = reinterpret_cast<uint16_t*>(smem_raw);
  s.W = s.A + p.mtiles * 16 * pitchA;
  s.red = reinterpret_cast<float*>(s.W + 2 * DS_BN * pitchA);
  DsRing ring;
  {
    // (red scratch, then the ring's barriers, then -- 1024-byte aligned -- its stages)
    uint8_t* after = reinterpret_cast<uint8_t*>(s.red) + ds_scratch_bytes(p.mtiles, p.V);
    ring.full = reinterpret_cast<uint64_t*>(after);
    ring.empty = ring.full + DS_STAGES;
    const uintptr_t base = (reinterpret_cast<uintptr_t>(after) + 64 + 1023) & ~(uintptr_t)1023;
    ring.buf = reinterpret_cast<uint8_t*>(base);
    ring.issued = ring.consumed = 0; ring.cu = 0; ring.cblk = 0;
    if (threadIdx.x == 0) {
      for (int q = 0; q < DS_STAGES; ++q) { tc::mbar_init(&ring.full[q], 1); tc::mbar_init(&ring.empty[q], DS_WARPS); }
      tc::fence_barrier_init();
    }
    __syncthreads();
  }
  unsigned epoch = 0;
  const int G = gridDim.x;
  const int nl = p.layers;
  const int ns_qkv = (3 * d + DS_BN - 1) / DS_BN, ns_d = (d + DS_BN - 1) / DS_BN, ns_h = (d / 2 + DS_BN - 1) / DS_BN,
            ns_v = (p.V + DS_BN - 1) / DS_BN;

  long long prof_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, prof_t = 0;
  const bool prof_on = p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (prof_on) prof_t = gtime();
#define DS_PROF(k) do { if (prof_on) { const long long now = gtime(); prof_acc[k] += now - prof_t; prof_t = now; } } while (0)
  prefetch_strips(s, p.L[0].Wqkv, 3 * d, d, ns_qkv);
  for (int step = 0; step < p.n_steps; ++step) {
    const int t = p.t0 + step;
    for (int li = 0; li < nl; ++li) {
      const DsLayer& Ly = p.L[li];
      const bool f16 = Ly.f16 != 0;
      // ---- QKV projection (its weight strips were requested before the last barrier)
      if (li == 0) stage_embed(s, p, t, f16);
      else stage_ln<NV>(s, p.f, p.out1, p.L[li - 1].g2, p.L[li - 1].b2, p.x, p.B, d, p.mtiles, f16);
      cp_async_wait_all();
      __syncthreads();
      DS_PROF(0);
      run_strips(s, p, Ly.Wqkv, Ly.bqkv, 3 * d, d, ns_qkv, f16, false, OUT_16, f16, p.qkv, nullptr);
      __syncthreads();
      prefetch_strips(s, Ly.Wfc, d, d, ns_d);
      ring.cu = blockIdx.x; ring.cblk = 0;
      if (threadIdx.x == 0) ring_fill(ring, p, Ly, t);      // the first blocks of the attention phase, under the barrier
      DS_PROF(1);
      grid_sync(p.barrier, epoch);
      DS_PROF(9);
      // ---- attention (+ KV append)
      for (int u = blockIdx.x; u < p.B * p.h; u += G) attend_unit(s, ring, p, Ly, u / p.h, u % p.h, t);
      DS_PROF(2);
      grid_sync(p.barrier, epoch);
      DS_PROF(9);
      // ---- fc
      if ((int)blockIdx.x < ns_d) stage_plain(s, p.o, p.B, d, p.mtiles);
      cp_async_wait_all();
      __syncthreads();
      run_strips(s, p, Ly.Wfc, Ly.bfc, d, d, ns_d, f16, false, OUT_F32, false, p.a, nullptr);
      __syncthreads();
      prefetch_strips(s, Ly.Wpre, d / 2, d, ns_h);
      DS_PROF(3);
      grid_sync(p.barrier, epoch);
      DS_PROF(9);
      // ---- FFN_pre + ReLU on LayerNorm(a + x)
      if ((int)blockIdx.x < max(ns_h, p.B))       // (CTAs without a strip or a row of their own skip the prologue: L2 traffic)
        stage_ln<NV>(s, p.a, p.x, Ly.g1, Ly.b1, p.out1, p.B, d, p.mtiles, false);
      cp_async_wait_all();
      __syncthreads();
      DS_PROF(4);
      run_strips(s, p, Ly.Wpre, Ly.bpre, d / 2, d, ns_h, false, true, OUT_16, false, p.hmid, nullptr);
      __syncthreads();
      prefetch_strips(s, Ly.Wsuf, d, d / 2, ns_d);
      DS_PROF(5);
      grid_sync(p.barrier, epoch);
      DS_PROF(9);
      // ---- FFN_suf
      if ((int)blockIdx.x < ns_d) stage_plain(s, p.hmid, p.B, d / 2, p.mtiles);
      cp_async_wait_all();
      __syncthreads();
      run_strips(s, p, Ly.Wsuf, Ly.bsuf, d, d / 2, ns_d, false, false, OUT_F32, false, p.f, nullptr);
      __syncthreads();
      if (li + 1 < nl) prefetch_strips(s, p.L[li + 1].Wqkv, 3 * d, d, ns_qkv);
      else prefetch_strips(s, p.Wv, p.V, d, ns_v);
      DS_PROF(6);
      grid_sync(p.barrier, epoch);
      DS_PROF(9);
    }
    // ---- vocabulary projection on the last LayerNorm(f + out1)
    if ((int)blockIdx.x < max(ns_v, p.B))
      stage_ln<NV>(s, p.f, p.out1, p.L[nl - 1].g2, p.L[nl - 1].b2, p.x, p.B, d, p.mtiles, false);
    cp_async_wait_all();
    __syncthreads();
    run_strips(s, p, p.Wv, p.bv, p.V, d, ns_v, false, false, OUT_F32, false, p.logits,
               p.logits_out ? p.logits_out + (int64_t)step * p.B * p.V : nullptr);
    __syncthreads();
    if (step + 1 < p.n_steps) prefetch_strips(s, p.L[0].Wqkv, 3 * d, d, ns_qkv);
    DS_PROF(7);
    grid_sync(p.barrier, epoch);
    DS_PROF(9);
    // ---- sampler: the event drawn from position t's logits becomes the token at t + 1 unless that position still
    // belongs to the prior
    if (t + 1 >= p.prior_len) {             // one warp per row
      const int wrow = (int)blockIdx.x * DS_WARPS + (threadIdx.x >> 5);
      if (wrow < p.B) sample_row(p, wrow, t);
    }
    DS_PROF(8);
    grid_sync(p.barrier, epoch);
    DS_PROF(9);
  }
  cp_async_wait_all();
  if (prof_on)
    for (int k = 0; k < 10; ++k) p.prof[k] = prof_acc[k];
#undef DS_PROF
}

size_t ds_smem_bytes(int mtiles, int d, int V) {
  const size_t panel = (size_t)(mtiles * 16 + 2 * DS_BN) * (d + DS_PAD) * 2;
  return panel + ds_scratch_bytes(mtiles, V) + 64 + 1024 + (size_t)DS_STAGES * DS_STAGE_BYTES + 16;
}

struct DsWs { size_t off_x, off_a, off_out1, off_f, off_logits, off_qkv, off_o, off_hmid, total; };
DsWs ds_layout(int64_t B, int64_t d, int64_t V) {
  DsWs w;
  size_t o = 256;
  auto take = [&](size_t n) { size_t r = o; o += (n + 255) / 256 * 256; return r; };
  w.off_x = take(B * d * 4); w.off_a = take(B * d * 4); w.off_out1 = take(B * d * 4); w.off_f = take(B * d * 4);
  w.off_logits = take(B * V * 4); w.off_qkv = take(B * 3 * d * 2); w.off_o = take(B * d * 2); w.off_hmid = take(B * (d / 2) * 2);
  w.total = o;
  return w;
}

}  // namespace
}  // namespace mt

using namespace mt;

extern "C" {

size_t mt_decode_run_workspace_bytes(int64_t B, int64_t d, int64_t V) {
  if (B <= 0 || d <= 0 || V <= 0) return 0;
  return ds_layout(B, d, V).total;
}

int mt_decode_run_supported(int64_t B, int64_t d, int64_t h, int64_t V, int64_t layers) {
  if (B < 1 || B > 64 || layers < 1 || layers > DS_MAXL) return 0;
  if (h < 1 || d != h * 64 || d % 32 != 0 || d > 1024) return 0;           // head dim 64; K = d and d / 2 in mma k-steps of 16
  if (V < 1 || V > 512) return 0;                                          // (the sampler holds a row in one warp's registers)
  if (ds_smem_bytes((int)((B + 15) / 16), (int)d, (int)V) > 232448) return 0;     // 227 KB of dynamic shared memory per CTA
  return mt_device_ok() != 0;
}

int mt_decode_run(int32_t* ids, int64_t ld_ids, int64_t B, int64_t t0, int64_t n_steps, int64_t prior_len,
                  const float* emb, const float* pe, const void* const* layer_ptrs, const int32_t* layer_f16,
                  int64_t layers, const void* Wv, const float* bv, int64_t d, int64_t h, int64_t V, int64_t max_seq,
                  int32_t pad_token, uint8_t* pad_bits, const float* uniforms, float temperature, int32_t top_k,
                  int greedy, float* logits_out, void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(ids && emb && pe && layer_ptrs && layer_f16 && Wv && bv && pad_bits && workspace, "decode_run: null pointer");
  MT_REQUIRE(mt_decode_run_supported(B, d, h, V, layers), "decode_run: shape not served by the persistent decode kernel (B=%ld d=%ld h=%ld V=%ld layers=%ld)",
             (long)B, (long)d, (long)h, (long)V, (long)layers);
  MT_REQUIRE(t0 >= 0 && n_steps >= 1 && t0 + n_steps <= max_seq && ld_ids > t0 + n_steps && prior_len >= 1, "decode_run: bad positions (t0=%ld n_steps=%ld max_seq=%ld ld_ids=%ld)",
             (long)t0, (long)n_steps, (long)max_seq, (long)ld_ids);
  MT_REQUIRE(greedy || (uniforms != nullptr && temperature > 0.f), "decode_run: need uniforms and temperature > 0");
  const DsWs w = ds_layout(B, d, V);
  MT_REQUIRE(workspace_bytes >= w.total && aligned(workspace, 256), "decode_run: workspace of %zu bytes (256-byte aligned) needed", w.total);
  DsParams p;
  for (int li = 0; li < (int)layers; ++li) {
    const void* const* q = layer_ptrs + (size_t)li * 15;
    for (int x = 0; x < 15; ++x) MT_REQUIRE(q[x] != nullptr, "decode_run: null layer pointer (layer %d, slot %d)", li, x);
    DsLayer& Ly = p.L[li];
    Ly.Wqkv = q[0]; Ly.bqkv = (const float*)q[1]; Ly.Wfc = q[2]; Ly.bfc = (const float*)q[3];
    Ly.Wpre = q[4]; Ly.bpre = (const float*)q[5]; Ly.Wsuf = q[6]; Ly.bsuf = (const float*)q[7];
    Ly.g1 = (const float*)q[8]; Ly.b1 = (const float*)q[9]; Ly.g2 = (const float*)q[10]; Ly.b2 = (const float*)q[11];
    Ly.E = q[12]; Ly.kc = const_cast<void*>(q[13]); Ly.vc = const_cast<void*>(q[14]);
    Ly.f16 = layer_f16[li];
    MT_REQUIRE(aligned(Ly.Wqkv, 16) && aligned(Ly.Wfc, 16) && aligned(Ly.Wpre, 16) && aligned(Ly.Wsuf, 16) && aligned(Ly.E, 16) &&
               aligned(Ly.kc, 16) && aligned(Ly.vc, 16) && aligned(Ly.g1, 16) && aligned(Ly.b1, 16) && aligned(Ly.g2, 16) && aligned(Ly.b2, 16),
               "decode_run: misaligned layer operand (layer %d)", li);
  }
  MT_REQUIRE(aligned(Wv, 16) && aligned(emb, 16) && aligned(pe, 16), "decode_run: misaligned operand");
  p.layers = (int)layers;
  p.ids = ids; p.ld_ids = ld_ids;
  p.B = (int)B; p.d = (int)d; p.h = (int)h; p.V = (int)V; p.max_seq = (int)max_seq; p.mtiles = (int)((B + 15) / 16);
  p.t0 = (int)t0; p.n_steps = (int)n_steps; p.prior_len = (int)prior_len;
  p.emb = emb; p.pe = pe; p.Wv = Wv; p.bv = bv;
  p.pad_token = pad_token; p.pad_bits = pad_bits;
  p.uniforms = uniforms; p.temperature = temperature; p.top_k = top_k; p.greedy = greedy;
  p.logits_out = logits_out;
  uint8_t* base = static_cast<uint8_t*>(workspace);
  p.barrier = reinterpret_cast<unsigned*>(base);
  p.x = reinterpret_cast<float*>(base + w.off_x); p.a = reinterpret_cast<float*>(base + w.off_a);
  p.out1 = reinterpret_cast<float*>(base + w.off_out1); p.f = reinterpret_cast<float*>(base + w.off_f);
  p.logits = reinterpret_cast<float*>(base + w.off_logits);
  p.qkv = base + w.off_qkv; p.o = base + w.off_o; p.hmid = base + w.off_hmid;
  cudaStream_t st = as_stream(stream);
  static const bool want_prof = getenv("MT_DECODE_PROF") != nullptr;
  static long long* prof_dev = nullptr;
  p.prof = nullptr;
  if (want_prof) {
    if (!prof_dev) cudaMalloc(&prof_dev, 16 * sizeof(long long));
    p.prof = prof_dev;
  }
  cudaError_t e = cudaMemsetAsync(p.barrier, 0, 256, st);
  if (e != cudaSuccess) { set_error("decode_run: memset: %s", cudaGetErrorString(e)); return (int)e; }
  const size_t smem = ds_smem_bytes(p.mtiles, p.d, p.V);
  const int d4 = p.d / 4, nv = (d4 + 31) / 32;
  // every CTA must be resident at once (the grid barrier spins): one CTA per SM, cooperative launch
  void* args[] = {(void*)&p};
  const dim3 grid((unsigned)sm_count()), block(DS_THREADS);
#define MT_DS_LAUNCH(NVC)                                                                                           \
  {                                                                                                                 \
    auto kern = decode_run_kernel<NVC>;                                                                             \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                         \
    if (e != cudaSuccess) { set_error("decode_run: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; } \
    e = cudaLaunchCooperativeKernel((const void*)kern, grid, block, args, smem, st);                                \
    if (e != cudaSuccess) { set_error("decode_run: cooperative launch: %s", cudaGetErrorString(e)); return (int)e; } \
  }
  if (nv <= 2) MT_DS_LAUNCH(2)
  else if (nv <= 4) MT_DS_LAUNCH(4)
  else if (nv <= 6) MT_DS_LAUNCH(6)
  else MT_DS_LAUNCH(8)
#undef MT_DS_LAUNCH
  if (want_prof) {
    long long host[10];
    cudaMemcpyAsync(host, prof_dev, sizeof(host), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    static const char* nm[10] = {"qkv prologue", "qkv strips", "attention", "fc", "pre prologue (LN)", "pre strips", "suf", "vocab", "sampler", "barrier wait"};
    for (int k = 0; k < 10; ++k) fprintf(stderr, "decode_run prof (CTA 0, %ld steps): %-18s %9.1f us/step\n", (long)n_steps, nm[k], host[k] / 1e3 / (double)n_steps);
  }
  return check_launch("decode_run");
}

}  // extern "C"
