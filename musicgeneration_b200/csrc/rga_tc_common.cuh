// Pieces shared by the tcgen05 relative-attention kernels (forward and backward).
//
// Thread organisation of the "softmax" part of both kernels: 8 warps = 2 warpgroups.  Warp w
// works on TMEM lanes 32*(w&3) .. +31 (the only lanes it may address), i.e. query row
// a = 32*(w&3) + lane of the 128-row tile; warpgroup wg = w>>2 owns key columns [64*wg, 64*wg+64)
// of that row.  Two warps per scheduler hide each other's TMEM / shared-memory / MUFU latency.
#pragma once

#include "tc_common.cuh"

namespace mt {
namespace rga {

constexpr int TT = 128;                 // tile edge (queries and keys)
constexpr int DHC = 64;                 // head dim
constexpr int TILE = TT * DHC * 2;      // 16 KB: one [128 x 64] 16-bit operand tile
constexpr int SM_WARPS = 8;             // softmax warps
constexpr int SM_THREADS = SM_WARPS * 32;
constexpr int NTHREADS = SM_THREADS + 64;   // + TMA producer warp + MMA issuer warp
constexpr int SCR_WORDS = 28;           // per-thread skew scratch (== 28 mod 32: conflict-free v4 stores)
constexpr int SCR_BYTES = SM_THREADS * SCR_WORDS * 4;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack16(float a, float b, int fmt) {
  return fmt == 1 ? pack_bf16x2(a, b) : pack_f16x2(a, b);
}

// The skew (MT/layers.py:116-125) as index arithmetic.  Adds
//     Srel[a][64*wg + x] = [G_lo | G_hi][a][127 - a + 64*wg + x],   x in [0, 64)
// to sv[x].  TMEM column addresses are warp-uniform, so per pass of 16 output columns the warp
// loads the 48-column window common to its 32 rows, parks it (packed to f16 pairs) in a private
// 28-word shared-memory scratch and reads it back at the per-lane offset 31 - lane; an odd offset
// is a 16-bit funnel shift of adjacent words.
__device__ __forceinline__ void skew_add_64(float (&sv)[64], uint32_t g_lo, uint32_t g_hi, uint32_t lane_base,
                                            int w4, int wg, int lane, uint32_t* scr) {
  const int o = 31 - lane;
  const uint32_t sh = (uint32_t)(o & 1) * 16u;
  const uint32_t* rd = scr + (o >> 1);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int w0 = 96 - 32 * w4 + 64 * wg + 16 * q;     // multiple of 16: each x16 load is in lo or in hi
    uint32_t r[3][16];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int cc = w0 + 16 * c;
      tc::tmem_ld_32x16((cc < 128 ? g_lo + cc : g_hi + (cc - 128)) + lane_base, r[c]);
    }
    tc::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int x = 0; x < 16; x += 8)
        *reinterpret_cast<uint4*>(scr + c * 8 + x / 2) =
            make_uint4(pack_f16x2(__uint_as_float(r[c][x]), __uint_as_float(r[c][x + 1])),
                       pack_f16x2(__uint_as_float(r[c][x + 2]), __uint_as_float(r[c][x + 3])),
                       pack_f16x2(__uint_as_float(r[c][x + 4]), __uint_as_float(r[c][x + 5])),
                       pack_f16x2(__uint_as_float(r[c][x + 6]), __uint_as_float(r[c][x + 7])));
    uint32_t w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = rd[k];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint32_t u = __funnelshift_r(w[k], w[k + 1], sh);
      float2 f = __half22float2(*reinterpret_cast<__half2*>(&u));
      sv[q * 16 + 2 * k] += f.x;
      sv[q * 16 + 2 * k + 1] += f.y;
    }
  }
}

// ---- 32-column unit of the skew (used by the 16-warp forward and the P-producer warps of the
// backward).  A thread owns row a = 32*w4 + lane and the 32 key columns [c_first, c_first+32) with
// c_first a multiple of 32.  Its band columns are 127 - a + c_first + x = w0 + o + x with the
// warp-uniform w0 = 96 - 32*w4 + c_first (a multiple of 32, so each 32-column TMEM load lies in
// G_lo or in G_hi) and the per-lane o = 31 - lane.  skew_park_64 parks the 64-column window
// [w0, w0+64) as 32 f16 pairs in the thread's private scratch (SCR32_WORDS apart: 16-byte stores
// conflict-free); skew_fetch_32 reads it back at offset o (odd o = 16-bit funnel shift).
constexpr int SCR32_WORDS = 36;
__device__ __forceinline__ void skew_park_64(uint32_t g_lo, uint32_t g_hi, uint32_t lane_base, int w0,
                                             uint32_t* scr) {
  uint32_t r0[32], r1[32];
  const int c0 = w0, c1 = w0 + 32;
  tc::tmem_ld_32x32((c0 < 128 ? g_lo + c0 : g_hi + (c0 - 128)) + lane_base, r0);
  tc::tmem_ld_32x32((c1 < 128 ? g_lo + c1 : g_hi + (c1 - 128)) + lane_base, r1);
  tc::tmem_ld_wait();
#pragma unroll
  for (int x = 0; x < 32; x += 8)
    *reinterpret_cast<uint4*>(scr + x / 2) =
        make_uint4(pack_f16x2(__uint_as_float(r0[x]), __uint_as_float(r0[x + 1])),
                   pack_f16x2(__uint_as_float(r0[x + 2]), __uint_as_float(r0[x + 3])),
                   pack_f16x2(__uint_as_float(r0[x + 4]), __uint_as_float(r0[x + 5])),
                   pack_f16x2(__uint_as_float(r0[x + 6]), __uint_as_float(r0[x + 7])));
#pragma unroll
  for (int x = 0; x < 32; x += 8)
    *reinterpret_cast<uint4*>(scr + 16 + x / 2) =
        make_uint4(pack_f16x2(__uint_as_float(r1[x]), __uint_as_float(r1[x + 1])),
                   pack_f16x2(__uint_as_float(r1[x + 2]), __uint_as_float(r1[x + 3])),
                   pack_f16x2(__uint_as_float(r1[x + 4]), __uint_as_float(r1[x + 5])),
                   pack_f16x2(__uint_as_float(r1[x + 6]), __uint_as_float(r1[x + 7])));
}
// sv[x] += window[o + x], x in [0, 32)
__device__ __forceinline__ void skew_fetch_add_32(float (&sv)[32], const uint32_t* scr, int lane) {
  const int o = 31 - lane;
  const uint32_t sh = (uint32_t)(o & 1) * 16u;
  const uint32_t* rd = scr + (o >> 1);
  uint32_t w[17];
#pragma unroll
  for (int k = 0; k < 17; ++k) w[k] = rd[k];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    uint32_t u = __funnelshift_r(w[k], w[k + 1], sh);
    float2 f = __half22float2(*reinterpret_cast<__half2*>(&u));
    sv[2 * k] += f.x;
    sv[2 * k + 1] += f.y;
  }
}

// ---- the same 32-column unit WITHOUT shared memory: the 64-column window stays in registers as 32
// f16 pairs and is moved by the per-lane offset o = 31 - lane with a barrel shifter -- four stages of
// word selects (8, 4, 2, 1 words; the predicates are constants of the thread) and one 16-bit funnel
// shift for odd o.  ~80 SEL per 32 outputs instead of 8 STS.64 + 17 LDS: the attention kernels are
// bound by shared-memory bandwidth (UMMA operand reads alone take most of it), not by issue slots.
__device__ __forceinline__ void skew_window_64(uint32_t g_lo, uint32_t g_hi, uint32_t lane_base, int w0,
                                               uint32_t (&W)[32]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    const int cc = w0 + 32 * c;
    tc::tmem_ld_32x32((cc < 128 ? g_lo + cc : g_hi + (cc - 128)) + lane_base, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int x = 0; x < 32; x += 2) W[16 * c + x / 2] = pack_f16x2(__uint_as_float(r[x]), __uint_as_float(r[x + 1]));
  }
}
// sv[x] += window[o + x], x in [0, 32), o = 31 - lane
__device__ __forceinline__ void skew_shift_add_32(float (&sv)[32], const uint32_t (&W)[32], int lane) {
  const int o = 31 - lane;
  const bool b8 = o & 16, b4 = o & 8, b2 = o & 4, b1 = o & 2;      // word offset o >> 1 = 8*b8 + 4*b4 + 2*b2 + b1
  const uint32_t sh = (uint32_t)(o & 1) * 16u;
  uint32_t A[24], B[20], C[18], D[17];
#pragma unroll
  for (int i = 0; i < 24; ++i) A[i] = b8 ? W[i + 8] : W[i];
#pragma unroll
  for (int i = 0; i < 20; ++i) B[i] = b4 ? A[i + 4] : A[i];
#pragma unroll
  for (int i = 0; i < 18; ++i) C[i] = b2 ? B[i + 2] : B[i];
#pragma unroll
  for (int i = 0; i < 17; ++i) D[i] = b1 ? C[i + 1] : C[i];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    uint32_t u = __funnelshift_r(D[k], D[k + 1], sh);
    float2 f = __half22float2(*reinterpret_cast<__half2*>(&u));
    sv[2 * k] += f.x;
    sv[2 * k + 1] += f.y;
  }
}

// Byte offset of 16-byte chunk `chunk` (0..7) of row `a` inside a [128 x 64] 16-bit tile stored in
// the UMMA 128B-swizzled layout (rows of 128 B, chunk index XOR-ed with row & 7).
__device__ __forceinline__ int swz_chunk(int a, int chunk) { return a * 128 + ((chunk ^ (a & 7)) << 4); }
// Byte offset of 32-bit word `win` (0..31) of row `a` in such a tile.
__device__ __forceinline__ int swz_word(int a, int win) {
  return a * 128 + ((((win >> 2) ^ (a & 7)) << 4) | ((win & 3) << 2));
}

// dG band store.  A thread holds the 64 dS values of row a / key columns [64*half, +64) as 32 packed
// bf16 words A[]; in band coordinates they are columns 127-a+64*half .. +63 of the [128 x 256] dG
// operand (4 sub-tiles of 64 columns, 128B-swizzled rows).  base_w = ((127-a)>>1) + 32*half is the
// first 32-bit word of the run; for even a the run starts at an odd column, so every word takes a
// half from two neighbours and the first / last element are 16-bit stores.  The word addresses are
// chunk address (9 per thread, 16-byte granules) + one of 4 in-chunk offsets: two integer ops per
// store instead of the full swizzle arithmetic.
template <int NW>
__device__ __forceinline__ void band_store_n(uint8_t* dg_base, int a, int base_w, const uint32_t (&A)[NW]) {
  constexpr int NC = NW / 4 + 1;
  const int cb = base_w >> 2, r0 = base_w & 3, a7 = a & 7;
  uint8_t* const rowbase = dg_base + a * 128;
  uint8_t* ca[NC];
#pragma unroll
  for (int q = 0; q < NC; ++q) {
    const int c = cb + q;
    ca[q] = rowbase + (c >> 3) * TILE + (((c & 7) ^ a7) << 4);
  }
  int offs[4];
  bool carry[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { carry[j] = (r0 + j) >= 4; offs[j] = ((r0 + j) & 3) * 4; }
  auto wp = [&](int k) -> uint8_t* {
    const int j = k & 3, q = k >> 2;
    return ((q < NC - 1 && carry[j]) ? ca[q + 1 < NC ? q + 1 : NC - 1] : ca[q]) + offs[j];
  };
  if (a & 1) {
#pragma unroll
    for (int k = 0; k < NW; ++k) *reinterpret_cast<uint32_t*>(wp(k)) = A[k];
  } else {
    *reinterpret_cast<uint16_t*>(wp(0) + 2) = (uint16_t)(A[0] & 0xffffu);
#pragma unroll
    for (int k = 1; k < NW; ++k) *reinterpret_cast<uint32_t*>(wp(k)) = __byte_perm(A[k - 1], A[k], 0x5432);
    *reinterpret_cast<uint16_t*>(wp(NW)) = (uint16_t)(A[NW - 1] >> 16);
  }
}
// (a thread may also hold only 32 of the values -- NW = 16 words, base_w = ((127-a)>>1) + 16*quarter)
__device__ __forceinline__ void band_store(uint8_t* dg_base, int a, int base_w, const uint32_t (&A)[32]) {
  band_store_n<32>(dg_base, a, base_w, A);
}

}  // namespace rga
}  // namespace mt
