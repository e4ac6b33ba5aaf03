// Data feed: HBM-resident token arena and the window gather of the reference's batch builders.
//
// Reference: MT/data.py:41-67 (`batch` draws `batch_size` files with random.sample and one window per
// file with random.randrange, `_get_seq` :96-107; `slide_seq2seq_batch` returns x = w[:, :-1],
// y = w[:, 1:]; `seq2seq_batch` x = w[:, :L], y = w[:, L:]) followed by MT/train.py:258-260 (numpy int16 ->
// torch int32 on the device).  Here every `.data` file lives once in one uint16 / uint8 arena in HBM
// (file f occupies arena[file_off[f] .. file_off[f+1])) and a batch is ONE launch that writes the int32
// x and y windows directly; the per-step host->device traffic is the B window starts (or nothing with
// the on-device sampler).  HBM-bound integer work: B*(L+shift)*s bytes read, 2*B*L*4 bytes written.
#include "common.cuh"
#include "../../include/mt_b200.h"

namespace mt {
namespace {

// one CTA row-chunk: thread t handles 4 consecutive tokens of row b (16-byte int32 stores; the 2-byte
// source loads of a warp cover 256 contiguous bytes)
template <typename TOK>
__global__ void window_gather_kernel(const TOK* __restrict__ arena, const int64_t* __restrict__ starts,
                                     int32_t* __restrict__ x, int32_t* __restrict__ y, int64_t L,
                                     int64_t y_len, int64_t y_shift) {
  const int64_t b = blockIdx.y;
  const int64_t t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const TOK* src = arena + starts[b];
  if (t0 < L) {
    int32_t v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (t0 + i < L) ? (int32_t)src[t0 + i] : 0;
    int32_t* dst = x + b * L + t0;
    if (t0 + 4 <= L && ((L & 3) == 0)) {
      *reinterpret_cast<int4*>(dst) = make_int4(v[0], v[1], v[2], v[3]);
    } else {
      for (int i = 0; i < 4 && t0 + i < L; ++i) dst[i] = v[i];
    }
  }
  if (y != nullptr && t0 < y_len) {
    int32_t v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (t0 + i < y_len) ? (int32_t)src[y_shift + t0 + i] : 0;
    int32_t* dst = y + b * y_len + t0;
    if (t0 + 4 <= y_len && ((y_len & 3) == 0)) {
      *reinterpret_cast<int4*>(dst) = make_int4(v[0], v[1], v[2], v[3]);
    } else {
      for (int i = 0; i < 4 && t0 + i < y_len; ++i) dst[i] = v[i];
    }
  }
}

__device__ __forceinline__ uint32_t mix32(uint32_t v) {   // murmur3 finaliser
  v ^= v >> 16; v *= 0x85ebca6bu; v ^= v >> 13; v *= 0xc2b2ae35u; v ^= v >> 16;
  return v;
}

// Keyed bijection of [0, n): 4-round balanced Feistel network on 2*hb bits + cycle walking.  Row b of
// step s takes file perm(b): distinct rows get distinct files (random.sample draws without replacement).
__device__ int64_t feistel_perm(int64_t i, int64_t n, uint32_t k0, uint32_t k1) {
  int bits = 2;
  while (((int64_t)1 << bits) < n) bits += 2;
  const int hb = bits / 2;
  const uint32_t hm = (1u << hb) - 1u;
  uint64_t v = (uint64_t)i;
  do {
    uint32_t l = (uint32_t)(v >> hb) & hm, r = (uint32_t)v & hm;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
      uint32_t f = mix32(r ^ (k0 + 0x9e3779b9u * round) ^ mix32(k1 + round)) & hm;
      uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    v = ((uint64_t)l << hb) | r;
  } while ((int64_t)v >= n);
  return (int64_t)v;
}

// starts[b] = file_off[f] + u,  f = eligible[perm(b)],  u uniform in [0, len_f - need)
__global__ void window_sample_kernel(const int64_t* __restrict__ file_off, const int64_t* __restrict__ eligible,
                                     int64_t n_eligible, int64_t need, uint64_t seed, uint64_t step,
                                     int64_t* __restrict__ starts, int64_t* __restrict__ files, int64_t B) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  uint4 r = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)(step >> 32), 0x66656564u, 0u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const int64_t f = eligible[feistel_perm(b, n_eligible, r.x, r.y)];
  const int64_t len = file_off[f + 1] - file_off[f];
  uint4 q = philox4x32_10(make_uint4((uint32_t)step, (uint32_t)(step >> 32), 0x66656564u, (uint32_t)b + 1u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint64_t span = (uint64_t)(len - need);          // > 0: the host filtered on len > need
  const uint64_t u64 = ((uint64_t)q.x << 32) | q.y;
  const int64_t u = (int64_t)__umul64hi(u64, span);       // floor(u64 * span / 2^64), bias < 2^-40
  starts[b] = file_off[f] + u;
  if (files) files[b] = f;
}

}  // namespace
}  // namespace mt

using namespace mt;

extern "C" {

int mt_window_gather(const void* arena, int token_bytes, const int64_t* starts, int32_t* x, int32_t* y,
                     int64_t B, int64_t L, int64_t y_len, int64_t y_shift, void* stream) {
  MT_REQUIRE(arena && starts && x && B > 0 && L > 0, "window_gather: bad args");
  MT_REQUIRE(token_bytes == 1 || token_bytes == 2, "window_gather: tokens are uint8 or uint16");
  MT_REQUIRE(y == nullptr || (y_len > 0 && y_shift >= 0), "window_gather: bad y window");
  MT_REQUIRE(B <= 65535, "window_gather: more than 65535 rows");
  MT_REQUIRE(aligned(x, 16) && (y == nullptr || aligned(y, 16)), "window_gather: misaligned output");
  const int64_t longest = (y && y_len > L) ? y_len : L;
  dim3 grid((unsigned)((longest + 4 * 256 - 1) / (4 * 256)), (unsigned)B);
  if (token_bytes == 2)
    window_gather_kernel<uint16_t><<<grid, 256, 0, as_stream(stream)>>>((const uint16_t*)arena, starts, x, y, L,
                                                                        y ? y_len : 0, y_shift);
  else
    window_gather_kernel<uint8_t><<<grid, 256, 0, as_stream(stream)>>>((const uint8_t*)arena, starts, x, y, L,
                                                                       y ? y_len : 0, y_shift);
  return check_launch("window_gather");
}

int mt_window_sample(const int64_t* file_off, const int64_t* eligible, int64_t n_eligible, int64_t need,
                     uint64_t seed, uint64_t step, int64_t* starts, int64_t* files, int64_t B, void* stream) {
  MT_REQUIRE(file_off && eligible && starts && B > 0 && need > 0, "window_sample: bad args");
  MT_REQUIRE(n_eligible >= B, "window_sample: fewer eligible files than rows (random.sample would raise)");
  MT_REQUIRE(n_eligible < ((int64_t)1 << 40), "window_sample: too many files");
  window_sample_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(
      file_off, eligible, n_eligible, need, seed, step, starts, files, B);
  return check_launch("window_sample");
}

}  // extern "C"
