// Counter-based Philox4x32-10 used when the caller passes no explicit noise tensors.
// counter = (low 32 bits of the global step | episode, global row, stream,
// 'METR' ^ high 32 bits of the global step), key = 64-bit seed.  The step counter is 64 bits wide so
// that callers spacing their launches by a fixed stride (VectorizedSampler: 2^20 per call) never
// replay a stream, however long the run.
// oracle/rollout.py implements the same integer pipeline; the Box-Muller transform there is
// evaluated in float64 and rounded, here in fp32 (agreement ~1e-6).
#pragma once
#include <stdint.h>

namespace metrpo {

constexpr uint32_t PHILOX_STREAM_EPS = 0u;        // + action block (4 normals per block)
constexpr uint32_t PHILOX_STREAM_IDX = 0x10000u;  // step_rand model index
constexpr uint32_t PHILOX_STREAM_EIDX = 0x10001u; // eps_rand model index (c0 = episode of row)
constexpr uint32_t PHILOX_STREAM_STD = 0x20000u;  // + state block (model_mean_std noise)
constexpr uint32_t PHILOX_C3 = 0x4D455452u;       // 'METR'

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& n0, float& n1) {
  float u1 = (static_cast<float>(xa >> 8) + 1.0f) * 5.9604644775390625e-08f;  // (0,1]
  float u2 = static_cast<float>(xb >> 8) * 5.9604644775390625e-08f;           // [0,1)
  float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);   // exact argument reduction for [0, 2): cheap and accurate
  n0 = r * c;
  n1 = r * s;
}

// 4 N(0,1) values of block `blk` of a stream
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t step, uint32_t row,
                                               uint32_t stream, float (&n)[4]) {
  uint4 x = philox4x32_10(static_cast<uint32_t>(step), row, stream,
                          PHILOX_C3 ^ static_cast<uint32_t>(step >> 32), static_cast<uint32_t>(seed),
                          static_cast<uint32_t>(seed >> 32));
  box_muller(x.x, x.y, n[0], n[1]);
  box_muller(x.z, x.w, n[2], n[3]);
}

__device__ __forceinline__ int philox_index(uint64_t seed, uint64_t counter, uint32_t row,
                                            uint32_t stream, int K) {
  uint4 x = philox4x32_10(static_cast<uint32_t>(counter), row, stream,
                          PHILOX_C3 ^ static_cast<uint32_t>(counter >> 32), static_cast<uint32_t>(seed),
                          static_cast<uint32_t>(seed >> 32));
  return static_cast<int>(__umulhi(x.x, static_cast<uint32_t>(K)));
}

}  // namespace metrpo
