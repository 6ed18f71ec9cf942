// Counter-based dropout masks (nn.Dropout at nanoGPT/model.py:39-40,85,129 and the dropout_p of SDPA at :64).
// A mask bit is a pure function of (site key, row, column), so the backward pass regenerates it instead of storing it:
//   row_key  = fmix32(site_key + row * 0x85EBCA77)
//   h(pair)  = fmix32(row_key ^ (pair * 0x27D4EB2F)),  pair = column >> 1
//   keep(col) = 16-bit lane (col & 1) of h  >=  thr16,  thr16 = round(p * 65536)
// (murmur3 finaliser; not torch's Philox stream — dropout parity is therefore checked against the oracle run with the
// SAME masks, regenerated on the host by ai_music_generation_b200/dropout.py.)
#pragma once
#include <stdint.h>

namespace abcgpt {

struct DropCfg {
  uint32_t key;    // site key
  uint32_t thr16;  // 0 => dropout disabled
  float inv_keep;  // 1 / (1 - p)
};

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t drop_row_key(uint32_t site_key, uint32_t row) {
  return fmix32(site_key + row * 0x85EBCA77u);
}
// bits for columns (2*pair, 2*pair+1) of a row
__host__ __device__ __forceinline__ uint32_t drop_pair_bits(uint32_t row_key, uint32_t pair) {
  return fmix32(row_key ^ (pair * 0x27D4EB2Fu));
}
__host__ __device__ __forceinline__ bool drop_keep_lo(uint32_t bits, uint32_t thr16) { return (bits & 0xFFFFu) >= thr16; }
__host__ __device__ __forceinline__ bool drop_keep_hi(uint32_t bits, uint32_t thr16) { return (bits >> 16) >= thr16; }

inline DropCfg make_drop(float p, uint32_t key) {
  DropCfg d;
  d.key = key;
  if (p <= 0.f) {
    d.thr16 = 0;
    d.inv_keep = 1.f;
  } else {
    long t = static_cast<long>(p * 65536.0f + 0.5f);
    d.thr16 = static_cast<uint32_t>(t > 65535 ? 65535 : t);
    d.inv_keep = 1.0f / (1.0f - p);
  }
  return d;
}

}  // namespace abcgpt
