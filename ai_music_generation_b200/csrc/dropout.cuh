// Counter-based dropout masks (nn.Dropout at nanoGPT/model.py:39-40,85,129 and the dropout_p of SDPA at :64).
// A mask bit is a pure function of (site key, row, column), so the backward pass regenerates it instead of storing it:
//   row_key  = fmix32(site_key + row * 0x85EBCA77)
//   h(pair)  = fmix32(row_key ^ (pair * 0x27D4EB2F)),  pair = column >> 1
//   keep(col) = 16-bit lane (col & 1) of h  >=  thr16,  thr16 = round(p * 65536)
// The attention-probability site (one decision per score, inside instruction-issue-bound kernels) uses a cheaper
// generator with the same (site key, row, column) contract — a Weyl step along the row and one squaring-type folded
// 32x32->64 multiply ("wyhash32" construction), two 15-bit lanes per word:
//   a = row_key (above),  b = a * 0x9E3779B1 + 0x7F4A7C15
//   s = a + pair * 0x53C5CA59,  c = (u64)(s ^ b) * s,  h(pair) = lo32(c) ^ hi32(c)
//   keep(col) = 15-bit lane (col & 1) of h  >=  thr15,  thr15 = round(p * 32768)
// The comparison is done for both lanes at once: (h & 0x7FFF7FFF) + k15 carries into bit 15 / bit 31 exactly when the
// lane passes (k15 = (0x8000 - thr15) in both halves), and one PRMT with sign replication turns those bits into AND masks.
// The 1/(1-p) factor is applied once per output (O, dV) or folded into the softmax scale (dP), not per probability.
// (murmur3 finaliser; not torch's Philox stream — dropout parity is therefore checked against the oracle run with the
// SAME masks, regenerated on the host by ai_music_generation_b200/dropout.py.)
#pragma once
#include <stdint.h>

namespace abcgpt {

struct DropCfg {
  uint32_t key;    // site key
  uint32_t thr16;  // 0 => dropout disabled
  float inv_keep;  // 1 / (1 - p)
  uint32_t k15;    // attention site: (0x8000 - thr15) replicated in both 16-bit halves
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

// ---- attention-probability site ----
constexpr uint32_t kDropWeyl = 0x53C5CA59u;
__host__ __device__ __forceinline__ uint32_t drop_row_key2(uint32_t a) { return a * 0x9E3779B1u + 0x7F4A7C15u; }
// s = a + pair * kDropWeyl (callers step it along the row)
__host__ __device__ __forceinline__ uint32_t attn_drop_fold(uint32_t s, uint32_t b) {
  const unsigned long long c = static_cast<unsigned long long>(s ^ b) * s;
  return static_cast<uint32_t>(c) ^ static_cast<uint32_t>(c >> 32);
}
// bit 15 (even column) / bit 31 (odd column) set <=> keep
__host__ __device__ __forceinline__ uint32_t attn_drop_signs(uint32_t h, uint32_t k15) { return (h & 0x7FFF7FFFu) + k15; }

inline DropCfg make_drop(float p, uint32_t key) {
  DropCfg d;
  d.key = key;
  if (p <= 0.f) {
    d.thr16 = 0;
    d.inv_keep = 1.f;
    d.k15 = 0x80008000u;
  } else {
    long t = static_cast<long>(p * 65536.0f + 0.5f);
    d.thr16 = static_cast<uint32_t>(t > 65535 ? 65535 : t);
    d.inv_keep = 1.0f / (1.0f - p);
    long t15 = static_cast<long>(p * 32768.0f + 0.5f);
    if (t15 > 32767) t15 = 32767;
    const uint32_t k = static_cast<uint32_t>(0x8000 - t15);
    d.k15 = (k << 16) | k;
  }
  return d;
}

}  // namespace abcgpt
