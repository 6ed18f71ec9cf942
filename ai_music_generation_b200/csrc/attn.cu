// Causal flash attention, forward and backward, head size 64, on tcgen05 tensor cores with TMEM accumulators.
// Replaces F.scaled_dot_product_attention(q, k, v, is_causal=True) at nanoGPT/model.py:64 together with the head
// split / merge copies at model.py:56-59,72: the kernels read q|k|v straight out of the packed c_attn output
// [B*T, 3C] through one TMA tensor map (column offset selects q/k/v and the head) and write [B*T, C] directly.
//
//   attn_fwd_kernel   one CTA per (128 query rows, batch*head); K/V streamed in 128-row tiles
//                     S = Q K^T (UMMA 128x128x64) -> TMEM -> online softmax, one thread per query row
//                     -> P (bf16) into 128B-swizzled smem -> O_tile = P V (UMMA 128x64x128) -> TMEM -> registers
//   attn_bwd_dq_kernel    one CTA per 128 query rows, K/V in 64-row tiles:   S, dP -> dS -> dQ += dS K   (TMEM acc)
//   attn_bwd_dkv_kernel   one CTA per 128 key rows,  Q/dO in 64-row tiles:  S^T, dP^T -> P^T, dS^T -> dV += P^T dO,
//                         dK += dS^T Q (TMEM acc).  Two kernels instead of atomics on dQ: deterministic gradients.
// Two CTAs are co-resident per SM so that one CTA's softmax (MUFU-bound) overlaps the other's MMAs.
#include "common.h"
#include "dropout.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace abcgpt {

long long* g_attn_trace = nullptr;  // debug only (abcgpt_debug_attn_trace): per-phase clock64 stamps of one CTA

namespace {

constexpr int HS = 64;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kScale = 0.125f;  // 1/sqrt(64)
constexpr float kSl2 = kScale * kLog2e;
constexpr int kThreads = 192;  // warp 0 TMA, warp 1 MMA, warps 2..5 one thread per tile row

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// write 32 consecutive bf16 columns (16 packed words) of one row into a [rows x 64] K-major SWIZZLE_128B slab
__device__ __forceinline__ void st_slab32(uint32_t slab_addr, int row, int c32, const uint32_t* pk) {
  const uint32_t base = slab_addr + row * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t addr = base + (((c32 * 4 + q) ^ (row & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                 "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                 : "memory");
  }
}

// K-major SW128 operand: rows x 64 bf16 slab, 16-column K step k16
__device__ __forceinline__ uint64_t desc_k(uint32_t slab_addr, int k16) {
  return ptx::umma_smem_desc(slab_addr + k16 * 32, 0, 1024);
}
// MN-major SW128 operand over a [k rows x 64] slab (64 contiguous MN elements per row), K step of 16 rows
__device__ __forceinline__ uint64_t desc_mn(uint32_t slab_addr, int k16) {
  return ptx::umma_smem_desc(slab_addr + k16 * 2048, 8192, 1024);
}


__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 f2(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

// Per-chunk mask classes (a chunk = 32 consecutive score columns seen by one warp = 32 consecutive rows).  Row and
// column ranges are both 32-aligned, so a chunk is either entirely visible, entirely masked, or the diagonal one.
constexpr int kFull = 0, kDiag = 1, kMasked = 2;

// forward pass 1: row max of one chunk
template <int MODE>
__device__ __forceinline__ float fwd_chunk_max(uint32_t taddr, int lane) {
  if (MODE == kMasked) return -1e30f;
  uint32_t v[32];
  ptx::tmem_ld32(taddr, v);
  ptx::tmem_ld_wait();
  float m0 = -1e30f, m1 = -1e30f, m2 = -1e30f, m3 = -1e30f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float a = __uint_as_float(v[i]), b = __uint_as_float(v[i + 1]), c = __uint_as_float(v[i + 2]), d = __uint_as_float(v[i + 3]);
    if (MODE == kDiag) {
      a = (i <= lane) ? a : -1e30f;
      b = (i + 1 <= lane) ? b : -1e30f;
      c = (i + 2 <= lane) ? c : -1e30f;
      d = (i + 3 <= lane) ? d : -1e30f;
    }
    m0 = fmaxf(m0, a); m1 = fmaxf(m1, b); m2 = fmaxf(m2, c); m3 = fmaxf(m3, d);
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// forward pass 2: p = 2^(s*sl2 - m), bf16 P into the swizzled slab, returns the fp32 row-sum contribution
template <int MODE>
__device__ __forceinline__ float fwd_chunk_exp(uint32_t taddr, int lane, float neg_m, uint32_t slab, int r, int c32) {
  uint32_t pk[16];
  if (MODE == kMasked) {
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = 0u;
    st_slab32(slab, r, c32, pk);
    return 0.f;
  }
  uint32_t v[32];
  ptx::tmem_ld32(taddr, v);
  ptx::tmem_ld_wait();
  const float2 sl = make_float2(kSl2, kSl2), nm = make_float2(neg_m, neg_m);
  float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float2 t0 = __ffma2_rn(f2(v[2 * i], v[2 * i + 1]), sl, nm);
    const float2 t1 = __ffma2_rn(f2(v[2 * i + 2], v[2 * i + 3]), sl, nm);
    float2 p0 = make_float2(ex2(t0.x), ex2(t0.y));
    float2 p1 = make_float2(ex2(t1.x), ex2(t1.y));
    if (MODE == kDiag) {
      p0.x = (2 * i <= lane) ? p0.x : 0.f;
      p0.y = (2 * i + 1 <= lane) ? p0.y : 0.f;
      p1.x = (2 * i + 2 <= lane) ? p1.x : 0.f;
      p1.y = (2 * i + 3 <= lane) ? p1.y : 0.f;
    }
    acc0 = __fadd2_rn(acc0, p0);
    acc1 = __fadd2_rn(acc1, p1);
    pk[i] = ptx::pack_bf16x2(p0.x, p0.y);
    pk[i + 1] = ptx::pack_bf16x2(p1.x, p1.y);
  }
  st_slab32(slab, r, c32, pk);
  return (acc0.x + acc0.y) + (acc1.x + acc1.y);
}

// backward (dQ kernel): dS = P * (dP*scale - delta*scale) for one chunk; row statistics are per thread
template <int MODE, bool DROP>
__device__ __forceinline__ void dq_chunk(uint32_t taddr_s, uint32_t taddr_dp, int lane, float neg_lse2, float neg_delta8,
                                         uint32_t* pk, const DropCfg& dcfg, uint32_t drop_rk, int kv0) {
  if (MODE == kMasked) {
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = 0u;
    return;
  }
  uint32_t s[32], dp[32];
  ptx::tmem_ld32(taddr_s, s);
  ptx::tmem_ld32(taddr_dp, dp);
  ptx::tmem_ld_wait();
  const float2 sl = make_float2(kSl2, kSl2), nl = make_float2(neg_lse2, neg_lse2);
  const float2 sc = make_float2(kScale, kScale), nd = make_float2(neg_delta8, neg_delta8);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 t = __ffma2_rn(f2(s[2 * i], s[2 * i + 1]), sl, nl);
    const float2 p = make_float2(ex2(t.x), ex2(t.y));
    float2 dpe = f2(dp[2 * i], dp[2 * i + 1]);
    if (DROP) {  // dP = (dO V^T) o mask / (1-p)
      const uint32_t bits = drop_pair_bits(drop_rk, static_cast<uint32_t>(kv0 + 2 * i) >> 1);
      dpe.x = drop_keep_lo(bits, dcfg.thr16) ? dpe.x * dcfg.inv_keep : 0.f;
      dpe.y = drop_keep_hi(bits, dcfg.thr16) ? dpe.y * dcfg.inv_keep : 0.f;
    }
    const float2 u = __ffma2_rn(dpe, sc, nd);
    float2 d = __fmul2_rn(p, u);
    if (MODE == kDiag) {  // keep column <= row
      d.x = (2 * i <= lane) ? d.x : 0.f;
      d.y = (2 * i + 1 <= lane) ? d.y : 0.f;
    }
    pk[i] = ptx::pack_bf16x2(d.x, d.y);
  }
}

// backward (dK/dV kernel): P^T and dS^T for one chunk; statistics are per COLUMN (query), read from shared memory.
// MODE kDiag here is the general path: keep iff (q >= kv) && (q < T).
template <int MODE, bool DROP>
__device__ __forceinline__ void dkv_chunk(uint32_t taddr_s, uint32_t taddr_dp, uint32_t st_lse2, uint32_t st_delta8,
                                          int q_base, int kv_t, int T, uint32_t* pk_p, uint32_t* pk_ds, const DropCfg& dcfg,
                                          uint32_t st_rowkey) {
  if (MODE == kMasked) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { pk_p[i] = 0u; pk_ds[i] = 0u; }
    return;
  }
  uint32_t s[32], dp[32];
  ptx::tmem_ld32(taddr_s, s);
  ptx::tmem_ld32(taddr_dp, dp);
  ptx::tmem_ld_wait();
  const float2 sl = make_float2(kSl2, kSl2), sc = make_float2(kScale, kScale);
#pragma unroll
  for (int i4 = 0; i4 < 8; ++i4) {
    const float4 l4 = lds128(st_lse2 + 16 * i4);     // already negated: -lse*log2e   (ld.shared, broadcast)
    const float4 d4 = lds128(st_delta8 + 16 * i4);   // already negated: -delta*scale
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * i4 + h;  // pair index: columns 2i, 2i+1
      const float2 nl = h == 0 ? make_float2(l4.x, l4.y) : make_float2(l4.z, l4.w);
      const float2 nd = h == 0 ? make_float2(d4.x, d4.y) : make_float2(d4.z, d4.w);
      const float2 t = __ffma2_rn(f2(s[2 * i], s[2 * i + 1]), sl, nl);
      float2 p = make_float2(ex2(t.x), ex2(t.y));
      float2 dpe = f2(dp[2 * i], dp[2 * i + 1]);
      float2 pd = p;  // the (dropped) probabilities that multiply dO in dV
      if (DROP) {
        // the mask row is the QUERY (a column here), so every element needs its own hash: lane (kv & 1) of pair kv >> 1
        uint32_t rk0, rk1;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(rk0), "=r"(rk1) : "r"(st_rowkey + 8 * i));
        const uint32_t sh = (kv_t & 1) * 16, pr = static_cast<uint32_t>(kv_t) >> 1;
        const bool k0 = ((drop_pair_bits(rk0, pr) >> sh) & 0xFFFFu) >= dcfg.thr16;
        const bool k1 = ((drop_pair_bits(rk1, pr) >> sh) & 0xFFFFu) >= dcfg.thr16;
        pd.x = k0 ? p.x * dcfg.inv_keep : 0.f;
        pd.y = k1 ? p.y * dcfg.inv_keep : 0.f;
        dpe.x = k0 ? dpe.x * dcfg.inv_keep : 0.f;
        dpe.y = k1 ? dpe.y * dcfg.inv_keep : 0.f;
      }
      const float2 u = __ffma2_rn(dpe, sc, nd);
      float2 d = __fmul2_rn(p, u);
      if (MODE == kDiag) {
        const int q0 = q_base + 2 * i;
        const bool k0 = (q0 >= kv_t) && (q0 < T), k1 = (q0 + 1 >= kv_t) && (q0 + 1 < T);
        pd.x = k0 ? pd.x : 0.f; d.x = k0 ? d.x : 0.f;
        pd.y = k1 ? pd.y : 0.f; d.y = k1 ? d.y : 0.f;
      }
      pk_p[i] = ptx::pack_bf16x2(pd.x, pd.y);
      pk_ds[i] = ptx::pack_bf16x2(d.x, d.y);
    }
  }
}

// ======================================================================================================
// forward
// ======================================================================================================
// 128 query rows per CTA, K/V streamed in 64-row tiles through a 3-stage ring.  The score tile S (128 x 64 fp32) is
// double-buffered in TMEM so the tensor core computes S_{j+1} while the softmax threads work on S_j; P is
// double-buffered in shared memory; O accumulates in TMEM across all tiles (tcgen05.mma accumulate), so the softmax
// threads never touch O until the end.  That needs a per-row reference exponent that is NOT the running max:
// m_ref is the true max of the first tile and is only raised (with an in-TMEM rescale of O) when a later tile exceeds
// it by more than 2^64 -- exact in floating point, because a common power-of-two factor cancels in O / l.
struct FwdSmem {
  static constexpr int Q = 0;                 // 128 x 64 bf16
  static constexpr int KV = 16384;            // 3 stages x (K 64x64 | V 64x64)
  static constexpr int BAR = 16384 + 3 * 16384;  // P never touches shared memory: it is written back into TMEM
  static constexpr int TOTAL = BAR + 256 + 1024;
};
constexpr float kRescaleThreshold = 64.0f;  // log2 units

// p = 2^(s*sl2 - m_ref) for one 32-column chunk, also tracks the raw row max; MODE as for the other chunk helpers
// DROP: attention dropout (SDPA dropout_p, model.py:64): the row sum uses the undropped probabilities, the P fed to P V is
// masked and scaled by 1/(1-p); mask bit = f(site key, row (b,h,q), key position), regenerated in the backward kernels.
template <int MODE, bool DROP>
__device__ __forceinline__ void fwd_chunk(uint32_t taddr, int lane, float neg_m, float& tmax, float& rowsum, uint32_t* pk,
                                          const DropCfg& dcfg, uint32_t drop_rk, int kv0) {
  if (MODE == kMasked) {
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = 0u;
    return;
  }
  uint32_t v[32];
  ptx::tmem_ld32(taddr, v);
  ptx::tmem_ld_wait();
  const float2 sl = make_float2(kSl2, kSl2), nm = make_float2(neg_m, neg_m);
  float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
  float m0 = tmax, m1 = -1e30f;
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    float s0 = __uint_as_float(v[2 * i]), s1 = __uint_as_float(v[2 * i + 1]);
    float s2 = __uint_as_float(v[2 * i + 2]), s3 = __uint_as_float(v[2 * i + 3]);
    if (MODE == kDiag) {
      s0 = (2 * i <= lane) ? s0 : -1e30f;
      s1 = (2 * i + 1 <= lane) ? s1 : -1e30f;
      s2 = (2 * i + 2 <= lane) ? s2 : -1e30f;
      s3 = (2 * i + 3 <= lane) ? s3 : -1e30f;
    }
    m0 = fmaxf(m0, fmaxf(s0, s1));
    m1 = fmaxf(m1, fmaxf(s2, s3));
    const float2 t0 = __ffma2_rn(make_float2(s0, s1), sl, nm);
    const float2 t1 = __ffma2_rn(make_float2(s2, s3), sl, nm);
    float2 p0 = make_float2(ex2(t0.x), ex2(t0.y));  // masked entries: 2^(-huge) = 0
    float2 p1 = make_float2(ex2(t1.x), ex2(t1.y));
    acc0 = __fadd2_rn(acc0, p0);
    acc1 = __fadd2_rn(acc1, p1);
    if (DROP) {
      const uint32_t b0 = drop_pair_bits(drop_rk, static_cast<uint32_t>(kv0 + 2 * i) >> 1);
      const uint32_t b1 = drop_pair_bits(drop_rk, static_cast<uint32_t>(kv0 + 2 * i + 2) >> 1);
      p0.x = drop_keep_lo(b0, dcfg.thr16) ? p0.x * dcfg.inv_keep : 0.f;
      p0.y = drop_keep_hi(b0, dcfg.thr16) ? p0.y * dcfg.inv_keep : 0.f;
      p1.x = drop_keep_lo(b1, dcfg.thr16) ? p1.x * dcfg.inv_keep : 0.f;
      p1.y = drop_keep_hi(b1, dcfg.thr16) ? p1.y * dcfg.inv_keep : 0.f;
    }
    pk[i] = ptx::pack_bf16x2(p0.x, p0.y);
    pk[i + 1] = ptx::pack_bf16x2(p1.x, p1.y);
  }
  tmax = fmaxf(m0, m1);
  rowsum += (acc0.x + acc0.y) + (acc1.x + acc1.y);
}

template <bool DROP>
__device__ __forceinline__ void fwd_tile(uint32_t tm_s, int lane, int cls0, int cls1, float neg_m, float& tmax, float& rowsum,
                                         uint32_t* pk, const DropCfg& dcfg, uint32_t drop_rk, int kv_tile0) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int cls = c == 0 ? cls0 : cls1;
    const int kv0 = kv_tile0 + c * 32;
    if (cls == kFull) fwd_chunk<kFull, DROP>(tm_s + c * 32, lane, neg_m, tmax, rowsum, pk + c * 16, dcfg, drop_rk, kv0);
    else if (cls == kDiag) fwd_chunk<kDiag, DROP>(tm_s + c * 32, lane, neg_m, tmax, rowsum, pk + c * 16, dcfg, drop_rk, kv0);
    else fwd_chunk<kMasked, DROP>(tm_s + c * 32, lane, neg_m, tmax, rowsum, pk + c * 16, dcfg, drop_rk, kv0);
  }
}

template <bool DROP>
__global__ void __launch_bounds__(kThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T, int H, int C, long long* trace,
                const DropCfg dcfg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [3]
  uint64_t* kv_empty = bars + 4;   // [3]
  uint64_t* s_full = bars + 7;     // [2]
  uint64_t* p_full = bars + 9;     // [2]
  uint64_t* p_empty = bars + 11;   // [2]
  uint64_t* o_full = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = gridDim.x - 1 - blockIdx.x;
  const int b = blockIdx.y / H, h = blockIdx.y % H;
  const int row0 = b * T + qt * 128;
  const int kv_end = min(T, qt * 128 + 128);
  const int num_kv = (kv_end + 63) / 64;
  const bool tr = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64;
#define ATTN_STAMP(k) do { if (tr) trace[j * 8 + (k)] = clock64(); } while (0)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmKV);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 3; ++s) {
      ptx::mbar_init(&kv_full[s], 1);
      ptx::mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&p_full[s], 128);
      ptx::mbar_init(&p_empty[s], 1);
    }
    ptx::mbar_init(o_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_O = tmem_base + 128;  // S buffers at columns [0,64) and [64,128)

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(q_full, 16384);
      ptx::tma_load_2d(smem + FwdSmem::Q, &tmQ, q_full, h * HS, row0);
      for (int j = 0; j < num_kv; ++j) {
        const int st = j % 3;
        const uint32_t ph = (j / 3) & 1;
        ptx::mbar_wait(&kv_empty[st], ph ^ 1, 10);
        ptx::mbar_expect_tx(&kv_full[st], 16384);
        uint8_t* dst = smem + FwdSmem::KV + st * 16384;
        ptx::tma_load_2d(dst, &tmKV, &kv_full[st], C + h * HS, b * T + j * 64);
        ptx::tma_load_2d(dst + 8192, &tmKV, &kv_full[st], 2 * C + h * HS, b * T + j * 64);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16(128, 64, 0, 1);
      const uint32_t sQ = ptx::smem_u32(smem + FwdSmem::Q);
      const uint32_t sKV = ptx::smem_u32(smem + FwdSmem::KV);
      ptx::mbar_wait(q_full, 0, 12);
      ptx::mbar_wait(&kv_full[0], 0, 13);
      ptx::tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) ptx::umma_ss(tmem_base, desc_k(sQ, k), desc_k(sKV, k), idesc_s, k > 0);
      ptx::umma_commit(&s_full[0]);
      for (int j = 0; j < num_kv; ++j) {
        if (j + 1 < num_kv) {  // S_{j+1} runs on the tensor core while the softmax threads work on S_j
          const int st = (j + 1) % 3;
          ptx::mbar_wait(&kv_full[st], ((j + 1) / 3) & 1, 14);
          ptx::tc_fence_after();
          const uint32_t sK = sKV + st * 16384;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_ss(tmem_base + ((j + 1) & 1) * 64, desc_k(sQ, k), desc_k(sK, k), idesc_s, k > 0);
          ptx::umma_commit(&s_full[(j + 1) & 1]);
        }
        ptx::mbar_wait(&p_full[j & 1], (j >> 1) & 1, 15);
        ptx::tc_fence_after();
        const uint32_t sV = sKV + (j % 3) * 16384 + 8192;
        const uint32_t tP = tmem_base + (j & 1) * 64;  // bf16 P (32 packed columns) written over the consumed S tile
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ts(tm_O, tP + 8 * k, desc_mn(sV, k), idesc_o, (j > 0 || k > 0));
        ptx::umma_commit(&kv_empty[j % 3]);
        ptx::umma_commit(&p_empty[j & 1]);
      }
      ptx::umma_commit(o_full);
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int r0 = qt * 128 + quarter * 32;  // first query row of this warp (relative to the sequence)
    float m_ref = 0.f, l = 0.f;
    const uint32_t drop_rk = drop_row_key(dcfg.key, static_cast<uint32_t>(blockIdx.y * T + qt * 128 + r));
    for (int j = 0; j < num_kv; ++j) {
      const int bsel = j & 1;
      const uint32_t tm_s = tmem_base + lane_off + bsel * 64;
      // chunk classes of this warp for key columns [64j, 64j+32) and [64j+32, 64j+64)
      const int c0 = j * 64, c1 = j * 64 + 32;
      const int cls0 = (c0 + 31 <= r0) ? kFull : ((c0 > r0 + 31) ? kMasked : kDiag);
      const int cls1 = (c1 + 31 <= r0) ? kFull : ((c1 > r0 + 31) ? kMasked : kDiag);
      ATTN_STAMP(0);
      ptx::mbar_wait(&s_full[bsel], (j >> 1) & 1, 17);
      ptx::tc_fence_after();
      ATTN_STAMP(1);
      uint32_t pk[32];
      float tmax = -1e30f, rowsum = 0.f;
      if (j == 0) {
        // first tile: the reference exponent is its true row max (cheap max-only pass, then the exp pass)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int cls = c == 0 ? cls0 : cls1;
          if (cls == kFull) tmax = fmaxf(tmax, fwd_chunk_max<kFull>(tm_s + c * 32, lane));
          else if (cls == kDiag) tmax = fmaxf(tmax, fwd_chunk_max<kDiag>(tm_s + c * 32, lane));
        }
        m_ref = tmax * kSl2;
        fwd_tile<DROP>(tm_s, lane, cls0, cls1, -m_ref, tmax, rowsum, pk, dcfg, drop_rk, j * 64);
      } else {
        fwd_tile<DROP>(tm_s, lane, cls0, cls1, -m_ref, tmax, rowsum, pk, dcfg, drop_rk, j * 64);
        const bool need = tmax * kSl2 - m_ref > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // rare: raise the reference, rescale the O accumulator in TMEM, recompute this tile's P
          const float m_new = need ? tmax * kSl2 : m_ref;
          const float alpha = ex2(m_ref - m_new);
          ptx::mbar_wait(&p_empty[(j - 1) & 1], ((j - 1) >> 1) & 1, 19);  // every earlier P V product has landed in O
          ptx::tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            ptx::tmem_ld32(tm_O + lane_off + c * 32, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            ptx::tmem_st32(tm_O + lane_off + c * 32, o);
          }
          ptx::tmem_st_wait();
          l *= alpha;
          m_ref = m_new;
          tmax = -1e30f;
          rowsum = 0.f;
          fwd_tile<DROP>(tm_s, lane, cls0, cls1, -m_ref, tmax, rowsum, pk, dcfg, drop_rk, j * 64);
        }
      }
      l += rowsum;
      ATTN_STAMP(2);
      ATTN_STAMP(3);
      // P (64 key columns = 32 packed words) overwrites the first half of this tile's S buffer; the tensor pipe runs in
      // issue order, so S_{j+2} cannot overwrite it before P V of this tile has consumed it
      ptx::tmem_st16(tm_s, pk);
      ptx::tmem_st16(tm_s + 16, pk + 16);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_full[bsel]);
      ATTN_STAMP(4);
    }
#undef ATTN_STAMP
    ptx::mbar_wait(o_full, 0, 20);
    ptx::tc_fence_after();
    const int t = qt * 128 + r;
    const float inv = 1.0f / l;
    __nv_bfloat16* o = out + static_cast<long long>(b * T + t) * C + h * HS;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      ptx::tmem_ld32(tm_O + lane_off + c * 32, v);
      ptx::tmem_ld_wait();
      if (t < T) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 w;
          w.x = ptx::pack_bf16x2(__uint_as_float(v[8 * q + 0]) * inv, __uint_as_float(v[8 * q + 1]) * inv);
          w.y = ptx::pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv, __uint_as_float(v[8 * q + 3]) * inv);
          w.z = ptx::pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv, __uint_as_float(v[8 * q + 5]) * inv);
          w.w = ptx::pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv, __uint_as_float(v[8 * q + 7]) * inv);
          reinterpret_cast<uint4*>(o)[c * 4 + q] = w;
        }
      }
    }
    if (t < T) lse[(static_cast<long long>(b) * H + h) * T + t] = (m_ref + log2f(l)) * kLn2;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 256);
}

// ======================================================================================================
// backward: delta = rowsum(dO * O) per (token, head)
// ======================================================================================================
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, float* __restrict__ delta,
                  int B, int T, int H, int C) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(B) * T) return;
  const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
  const uint32_t* orow = reinterpret_cast<const uint32_t*>(o + row * C);
  const uint32_t* drow = reinterpret_cast<const uint32_t*>(dout + row * C);
  for (int h = 0; h < H; ++h) {
    const uint32_t a = __ldg(orow + h * 32 + lane), d = __ldg(drow + h * 32 + lane);
    float s = ptx::bf16lo(a) * ptx::bf16lo(d) + ptx::bf16hi(a) * ptx::bf16hi(d);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) delta[(static_cast<long long>(b) * H + h) * T + t] = s;
  }
}

// ======================================================================================================
// backward: dQ and dK/dV
// ======================================================================================================
// Both backward kernels run ONE CTA per SM with two compute groups of 128 threads.  The score tiles S and dP
// (128 x 64 fp32 each) are double-buffered in TMEM and buffer g belongs to group g: while group 0 turns S/dP of step j
// into dS, the tensor core already produces S/dP of step j+1 for group 1, and the accumulating MMAs (dQ, or dV and dK)
// of finished steps interleave in between.  Without this, every step serialises "MMA -> exp/dS -> MMA" (measured:
// ~1400 of ~2700 cycles per step spent waiting for the score MMAs).
constexpr int kSBuf = 3;          // S/dP score buffers in TMEM: two are being consumed by the two groups, one is being produced
constexpr int kRing = 6;          // K/V (dQ kernel) or Q/dO (dK/dV kernel) tiles in flight: TMA latency >> one step
constexpr int kBwdThreads = 352;  // warp 0 TMA, warp 1 score MMAs, warps 2..5 group 0, warps 6..9 group 1, warp 10 accumulating MMAs

struct DqSmem {
  static constexpr int Q = 0;         // 128 x 64
  static constexpr int DO = 16384;    // 128 x 64
  static constexpr int KV = 32768;    // kRing stages x (K 64x64 | V 64x64)
  static constexpr int BAR = KV + kRing * 16384;  // dS never touches shared memory: it is written back into TMEM
  static constexpr int TOTAL = BAR + 256 + 1024;
};

template <bool DROP>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                   const __grid_constant__ CUtensorMap tmDO128, const float* __restrict__ lse,
                   const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int T, int H, int C, long long* trace,
                   const DropCfg dcfg) {
  const bool tr = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64;
#define DQ_STAMP(k) do { if (tr) trace[j * 8 + (k)] = clock64(); } while (0)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DqSmem::BAR);
  uint64_t* qdo_full = bars + 0;
  uint64_t* kv_full = bars + 1;                 // [kRing]
  uint64_t* kv_empty = kv_full + kRing;         // [kRing]
  uint64_t* s_full = kv_empty + kRing;          // [kSBuf]
  uint64_t* s_free = s_full + kSBuf;            // [kSBuf] committed after the dQ MMA that read dS out of this buffer
  uint64_t* ds_full = s_free + kSBuf;           // [2]
  uint64_t* all_done = ds_full + 2;  // dedicated: a parity wait is only meaningful to a thread that followed every phase
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(all_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = gridDim.x - 1 - blockIdx.x;
  const int b = blockIdx.y / H, h = blockIdx.y % H;
  const int row0 = b * T + qt * 128;
  const int kv_end = min(T, qt * 128 + 128);
  const int num_kv = (kv_end + 63) / 64;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQKV128);
    ptx::prefetch_tmap(&tmQKV64);
    ptx::prefetch_tmap(&tmDO128);
    ptx::mbar_init(qdo_full, 1);
    for (int s = 0; s < kRing; ++s) {
      ptx::mbar_init(&kv_full[s], 1);
      ptx::mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < kSBuf; ++s) {
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&s_free[s], 1);
    }
    for (int s = 0; s < 2; ++s) ptx::mbar_init(&ds_full[s], 128);
    ptx::mbar_init(all_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_dQ = tmem_base + 128 * kSBuf;  // score buffer b: S at 128 b, dP at 128 b + 64

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(qdo_full, 32768);
      ptx::tma_load_2d(smem + DqSmem::Q, &tmQKV128, qdo_full, h * HS, row0);
      ptx::tma_load_2d(smem + DqSmem::DO, &tmDO128, qdo_full, h * HS, row0);
      for (int j = 0; j < num_kv; ++j) {
        const int st = j % kRing;
        const uint32_t ph = (j / kRing) & 1;
        ptx::mbar_wait(&kv_empty[st], ph ^ 1, 20);
        ptx::mbar_expect_tx(&kv_full[st], 16384);
        uint8_t* dst = smem + DqSmem::KV + st * 16384;
        ptx::tma_load_2d(dst, &tmQKV64, &kv_full[st], C + h * HS, b * T + j * 64);
        ptx::tma_load_2d(dst + 8192, &tmQKV64, &kv_full[st], 2 * C + h * HS, b * T + j * 64);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- score MMAs: S_j = Q K_j^T, dP_j = dO V_j^T into TMEM buffer j & 1 (one issuing thread per MMA family:
    // a single thread issuing all twelve MMAs of a step plus its barrier traffic was the bottleneck)
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, 64, 0, 0);
      const uint64_t dQ0 = desc_k(ptx::smem_u32(smem + DqSmem::Q), 0), dDO0 = desc_k(ptx::smem_u32(smem + DqSmem::DO), 0);
      const uint64_t dKV0 = desc_k(ptx::smem_u32(smem + DqSmem::KV), 0);
      ptx::mbar_wait(qdo_full, 0, 21);
      for (int j = 0; j < num_kv; ++j) {
        ptx::mbar_wait(&kv_full[j % kRing], (j / kRing) & 1, 22);
        if (j >= kSBuf) ptx::mbar_wait(&s_free[j % kSBuf], ((j / kSBuf) - 1) & 1, 23);  // the dQ MMA of step j-3 has read its dS
        ptx::tc_fence_after();
        const uint64_t dK = dKV0 + static_cast<uint64_t>((j % kRing) * (16384 >> 4)), dV = dK + (8192 >> 4);
        const uint32_t tS = tmem_base + (j % kSBuf) * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ss(tS, dQ0 + 2 * k, dK + 2 * k, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ss(tS + 64, dDO0 + 2 * k, dV + 2 * k, idesc_s, k > 0);
        ptx::umma_commit(&s_full[j % kSBuf]);
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // ---- accumulating MMAs: dQ += dS_j K_j
    if (lane == 0) {
      constexpr uint32_t idesc_dq = ptx::umma_idesc_bf16(128, 64, 0, 1);
      const uint64_t dKmn0 = desc_mn(ptx::smem_u32(smem + DqSmem::KV), 0);
      for (int j = 0; j < num_kv; ++j) {
        ptx::mbar_wait(&ds_full[j & 1], (j >> 1) & 1, 23);
        ptx::tc_fence_after();
        // A = dS (bf16 pairs, 32 columns) sits in the first columns of this step's score buffer
        const uint32_t tA = tmem_base + (j % kSBuf) * 128;
        const uint64_t dK = dKmn0 + static_cast<uint64_t>((j % kRing) * (16384 >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ts(tm_dQ, tA + 8 * k, dK + (2048 >> 4) * k, idesc_dq, (j > 0 || k > 0));
        ptx::umma_commit(&kv_empty[j % kRing]);  // S_j / dP_j (other issuer) completed before ds_full(j) could complete
        ptx::umma_commit(&s_free[j % kSBuf]);    // the score buffer (now holding dS) may be overwritten
      }
      ptx::umma_commit(all_done);
    }
    __syncwarp();
  } else {
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int t = qt * 128 + r;
    const bool valid = t < T;
    const long long stat_idx = (static_cast<long long>(b) * H + h) * T + t;
    const float neg_lse2 = valid ? -__ldg(lse + stat_idx) * kLog2e : 0.f;
    const float neg_delta8 = valid ? -__ldg(delta + stat_idx) * kScale : 0.f;
    const int r0 = qt * 128 + quarter * 32;
    const uint32_t drop_rk = drop_row_key(dcfg.key, static_cast<uint32_t>(blockIdx.y * T + t));
    for (int j = g; j < num_kv; j += 2) {
      const int use = j >> 1;
      DQ_STAMP(0);
      ptx::mbar_wait(&s_full[j % kSBuf], (j / kSBuf) & 1, 24);
      ptx::tc_fence_after();
      DQ_STAMP(1);
      const uint32_t tbuf = tmem_base + lane_off + (j % kSBuf) * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int c0 = j * 64 + c * 32;
        const uint32_t ta_s = tbuf + c * 32, ta_dp = ta_s + 64;
        uint32_t pk[16];
        if (c0 + 31 <= r0) dq_chunk<kFull, DROP>(ta_s, ta_dp, lane, neg_lse2, neg_delta8, pk, dcfg, drop_rk, c0);
        else if (c0 > r0 + 31) dq_chunk<kMasked, DROP>(ta_s, ta_dp, lane, neg_lse2, neg_delta8, pk, dcfg, drop_rk, c0);
        else dq_chunk<kDiag, DROP>(ta_s, ta_dp, lane, neg_lse2, neg_delta8, pk, dcfg, drop_rk, c0);
        // dS chunk c (32 key columns = 16 packed words) overwrites score columns that this thread has already consumed
        ptx::tmem_st16(tbuf + c * 16, pk);
      }
      DQ_STAMP(2);
      DQ_STAMP(3);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ds_full[g]);
      DQ_STAMP(4);
    }
#undef DQ_STAMP
    ptx::mbar_wait(all_done, 0, 26);  // committed after the last MMA
    ptx::tc_fence_after();
    // group g writes columns [32 g, 32 g + 32) of dQ
    __nv_bfloat16* o = dqkv + static_cast<long long>(b * T + t) * (3 * C) + h * HS + g * 32;
    uint32_t v[32];
    ptx::tmem_ld32(tm_dQ + lane_off + g * 32, v);
    ptx::tmem_ld_wait();
    if (valid) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 w;
        w.x = ptx::pack_bf16x2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
        w.y = ptx::pack_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
        w.z = ptx::pack_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
        w.w = ptx::pack_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
        reinterpret_cast<uint4*>(o)[q] = w;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

struct DkvSmem {
  static constexpr int K = 0;          // 128 x 64
  static constexpr int V = 16384;      // 128 x 64
  static constexpr int QDO = 32768;    // kRing stages x (Q 64x64 | dO 64x64)
  // P^T and dS^T never touch shared memory: they are written back into TMEM and consumed as the A operand from there
  static constexpr int STAT = QDO + kRing * 16384;  // 2 groups x 2 buffers x (-lse2[64] | -delta8[64] | dropout row key[64])
  static constexpr int BAR = STAT + 3072;
  static constexpr int TOTAL = BAR + 256 + 1024;
};

template <bool DROP>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                    const __grid_constant__ CUtensorMap tmDO64, const float* __restrict__ lse,
                    const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int T, int H, int C, long long* trace,
                    const DropCfg dcfg) {
  const bool tr = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64;
#define DKV_STAMP(k) do { if (tr) trace[128 + n * 8 + (k)] = clock64(); } while (0)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DkvSmem::BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;                // [kRing]
  uint64_t* qdo_empty = qdo_full + kRing;       // [kRing]
  uint64_t* s_full = qdo_empty + kRing;         // [kSBuf]
  uint64_t* s_free = s_full + kSBuf;            // [kSBuf] committed after the dV / dK MMAs that read P^T / dS^T out of the buffer
  uint64_t* pds_full = s_free + kSBuf;          // [2]
  uint64_t* all_done = pds_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(all_done + 1);
  float* stat = reinterpret_cast<float*>(smem + DkvSmem::STAT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x;  // key tile; tile 0 is the heaviest and is scheduled first
  const int b = blockIdx.y / H, h = blockIdx.y % H;
  const int i0 = kt * 2;                // first 64-row query tile that can see this key tile
  const int nq = (T + 63) / 64 - i0;    // >= 1

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQKV128);
    ptx::prefetch_tmap(&tmQKV64);
    ptx::prefetch_tmap(&tmDO64);
    ptx::mbar_init(kv_full, 1);
    for (int s = 0; s < kRing; ++s) {
      ptx::mbar_init(&qdo_full[s], 1);
      ptx::mbar_init(&qdo_empty[s], 1);
    }
    for (int s = 0; s < kSBuf; ++s) {
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&s_free[s], 1);
    }
    for (int s = 0; s < 2; ++s) ptx::mbar_init(&pds_full[s], 128);
    ptx::mbar_init(all_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_dV = tmem_base + 128 * kSBuf, tm_dK = tm_dV + 64;  // score buffer b: S^T at 128 b, dP^T at 128 b + 64

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(kv_full, 32768);
      ptx::tma_load_2d(smem + DkvSmem::K, &tmQKV128, kv_full, C + h * HS, b * T + kt * 128);
      ptx::tma_load_2d(smem + DkvSmem::V, &tmQKV128, kv_full, 2 * C + h * HS, b * T + kt * 128);
      for (int n = 0; n < nq; ++n) {
        const int st = n % kRing;
        const uint32_t ph = (n / kRing) & 1;
        ptx::mbar_wait(&qdo_empty[st], ph ^ 1, 30);
        ptx::mbar_expect_tx(&qdo_full[st], 16384);
        uint8_t* dst = smem + DkvSmem::QDO + st * 16384;
        ptx::tma_load_2d(dst, &tmQKV64, &qdo_full[st], h * HS, b * T + (i0 + n) * 64);
        ptx::tma_load_2d(dst + 8192, &tmDO64, &qdo_full[st], h * HS, b * T + (i0 + n) * 64);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- score MMAs: S^T_n = K Q_n^T, dP^T_n = V dO_n^T into TMEM buffer n & 1
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, 64, 0, 0);
      const uint64_t dK0 = desc_k(ptx::smem_u32(smem + DkvSmem::K), 0), dV0 = desc_k(ptx::smem_u32(smem + DkvSmem::V), 0);
      const uint64_t dQDO0 = desc_k(ptx::smem_u32(smem + DkvSmem::QDO), 0);
      ptx::mbar_wait(kv_full, 0, 31);
      for (int n = 0; n < nq; ++n) {
        ptx::mbar_wait(&qdo_full[n % kRing], (n / kRing) & 1, 32);
        if (n >= kSBuf) ptx::mbar_wait(&s_free[n % kSBuf], ((n / kSBuf) - 1) & 1, 33);  // dV / dK of step n-3 have read it
        ptx::tc_fence_after();
        const uint64_t dQ = dQDO0 + static_cast<uint64_t>((n % kRing) * (16384 >> 4)), dDO = dQ + (8192 >> 4);
        const uint32_t tS = tmem_base + (n % kSBuf) * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ss(tS, dK0 + 2 * k, dQ + 2 * k, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ss(tS + 64, dV0 + 2 * k, dDO + 2 * k, idesc_s, k > 0);
        ptx::umma_commit(&s_full[n % kSBuf]);
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // ---- accumulating MMAs: dV += P^T_n dO_n, dK += dS^T_n Q_n
    if (lane == 0) {
      constexpr uint32_t idesc_g = ptx::umma_idesc_bf16(128, 64, 0, 1);
      const uint64_t dQmn0 = desc_mn(ptx::smem_u32(smem + DkvSmem::QDO), 0);
      for (int n = 0; n < nq; ++n) {
        ptx::mbar_wait(&pds_full[n & 1], (n >> 1) & 1, 33);
        ptx::tc_fence_after();
        const uint32_t tP = tmem_base + (n % kSBuf) * 128, tDS = tP + 64;  // bf16 pairs written over the consumed scores
        const uint64_t dQ = dQmn0 + static_cast<uint64_t>((n % kRing) * (16384 >> 4)), dDO = dQ + (8192 >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ts(tm_dV, tP + 8 * k, dDO + (2048 >> 4) * k, idesc_g, (n > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ts(tm_dK, tDS + 8 * k, dQ + (2048 >> 4) * k, idesc_g, (n > 0 || k > 0));
        ptx::umma_commit(&qdo_empty[n % kRing]);
        ptx::umma_commit(&s_free[n % kSBuf]);
      }
      ptx::umma_commit(all_done);
    }
    __syncwarp();
  } else {
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;      // key row inside the tile
    const int tid = (warp - 2 - 4 * g) * 32 + lane;  // 0..127 within the group
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int kv_t = kt * 128 + r;
    const bool valid = kv_t < T;
    const long long stat_base = (static_cast<long long>(b) * H + h) * T;
    const int r0 = kt * 128 + quarter * 32;
    for (int n = g; n < nq; n += 2) {
      const int use = n >> 1;
      const int q0 = (i0 + n) * 64;
      float* st_lse = stat + (g * 2 + (use & 1)) * 192;
      {
        const int qi = q0 + (tid & 63);
        float v = 0.f;
        if (qi < T) v = (tid < 64) ? -__ldg(lse + stat_base + qi) * kLog2e : -__ldg(delta + stat_base + qi) * kScale;
        st_lse[tid] = v;
        if (DROP && tid < 64)
          st_lse[128 + tid] = __uint_as_float(drop_row_key(dcfg.key, static_cast<uint32_t>(blockIdx.y * T + qi)));
      }
      DKV_STAMP(0);
      ptx::bar_sync(1 + g, 128);
      ptx::mbar_wait(&s_full[n % kSBuf], (n / kSBuf) & 1, 34);
      ptx::tc_fence_after();
      DKV_STAMP(1);
      const uint32_t tbuf = tmem_base + lane_off + (n % kSBuf) * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int c0 = q0 + c * 32;
        const uint32_t ta_s = tbuf + c * 32, ta_dp = ta_s + 64;
        const uint32_t l2 = ptx::smem_u32(st_lse) + c * 128, d8 = l2 + 256, rkeys = l2 + 512;
        uint32_t pk_p[16], pk_ds[16];
        if (c0 > r0 && c0 + 31 < T) dkv_chunk<kFull, DROP>(ta_s, ta_dp, l2, d8, c0, kv_t, T, pk_p, pk_ds, dcfg, rkeys);
        else if (c0 + 31 < r0 || c0 >= T) dkv_chunk<kMasked, DROP>(ta_s, ta_dp, l2, d8, c0, kv_t, T, pk_p, pk_ds, dcfg, rkeys);
        else dkv_chunk<kDiag, DROP>(ta_s, ta_dp, l2, d8, c0, kv_t, T, pk_p, pk_ds, dcfg, rkeys);
        // P^T / dS^T chunk c (32 query columns = 16 packed words each) overwrite score columns already consumed
        ptx::tmem_st16(tbuf + c * 16, pk_p);
        ptx::tmem_st16(tbuf + 64 + c * 16, pk_ds);
      }
      DKV_STAMP(2);
      DKV_STAMP(3);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&pds_full[g]);
      DKV_STAMP(4);
    }
#undef DKV_STAMP
    ptx::mbar_wait(all_done, 0, 36);  // committed after the last MMA
    ptx::tc_fence_after();
    // group 0 writes dV, group 1 writes dK
    __nv_bfloat16* o = dqkv + static_cast<long long>(b * T + kv_t) * (3 * C) + (g == 0 ? 2 * C : C) + h * HS;
    const uint32_t tm = g == 0 ? tm_dV : tm_dK;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      ptx::tmem_ld32(tm + lane_off + c * 32, v);
      ptx::tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 w;
          w.x = ptx::pack_bf16x2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
          w.y = ptx::pack_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
          w.z = ptx::pack_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
          w.w = ptx::pack_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
          reinterpret_cast<uint4*>(o)[c * 4 + q] = w;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

template <typename K>
int set_smem(K kern, int bytes) {
  ABCGPT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  // two CTAs per SM need the full shared-memory carve-out
  ABCGPT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  return 0;
}

int check_shape(const char* who, int B, int T, int H) {
  ABCGPT_CHECK_ARG(B > 0 && T > 0 && H > 0, "%s: bad shape B=%d T=%d H=%d", who, B, T, H);
  ABCGPT_CHECK_ARG(static_cast<long long>(B) * H <= 65535, "%s: B*H must be <= 65535 (grid.y)", who);
  return 0;
}

}  // namespace

int attn_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, float drop_p, uint32_t drop_key,
             cudaStream_t stream) {
  int rc = check_shape("attn_fwd", B, T, H);
  if (rc) return rc;
  ABCGPT_CHECK_ARG(qkv && out && lse, "attn_fwd: null pointer");
  ABCGPT_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "attn_fwd: dropout p must be in [0, 1)");
  const int C = H * HS;
  const DropCfg dcfg = make_drop(drop_p, drop_key);
  CUtensorMap tmQ, tmKV;
  if ((rc = encode_tmap_2d(&tmQ, qkv, 2, 3ull * C, static_cast<uint64_t>(B) * T, 3ull * C * 2, 64, 128, true))) return rc;
  if ((rc = encode_tmap_2d(&tmKV, qkv, 2, 3ull * C, static_cast<uint64_t>(B) * T, 3ull * C * 2, 64, 64, true))) return rc;
  static bool done = false;
  if (!done) {
    if ((rc = set_smem(attn_fwd_kernel<false>, FwdSmem::TOTAL))) return rc;
    if ((rc = set_smem(attn_fwd_kernel<true>, FwdSmem::TOTAL))) return rc;
    done = true;
  }
  dim3 grid((T + 127) / 128, B * H);
  if (dcfg.thr16 == 0)
    attn_fwd_kernel<false><<<grid, kThreads, FwdSmem::TOTAL, stream>>>(tmQ, tmKV, reinterpret_cast<__nv_bfloat16*>(out), lse,
                                                                       T, H, C, g_attn_trace, dcfg);
  else
    attn_fwd_kernel<true><<<grid, kThreads, FwdSmem::TOTAL, stream>>>(tmQ, tmKV, reinterpret_cast<__nv_bfloat16*>(out), lse,
                                                                      T, H, C, g_attn_trace, dcfg);
  return launch_status("attn_fwd_kernel");
}

int attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, int B,
             int T, int H, float drop_p, uint32_t drop_key, cudaStream_t stream) {
  int rc = check_shape("attn_bwd", B, T, H);
  if (rc) return rc;
  ABCGPT_CHECK_ARG(qkv && out && dout && lse && delta && dqkv, "attn_bwd: null pointer");
  ABCGPT_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "attn_bwd: dropout p must be in [0, 1)");
  const int C = H * HS;
  const DropCfg dcfg = make_drop(drop_p, drop_key);
  const uint64_t rows = static_cast<uint64_t>(B) * T;
  CUtensorMap tmQKV128, tmQKV64, tmDO128, tmDO64;
  if ((rc = encode_tmap_2d(&tmQKV128, qkv, 2, 3ull * C, rows, 3ull * C * 2, 64, 128, true))) return rc;
  if ((rc = encode_tmap_2d(&tmQKV64, qkv, 2, 3ull * C, rows, 3ull * C * 2, 64, 64, true))) return rc;
  if ((rc = encode_tmap_2d(&tmDO128, dout, 2, static_cast<uint64_t>(C), rows, static_cast<uint64_t>(C) * 2, 64, 128, true))) return rc;
  if ((rc = encode_tmap_2d(&tmDO64, dout, 2, static_cast<uint64_t>(C), rows, static_cast<uint64_t>(C) * 2, 64, 64, true))) return rc;
  static bool done = false;
  if (!done) {
    if ((rc = set_smem(attn_bwd_dq_kernel<false>, DqSmem::TOTAL))) return rc;
    if ((rc = set_smem(attn_bwd_dq_kernel<true>, DqSmem::TOTAL))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel<false>, DkvSmem::TOTAL))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel<true>, DkvSmem::TOTAL))) return rc;
    done = true;
  }
  {
    const long long nrows = static_cast<long long>(B) * T;
    attn_delta_kernel<<<static_cast<int>((nrows + 7) / 8), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), delta, B, T, H, C);
    if ((rc = launch_status("attn_delta_kernel"))) return rc;
  }
  dim3 grid((T + 127) / 128, B * H);
  __nv_bfloat16* dq = reinterpret_cast<__nv_bfloat16*>(dqkv);
  if (dcfg.thr16 == 0) {
    attn_bwd_dkv_kernel<false><<<grid, kBwdThreads, DkvSmem::TOTAL, stream>>>(tmQKV128, tmQKV64, tmDO64, lse, delta, dq, T, H, C,
                                                                              g_attn_trace, dcfg);
    if ((rc = launch_status("attn_bwd_dkv_kernel"))) return rc;
    attn_bwd_dq_kernel<false><<<grid, kBwdThreads, DqSmem::TOTAL, stream>>>(tmQKV128, tmQKV64, tmDO128, lse, delta, dq, T, H, C,
                                                                            g_attn_trace, dcfg);
  } else {
    attn_bwd_dkv_kernel<true><<<grid, kBwdThreads, DkvSmem::TOTAL, stream>>>(tmQKV128, tmQKV64, tmDO64, lse, delta, dq, T, H, C,
                                                                             g_attn_trace, dcfg);
    if ((rc = launch_status("attn_bwd_dkv_kernel"))) return rc;
    attn_bwd_dq_kernel<true><<<grid, kBwdThreads, DqSmem::TOTAL, stream>>>(tmQKV128, tmQKV64, tmDO128, lse, delta, dq, T, H, C,
                                                                           g_attn_trace, dcfg);
  }
  return launch_status("attn_bwd_dq_kernel");
}

}  // namespace abcgpt
