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
#include "kernels.h"
#include "ptx.cuh"

namespace abcgpt {
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
template <int MODE>
__device__ __forceinline__ void dq_chunk(uint32_t taddr_s, uint32_t taddr_dp, int lane, float neg_lse2, float neg_delta8,
                                         uint32_t* pk) {
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
    const float2 u = __ffma2_rn(f2(dp[2 * i], dp[2 * i + 1]), sc, nd);
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
template <int MODE>
__device__ __forceinline__ void dkv_chunk(uint32_t taddr_s, uint32_t taddr_dp, const float* st_lse2, const float* st_delta8,
                                          int q_base, int kv_t, int T, uint32_t* pk_p, uint32_t* pk_ds) {
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
    const float4 l4 = reinterpret_cast<const float4*>(st_lse2)[i4];     // already negated: -lse*log2e
    const float4 d4 = reinterpret_cast<const float4*>(st_delta8)[i4];   // already negated: -delta*scale
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * i4 + h;  // pair index: columns 2i, 2i+1
      const float2 nl = h == 0 ? make_float2(l4.x, l4.y) : make_float2(l4.z, l4.w);
      const float2 nd = h == 0 ? make_float2(d4.x, d4.y) : make_float2(d4.z, d4.w);
      const float2 t = __ffma2_rn(f2(s[2 * i], s[2 * i + 1]), sl, nl);
      float2 p = make_float2(ex2(t.x), ex2(t.y));
      const float2 u = __ffma2_rn(f2(dp[2 * i], dp[2 * i + 1]), sc, nd);
      float2 d = __fmul2_rn(p, u);
      if (MODE == kDiag) {
        const int q0 = q_base + 2 * i;
        const bool k0 = (q0 >= kv_t) && (q0 < T), k1 = (q0 + 1 >= kv_t) && (q0 + 1 < T);
        p.x = k0 ? p.x : 0.f; d.x = k0 ? d.x : 0.f;
        p.y = k1 ? p.y : 0.f; d.y = k1 ? d.y : 0.f;
      }
      pk_p[i] = ptx::pack_bf16x2(p.x, p.y);
      pk_ds[i] = ptx::pack_bf16x2(d.x, d.y);
    }
  }
}

// ======================================================================================================
// forward
// ======================================================================================================
struct FwdSmem {
  static constexpr int Q = 0;                 // 128 x 64 bf16
  static constexpr int K = 16384;             // 128 x 64
  static constexpr int V = 32768;             // 128 x 64
  static constexpr int P = 49152;             // 2 slabs of 128 x 64
  static constexpr int BAR = 81920;
  static constexpr int TOTAL = BAR + 128 + 1024;
};

__global__ void __launch_bounds__(kThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                int T, int H, int C) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint64_t* o_empty = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_q_tiles = gridDim.x;
  const int qt = num_q_tiles - 1 - blockIdx.x;
  const int b = blockIdx.y / H, h = blockIdx.y % H;
  const int row0 = b * T + qt * 128;
  const int num_kv = qt + 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQKV);
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(k_full, 1);
    ptx::mbar_init(k_empty, 1);
    ptx::mbar_init(v_full, 1);
    ptx::mbar_init(v_empty, 1);
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 128);
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_empty, 128);
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
  const uint32_t tm_S = tmem_base;        // 128 columns
  const uint32_t tm_O = tmem_base + 128;  // 64 columns

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(q_full, 16384);
      ptx::tma_load_2d(smem + FwdSmem::Q, &tmQKV, q_full, h * HS, row0);
      for (int j = 0; j < num_kv; ++j) {
        const uint32_t ph = j & 1;
        ptx::mbar_wait(k_empty, ph ^ 1, 10);
        ptx::mbar_expect_tx(k_full, 16384);
        ptx::tma_load_2d(smem + FwdSmem::K, &tmQKV, k_full, C + h * HS, b * T + j * 128);
        ptx::mbar_wait(v_empty, ph ^ 1, 11);
        ptx::mbar_expect_tx(v_full, 16384);
        ptx::tma_load_2d(smem + FwdSmem::V, &tmQKV, v_full, 2 * C + h * HS, b * T + j * 128);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16(128, 64, 0, 1);
      const uint32_t sQ = ptx::smem_u32(smem + FwdSmem::Q), sK = ptx::smem_u32(smem + FwdSmem::K);
      const uint32_t sV = ptx::smem_u32(smem + FwdSmem::V), sP = ptx::smem_u32(smem + FwdSmem::P);
      ptx::mbar_wait(q_full, 0, 12);
      for (int j = 0; j < num_kv; ++j) {
        const uint32_t ph = j & 1;
        ptx::mbar_wait(k_full, ph, 13);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ss(tm_S, desc_k(sQ, k), desc_k(sK, k), idesc_s, k > 0);
        ptx::umma_commit(s_full);
        ptx::umma_commit(k_empty);
        ptx::mbar_wait(p_full, ph, 14);
        ptx::mbar_wait(v_full, ph, 15);
        ptx::mbar_wait(o_empty, ph ^ 1, 16);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          ptx::umma_ss(tm_O, desc_k(sP + (k >> 2) * 16384, k & 3), desc_mn(sV, k), idesc_o, k > 0);
        ptx::umma_commit(o_full);
        ptx::umma_commit(v_empty);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t sP = ptx::smem_u32(smem + FwdSmem::P);
    float O[HS];
#pragma unroll
    for (int i = 0; i < HS; ++i) O[i] = 0.f;
    float m = -1e30f, l = 0.f;
    for (int j = 0; j < num_kv; ++j) {
      const uint32_t ph = j & 1;
      const bool diag = (j == qt);
      ptx::mbar_wait(s_full, ph, 17);
      ptx::tc_fence_after();
      float mx = -1e30f;
      if (!diag) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) mx = fmaxf(mx, fwd_chunk_max<kFull>(tm_S + lane_off + c * 32, lane));
      } else {
#pragma unroll 1
        for (int c = 0; c <= quarter; ++c)
          mx = fmaxf(mx, c < quarter ? fwd_chunk_max<kFull>(tm_S + lane_off + c * 32, lane)
                                     : fwd_chunk_max<kDiag>(tm_S + lane_off + c * 32, lane));
      }
      const float m_new = fmaxf(m, mx * kSl2);
      const float alpha = ex2(m - m_new);
      float rowsum = 0.f;
      if (!diag) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c)
          rowsum += fwd_chunk_exp<kFull>(tm_S + lane_off + c * 32, lane, -m_new, sP + (c >> 1) * 16384, r, c & 1);
      } else {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const uint32_t ta = tm_S + lane_off + c * 32, slab = sP + (c >> 1) * 16384;
          if (c < quarter) rowsum += fwd_chunk_exp<kFull>(ta, lane, -m_new, slab, r, c & 1);
          else if (c == quarter) rowsum += fwd_chunk_exp<kDiag>(ta, lane, -m_new, slab, r, c & 1);
          else rowsum += fwd_chunk_exp<kMasked>(ta, lane, -m_new, slab, r, c & 1);
        }
      }
      l = l * alpha + rowsum;
      m = m_new;
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
      ptx::mbar_wait(o_full, ph, 18);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        ptx::tmem_ld32(tm_O + lane_off + c * 32, v);
        ptx::tmem_ld_wait();
        const float2 al = make_float2(alpha, alpha);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 o = __ffma2_rn(make_float2(O[c * 32 + i], O[c * 32 + i + 1]), al, f2(v[i], v[i + 1]));
          O[c * 32 + i] = o.x;
          O[c * 32 + i + 1] = o.y;
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(o_empty);
    }
    const int t = qt * 128 + r;
    if (t < T) {
      const float inv = 1.0f / l;
      __nv_bfloat16* o = out + static_cast<long long>(b * T + t) * C + h * HS;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 w;
        w.x = ptx::pack_bf16x2(O[8 * q + 0] * inv, O[8 * q + 1] * inv);
        w.y = ptx::pack_bf16x2(O[8 * q + 2] * inv, O[8 * q + 3] * inv);
        w.z = ptx::pack_bf16x2(O[8 * q + 4] * inv, O[8 * q + 5] * inv);
        w.w = ptx::pack_bf16x2(O[8 * q + 6] * inv, O[8 * q + 7] * inv);
        reinterpret_cast<uint4*>(o)[q] = w;
      }
      lse[(static_cast<long long>(b) * H + h) * T + t] = (m + log2f(l)) * kLn2;
    }
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
// backward: dQ
// ======================================================================================================
struct DqSmem {
  static constexpr int Q = 0;        // 128 x 64
  static constexpr int DO = 16384;   // 128 x 64
  static constexpr int KV = 32768;   // 2 stages x (K 64x64 | V 64x64) = 2 x 16 KB
  static constexpr int DS = 65536;   // 128 x 64
  static constexpr int BAR = 81920;
  static constexpr int TOTAL = BAR + 128 + 1024;
};

__global__ void __launch_bounds__(kThreads, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                   const __grid_constant__ CUtensorMap tmDO128, const float* __restrict__ lse,
                   const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int T, int H, int C) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DqSmem::BAR);
  uint64_t* qdo_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* ds_full = bars + 6;
  uint64_t* dq_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

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
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&kv_full[s], 1);
      ptx::mbar_init(&kv_empty[s], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(ds_full, 128);
    ptx::mbar_init(dq_done, 1);
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
  const uint32_t tm_S = tmem_base, tm_dP = tmem_base + 64, tm_dQ = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(qdo_full, 32768);
      ptx::tma_load_2d(smem + DqSmem::Q, &tmQKV128, qdo_full, h * HS, row0);
      ptx::tma_load_2d(smem + DqSmem::DO, &tmDO128, qdo_full, h * HS, row0);
      for (int j = 0; j < num_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        ptx::mbar_wait(&kv_empty[st], ph ^ 1, 20);
        ptx::mbar_expect_tx(&kv_full[st], 16384);
        uint8_t* dst = smem + DqSmem::KV + st * 16384;
        ptx::tma_load_2d(dst, &tmQKV64, &kv_full[st], C + h * HS, b * T + j * 64);
        ptx::tma_load_2d(dst + 8192, &tmQKV64, &kv_full[st], 2 * C + h * HS, b * T + j * 64);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_dq = ptx::umma_idesc_bf16(128, 64, 0, 1);
      const uint32_t sQ = ptx::smem_u32(smem + DqSmem::Q), sDO = ptx::smem_u32(smem + DqSmem::DO);
      const uint32_t sDS = ptx::smem_u32(smem + DqSmem::DS);
      ptx::mbar_wait(qdo_full, 0, 21);
      for (int j = 0; j < num_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const uint32_t sK = ptx::smem_u32(smem + DqSmem::KV + st * 16384), sV = sK + 8192;
        ptx::mbar_wait(&kv_full[st], ph, 22);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ss(tm_S, desc_k(sQ, k), desc_k(sK, k), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ss(tm_dP, desc_k(sDO, k), desc_k(sV, k), idesc_s, k > 0);
        ptx::umma_commit(s_full);
        ptx::mbar_wait(ds_full, j & 1, 23);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ss(tm_dQ, desc_k(sDS, k), desc_mn(sK, k), idesc_dq, (j > 0 || k > 0));
        ptx::umma_commit(&kv_empty[st]);
        ptx::umma_commit(dq_done);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t sDS = ptx::smem_u32(smem + DqSmem::DS);
    const int t = qt * 128 + r;
    const bool valid = t < T;
    const long long stat_idx = (static_cast<long long>(b) * H + h) * T + t;
    const float neg_lse2 = valid ? -__ldg(lse + stat_idx) * kLog2e : 0.f;
    const float neg_delta8 = valid ? -__ldg(delta + stat_idx) * kScale : 0.f;
    for (int j = 0; j < num_kv; ++j) {
      ptx::mbar_wait(s_full, j & 1, 24);
      ptx::tc_fence_after();
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int c0 = j * 64 + c * 32;      // first key column of the chunk
        const int r0 = qt * 128 + quarter * 32;  // first query row of this warp
        const uint32_t ta_s = tm_S + lane_off + c * 32, ta_dp = tm_dP + lane_off + c * 32;
        if (c0 + 31 <= r0) dq_chunk<kFull>(ta_s, ta_dp, lane, neg_lse2, neg_delta8, pk + c * 16);
        else if (c0 > r0 + 31) dq_chunk<kMasked>(ta_s, ta_dp, lane, neg_lse2, neg_delta8, pk + c * 16);
        else dq_chunk<kDiag>(ta_s, ta_dp, lane, neg_lse2, neg_delta8, pk + c * 16);
      }
      if (j > 0) ptx::mbar_wait(dq_done, (j - 1) & 1, 25);  // previous dQ MMA has finished reading dS
      st_slab32(sDS, r, 0, pk);
      st_slab32(sDS, r, 1, pk + 16);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(ds_full);
    }
    ptx::mbar_wait(dq_done, (num_kv - 1) & 1, 26);
    ptx::tc_fence_after();
    __nv_bfloat16* o = dqkv + static_cast<long long>(b * T + t) * (3 * C) + h * HS;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      ptx::tmem_ld32(tm_dQ + lane_off + c * 32, v);
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
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 256);
}

// ======================================================================================================
// backward: dK, dV
// ======================================================================================================
struct DkvSmem {
  static constexpr int K = 0;        // 128 x 64
  static constexpr int V = 16384;    // 128 x 64
  static constexpr int QDO = 32768;  // 2 stages x (Q 64x64 | dO 64x64)
  static constexpr int PT = 65536;   // 128 x 64  P^T
  static constexpr int DST = 81920;  // 128 x 64  dS^T
  static constexpr int STAT = 98304; // 2 stages x (lse2[64] | delta[64]) fp32
  static constexpr int BAR = 99328;
  static constexpr int TOTAL = BAR + 128 + 1024;
};

__global__ void __launch_bounds__(kThreads, 2)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                    const __grid_constant__ CUtensorMap tmDO64, const float* __restrict__ lse,
                    const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int T, int H, int C) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DkvSmem::BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;   // [2]
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* pds_full = bars + 6;
  uint64_t* pds_empty = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
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
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&qdo_full[s], 1);
      ptx::mbar_init(&qdo_empty[s], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(pds_full, 128);
    ptx::mbar_init(pds_empty, 1);
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
  const uint32_t tm_S = tmem_base, tm_dP = tmem_base + 64, tm_dV = tmem_base + 128, tm_dK = tmem_base + 192;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(kv_full, 32768);
      ptx::tma_load_2d(smem + DkvSmem::K, &tmQKV128, kv_full, C + h * HS, b * T + kt * 128);
      ptx::tma_load_2d(smem + DkvSmem::V, &tmQKV128, kv_full, 2 * C + h * HS, b * T + kt * 128);
      for (int n = 0; n < nq; ++n) {
        const int st = n & 1;
        const uint32_t ph = (n >> 1) & 1;
        ptx::mbar_wait(&qdo_empty[st], ph ^ 1, 30);
        ptx::mbar_expect_tx(&qdo_full[st], 16384);
        uint8_t* dst = smem + DkvSmem::QDO + st * 16384;
        ptx::tma_load_2d(dst, &tmQKV64, &qdo_full[st], h * HS, b * T + (i0 + n) * 64);
        ptx::tma_load_2d(dst + 8192, &tmDO64, &qdo_full[st], h * HS, b * T + (i0 + n) * 64);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_g = ptx::umma_idesc_bf16(128, 64, 0, 1);
      const uint32_t sK = ptx::smem_u32(smem + DkvSmem::K), sV = ptx::smem_u32(smem + DkvSmem::V);
      const uint32_t sPT = ptx::smem_u32(smem + DkvSmem::PT), sDST = ptx::smem_u32(smem + DkvSmem::DST);
      ptx::mbar_wait(kv_full, 0, 31);
      for (int n = 0; n < nq; ++n) {
        const int st = n & 1;
        const uint32_t ph = (n >> 1) & 1;
        const uint32_t sQ = ptx::smem_u32(smem + DkvSmem::QDO + st * 16384), sDO = sQ + 8192;
        ptx::mbar_wait(&qdo_full[st], ph, 32);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ss(tm_S, desc_k(sK, k), desc_k(sQ, k), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ss(tm_dP, desc_k(sV, k), desc_k(sDO, k), idesc_s, k > 0);
        ptx::umma_commit(s_full);
        ptx::mbar_wait(pds_full, n & 1, 33);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ss(tm_dV, desc_k(sPT, k), desc_mn(sDO, k), idesc_g, (n > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_ss(tm_dK, desc_k(sDST, k), desc_mn(sQ, k), idesc_g, (n > 0 || k > 0));
        ptx::umma_commit(&qdo_empty[st]);
        ptx::umma_commit(pds_empty);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // key row inside the tile
    const int tid = threadIdx.x - 64;   // 0..127 within the compute group
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t sPT = ptx::smem_u32(smem + DkvSmem::PT), sDST = ptx::smem_u32(smem + DkvSmem::DST);
    const int kv_t = kt * 128 + r;
    const bool valid = kv_t < T;
    const long long stat_base = (static_cast<long long>(b) * H + h) * T;
    for (int n = 0; n < nq; ++n) {
      const int q0 = (i0 + n) * 64;
      float* st_lse = stat + (n & 1) * 128;
      {
        const int qi = q0 + (tid & 63);
        float v = 0.f;
        if (qi < T) v = (tid < 64) ? -__ldg(lse + stat_base + qi) * kLog2e : -__ldg(delta + stat_base + qi) * kScale;
        st_lse[tid] = v;
      }
      ptx::bar_sync(1, 128);
      ptx::mbar_wait(s_full, n & 1, 34);
      ptx::tc_fence_after();
      uint32_t pk_p[32], pk_ds[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int c0 = q0 + c * 32;                 // first query column of the chunk
        const int r0 = kt * 128 + quarter * 32;     // first key row of this warp
        const uint32_t ta_s = tm_S + lane_off + c * 32, ta_dp = tm_dP + lane_off + c * 32;
        const float* l2 = st_lse + c * 32;
        const float* d8 = st_lse + 64 + c * 32;
        if (c0 > r0 && c0 + 31 < T) dkv_chunk<kFull>(ta_s, ta_dp, l2, d8, c0, kv_t, T, pk_p + c * 16, pk_ds + c * 16);
        else if (c0 + 31 < r0 || c0 >= T) dkv_chunk<kMasked>(ta_s, ta_dp, l2, d8, c0, kv_t, T, pk_p + c * 16, pk_ds + c * 16);
        else dkv_chunk<kDiag>(ta_s, ta_dp, l2, d8, c0, kv_t, T, pk_p + c * 16, pk_ds + c * 16);
      }
      if (n > 0) ptx::mbar_wait(pds_empty, (n - 1) & 1, 35);
      st_slab32(sPT, r, 0, pk_p);
      st_slab32(sPT, r, 1, pk_p + 16);
      st_slab32(sDST, r, 0, pk_ds);
      st_slab32(sDST, r, 1, pk_ds + 16);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(pds_full);
    }
    ptx::mbar_wait(pds_empty, (nq - 1) & 1, 36);
    ptx::tc_fence_after();
    __nv_bfloat16* ok = dqkv + static_cast<long long>(b * T + kv_t) * (3 * C) + C + h * HS;
    __nv_bfloat16* ov = ok + C;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* o = which == 0 ? ov : ok;
      const uint32_t tm = which == 0 ? tm_dV : tm_dK;
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
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 256);
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

int attn_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, cudaStream_t stream) {
  int rc = check_shape("attn_fwd", B, T, H);
  if (rc) return rc;
  ABCGPT_CHECK_ARG(qkv && out && lse, "attn_fwd: null pointer");
  const int C = H * HS;
  CUtensorMap tm;
  rc = encode_tmap_2d(&tm, qkv, 2, 3ull * C, static_cast<uint64_t>(B) * T, 3ull * C * 2, 64, 128, true);
  if (rc) return rc;
  static bool done = false;
  if (!done) {
    rc = set_smem(attn_fwd_kernel, FwdSmem::TOTAL);
    if (rc) return rc;
    done = true;
  }
  dim3 grid((T + 127) / 128, B * H);
  attn_fwd_kernel<<<grid, kThreads, FwdSmem::TOTAL, stream>>>(tm, reinterpret_cast<__nv_bfloat16*>(out), lse, T, H, C);
  return launch_status("attn_fwd_kernel");
}

int attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv, int B,
             int T, int H, cudaStream_t stream) {
  int rc = check_shape("attn_bwd", B, T, H);
  if (rc) return rc;
  ABCGPT_CHECK_ARG(qkv && out && dout && lse && delta && dqkv, "attn_bwd: null pointer");
  const int C = H * HS;
  const uint64_t rows = static_cast<uint64_t>(B) * T;
  CUtensorMap tmQKV128, tmQKV64, tmDO128, tmDO64;
  if ((rc = encode_tmap_2d(&tmQKV128, qkv, 2, 3ull * C, rows, 3ull * C * 2, 64, 128, true))) return rc;
  if ((rc = encode_tmap_2d(&tmQKV64, qkv, 2, 3ull * C, rows, 3ull * C * 2, 64, 64, true))) return rc;
  if ((rc = encode_tmap_2d(&tmDO128, dout, 2, static_cast<uint64_t>(C), rows, static_cast<uint64_t>(C) * 2, 64, 128, true))) return rc;
  if ((rc = encode_tmap_2d(&tmDO64, dout, 2, static_cast<uint64_t>(C), rows, static_cast<uint64_t>(C) * 2, 64, 64, true))) return rc;
  static bool done = false;
  if (!done) {
    if ((rc = set_smem(attn_bwd_dq_kernel, DqSmem::TOTAL))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_kernel, DkvSmem::TOTAL))) return rc;
    done = true;
  }
  {
    const long long nrows = static_cast<long long>(B) * T;
    attn_delta_kernel<<<static_cast<int>((nrows + 7) / 8), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), delta, B, T, H, C);
    if ((rc = launch_status("attn_delta_kernel"))) return rc;
  }
  dim3 grid((T + 127) / 128, B * H);
  attn_bwd_dkv_kernel<<<grid, kThreads, DkvSmem::TOTAL, stream>>>(tmQKV128, tmQKV64, tmDO64, lse, delta,
                                                                 reinterpret_cast<__nv_bfloat16*>(dqkv), T, H, C);
  if ((rc = launch_status("attn_bwd_dkv_kernel"))) return rc;
  attn_bwd_dq_kernel<<<grid, kThreads, DqSmem::TOTAL, stream>>>(tmQKV128, tmQKV64, tmDO128, lse, delta,
                                                               reinterpret_cast<__nv_bfloat16*>(dqkv), T, H, C);
  return launch_status("attn_bwd_dq_kernel");
}

}  // namespace abcgpt
