// bf16 GEMM for sm_100a: C[M,N] = sum_k A[m,k] * B[n,k], fp32 accumulation in TMEM.
//
// Covers every nn.Linear of the reference's hot path (nanoGPT/model.py:56,75,88,90,186) in its three
// autograd roles, selected by operand majorness instead of materialised transposes:
//   forward  Y  = X  * W^T        A = X  (K-major)   B = W  (K-major)
//   dgrad    dX = dY * W          A = dY (K-major)   B = W  (MN-major: W is [k_contract, n_out] row-major)
//   wgrad    dW = dY^T * X        A = dY (MN-major)  B = X  (MN-major), fp32 reduction into the grad arena
//
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer   cp.async.bulk.tensor 2D, 128B swizzle, STAGES-deep mbarrier ring
//   warp 1      MMA issuer     one thread issues tcgen05.mma.cta_group::1.kind::f16 (128 x BN x 16)
//   warps 2..9  epilogue       tcgen05.ld TMEM -> registers -> fused epilogue -> global
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace abcgpt {

unsigned long long* g_gemm_stats = nullptr;  // debug only: set through abcgpt_debug_gemm_stats

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kNumThreads = 320;  // 1 producer warp + 1 MMA warp + 8 epilogue warps
constexpr int kNumEpiWarps = 8;

struct GemmParams {
  int M, N, K;
  int num_m_blk, num_n_blk, num_k_blk;
  int splits, kb_per_split;
  void* c;
  long long ldc;
  void* c2;
  long long ldc2;
  const void* aux;
  long long ldaux;
  const float* bias;
  unsigned long long* stats;  // optional debug counters (cycles): [0] producer empty-wait, [1] mma full-wait,
                              // [2] mma tmem-empty wait, [3] epilogue tmem-full wait, [4] epilogue busy, [5] cta total
};

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 192 ? 5 : (BN >= 128 ? 6 : 8));
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : ((2 * BN <= 256) ? 256 : 512);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;
};

// ---- GELU (exact-erf form of nn.GELU(), model.py:83) ---------------------------------------------------
// erf via Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7, i.e. fp32-erff class accuracy) so that erf and the
// Gaussian pdf share ONE exponential: erf(x/sqrt2) = 1 - poly(t) * exp(-x^2/2).
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf) {
  const float ax = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = __expf(-0.5f * x * x);
  const float erf_abs = 1.0f - poly * e;
  const float erfv = copysignf(erf_abs, x);
  cdf = 0.5f * (1.0f + erfv);
  pdf = e * 0.39894228040143268f;
}
__device__ __forceinline__ float gelu_fwd(float x) {
  float cdf, pdf;
  gelu_parts(x, cdf, pdf);
  return x * cdf;
}
__device__ __forceinline__ float gelu_bwd(float x) {
  float cdf, pdf;
  gelu_parts(x, cdf, pdf);
  return fmaf(x, pdf, cdf);
}

// ---- epilogues: one thread owns 32 consecutive columns of one output row -------------------------------
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int row, int col0, uint32_t (&r)[32]) {
  if (row >= p.M || col0 >= p.N) return;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (col0 + i < p.N) v[i] += __ldg(p.bias + col0 + i);
  }
  const int ncols = min(32, p.N - col0);  // multiple of 8 (host-checked)

  if constexpr (EPI == ABCGPT_EPI_BF16 || EPI == ABCGPT_EPI_GELU || EPI == ABCGPT_EPI_DGELU) {
    __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.c) + static_cast<long long>(row) * p.ldc + col0;
    if constexpr (EPI == ABCGPT_EPI_DGELU) {
      const uint4* hp =
          reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.aux) + static_cast<long long>(row) * p.ldaux + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (8 * j < ncols) {
          const uint4 h = __ldg(hp + j);
          const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            v[8 * j + 2 * q] = ptx::bf16_round(v[8 * j + 2 * q]) * gelu_bwd(ptx::bf16lo(hw[q]));
            v[8 * j + 2 * q + 1] = ptx::bf16_round(v[8 * j + 2 * q + 1]) * gelu_bwd(ptx::bf16hi(hw[q]));
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (8 * j < ncols) {
        uint4 o;
        o.x = ptx::pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
        o.y = ptx::pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        o.z = ptx::pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
        o.w = ptx::pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        reinterpret_cast<uint4*>(c)[j] = o;
      }
    }
    if constexpr (EPI == ABCGPT_EPI_GELU) {
      __nv_bfloat16* g = reinterpret_cast<__nv_bfloat16*>(p.c2) + static_cast<long long>(row) * p.ldc2 + col0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (8 * j < ncols) {
          float a[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) a[q] = gelu_fwd(ptx::bf16_round(v[8 * j + q]));
          uint4 o;
          o.x = ptx::pack_bf16x2(a[0], a[1]);
          o.y = ptx::pack_bf16x2(a[2], a[3]);
          o.z = ptx::pack_bf16x2(a[4], a[5]);
          o.w = ptx::pack_bf16x2(a[6], a[7]);
          reinterpret_cast<uint4*>(g)[j] = o;
        }
      }
    }
  } else if constexpr (EPI == ABCGPT_EPI_RESID) {
    // x_out(fp32) = x_in(fp32) + bf16(acc): the reference adds the bf16 Linear output into the fp32 stream
    const float4* xin = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) + static_cast<long long>(row) * p.ldaux + col0);
    float4* xout = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.c) + static_cast<long long>(row) * p.ldc + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (4 * j < ncols) {
        float4 x = __ldg(xin + j);
        x.x += ptx::bf16_round(v[4 * j + 0]);
        x.y += ptx::bf16_round(v[4 * j + 1]);
        x.z += ptx::bf16_round(v[4 * j + 2]);
        x.w += ptx::bf16_round(v[4 * j + 3]);
        xout[j] = x;
      }
    }
  } else if constexpr (EPI == ABCGPT_EPI_F32_RED) {
    float* c = reinterpret_cast<float*>(p.c) + static_cast<long long>(row) * p.ldc + col0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (4 * j < ncols) ptx::red_add_v4(c + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {  // ABCGPT_EPI_F32
    float4* c = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.c) + static_cast<long long>(row) * p.ldc + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (4 * j < ncols) c[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
}

__device__ __forceinline__ void timed_wait(uint64_t* bar, uint32_t parity, int tag, unsigned long long* stats, int slot,
                                           long long& acc) {
  if (stats == nullptr) {
    ptx::mbar_wait(bar, parity, tag);
  } else {
    const long long t0 = clock64();
    ptx::mbar_wait(bar, parity, tag);
    acc += clock64() - t0;
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tfull[s], 1);
      ptx::mbar_init(&tempty[s], kNumEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_work = p.num_m_blk * p.num_n_blk * p.splits;
  const long long t_start = p.stats ? clock64() : 0;
  long long w0 = 0, w1 = 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w % p.splits;
        const int tile = w / p.splits;
        const int m_blk = tile / p.num_n_blk;
        const int n_blk = tile % p.num_n_blk;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_k_blk, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          timed_wait(&empty[stage], phase ^ 1, 1, p.stats, 0, w0);
          ptx::mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          if constexpr (!A_MN) {
            ptx::tma_load_2d(sa, &tmA, &full[stage], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int a = 0; a < BM / 64; ++a)
              ptx::tma_load_2d(sa + a * (BK * 128), &tmA, &full[stage], m_blk * BM + a * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            ptx::tma_load_2d(sb, &tmB, &full[stage], kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int b = 0; b < BN / 64; ++b)
              ptx::tma_load_2d(sb + b * (BK * 128), &tmB, &full[stage], n_blk * BN + b * 64, kb * BK);
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (p.stats) atomicAdd(&p.stats[0], static_cast<unsigned long long>(w0));
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
        const int split = w % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_k_blk, kb0 + p.kb_per_split);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        timed_wait(&tempty[as], aphase ^ 1, 2, p.stats, 2, w1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          timed_wait(&full[stage], phase, 3, p.stats, 1, w0);
          ptx::tc_fence_after();
          const uint32_t a_base = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t b_base = a_base + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major SW128: 16 bf16 of K = 32 bytes inside the 128-byte swizzle row; SBO = 8 rows * 128 B.
            // MN-major SW128: 16 K-rows = 2048 B; LBO = one 64-wide MN atom (BK rows * 128 B), SBO = 1024 B.
            const uint64_t adesc = A_MN ? ptx::umma_smem_desc(a_base + k * 2048, BK * 128, 1024)
                                        : ptx::umma_smem_desc(a_base + k * 32, 0, 1024);
            const uint64_t bdesc = B_MN ? ptx::umma_smem_desc(b_base + k * 2048, BK * 128, 1024)
                                        : ptx::umma_smem_desc(b_base + k * 32, 0, 1024);
            ptx::umma_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::umma_commit(&tfull[as]);  // accumulator complete
      }
      if (p.stats) {
        atomicAdd(&p.stats[1], static_cast<unsigned long long>(w0));
        atomicAdd(&p.stats[2], static_cast<unsigned long long>(w1));
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, 32*quarter+32) are visible to this warp
    const int half = (warp - 2) >> 2;      // which half of the BN columns
    constexpr int COLS_PER_WARP = BN / 2;
    int it = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
      const int tile = w / p.splits;
      const int m_blk = tile / p.num_n_blk;
      const int n_blk = tile % p.num_n_blk;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      timed_wait(&tfull[as], aphase, 4, p.stats, 3, w0);
      ptx::tc_fence_after();
      const int row = m_blk * BM + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + half * COLS_PER_WARP;
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld32(taddr + c, r);
        ptx::tmem_ld_wait();
        epilogue_chunk<EPI>(p, row, n_blk * BN + half * COLS_PER_WARP + c, r);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[as]);
    }
    if (p.stats && warp == 2 && lane == 0) {
      atomicAdd(&p.stats[3], static_cast<unsigned long long>(w0));
      atomicAdd(&p.stats[5], static_cast<unsigned long long>(clock64() - t_start));
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int BN, bool A_MN, bool B_MN, int EPI>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid, cudaStream_t stream) {
  using C = Cfg<BN>;
  auto kern = gemm_kernel<BN, A_MN, B_MN, EPI>;
  static bool configured = false;  // per instantiation; benign race (idempotent)
  if (!configured) {
    ABCGPT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  kern<<<grid, kNumThreads, C::SMEM_BYTES, stream>>>(tmA, tmB, p);
  return launch_status("gemm_kernel");
}

template <int BN, bool A_MN, bool B_MN>
int dispatch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid,
                 cudaStream_t stream) {
  switch (epi) {
    case ABCGPT_EPI_BF16: return launch<BN, A_MN, B_MN, ABCGPT_EPI_BF16>(tmA, tmB, p, grid, stream);
    case ABCGPT_EPI_GELU: return launch<BN, A_MN, B_MN, ABCGPT_EPI_GELU>(tmA, tmB, p, grid, stream);
    case ABCGPT_EPI_RESID: return launch<BN, A_MN, B_MN, ABCGPT_EPI_RESID>(tmA, tmB, p, grid, stream);
    case ABCGPT_EPI_DGELU: return launch<BN, A_MN, B_MN, ABCGPT_EPI_DGELU>(tmA, tmB, p, grid, stream);
    case ABCGPT_EPI_F32_RED: return launch<BN, A_MN, B_MN, ABCGPT_EPI_F32_RED>(tmA, tmB, p, grid, stream);
    case ABCGPT_EPI_F32: return launch<BN, A_MN, B_MN, ABCGPT_EPI_F32>(tmA, tmB, p, grid, stream);
  }
  return fail(-1, "unknown GEMM epilogue %d", epi);
}

template <int BN>
int dispatch_major(int a_mn, int b_mn, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                   int grid, cudaStream_t stream) {
  if (!a_mn && !b_mn) return dispatch_epi<BN, false, false>(epi, tmA, tmB, p, grid, stream);
  if (!a_mn && b_mn) return dispatch_epi<BN, false, true>(epi, tmA, tmB, p, grid, stream);
  if (a_mn && b_mn) return dispatch_epi<BN, true, true>(epi, tmA, tmB, p, grid, stream);
  return fail(-1, "GEMM operand combination A=MN-major,B=K-major is not instantiated");
}

}  // namespace

int gemm_bf16(const void* a, int a_mn, long long lda, const void* b, int b_mn, long long ldb, int M, int N, int K,
              int epi, void* c, long long ldc, void* c2, long long ldc2, const void* aux, long long ldaux,
              const float* bias, int bn_hint, int splits_hint, cudaStream_t stream) {

  ABCGPT_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: M,N,K must be positive (got %d,%d,%d)", M, N, K);
  ABCGPT_CHECK_ARG(N % 8 == 0, "gemm: N must be a multiple of 8 (pad the output; got %d)", N);
  ABCGPT_CHECK_ARG(c != nullptr, "gemm: null output");
  ABCGPT_CHECK_ARG(epi != ABCGPT_EPI_GELU || c2 != nullptr, "gemm: GELU epilogue needs the second output");
  ABCGPT_CHECK_ARG((epi != ABCGPT_EPI_RESID && epi != ABCGPT_EPI_DGELU) || aux != nullptr,
                   "gemm: epilogue %d needs the aux input", epi);

  // tile-N choice: the widest tile that keeps the persistent schedule's last wave reasonably full
  const int sms = sm_count();
  const int num_m_blk = (M + BM - 1) / BM;
  const int num_k_blk = (K + BK - 1) / BK;
  int bn = bn_hint;
  if (bn == 0) {
    double best = -1.0;
    const int cands[2] = {256, 128};
    for (int ci = 0; ci < 2; ++ci) {
      const int cbn = cands[ci];
      if (cbn == 256 && N <= 128) continue;
      const long long tiles = static_cast<long long>(num_m_blk) * ((N + cbn - 1) / cbn);
      const long long waves = (tiles + sms - 1) / sms;
      // useful fraction of MMA issue slots: quantisation x padded-N waste, with a mild preference for 256
      const double eff = (static_cast<double>(tiles) / (waves * sms)) * (static_cast<double>(N) / (((N + cbn - 1) / cbn) * cbn)) *
                         (cbn == 256 ? 1.0 : 0.93);
      if (eff > best) {
        best = eff;
        bn = cbn;
      }
    }
  }
  ABCGPT_CHECK_ARG(bn == 128 || bn == 256, "gemm: unsupported tile N %d", bn);
  const int num_n_blk = (N + bn - 1) / bn;

  int splits = 1;
  if (epi == ABCGPT_EPI_F32_RED) {
    const long long tiles = static_cast<long long>(num_m_blk) * num_n_blk;
    splits = splits_hint > 0 ? splits_hint : static_cast<int>((2LL * sms + tiles - 1) / tiles);
    if (splits > num_k_blk) splits = num_k_blk;
    if (splits < 1) splits = 1;
  }
  const int kb_per_split = (num_k_blk + splits - 1) / splits;
  splits = (num_k_blk + kb_per_split - 1) / kb_per_split;

  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn)
    rc = encode_tmap_2d(&tmA, a, 2, K, M, lda * 2, BK, BM, true);
  else
    rc = encode_tmap_2d(&tmA, a, 2, M, K, lda * 2, 64, BK, true);
  if (rc) return rc;
  if (!b_mn)
    rc = encode_tmap_2d(&tmB, b, 2, K, N, ldb * 2, BK, bn, true);
  else
    rc = encode_tmap_2d(&tmB, b, 2, N, K, ldb * 2, 64, BK, true);
  if (rc) return rc;

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.num_m_blk = num_m_blk; p.num_n_blk = num_n_blk; p.num_k_blk = num_k_blk;
  p.splits = splits; p.kb_per_split = kb_per_split;
  p.c = c; p.ldc = ldc; p.c2 = c2; p.ldc2 = ldc2; p.aux = aux; p.ldaux = ldaux; p.bias = bias; p.stats = g_gemm_stats;

  const long long total = static_cast<long long>(num_m_blk) * num_n_blk * splits;
  const int grid = static_cast<int>(total < sms ? total : sms);
  if (bn == 256) return dispatch_major<256>(a_mn, b_mn, epi, tmA, tmB, p, grid, stream);
  return dispatch_major<128>(a_mn, b_mn, epi, tmA, tmB, p, grid, stream);
}

}  // namespace abcgpt
