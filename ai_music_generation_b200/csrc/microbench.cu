// Hardware micro-benchmarks behind abcgpt_debug_tmem_ld_bench / abcgpt_debug_mma_bench (tools/tmem_bench.py,
// tools/mma_bench.py): tcgen05.ld throughput / latency and tcgen05.mma cost by shape, operand source and CTA group.  Not on
// any product path; their results decided the kernel designs.
#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

// ---- debug micro-benchmark: tcgen05.ld throughput / latency (tools/tmem_bench.py) -------------------------------------
namespace abcgpt {
namespace {
template <int INFLIGHT>
__global__ void __launch_bounds__(256, 1) tmem_ld_bench_kernel(long long* out, int iters, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    ptx::tmem_alloc(&slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    uint32_t v[32], w[32], x[32], y[32];
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      ptx::tmem_ld32(base, v);
      if (INFLIGHT >= 2) ptx::tmem_ld32(base + 32, w);
      if (INFLIGHT >= 4) {
        ptx::tmem_ld32(base + 64, x);
        ptx::tmem_ld32(base + 96, y);
      }
      ptx::tmem_ld_wait();
      acc += v[0] ^ v[13] ^ v[31];
      if (INFLIGHT >= 2) acc += w[0] ^ w[13] ^ w[31];
      if (INFLIGHT >= 4) acc += (x[0] ^ x[13] ^ x[31]) + (y[0] ^ y[13] ^ y[31]);
    }
    t1 = clock64();
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && warp < nwarps) {
    out[blockIdx.x * 8 + warp] = t1 - t0 + (acc == 0x12345u ? 1 : 0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(slot, 512);
}
}  // namespace
int tmem_ld_bench(long long* out, int iters, int nwarps, int inflight, cudaStream_t stream) {
  if (inflight >= 4) tmem_ld_bench_kernel<4><<<1, 256, 0, stream>>>(out, iters, nwarps);
  else if (inflight >= 2) tmem_ld_bench_kernel<2><<<1, 256, 0, stream>>>(out, iters, nwarps);
  else tmem_ld_bench_kernel<1><<<1, 256, 0, stream>>>(out, iters, nwarps);
  return launch_status("tmem_ld_bench_kernel");
}
}  // namespace abcgpt

// ---- debug micro-benchmark: tcgen05.mma cost by shape / operand source (tools/mma_bench.py) ---------------------------
namespace abcgpt {
namespace {
// mode bits: 0 = A from smem K-major, 1 = A from smem MN-major, 2 = A from TMEM; +4 = B MN-major (else K-major)
template <int N>
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(long long* out, int iters, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = ptx::uniform(threadIdx.x >> 5);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ptx::uniform(slot);
  if (warp == 1) {
    const bool issue = ptx::elect_one();
    const int amode = mode & 3, b_mn = (mode >> 2) & 1;
    const uint32_t sA = ptx::smem_u32(smem), sB = sA + 32768;
    const uint32_t idesc = ptx::umma_idesc_bf16(128, N, amode == 1 ? 1 : 0, b_mn);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adesc = amode == 1 ? ptx::umma_smem_desc(sA + k * 2048, 8192, 1024) : ptx::umma_smem_desc(sA + k * 32, 0, 1024);
        const uint64_t bdesc = b_mn ? ptx::umma_smem_desc(sB + k * 2048, 8192, 1024) : ptx::umma_smem_desc(sB + k * 32, 0, 1024);
        const uint32_t dcol = (mode & 8) ? ((k & 1) ? 256u : 0u) : 256u;  // bit 3: alternate between two accumulators
        if (issue) {
          if (amode == 2) ptx::umma_ts(tmem_base + dcol, tmem_base + 480 + 8 * (k & 3), bdesc, idesc, 1);
          else ptx::umma_ss(tmem_base + dcol, adesc, bdesc, idesc, 1);
        }
      }
    }
    if (issue) ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0, 99);
    const long long t1 = clock64();
    if (issue) out[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}
}  // namespace
int mma_bench(long long* out, int iters, int n, int mode, cudaStream_t stream) {
  const int smem = 100 * 1024;
  static bool done = false;
  if (!done) {
    ABCGPT_CUDA(cudaFuncSetAttribute(mma_bench_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ABCGPT_CUDA(cudaFuncSetAttribute(mma_bench_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ABCGPT_CUDA(cudaFuncSetAttribute(mma_bench_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    done = true;
  }
  if (n == 64) mma_bench_kernel<64><<<1, 128, smem, stream>>>(out, iters, mode);
  else if (n == 128) mma_bench_kernel<128><<<1, 128, smem, stream>>>(out, iters, mode);
  else mma_bench_kernel<256><<<1, 128, smem, stream>>>(out, iters, mode);
  return launch_status("mma_bench_kernel");
}
}  // namespace abcgpt

// ---- debug micro-benchmark: cta_group::2 tcgen05.mma 256 x N x 16 (tools/mma_bench.py) ---------------------------------
namespace abcgpt {
namespace {
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma2_bench_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = ptx::uniform(threadIdx.x >> 5);
  const uint32_t rank = ptx::cluster_ctarank();
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc_2sm(&slot, 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ptx::uniform(slot);
  if (warp == 1 && rank == 0) {
    const bool issue = ptx::elect_one();
    const uint32_t sA = ptx::smem_u32(smem), sB = sA + 32768;
    const uint32_t idesc = ptx::umma_idesc_bf16(256, N, 0, 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adesc = ptx::umma_smem_desc(sA + k * 32, 0, 1024);
        const uint64_t bdesc = ptx::umma_smem_desc(sB + k * 32, 0, 1024);
        if (issue) ptx::umma_ss_2sm(tmem_base, adesc, bdesc, idesc, 1);
      }
    }
    if (issue) ptx::umma_commit_2sm(&bar, 1);
    ptx::mbar_wait(&bar, 0, 98);
    const long long t1 = clock64();
    if (issue) out[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 0) ptx::tmem_dealloc_2sm(tmem_base, 512);
}
}  // namespace
int mma2_bench(long long* out, int iters, int n, cudaStream_t stream) {
  const int smem = 100 * 1024;
  static bool done = false;
  if (!done) {
    ABCGPT_CUDA(cudaFuncSetAttribute(mma2_bench_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ABCGPT_CUDA(cudaFuncSetAttribute(mma2_bench_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ABCGPT_CUDA(cudaFuncSetAttribute(mma2_bench_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    done = true;
  }
  if (n == 64) mma2_bench_kernel<64><<<2, 128, smem, stream>>>(out, iters);
  else if (n == 128) mma2_bench_kernel<128><<<2, 128, smem, stream>>>(out, iters);
  else mma2_bench_kernel<256><<<2, 128, smem, stream>>>(out, iters);
  return launch_status("mma2_bench_kernel");
}
}  // namespace abcgpt

// ---- debug probe: operand formats of the CTA-pair MMA used by the pair attention kernels (tools/pair_probe.py) ---------
// D[256 x 64] = A[256 x 64] * B[64 x 64] with one tcgen05.mma.cta_group::2 chain (K = 64 = 4 instructions), A rows split over
// the two CTAs, B split by N (32 columns per CTA).  mode bits:
//   bit 0  B layout: 0 = MN-major [64 k x 32 n] slab, 64-byte rows, SWIZZLE_64B (TMA box 32 x 64)
//                    1 = K-major  [32 n x 64 k] slab, 128-byte rows, SWIZZLE_128B (B given transposed: bt[n][k])
//   bit 1  A source: 0 = shared memory (K-major SW128), 1 = tensor memory (bf16 pairs written by the CTA's threads)
namespace abcgpt {
namespace {
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __nv_bfloat16* __restrict__ a,
                  float* __restrict__ d, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 32768);
  uint64_t* done = full + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = ptx::uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool b_kmajor = mode & 1, a_tmem = mode & 2;
  if (threadIdx.x == 0) {
    ptx::mbar_init(full, 1);
    ptx::mbar_init(done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(slot, 128);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ptx::uniform(*slot);
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  if (a_tmem) {  // A row (128 * rank + tid) as 32 packed bf16 pairs into TMEM columns [64, 96)
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a + (128 * rank + threadIdx.x) * 64);
    uint32_t w[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) w[i] = src[i];
    ptx::tmem_st32(tmem_base + lane_off + 64, w);
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
  }
  ptx::cluster_sync_all();  // both CTAs' TMEM operands are in place
  ptx::tc_fence_after();
  if (warp == 0) {
    const bool issue = ptx::elect_one();
    const uint32_t full_leader = ptx::mapa(ptx::smem_u32(full), 0);
    const uint32_t b_bytes = b_kmajor ? 4096 : 4096;
    if (rank == 0 && issue) ptx::mbar_expect_tx(full, 2 * (16384 + b_bytes));
    if (issue) ptx::tma_load_2d_2sm(smem, &tmA, full_leader, 0, 128 * static_cast<int>(rank));
    if (b_kmajor) {
      if (issue) ptx::tma_load_2d_2sm(smem + 16384, &tmB, full_leader, 0, 32 * static_cast<int>(rank));   // rows [32 r, +32) of bt[n][k]
    } else {
      if (issue) ptx::tma_load_2d_2sm(smem + 16384, &tmB, full_leader, 32 * static_cast<int>(rank), 0);   // columns [32 r, +32) of b[k][n]
    }
    __syncwarp();
  } else if (warp == 1 && rank == 0) {
    const bool issue = ptx::elect_one();
    ptx::mbar_wait(full, 0, 97);
    ptx::tc_fence_after();
    const uint32_t sA = ptx::smem_u32(smem), sB = sA + 16384;
    const uint32_t idesc = ptx::umma_idesc_bf16(256, 64, 0, b_kmajor ? 0 : 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t adesc = ptx::umma_smem_desc(sA + k * 32, 0, 1024);
      // MN-major SW64: 16 k-rows of 64 bytes per instruction = 1024 bytes; 8-row groups 512 bytes apart (SBO); one 32-wide MN atom
      const uint64_t bdesc = b_kmajor ? ptx::umma_smem_desc(sB + k * 32, 0, 1024) : ptx::umma_smem_desc_lt(sB + k * 1024, 512, 512, 4);
      if (issue) {
        if (a_tmem) ptx::umma_ts_2sm(tmem_base, tmem_base + 64 + 8 * k, bdesc, idesc, k > 0);
        else ptx::umma_ss_2sm(tmem_base, adesc, bdesc, idesc, k > 0);
      }
    }
    if (issue) ptx::umma_commit_2sm(done, 0x3);
    __syncwarp();
  }
  ptx::mbar_wait(done, 0, 96);
  ptx::tc_fence_after();
  uint32_t v0[32], v1[32];
  ptx::tmem_ld32(tmem_base + lane_off, v0);
  ptx::tmem_ld32(tmem_base + lane_off + 32, v1);
  ptx::tmem_ld_wait();
  float* o = d + (128 * rank + threadIdx.x) * 64;
#pragma unroll
  for (int i = 0; i < 32; ++i) { o[i] = __uint_as_float(v0[i]); o[32 + i] = __uint_as_float(v1[i]); }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc_2sm(tmem_base, 128);
}
}  // namespace
int pair_probe(const void* a, const void* b, float* d, int mode, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = encode_tmap_2d(&tmA, a, 2, 64, 256, 128, 64, 128, true))) return rc;
  if (mode & 1) rc = encode_tmap_2d(&tmB, b, 2, 64, 64, 128, 64, 32, true);          // bt[n][k]: rows-half, SW128
  else rc = encode_tmap_2d_sw(&tmB, b, 2, 64, 64, 128, 32, 64, 64);                   // b[k][n]: columns-half, SW64
  if (rc) return rc;
  const int smem = 40 * 1024;
  static bool done = false;
  if (!done) {
    ABCGPT_CUDA(cudaFuncSetAttribute(pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    done = true;
  }
  pair_probe_kernel<<<2, 128, smem, stream>>>(tmA, tmB, reinterpret_cast<const __nv_bfloat16*>(a), d, mode);
  return launch_status("pair_probe_kernel");
}
}  // namespace abcgpt

// ---- debug micro-benchmark: MUFU.EX2 issue rate per scheduler (tools/mufu_bench.py) -------------------------------------
namespace abcgpt {
namespace {
__global__ void __launch_bounds__(1024, 1) mufu_bench_kernel(long long* out, float* sink, int iters, int mode) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * static_cast<float>(threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (mode == 0) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      } else if (mode == 1) {  // ex2 with an FMA in front (the softmax pattern)
        x[i] = fmaf(x[i], 0.999f, -0.0001f);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      } else if (mode == 3) {  // packed half: two results per MUFU instruction?
        uint32_t h = __float_as_uint(x[i]);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h));
        x[i] = __uint_as_float(h);
      } else if (mode == 4) {
        uint32_t h = __float_as_uint(x[i]);
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h));
        x[i] = __uint_as_float(h);
      } else {  // FMA-pipe only: Cody-Waite exp2 with a cubic (no MUFU)
        const float t = x[i] + 12582912.f;
        const float fr = x[i] - (t - 12582912.f);
        float p = fmaf(fr, 0.0555041f, 0.2402265f);
        p = fmaf(p, fr, 0.6931472f);
        p = fmaf(p, fr, 1.0f);
        x[i] = __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23)) * -0.5f;
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += x[i];
  if (acc == 1234.5f) sink[0] = acc;
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;
}
}  // namespace
int mufu_bench(long long* out, float* sink, int iters, int warps, int mode, cudaStream_t stream) {
  mufu_bench_kernel<<<1, warps * 32, 0, stream>>>(out, sink, iters, mode);
  return launch_status("mufu_bench_kernel");
}
}  // namespace abcgpt

// ---- debug micro-benchmark: tcgen05.ld while the tensor core is busy (tools/tmem_mma_bench.py) --------------------------
// One CTA: warp 1 issues a continuous chain of tcgen05.mma 128 x N x 16 (accumulator in TMEM columns [0, N)) while `nwarps`
// other warps loop over tcgen05.ld 32x32b (x32 or x16, `inflight` loads per wait) on columns [256, 512).  Question: what does
// a TMEM -> register load cost per warp, and how many bytes per clock does the SM deliver, when the MMA's own accumulator
// traffic competes for tensor memory — the situation of every epilogue / softmax warp of the GEMM and attention kernels.
namespace abcgpt {
namespace {
template <int N>
__global__ void __launch_bounds__(576, 1) tmem_mma_bench_kernel(long long* out, int iters, int nwarps, int x16, int inflight, int mma_iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = ptx::uniform(threadIdx.x >> 5);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ptx::uniform(slot);
  if (warp == 1) {
    const bool issue = ptx::elect_one();
    const uint32_t sA = ptx::smem_u32(smem), sB = sA + 32768;
    const uint32_t idesc = ptx::umma_idesc_bf16(128, N, 0, 0);
    const long long t0 = clock64();
    for (int i = 0; i < mma_iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adesc = ptx::umma_smem_desc(sA + k * 32, 0, 1024);
        const uint64_t bdesc = ptx::umma_smem_desc(sB + k * 32, 0, 1024);
        if (issue) ptx::umma_ss(tmem_base, adesc, bdesc, idesc, 1);
      }
    }
    if (issue) ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0, 95);
    const long long t1 = clock64();
    if (issue) out[0] = t1 - t0;
  } else if (warp >= 2 && warp < 2 + nwarps) {
    const uint32_t base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 256 + ((warp - 2) >> 2) * 64;
    uint32_t acc = 0;
    uint32_t v[32], w[32];
    uint32_t a16[16], b16[16];
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (x16) {
        ptx::tmem_ld16(base, a16);
        if (inflight >= 2) ptx::tmem_ld16(base + 16, b16);
        ptx::tmem_ld_wait();
        acc += a16[0] ^ a16[7] ^ a16[15];
        if (inflight >= 2) acc += b16[0] ^ b16[7] ^ b16[15];
      } else {
        ptx::tmem_ld32(base, v);
        if (inflight >= 2) ptx::tmem_ld32(base + 32, w);
        ptx::tmem_ld_wait();
        acc += v[0] ^ v[13] ^ v[31];
        if (inflight >= 2) acc += w[0] ^ w[13] ^ w[31];
      }
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0 + (acc == 0x12345u ? 1 : 0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}
}  // namespace
// out[0] = cycles of the MMA chain (4 * mma_iters instructions), out[2 + w] = cycles of load warp w
int tmem_mma_bench(long long* out, int iters, int nwarps, int x16, int inflight, int mma_n, int mma_iters, cudaStream_t stream) {
  const int smem = 100 * 1024;
  static bool done = false;
  if (!done) {
    ABCGPT_CUDA(cudaFuncSetAttribute(tmem_mma_bench_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ABCGPT_CUDA(cudaFuncSetAttribute(tmem_mma_bench_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ABCGPT_CUDA(cudaFuncSetAttribute(tmem_mma_bench_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    done = true;
  }
  if (nwarps < 0 || nwarps > 16) return fail(-1, "tmem_mma_bench: 0..16 load warps");
  if (mma_n == 64) tmem_mma_bench_kernel<64><<<1, 576, smem, stream>>>(out, iters, nwarps, x16, inflight, mma_iters);
  else if (mma_n == 128) tmem_mma_bench_kernel<128><<<1, 576, smem, stream>>>(out, iters, nwarps, x16, inflight, mma_iters);
  else tmem_mma_bench_kernel<256><<<1, 576, smem, stream>>>(out, iters, nwarps, x16, inflight, mma_iters);
  return launch_status("tmem_mma_bench_kernel");
}
}  // namespace abcgpt

// ---- debug micro-benchmark: MUFU.EX2 rate by operand format, compile-time modes (tools/mufu2_bench.py) -------------------
// The run-time `mode` dispatch of mufu_bench_kernel sits inside its inner loop and dominates its timing; here every format is
// its own instantiation: 16 independent dependency chains per thread, `iters` rounds.  MODE 0: ex2.approx.ftz.f32 (one result
// per lane-op); 1: ex2.approx.f16x2; 2: ex2.approx.ftz.bf16x2 (two results per lane-op if the unit is really packed);
// 3: the softmax pattern fma.f32x2 + 2 x ex2.f32 + cvt.bf16x2; 4: fma.f32x2 + cvt.f16x2 + ex2.f16x2 (the packed alternative).
namespace abcgpt {
namespace {
template <int MODE>
__global__ void __launch_bounds__(1024, 1) mufu2_bench_kernel(long long* out, float* sink, int iters) {
  float x[16];
  uint32_t h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    x[i] = -0.001f * static_cast<float>((threadIdx.x & 63) + i);
    h[i] = 0xB800B400u + i;  // two small negative halves
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if constexpr (MODE == 0) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        x[i] = -x[i];
      } else if constexpr (MODE == 1) {
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
        h[i] ^= 0x80008000u;
      } else if constexpr (MODE == 2) {
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
        h[i] ^= 0x80008000u;
      } else if constexpr (MODE == 3) {   // two fp32 exponentials + pack per pair (i, i^1 share a pair: 8 pairs)
        if (i < 8) {
          float2 t = __ffma2_rn(make_float2(x[2 * i], x[2 * i + 1]), make_float2(0.18f, 0.18f), make_float2(-1.f, -1.f));
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(t.x));
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(t.y));
          h[i] = ptx::pack_bf16x2(t.x, t.y);
          x[2 * i] = -t.x; x[2 * i + 1] = -t.y;
        }
      } else {                             // fma pair + cvt to f16x2 + ONE packed exponential per pair
        if (i < 8) {
          const float2 t = __ffma2_rn(make_float2(x[2 * i], x[2 * i + 1]), make_float2(0.18f, 0.18f), make_float2(-1.f, -1.f));
          uint32_t hh;
          asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hh) : "f"(t.y), "f"(t.x));
          asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(hh));
          h[i] = hh;
          x[2 * i] = -t.x * 0.5f; x[2 * i + 1] = -t.y * 0.5f;
        }
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += x[i] + __uint_as_float(h[i]);
  if (acc == 1234.5f) sink[0] = acc;
  if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
}
}  // namespace
int mufu2_bench(long long* out, float* sink, int iters, int warps, int mode, cudaStream_t stream) {
  switch (mode) {
    case 0: mufu2_bench_kernel<0><<<1, warps * 32, 0, stream>>>(out, sink, iters); break;
    case 1: mufu2_bench_kernel<1><<<1, warps * 32, 0, stream>>>(out, sink, iters); break;
    case 2: mufu2_bench_kernel<2><<<1, warps * 32, 0, stream>>>(out, sink, iters); break;
    case 3: mufu2_bench_kernel<3><<<1, warps * 32, 0, stream>>>(out, sink, iters); break;
    default: mufu2_bench_kernel<4><<<1, warps * 32, 0, stream>>>(out, sink, iters); break;
  }
  return launch_status("mufu2_bench_kernel");
}
}  // namespace abcgpt
