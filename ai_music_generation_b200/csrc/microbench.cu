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
