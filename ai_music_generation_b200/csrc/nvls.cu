// Gradient exchange of the data-parallel step INSIDE our own kernel, over NVLink-switch multicast memory (NVLS):
// the mean of the replicas' gradients (torch DDP's all-reduce, nanoGPT/train.py:226-227) and the sum of squares that
// clip_grad_norm_ needs (train.py:352) in ONE pass, instead of ncclAllReduce kernels followed by a separate norm pass.
//
// Every rank owns a contiguous 1/world slice of the flat fp32 gradient arena.  For each 16 bytes of its slice a thread issues
//   multimem.ld_reduce.add.v4.f32   — the switch reads the element from EVERY replica and returns the sum (one NVLink read per
//                                     rank, no partial sums in HBM),
//   scales it by 1 / world, accumulates its square into the norm, and
//   multimem.st.v4.f32              — the switch writes the mean back into EVERY replica.
// So all replicas receive bit-identical gradients (each element is reduced exactly once, by its owner), the per-block partial
// norms are multicast into a [world x blocks] table that every rank then sums in the same order (abcgpt_sumsq_partials):
// identical clip coefficients, replicas stay bitwise in sync — the property the NCCL path has too (tests/_ddp_gpu_worker.py).
// The arena and the table live in symmetric memory (torch.distributed._symmetric_memory: allocation + handle exchange only);
// cross-rank ordering is the caller's: a device-side barrier before (all backward passes finished) and after (all slices
// written) the launch (ai_music_generation_b200/ddp.py).
#include "common.h"
#include "kernels.h"

namespace abcgpt {
namespace {

__device__ __forceinline__ float4 mm_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mm_st1(float* mc, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// world == 1 (no switch involved: plain loads / stores through the same code path; used by the co-residency probe and by
// single-rank process groups)
template <bool MC>
__device__ __forceinline__ float4 ld_sum(const float* p) {
  if constexpr (MC) return mm_ld_reduce_add(p);
  return *reinterpret_cast<const float4*>(p);
}
template <bool MC>
__device__ __forceinline__ void st_all(float* p, const float4& v) {
  if constexpr (MC) mm_st(p, v);
  else *reinterpret_cast<float4*>(p) = v;
}

constexpr int kUnroll = 4;   // 16-byte switch reductions in flight per thread (~3 us round trip: 128 blocks x 512 threads x 64 B = 4 MB)

template <bool MC>
__global__ void __launch_bounds__(512)
nvls_allreduce_sumsq_kernel(float* __restrict__ grad_mc, long long lo4, long long hi4, float scale, float* __restrict__ partials_mc,
                            int slot0) {
  __shared__ float sh[16];
  float s = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = lo4 + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  for (; i + (kUnroll - 1) * stride < hi4; i += kUnroll * stride) {
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) v[u] = ld_sum<MC>(grad_mc + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      v[u].x *= scale; v[u].y *= scale; v[u].z *= scale; v[u].w *= scale;
      s += (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
      st_all<MC>(grad_mc + 4 * (i + u * stride), v[u]);
    }
  }
  for (; i < hi4; i += stride) {
    float4 v = ld_sum<MC>(grad_mc + 4 * i);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    st_all<MC>(grad_mc + 4 * i, v);
  }
  s = warp_sum_f(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    s = warp_sum_f(s);
    if (threadIdx.x == 0) {   // this block's share of the norm, to every rank's table
      if constexpr (MC) mm_st1(partials_mc + slot0 + blockIdx.x, s);
      else partials_mc[slot0 + blockIdx.x] = s;
    }
  }
}

__global__ void __launch_bounds__(1024) sumsq_partials_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ out) {
  __shared__ float sh[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 1024) s += partial[i];   // fixed order: the same value on every rank
  s = warp_sum_f(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = sh[threadIdx.x];
    s = warp_sum_f(s);
    if (threadIdx.x == 0) out[0] += s;
  }
}

}  // namespace

// `threads`: 512 for a launch that has the GPU to itself; 128 (about 40 registers, no shared memory beyond 64 bytes) for launches
// that run BESIDE the backward's persistent GEMMs: one such CTA fits next to a 576-thread / 55 k-register / 225 KB GEMM CTA on the
// same SM, so the exchange takes issue slots and LSU bandwidth but never an SM the static tile schedule counts on — which is what
// an NCCL kernel (hundreds of threads, large shared-memory FIFOs) does to it.
int nvls_allreduce_sumsq(void* grad_mc, long long n, int rank, int world, float scale, void* partials_mc, int blocks_per_rank,
                         int threads, cudaStream_t stream) {
  ABCGPT_CHECK_ARG(threads >= 32 && threads <= 512 && threads % 32 == 0, "nvls_allreduce_sumsq: threads per block must be 32..512");
  ABCGPT_CHECK_ARG(grad_mc && partials_mc && n > 0 && n % 4 == 0, "nvls_allreduce_sumsq: the arena must hold a multiple of 4 floats");
  ABCGPT_CHECK_ARG(world >= 1 && rank >= 0 && rank < world && blocks_per_rank >= 1, "nvls_allreduce_sumsq: bad rank / world");
  ABCGPT_CHECK_ARG((reinterpret_cast<uintptr_t>(grad_mc) & 15) == 0 && (reinterpret_cast<uintptr_t>(partials_mc) & 3) == 0,
                   "nvls_allreduce_sumsq: multicast pointers must be 16-byte (arena) / 4-byte (table) aligned");
  const long long n4 = n / 4, per = (n4 + world - 1) / world;
  const long long lo = per * rank < n4 ? per * rank : n4, hi = lo + per < n4 ? lo + per : n4;
  if (world == 1)
    nvls_allreduce_sumsq_kernel<false><<<blocks_per_rank, threads, 0, stream>>>(reinterpret_cast<float*>(grad_mc), lo, hi, scale,
                                                                            reinterpret_cast<float*>(partials_mc), 0);
  else
    nvls_allreduce_sumsq_kernel<true><<<blocks_per_rank, threads, 0, stream>>>(reinterpret_cast<float*>(grad_mc), lo, hi, scale,
                                                                           reinterpret_cast<float*>(partials_mc), 0);
  return launch_status("nvls_allreduce_sumsq_kernel");
}

int sumsq_partials(const float* partials, int nparts, float* out, cudaStream_t stream) {
  ABCGPT_CHECK_ARG(partials && out && nparts > 0, "sumsq_partials: bad arguments");
  sumsq_partials_kernel<<<1, 1024, 0, stream>>>(partials, nparts, out);
  return launch_status("sumsq_partials_kernel");
}

}  // namespace abcgpt
