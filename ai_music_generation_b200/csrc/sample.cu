// Sampling head of GPT.generate (nanoGPT/model.py:316-328): logits / temperature, optional top-k crop, softmax, one
// multinomial draw per sequence — fused into ONE launch that writes the next token straight into the token column the
// following decode step reads, so sample.py's defaults (temperature 0.8, top_k 200; nanoGPT/sample.py:33-36) never leave the
// C ABI and the whole decode step stays a replayable launch list.
//
//   reference                                       here (one CTA per sequence, the row lives in shared memory)
//   logits = logits[:, -1, :] / temperature         x_i = bf16(float(l_i) / temperature)   (the reference's logits are bf16
//                                                   under autocast, and bf16_tensor / python_float rounds back to bf16)
//   v, _ = torch.topk(logits, min(top_k, V))        kth = exact k-th largest x (MSB-first radix select on order-preserving keys)
//   logits[logits < v[:, [-1]]] = -inf              keep x_i >= kth   (ties at the threshold are all kept, as in the reference)
//   probs = F.softmax(logits, dim=-1)               p_i = exp(x_i - max) in fp32 (autocast runs softmax in fp32)
//   idx_next = torch.multinomial(probs, 1)          inverse CDF: smallest i with cumsum(p)_i > u * sum(p)
// u is one Philox4x32-10 draw per (seed, sequence, decode position): a pure function of its counters, so a replayed launch
// list needs no generator state, and tests/oracle can regenerate the very same uniform on the host (oracle: philox_uniform).
// The stream of random numbers is NOT torch's (torch.multinomial draws per-category exponentials): parity with the reference
// is distributional (chi-square in tests/test_sampling_gpu.py) and exact given u (every sampled token is checked against the
// CDF interval that u falls into).
#include "common.h"
#include "kernels.h"

#include <cuda_bf16.h>

namespace abcgpt {
namespace {

constexpr int kSampleThreads = 128;

__device__ __forceinline__ uint32_t philox_uniform_bits(uint64_t seed, uint32_t row, uint64_t counter) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t c0 = row, c1 = static_cast<uint32_t>(counter), c2 = static_cast<uint32_t>(counter >> 32), c3 = 0x5A17u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c0;
}

// order-preserving map float -> uint32 (larger float <=> larger key)
__device__ __forceinline__ uint32_t order_key(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ int block_sum_int(int v, int* scratch) {
  v = __reduce_add_sync(0xffffffffu, v);
  __syncthreads();  // scratch reuse
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int w = 0; w < kSampleThreads / 32; ++w) t += scratch[w];
  return t;
}

__global__ void __launch_bounds__(kSampleThreads)
sample_topk_kernel(const __nv_bfloat16* __restrict__ logits, long long ldl, int V, float temperature, int top_k,
                   const unsigned long long* __restrict__ seed, long long counter, int64_t* __restrict__ out,
                   long long out_stride) {
  extern __shared__ float xs[];  // [V]
  __shared__ float red_f[kSampleThreads / 32];
  __shared__ int red_i[kSampleThreads / 32];
  __shared__ int red_idx[kSampleThreads / 32];
  __shared__ int chosen;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = blockIdx.x;
  const __nv_bfloat16* lr = logits + static_cast<long long>(row) * ldl;

  // ---- x = bf16(l / temperature); row max and its first index
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int i = tid; i < V; i += kSampleThreads) {
    const float x = __bfloat162float(__float2bfloat16_rn(__fdiv_rn(__bfloat162float(lr[i]), temperature)));
    xs[i] = x;
    if (x > mx) { mx = x; mi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > mx || (om == mx && oi < mi)) { mx = om; mi = oi; }
  }
  if (lane == 0) { red_f[warp] = mx; red_idx[warp] = mi; }
  __syncthreads();
  mx = red_f[0];
  mi = red_idx[0];
#pragma unroll
  for (int w = 1; w < kSampleThreads / 32; ++w)
    if (red_f[w] > mx || (red_f[w] == mx && red_idx[w] < mi)) { mx = red_f[w]; mi = red_idx[w]; }
  if (tid == 0) chosen = mi;  // fall-back if rounding leaves the target at the very end of the CDF

  // ---- k-th largest value (exact): build its key bit by bit, counting how many keys are >= the candidate
  uint32_t thr_key = 0u;  // everything passes
  if (top_k > 0 && top_k < V) {
    uint32_t prefix = 0u;
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = prefix | (1u << bit);
      int c = 0;
      for (int i = tid; i < V; i += kSampleThreads) c += (order_key(xs[i]) >= cand) ? 1 : 0;
      if (block_sum_int(c, red_i) >= top_k) prefix = cand;
    }
    thr_key = prefix;
  }

  // ---- softmax numerators over a contiguous chunk per thread (so that a prefix over threads is a prefix over tokens)
  const int chunk = (V + kSampleThreads - 1) / kSampleThreads;
  const int i0 = min(V, tid * chunk), i1 = min(V, i0 + chunk);
  float local = 0.f;
  for (int i = i0; i < i1; ++i) {
    const float x = xs[i];
    const float p = (order_key(x) >= thr_key) ? __expf(x - mx) : 0.f;
    xs[i] = p;  // only this thread touches [i0, i1)
    local += p;
  }
  // inclusive scan of `local` over the block
  float incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  __syncthreads();  // red_f reuse
  if (lane == 31) red_f[warp] = incl;
  __syncthreads();
  float warp_off = 0.f, total = 0.f;
#pragma unroll
  for (int w = 0; w < kSampleThreads / 32; ++w) {
    if (w < warp) warp_off += red_f[w];
    total += red_f[w];
  }
  const float excl = warp_off + incl - local;

  const uint32_t bits = philox_uniform_bits(*seed, static_cast<uint32_t>(row), static_cast<uint64_t>(counter));
  const float u = static_cast<float>(bits >> 8) * (1.0f / 16777216.0f);  // [0, 1)
  const float target = u * total;
  if (local > 0.f && target >= excl && target < excl + local) {
    float acc = excl;
    int pick = -1, last_pos = -1;
    for (int i = i0; i < i1; ++i) {
      const float p = xs[i];
      if (p > 0.f) {
        last_pos = i;
        acc += p;
        if (acc > target) { pick = i; break; }
      }
    }
    chosen = pick >= 0 ? pick : last_pos;  // exactly one thread's interval contains the target
  }
  __syncthreads();
  if (tid == 0) out[static_cast<long long>(row) * out_stride] = chosen;
}

}  // namespace

int sample_topk(const void* logits, long long ldl, int V, float temperature, int top_k, const void* seed, long long counter,
                int64_t* out, long long out_stride, int B, cudaStream_t stream) {
  ABCGPT_CHECK_ARG(logits && out && seed && V > 0 && B > 0, "sample_topk: bad arguments");
  ABCGPT_CHECK_ARG(temperature > 0.f, "sample_topk: temperature must be positive (got %g)", static_cast<double>(temperature));
  const size_t smem = static_cast<size_t>(V) * sizeof(float);
  ABCGPT_CHECK_ARG(smem <= 220 * 1024, "sample_topk: vocabulary of %d does not fit one CTA's shared memory", V);
  static bool configured = false;
  if (!configured && smem > 48 * 1024) {
    ABCGPT_CUDA(cudaFuncSetAttribute(sample_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  sample_topk_kernel<<<B, kSampleThreads, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(logits), ldl, V, temperature,
                                                          top_k, reinterpret_cast<const unsigned long long*>(seed), counter, out,
                                                          out_stride);
  return launch_status("sample_topk_kernel");
}

}  // namespace abcgpt
