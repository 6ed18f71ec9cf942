// LayerNorm forward / backward (nanoGPT/model.py:18-27: F.layer_norm(x, (C,), weight, bias, 1e-5)).
//
// HBM-bound: one warp owns one row, the row lives in registers (C/128 float4 per lane), all global traffic
// is 128-bit and coalesced, reductions are warp shuffles.  The forward emits the bf16 copy the next GEMM
// consumes (the reference rounds the fp32 LN output to bf16 at the autocast boundary of nn.Linear), the
// backward fuses the residual-gradient add and emits both the fp32 stream gradient and its bf16 copy.
#include "common.h"
#include "dropout.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace abcgpt {
namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ADD: the residual add of the block is fused in front of the normalisation (model.py:103 `x = x + self.attn(...)` followed
// by :104 `self.ln_2(x)`): x_sum = x + [dropout](branch), branch = the bf16 output of the preceding Linear; x_sum is written
// (fp32 residual stream) and normalised in the same pass.  Same arithmetic as the RESID epilogue of the GEMM (csrc/gemm.cu):
// the bf16 Linear output is added into the fp32 stream, the residual-branch dropout masks / rescales / re-rounds it first.
template <int NV, bool ADD>  // NV float4 per lane: C <= NV*128
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
              __nv_bfloat16* __restrict__ y, float* __restrict__ yf, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, int M, int C, const __nv_bfloat16* __restrict__ branch,
              float* __restrict__ xsum, DropCfg drop) {
  ptx::pdl_launch_dependents();  // programmatic dependent launch: see launch_k (common.h)
  ptx::pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nvec = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * C);
  float4 v[NV];
  float s = 0.f;
  if constexpr (ADD) {
    const uint2* br = reinterpret_cast<const uint2*>(branch + static_cast<long long>(row) * C);
    float4* xs = reinterpret_cast<float4*>(xsum + static_cast<long long>(row) * C);
    uint2 bv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {  // the whole row of both operands is requested up front
      const int idx = lane + 32 * i;
      v[i] = (idx < nvec) ? __ldg(xr + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      bv[i] = (idx < nvec) ? __ldg(br + idx) : make_uint2(0u, 0u);
    }
    const uint32_t rk = drop_row_key(drop.key, static_cast<uint32_t>(row));
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = lane + 32 * i;
      float y0 = ptx::bf16lo(bv[i].x), y1 = ptx::bf16hi(bv[i].x), y2 = ptx::bf16lo(bv[i].y), y3 = ptx::bf16hi(bv[i].y);
      if (drop.thr16 != 0) {  // bf16 dropout like nn.Dropout on the bf16 Linear output: scale, round, or zero
        const uint32_t b0 = drop_pair_bits(rk, 2 * idx), b1 = drop_pair_bits(rk, 2 * idx + 1);
        y0 = drop_keep_lo(b0, drop.thr16) ? ptx::bf16_round(y0 * drop.inv_keep) : 0.f;
        y1 = drop_keep_hi(b0, drop.thr16) ? ptx::bf16_round(y1 * drop.inv_keep) : 0.f;
        y2 = drop_keep_lo(b1, drop.thr16) ? ptx::bf16_round(y2 * drop.inv_keep) : 0.f;
        y3 = drop_keep_hi(b1, drop.thr16) ? ptx::bf16_round(y3 * drop.inv_keep) : 0.f;
      }
      v[i].x += y0; v[i].y += y1; v[i].z += y2; v[i].w += y3;
      if (idx < nvec) xs[idx] = v[i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = lane + 32 * i;
      v[i] = (idx < nvec) ? __ldg(xr + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const float a = v[i].x - mean, bq = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      ss += (a * a + bq * bq) + (c * c + d * d);
    }
  }
  const float var = warp_sum(ss) / static_cast<float>(C);
  const float rstd = rsqrtf(var + 1e-5f);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  uint2* yr = y ? reinterpret_cast<uint2*>(y + static_cast<long long>(row) * C) : nullptr;
  float4* yfr = yf ? reinterpret_cast<float4*>(yf + static_cast<long long>(row) * C) : nullptr;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(w) + idx);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x;
      o.y = (v[i].y - mean) * rstd * g.y;
      o.z = (v[i].z - mean) * rstd * g.z;
      o.w = (v[i].w - mean) * rstd * g.w;
      if (b) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + idx);
        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
      }
      if (yr) yr[idx] = make_uint2(ptx::pack_bf16x2(o.x, o.y), ptx::pack_bf16x2(o.z, o.w));
      if (yfr) yfr[idx] = o;
    }
  }
}

template <int NV, bool HAS_BIAS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 2)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres,
              float* __restrict__ dx, __nv_bfloat16* __restrict__ dxb, float* __restrict__ dw, float* __restrict__ db,
              int M, int C, DropCfg drop) {
  ptx::pdl_launch_dependents();  // programmatic dependent launch: see launch_k (common.h)
  ptx::pdl_wait();
  extern __shared__ float red[];  // [kWarpsPerBlock][C] dweight partials (+ [kWarpsPerBlock][C] dbias partials)
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nvec = C >> 2;
  // Register budget (<= 128 so that two 256-thread blocks are resident per SM): the whole row -- x fp32, dy packed bf16 AND
  // the incoming residual gradient -- is requested up front, so one warp keeps 10.5 KB (C = 768) in flight instead of
  // 4.5 KB followed by a second dependent round trip; the per-warp dweight / dbias partials live in this warp's private
  // slice of shared memory (conflict-free float4 read-modify-write, ~50 of the ~500 cycles a row takes at HBM speed), and
  // gamma is re-read from L1 in both passes instead of being held.
  float4* accw_s = reinterpret_cast<float4*>(red) + warp * nvec;
  float4* accb_s = accw_s + kWarpsPerBlock * nvec;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      accw_s[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (HAS_BIAS) accb_s[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float invC = 1.0f / static_cast<float>(C);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  for (int row = blockIdx.x * kWarpsPerBlock + warp; row < M; row += gridDim.x * kWarpsPerBlock) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * C);
    const uint2* dyr = reinterpret_cast<const uint2*>(dy + static_cast<long long>(row) * C);
    const float4* drr = dres ? reinterpret_cast<const float4*>(dres + static_cast<long long>(row) * C) : nullptr;
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float4 xv[NV], rv[NV];
    uint2 dv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nvec) {
        xv[i] = __ldg(xr + idx);
        dv[i] = __ldg(dyr + idx);
        rv[i] = drr ? __ldg(drr + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        xv[i] = make_float4(mu, mu, mu, mu);
        dv[i] = make_uint2(0u, 0u);
        rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = lane + 32 * i;
      const float4 gw = (idx < nvec) ? __ldg(w4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 dyv = make_float4(ptx::bf16lo(dv[i].x), ptx::bf16hi(dv[i].x), ptx::bf16lo(dv[i].y), ptx::bf16hi(dv[i].y));
      const float4 xh = make_float4((xv[i].x - mu) * rs, (xv[i].y - mu) * rs, (xv[i].z - mu) * rs, (xv[i].w - mu) * rs);
      const float4 g = make_float4(dyv.x * gw.x, dyv.y * gw.y, dyv.z * gw.z, dyv.w * gw.w);
      s1 += (g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w);
      s2 += (g.x + g.y) + (g.z + g.w);
      if (idx < nvec) {
        float4 a = accw_s[idx];
        a.x += dyv.x * xh.x; a.y += dyv.y * xh.y; a.z += dyv.z * xh.z; a.w += dyv.w * xh.w;
        accw_s[idx] = a;
        if (HAS_BIAS) {
          float4 bsum = accb_s[idx];
          bsum.x += dyv.x; bsum.y += dyv.y; bsum.z += dyv.z; bsum.w += dyv.w;
          accb_s[idx] = bsum;
        }
      }
    }
    const float c1 = warp_sum(s1) * invC;
    const float c2 = warp_sum(s2) * invC;
    float4* dxr = reinterpret_cast<float4*>(dx + static_cast<long long>(row) * C);
    uint2* dxbr = dxb ? reinterpret_cast<uint2*>(dxb + static_cast<long long>(row) * C) : nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nvec) {
        const float4 gw = __ldg(w4 + idx);
        const float4 dyv = make_float4(ptx::bf16lo(dv[i].x), ptx::bf16hi(dv[i].x), ptx::bf16lo(dv[i].y), ptx::bf16hi(dv[i].y));
        float4 o;
        o.x = (dyv.x * gw.x - c2 - (xv[i].x - mu) * rs * c1) * rs + rv[i].x;
        o.y = (dyv.y * gw.y - c2 - (xv[i].y - mu) * rs * c1) * rs + rv[i].y;
        o.z = (dyv.z * gw.z - c2 - (xv[i].z - mu) * rs * c1) * rs + rv[i].z;
        o.w = (dyv.w * gw.w - c2 - (xv[i].w - mu) * rs * c1) * rs + rv[i].w;
        dxr[idx] = o;
        if (dxbr) {
          // the bf16 copy feeds the backward of the Linear whose OUTPUT was dropped out in the forward: it carries that
          // site's mask (the fp32 stream gradient above does not: the skip connection is not dropped)
          uint32_t lo = ptx::pack_bf16x2(o.x, o.y), hi = ptx::pack_bf16x2(o.z, o.w);
          if (drop.thr16 != 0) {
            const uint32_t rk = drop_row_key(drop.key, static_cast<uint32_t>(row));
            const uint32_t b0 = drop_pair_bits(rk, 2 * idx), b1 = drop_pair_bits(rk, 2 * idx + 1);
            lo = ptx::pack_bf16x2(drop_keep_lo(b0, drop.thr16) ? ptx::bf16lo(lo) * drop.inv_keep : 0.f,
                                  drop_keep_hi(b0, drop.thr16) ? ptx::bf16hi(lo) * drop.inv_keep : 0.f);
            hi = ptx::pack_bf16x2(drop_keep_lo(b1, drop.thr16) ? ptx::bf16lo(hi) * drop.inv_keep : 0.f,
                                  drop_keep_hi(b1, drop.thr16) ? ptx::bf16hi(hi) * drop.inv_keep : 0.f);
          }
          dxbr[idx] = make_uint2(lo, hi);
        }
      }
    }
  }
  // block-level reduction of the per-warp dweight / dbias partials, then one atomic per column per block
  __syncthreads();
  for (int pass = 0; pass < (HAS_BIAS ? 2 : 1); ++pass) {
    float* dst = pass == 0 ? dw : db;
    if (dst == nullptr) continue;  // uniform across the block
    const float* src = red + pass * kWarpsPerBlock * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int ww = 0; ww < kWarpsPerBlock; ++ww) s += src[ww * C + c];
      atomicAdd(dst + c, s);
    }
  }
}

}  // namespace

#define LN_DISPATCH(NVV, ...) \
  switch (NVV) {              \
    case 1: { constexpr int NV = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int NV = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int NV = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int NV = 4; __VA_ARGS__; } break; \
    case 5: { constexpr int NV = 5; __VA_ARGS__; } break; \
    case 6: { constexpr int NV = 6; __VA_ARGS__; } break; \
    case 7: { constexpr int NV = 7; __VA_ARGS__; } break; \
    case 8: { constexpr int NV = 8; __VA_ARGS__; } break; \
    case 9: case 10: case 11: case 12: { constexpr int NV = 12; __VA_ARGS__; } break; \
    default: { constexpr int NV = 16; __VA_ARGS__; } break; \
  }

int layernorm_fwd(const float* x, const float* weight, const float* bias, void* y_bf16, float* y_f32, float* mean,
                  float* rstd, int M, int C, cudaStream_t stream) {
  ABCGPT_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0 && C <= 2048, "layernorm: need C %% 4 == 0 and C <= 2048 (got %d)", C);
  ABCGPT_CHECK_ARG(x && weight && (y_bf16 || y_f32), "layernorm_fwd: null pointer");
  const int nv = (C + 127) / 128;
  const int grid = (M + kWarpsPerBlock - 1) / kWarpsPerBlock;
  LN_DISPATCH(nv, (launch_k(ln_fwd_kernel<NV, false>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, stream,
                      x, weight, bias, reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32, mean, rstd, M, C,
                      static_cast<const __nv_bfloat16*>(nullptr), static_cast<float*>(nullptr), DropCfg{})));
  return launch_status("ln_fwd_kernel");
}

int layernorm_fwd_resid(const float* x_in, const void* branch_bf16, float* x_out, const float* weight, const float* bias,
                        void* y_bf16, float* mean, float* rstd, int M, int C, float drop_p, uint32_t drop_key,
                        cudaStream_t stream) {
  ABCGPT_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0 && C <= 2048, "layernorm: need C %% 4 == 0 and C <= 2048 (got %d)", C);
  ABCGPT_CHECK_ARG(x_in && branch_bf16 && x_out && weight && y_bf16, "layernorm_fwd_resid: null pointer");
  ABCGPT_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "layernorm_fwd_resid: dropout probability must be in [0, 1)");
  const DropCfg drop = make_drop(drop_p, drop_key);
  const int nv = (C + 127) / 128;
  const int grid = (M + kWarpsPerBlock - 1) / kWarpsPerBlock;
  LN_DISPATCH(nv, (launch_k(ln_fwd_kernel<NV, true>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, stream,
                      x_in, weight, bias, reinterpret_cast<__nv_bfloat16*>(y_bf16), static_cast<float*>(nullptr), mean, rstd, M, C,
                      reinterpret_cast<const __nv_bfloat16*>(branch_bf16), x_out, drop)));
  return launch_status("ln_fwd_kernel (residual add)");
}

int layernorm_bwd(const void* dy_bf16, const float* x, const float* weight, const float* mean, const float* rstd,
                  const float* dresid_in, float* dx_out, void* dx_bf16, float* dweight, float* dbias, int M, int C,
                  float drop_p, uint32_t drop_key, cudaStream_t stream) {
  const DropCfg drop = make_drop(drop_p, drop_key);
  ABCGPT_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0 && C <= 2048, "layernorm: need C %% 4 == 0 and C <= 2048 (got %d)", C);
  ABCGPT_CHECK_ARG(dy_bf16 && x && weight && mean && rstd && dx_out, "layernorm_bwd: null pointer");
  const int nv = (C + 127) / 128;
  int grid = (M + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int cap = sm_count() * 2;  // two resident blocks per SM, grid-stride over rows
  if (grid > cap) grid = cap;
  const bool has_bias = dbias != nullptr;
  const size_t smem = static_cast<size_t>(has_bias ? 2 : 1) * kWarpsPerBlock * C * sizeof(float);
  LN_DISPATCH(nv, {
    auto launch = [&](auto kern) -> int {
      if (smem > 48 * 1024) ABCGPT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      launch_k(kern, dim3(grid), dim3(kWarpsPerBlock * 32), smem, stream, reinterpret_cast<const __nv_bfloat16*>(dy_bf16), x,
               weight, mean, rstd, dresid_in, dx_out, reinterpret_cast<__nv_bfloat16*>(dx_bf16), dweight, dbias, M, C, drop);
      return 0;
    };
    int rc = has_bias ? launch(ln_bwd_kernel<NV, true>) : launch(ln_bwd_kernel<NV, false>);
    if (rc) return rc;
  });
  return launch_status("ln_bwd_kernel");
}

}  // namespace abcgpt
