// Host-side helpers shared by all translation units of libabcgpt: error reporting across the C ABI
// (return codes + thread-local message, never exceptions), TMA tensor-map encoding through the driver
// entry point (no link-time libcuda dependency), device attribute cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace abcgpt {

// ---- error channel -------------------------------------------------------------------------------
char* last_error_buf();  // thread-local, 512 bytes
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
#define ABCGPT_CHECK_ARG(cond, ...)                      \
  do {                                                   \
    if (!(cond)) return ::abcgpt::fail(-1, __VA_ARGS__); \
  } while (0)
#define ABCGPT_CUDA(expr)                                                                               \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      return ::abcgpt::fail(static_cast<int>(_e), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                                        \
  } while (0)
inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(static_cast<int>(e), "launch of %s failed: %s", what, cudaGetErrorString(e));
  return 0;
}

// ---- device info ---------------------------------------------------------------------------------
int sm_count();  // SM count of the current device (cached per device)

// ---- programmatic dependent launch -----------------------------------------------------------------
// The step is a chain of ~250 (training) / ~90 (one decoded token) dependent kernels on one stream.  Kernels launched
// through launch_k carry the programmatic-stream-serialization attribute: their CTAs may become resident as soon as every
// CTA of the previous kernel has executed griddepcontrol.launch_dependents (first instruction of our kernels) or exited,
// run their prologue (barrier init, TMEM allocation, descriptor prefetch) and then block in griddepcontrol.wait until the
// previous grid has completed and its memory is visible.  Launch latency and prologues leave the critical path; ordering
// of every global access is unchanged (all of them sit behind the wait).  Off unless requested (set_pdl, see common.cu).
bool pdl_enabled();
void set_pdl(bool on);
template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}

// ---- TMA descriptors -----------------------------------------------------------------------------
// 2D bf16/fp32 row-major matrix [outer, inner] with row pitch `row_bytes`; box = [box_outer, box_inner];
// 128-byte swizzle (box_inner * elem_bytes must be 128) or none.
int encode_tmap_2d(CUtensorMap* tm, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer,
                   uint64_t row_bytes, uint32_t box_inner, uint32_t box_outer, bool swizzle128);
// same with an explicit swizzle width: 0 (none), 64 or 128 bytes (= the inner box size in bytes)
int encode_tmap_2d_sw(CUtensorMap* tm, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer,
                      uint64_t row_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes);
// 3D variant: [d2, d1, d0] with byte strides for d1 and d2.
int encode_tmap_3d(CUtensorMap* tm, const void* ptr, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                   uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2,
                   bool swizzle128);

}  // namespace abcgpt
