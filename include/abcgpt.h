/*
 * libabcgpt — C ABI of the B200-native (sm_100a) kernels behind the nanoGPT training / sampling step of
 * Jakub-Kucinski/ai-music-generation (reference: nanoGPT/model.py, nanoGPT/train.py, nanoGPT/sample.py).
 *
 * The reference has no FFI of its own: its boundary for this path is the nn.Module / optimizer duck type
 * (GPT(GPTConfig), configure_optimizers, generate).  The Python host side in ai_music_generation_b200/ keeps
 * that surface and binds the entry points below with ctypes; each entry cites the reference call site whose
 * arithmetic it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory that must stay alive until `stream` has
 *     passed the call; the library never allocates, frees or retains device memory;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no host sync, graph-capturable);
 *   - return value: 0 ok; <0 invalid argument; >0 cudaError_t (or 1000+CUresult for TMA descriptor
 *     encoding); the message is available from abcgpt_last_error() (thread-local);
 *   - bf16 tensors are row-major with the leading dimension given in ELEMENTS; rows that feed a GEMM or
 *     the attention kernels must start 16-byte aligned (ld % 8 == 0).
 */
#ifndef ABCGPT_H_
#define ABCGPT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABCGPT_VERSION 100

/* GEMM epilogues */
#define ABCGPT_EPI_BF16 0    /* C(bf16) = acc (+bias)                                              */
#define ABCGPT_EPI_GELU 1    /* C(bf16) = h = acc (+bias);  C2(bf16) = gelu_erf(h)   model.py:88-89 */
#define ABCGPT_EPI_RESID 2   /* C(fp32) = AUX(fp32) + bf16(acc (+bias))             model.py:104-105 */
#define ABCGPT_EPI_DGELU 3   /* C(bf16) = bf16(acc) * gelu_erf'(AUX(bf16))          GELU backward    */
#define ABCGPT_EPI_F32_RED 4 /* C(fp32) += acc   (red.global.add; split-K; wgrad accumulation)      */
#define ABCGPT_EPI_F32 5     /* C(fp32) = acc                                                       */

int abcgpt_version(void);
const char* abcgpt_last_error(void);

/*
 * C[M,N] = sum_k A[m,k] * B[n,k]   bf16 operands, fp32 accumulation on tcgen05 tensor cores.
 * Replaces the cuBLAS GEMMs behind nn.Linear (model.py:56,75,88,90,186,190) and their autograd dgrad/wgrad.
 *   a_mn_major = 0: A is [M,K] row-major (lda);  1: A is stored [K,M] row-major (lda)  (i.e. transposed)
 *   b_mn_major = 0: B is [N,K] row-major (ldb);  1: B is stored [K,N] row-major (ldb)
 *   forward  Y=X W^T : a=X(0)  b=W(0)      dgrad dX=dY W : a=dY(0) b=W(1)      wgrad dW=dY^T X : a=dY(1) b=X(1)
 * N must be a multiple of 8.  bias (fp32[N]) may be NULL.  tile_n / splits = 0 lets the library choose.
 */
int abcgpt_gemm_bf16(const void* a, int a_mn_major, int64_t lda, const void* b, int b_mn_major, int64_t ldb, int M,
                     int N, int K, int epilogue, void* c, int64_t ldc, void* c2, int64_t ldc2, const void* aux,
                     int64_t ldaux, const float* bias, int tile_n, int splits, void* stream);

/*
 * Embedding: x[m,:] = wte[idx[m],:] + wpe[m % T,:]   (fp32)                    model.py:177-179
 * idx is int64 [M] (M = B*T).  Backward accumulates (+=) into dwte / dwpe (fp32).
 */
int abcgpt_embed_fwd(const int64_t* idx, const float* wte, const float* wpe, float* x, int M, int T, int C, int V,
                     void* stream);
int abcgpt_embed_bwd(const int64_t* idx, const float* dx, float* dwte, float* dwpe, int M, int T, int C, int V,
                     void* stream);

/*
 * LayerNorm over the last dim, eps = 1e-5, optional bias                           model.py:18-27
 * fwd: x fp32 [M,C] -> y bf16 [M,C] (the GEMM operand), saves mean/rstd fp32 [M].
 *      y_f32 (optional, may be NULL) additionally receives the fp32 result.
 * bwd: dx_out(fp32) = dresid_in(fp32, may be NULL) + LN'(dy bf16); also writes a bf16 copy of dx_out
 *      (dx_bf16, may be NULL) for the next dgrad/wgrad GEMM; dweight/dbias are accumulated (+=, fp32).
 */
int abcgpt_layernorm_fwd(const float* x, const float* weight, const float* bias, void* y_bf16, float* y_f32,
                         float* mean, float* rstd, int M, int C, void* stream);
int abcgpt_layernorm_bwd(const void* dy_bf16, const float* x, const float* weight, const float* mean,
                         const float* rstd, const float* dresid_in, float* dx_out, void* dx_bf16, float* dweight,
                         float* dbias, int M, int C, void* stream);

/*
 * Causal self-attention over a packed qkv buffer [B*T, 3C] bf16 (q | k | v, heads of 64)   model.py:56-72
 * fwd: out [B*T, C] bf16 (heads re-assembled side by side), lse fp32 [B, H, T] (natural log-sum-exp of the
 *      scaled scores).  bwd: dqkv [B*T, 3C] bf16 from dout, out, lse.  T must be a multiple of 128 ... or any
 *      T <= 128 that is a multiple of 16; head size is fixed at 64.  `delta` is fp32 scratch [B, H, T].
 */
int abcgpt_attn_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, void* stream);
int abcgpt_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv,
                    int B, int T, int H, void* stream);

/*
 * Fused softmax cross-entropy over logits bf16 [M, ldl] (first V columns valid)   model.py:187
 * fwd: row_loss[m] = -log softmax(logits[m])[target[m]] (0 when target == ignore_index -1).
 * finalize (deterministic single-block reduction): loss_sum_count[0] = sum of row losses,
 *      loss_sum_count[1] = number of valid rows, loss[0] = sum / count.
 * bwd: dlogits bf16 [M, ldl] = (softmax - onehot) * grad_loss[0] / count, zero for ignored rows and for the
 *      padding columns V..ldl.
 */
int abcgpt_ce_fwd(const void* logits, int64_t ldl, const int64_t* targets, float* row_loss, int M, int V,
                  void* stream);
int abcgpt_ce_finalize(const float* row_loss, const int64_t* targets, int M, float* loss_sum_count, float* loss,
                       void* stream);
int abcgpt_ce_bwd(const void* logits, int64_t ldl, const int64_t* targets, const float* loss_sum_count,
                  const float* grad_loss, void* dlogits, int M, int V, void* stream);

/*
 * Gradient norm + clipping + AdamW over flat fp32 arenas         train.py:350-354, model.py:263-287
 * sumsq: out[0] += sum(g^2) (caller zeroes out[0]); deterministic two-stage reduction through `workspace`
 *        (>= 1024 floats) so that every data-parallel rank derives the same clip coefficient.
 * adamw: torch.optim.AdamW semantics (decoupled decay p *= 1 - lr*wd; eps outside the bias-corrected sqrt),
 *        `step` is the 1-based step count.  If sumsq != NULL the gradient is first scaled by
 *        min(1, max_norm / (sqrt(sumsq[0]) + 1e-6))  (clip_grad_norm_).  If shadow_bf16 != NULL the updated
 *        parameter is also written there rounded to bf16 (the GEMM operand copy for the next step).
 */
int abcgpt_sumsq(const float* g, int64_t n, float* out, float* workspace, void* stream);
int abcgpt_adamw(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr, float beta1,
                 float beta2, float eps, float weight_decay, int step, const float* sumsq, float max_norm,
                 void* stream);
/* y(bf16) = x(fp32), elementwise; used to (re)build the bf16 weight shadow after load_state_dict */
int abcgpt_cast_f32_to_bf16(const float* x, void* y_bf16, int64_t n, void* stream);

/* out[n] += sum_m dy[m,n]: bias gradient of an nn.Linear with bias=True (dy bf16 [M, ld], out fp32 [N]) */
int abcgpt_colsum_bf16(const void* dy, int64_t ld, int M, int N, float* out, void* stream);

/*
 * Greedy / top-k sampling head for generate()                                    model.py:316-328
 * logits bf16 [B, ldl] (last position only); writes the argmax token id (int64) per row into out[b*out_stride].
 */
int abcgpt_argmax(const void* logits, int64_t ldl, int V, int64_t* out, int64_t out_stride, int B, void* stream);

/* Debug aid: device pointer to 8 uint64 cycle counters accumulated by subsequent GEMM launches (NULL disables):
 * [0] producer empty-wait [1] MMA full-wait [2] MMA tmem-empty wait [3] epilogue tmem-full wait [5] CTA total. */
int abcgpt_debug_gemm_stats(void* device_counters);
/* Debug aid: device pointer to int64[num_kv_tiles * 8]; one CTA of the next attention forward launches stamps clock64
 * at its phase boundaries (NULL disables). */
int abcgpt_debug_attn_trace(void* device_stamps);

#ifdef __cplusplus
}
#endif
#endif /* ABCGPT_H_ */
