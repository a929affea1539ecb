/*
 * vb_ops.cuh -- device-pointer launchers for the M>1 building blocks
 * (prefill, encoder, adapter, conv stem) and shared device helpers.
 */
#ifndef VB_OPS_CUH
#define VB_OPS_CUH

#include "vb_engine.h"
#include <cuda_runtime.h>

enum VbEpilogue { VB_EPI_STORE = 0, VB_EPI_GELU = 1, VB_EPI_RESIDUAL = 2, VB_EPI_SWIGLU = 3 };

/* C[M,N] = A[M,K] * W[N,K]^T (+bias) with epilogue.  A: f32 with row pitch lda.
 * W: bf16 row-major [N,K].  SWIGLU: W rows are (gate,up) interleaved and C is [M,N/2]. */
void vb_gemm_bf16w(VbEngine *e, const float *A, int lda, const uint16_t *W, const float *bias,
                   float *C, int ldc, int M, int N, int K, int epi);
/* Same with f32 weights (host-API parity seam only). */
void vb_gemm_f32w(VbEngine *e, const float *A, int lda, const float *W, const float *bias,
                  float *C, int ldc, int M, int N, int K, int epi);

void vb_rmsnorm_rows(VbEngine *e, float *out, const float *x, const float *w, const float *ada,
                     int rows, int hidden, float eps);

/* RoPE (interleaved pairs) on the q and k column blocks of qkv[M, ldq]; k (rotated)
 * and v are scattered to kdst/vdst rows (dst_row0+m) & slot_mask (slot_mask = -1: no wrap). */
void vb_rope_split(VbEngine *e, float *qkv, int ldq, int M, int n_q_heads, int n_kv_heads, int head_dim,
                   const float *inv_freq, int pos0, float *kdst, float *vdst, int dst_row0, int slot_mask);

void vb_attention_rows(VbEngine *e, float *out, int ldo, const float *Q, int ldq, const float *K,
                       const float *V, int ldkv, int seq_q, int seq_k, int n_heads, int n_kv_heads,
                       int head_dim, float scale, int window, int q_offset);

/* vb_attn_tc.cu: the same attention on tcgen05 (head_dim 64, MHA, >= 32 queries) */
int  vb_attn_tc_enabled(void);
int  vb_attn_tc_usable(int seq_q, int seq_k, int n_heads, int n_kv_heads, int head_dim, int ldq, int ldkv, int ldo);
void vb_attention_tc(VbEngine *e, float *out, int ldo, const float *Q, int ldq, const float *K, const float *V, int ldkv,
                     int seq_q, int seq_k, int n_heads, float scale, int window, int q_offset, uint16_t *oplanes);

/* vb_gemm_tc.cu: the tcgen05 GEMM on activations that a producer already wrote as bf16 planes [3][M][K] (fused path of the
 * long encoder calls: RMSNorm, attention and SwiGLU write planes instead of f32 rows + a separate split pass) */
uint16_t *vb_attn_tc_qplanes(VbEngine *e, int seq_q, int n_heads);
uint16_t *vb_attn_tc_kplanes(VbEngine *e, int seq_k, int n_heads);
void vb_attention_tc_pre(VbEngine *e, float *out, int ldo, const float *K, const float *V, int ldkv,
                         int seq_q, int seq_k, int n_heads, float scale, int window, int q_offset, uint16_t *oplanes);
int  vb_gemm_tc_qkv_ok(int M);
void vb_gemm_tc_qkv_rope(VbEngine *e, const uint16_t *planes, const uint16_t *W, const float *bias, int M, int K,
                         const float *inv_freq, int pos0, uint16_t *qplanes, uint16_t *kplanes, float *kdst, float *vdst,
                         int dst_row0, int seq_k);
int  vb_gemm_tc_fused_ok(int M);
void vb_gemm_tc_planes(VbEngine *e, const uint16_t *planes, const uint16_t *W, const float *bias, float *C, int ldc,
                       int M, int N, int K, int epi, uint16_t *oplanes);
void vb_rmsnorm_rows_planes(VbEngine *e, uint16_t *planes, const float *x, const float *w, int rows, int hidden, float eps);

void vb_launch_count(VbEngine *e, int n);

/* ---- device helpers ---- */
__device__ __forceinline__ float vb_bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float vb_bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

/* two f32 -> one word of two bf16 (round to nearest even): a in the low half, b in the high half */
__device__ __forceinline__ uint32_t vb_pack_bf16x2(float a, float b) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}

__device__ __forceinline__ float vb_gelu_tanh(float v) {
    /* voxtral_kernels.c:376-384 */
    float x3 = v * v * v;
    float inner = 0.7978845608028654f * (v + 0.044715f * x3);
    return 0.5f * v * (1.0f + tanhf(inner));
}
__device__ __forceinline__ float vb_silu(float v) { return v / (1.0f + expf(-v)); }   /* :369-374 */

__device__ __forceinline__ float vb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif
