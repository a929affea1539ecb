/*
 * vb_engine.h -- internal layout of the B200 engine (not part of the C ABI).
 *
 * One VbEngine per vox_ctx_t; the public struct is the first member so the
 * vox_ctx_t* handed to callers is also the engine pointer (the reference's
 * Metal backend keys its GPU state off the same struct, voxtral_metal.m:111-147).
 *
 * HBM layout (all allocations are made once in vox_load / first use):
 *   weights   bf16, row-major [out,in] exactly as in the checkpoint, except
 *             - decoder/encoder wq|wk|wv are stored back to back as one
 *               [q+k+v, in] matrix (one GEMV/GEMM instead of three),
 *             - w1|w3 are row-interleaved (g0,u0,g1,u1,...) so SiLU(g)*u is an
 *               epilogue of the producing kernel.
 *   small f32 tensors (norms, biases, conv weights, ada_scale, RoPE inv_freq)
 *   decoder KV   f32 ring  [26][8192][1024] x {K,V}   (slot = position & 8191)
 *   encoder KV   f32 tail  [32][750][2048]  x {K,V}   (last 750 positions per layer)
 *   activations  f32, sized on demand for the largest M seen.
 */
#ifndef VB_ENGINE_H
#define VB_ENGINE_H

#include "voxtral_b200.h"

#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define VB_ENC_QKV   (3 * VOX_ENC_HEADS * VOX_ENC_HEAD_DIM)            /* 6144 */
#define VB_ENC_ATT   (VOX_ENC_HEADS * VOX_ENC_HEAD_DIM)                /* 2048 */
#define VB_DEC_Q     (VOX_DEC_HEADS * VOX_DEC_HEAD_DIM)                /* 4096 */
#define VB_DEC_KV    (VOX_DEC_KV_HEADS * VOX_DEC_HEAD_DIM)             /* 1024 */
#define VB_DEC_QKV   (VB_DEC_Q + 2 * VB_DEC_KV)                        /* 6144 */
#define VB_KV_SLOTS  VOX_DEC_WINDOW                                    /* 8192 */
#define VB_TOKEN_EOS 2
#define VB_WS_SLOTS  32   /* 0-11: model blocks (see vb_encoder.cu), 12-19: stream pipeline (vb_stream.c), then: */
#define VB_WS_ALT    20   /* 8 floats: result of the alternatives kernel (vb_decode.cu) */
#define VB_WS_DIST_X 21   /* 21-23: sharded encoder (vb_dist.c): own rows, K and V with halo */
#define VB_WS_GEMM_PLANES 24  /* bf16 split planes of the tcgen05 GEMM (vb_gemm_tc.cu) */
#define VB_WS_ATT_QP 25      /* 25-27: bf16 planes of Q, K and V^T for the tcgen05 attention (vb_attn_tc.cu) */
#define VB_WS_ATT_KP 26
#define VB_WS_ATT_VT 27
#define VB_WS_ENC_XNP 28     /* 28-30: bf16 planes written by the fused encoder producers (RMSNorm, attention, SwiGLU) */
#define VB_WS_ENC_ATTP 29
#define VB_WS_ENC_GP 30
#define VB_WS_ENC_ROPE 31    /* [M][32] (cos, sin) table of the fused wq|wk|wv epilogue (vb_gemm_tc.cu) */

/* Error boundary.  Inside the library a failed CUDA call (cudaMalloc out of memory, a launch error...) reports through
 * vb_cuda_fail().  Public entry points that have an error return in the reference -- vox_load -> NULL (voxtral.c:132-158),
 * vox_stream_init -> NULL, vox_stream_feed/flush/finish -> -1, vox_decoder_forward -> 2 = EOS on out-of-memory
 * (voxtral_decoder.c:621,649), the malloc'ing vox_encoder_forward* / vox_adapter_forward / vox_transcribe* -> NULL -- open a
 * guard (VB_API_GUARD): a failure below them unwinds to the guard with longjmp and becomes that return value.  Entry points
 * without an error return (the void kernel-surface wrappers) still abort() with a message, as does any failure outside a guard.
 * Host memory held by the interrupted call is leaked; device state of the ctx/stream is undefined afterwards and the stream
 * is marked failed. */
#include <setjmp.h>
#ifdef __cplusplus
extern "C" {
#endif
extern __thread jmp_buf *vb_err_jmp;
void vb_cuda_fail(cudaError_t err, const char *file, int line);
#ifdef __cplusplus
}
#endif
#define VB_CUDA_OK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) vb_cuda_fail(e__, __FILE__, __LINE__); } while (0)
#define VB_FAIL(msg) do { fprintf(stderr, "voxtral_b200: %s\n", msg); vb_cuda_fail(cudaErrorUnknown, __FILE__, __LINE__); } while (0)
/* usage:  VB_API_GUARD({ cleanup; return -1; });  ...body...  VB_API_END;  (the block runs after a failure below) */
#define VB_API_GUARD(on_fail) jmp_buf vb_jb__; jmp_buf *vb_prev__ = vb_err_jmp;                              \
    if (setjmp(vb_jb__)) { vb_err_jmp = vb_prev__; cudaGetLastError(); on_fail }                               \
    vb_err_jmp = &vb_jb__
#define VB_API_END vb_err_jmp = vb_prev__

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    const uint16_t *wqkv, *wo, *w13, *w2;       /* bf16 device */
    const float *attn_norm, *ffn_norm;          /* f32 device  */
} VbDecLayerDev;

typedef struct {
    const uint16_t *wqkv, *wo, *w13, *w2;
    const float *bqkv;                          /* [6144] = [wq_bias | 0 | wv_bias] */
    const float *bo, *b2;
    const float *attn_norm, *ffn_norm;
} VbEncLayerDev;

/* Device-side autoregressive state: lives in HBM, advanced by the decode kernels. */
typedef struct {
    int pos;            /* logical position of the token about to be processed */
    int token;          /* previous token id (input of the next step) */
    int eos;            /* set once EOS was emitted: later steps are no-ops */
    int n_out;          /* tokens written to out_tokens so far in this call */
    int adapter_row;    /* row of d_adapter consumed by the next step */
    int pad[3];
} VbDecState;

/* scratch of the batched persistent decode kernel (vb_decode_v2.cu), owned by the engine that launches it */
typedef struct {
    float *x, *q, *attn_out, *gate, *part_m, *part_l, *part_o, *logits_extra;
    unsigned long long *argmax;
    unsigned int *bar, *ctr;
    int *err_host, *err_dev;                    /* pinned mapped word the kernel's wait guards report into */
    VbDecState *st;
    long long *prof;
} VbV2Scratch;

typedef struct VbHostMirror { const void *host; size_t bytes; void *dev; } VbHostMirror;

typedef struct VbEngine {
    vox_ctx_t pub;                              /* MUST be first */

    int device, sm_count, cc_major, cc_minor;
    struct VbEngine *parent;                    /* vox_cuda_ctx_fork: weights, host tensors and the CUDA stream belong to the parent */
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    cudaEvent_t ev_user0, ev_user1;             /* vox_cuda_timer_* */

    /* ---- weights ---- */
    uint16_t *d_tok_emb;
    VbDecLayerDev dec[VOX_DEC_LAYERS];
    float *d_dec_norm;
    float *d_ada_scale;                         /* [26][3072] */
    float *d_dec_inv_freq;                      /* [64]  */
    VbEncLayerDev enc[VOX_ENC_LAYERS];
    float *d_enc_norm;
    float *d_enc_inv_freq;                      /* [32]  */
    uint16_t *d_conv0_wk, *d_conv1_wk;          /* conv weights re-ordered to [cout][k][cin], bf16 (exact) */
    float *d_conv0_b, *d_conv1_b;
    uint16_t *d_adapter0, *d_adapter1;
    size_t weight_bytes;

    VbHostMirror *mirrors; int n_mirrors, cap_mirrors;
    void **owned; int n_owned, cap_owned;       /* device allocations freed in vox_free */

    /* ---- decoder state ---- */
    float *d_kv_k, *d_kv_v;                     /* [26][8192][1024] */
    size_t kv_bytes;
    VbDecState *d_state;
    float *d_x, *d_q, *d_attn_out, *d_gate, *d_logits;
    float *d_part_m, *d_part_l, *d_part_o;      /* split-S attention partials */
    unsigned long long *d_argmax;               /* per-CTA packed (value,index) */
    int *d_tokens;  int tokens_cap;
    int *h_tokens_pinned;
    float *d_embed_in;                          /* [3072] staging for the host-pointer API */
    unsigned int *d_mega_bar;                   /* grid-barrier counter + error word of the persistent kernel */
    int decode_mode;                            /* 0 = auto, 1 = CUDA-graph phases, 3 = persistent kernel with direct loads (round 1), 5 = v2 persistent kernel */
    cudaGraphExec_t step_graph;                 /* one decode step, device-state driven */
    int step_graph_ready;
    VbV2Scratch v2; int v2_checked, v2_ok;
    int persist_checked, persist_ok;            /* vb_decoder_persist_supported, cached per engine */
    int verify_depth;                           /* > 1: exact multi-token decoding with that many positions per weight pass (vb_decode_v2.cu) */
    long long verify_passes, verify_tokens;     /* weight passes spent / tokens emitted in verify mode */
    const uint8_t *pin_base; size_t pin_bytes;  /* vox_load: the mmap'd checkpoint while it is registered as pinned memory (async H2D) */
    double load_ms;                             /* wall time of vox_load */
    void *dist;                                 /* VbDist* (vb_dist.c): NCCL communicator of the sequence-sharded encoder */
    float *d_dist_adapter; int dist_adapter_cap;   /* gathered adapter rows of vox_cuda_encode_sharded */

    /* ---- M>1 scratch (prefill / encoder / adapter) ---- */
    float *ws[VB_WS_SLOTS]; size_t ws_bytes[VB_WS_SLOTS];         /* grow-on-demand workspaces */
    float *d_enc_tail_k, *d_enc_tail_v;         /* [32][750 + 2048][2048]: the encoder K / V cache (vb_encoder.cu) */
    int enc_tail_len;                           /* rows valid at the front of every layer's cache */

    /* ---- statistics ---- */
    unsigned long long launches;
    double last_decode_ms; int last_decode_steps;
    double last_encoder_ms; int last_encoder_positions;
    double last_mel_ms;
    double total_decode_ms; long long total_decode_steps;
    double total_encoder_ms; long long total_encoder_positions;
} VbEngine;

static inline VbEngine *vb_engine(vox_ctx_t *ctx) { return (VbEngine *)ctx; }

/* vb_runtime.cu */
int   vb_device_init(VbEngine *e);              /* picks the device, creates the stream; -1 if no sm_100 GPU */
void  vb_device_shutdown(VbEngine *e);
void *vb_dev_alloc(size_t bytes);
void *vb_dev_alloc_owned(VbEngine *e, size_t bytes);                    /* freed at shutdown */
void *vb_dev_upload(VbEngine *e, const void *host, size_t bytes);       /* alloc + H2D + register mirror */
void  vb_load_copy(VbEngine *e, void *dev, const void *host, size_t bytes);
void  vb_load_copy_2d(VbEngine *e, void *dev, size_t dpitch, const void *host, size_t spitch, size_t width, size_t height);
void  vb_register_mirror(VbEngine *e, const void *host, size_t bytes, void *dev);
void *vb_find_mirror(VbEngine *e, const void *host);
float *vb_ws(VbEngine *e, int slot, size_t bytes);                      /* workspace, grown on demand */
void  vb_require_gpu(const char *what);         /* aborts loudly if the CUDA device is missing */
VbEngine *vb_default_engine(void);              /* engine used by the host-pointer kernel wrappers */
void  vb_set_default_engine(VbEngine *e);

/* vb_decode.cu */
int  vb_decoder_alloc(VbEngine *e);
void vb_decoder_free(VbEngine *e);
void vb_decoder_reset(VbEngine *e);
void vb_decoder_set_state(VbEngine *e, int pos, int token, int adapter_row);
int  vb_decoder_run_steps(VbEngine *e, const float *d_adapter, int adapter_row, int n_steps,
                          int prev_token, int pos, int *out_tokens_host);
int  vb_decoder_step_from_embed(VbEngine *e, const float *d_embed, int pos, float *logits_host);
void vb_decoder_prefill_dev(VbEngine *e, const float *d_embeds, int n, int start_pos);

/* vb_decode_persist.cu */
void vb_alt_candidates(VbEngine *e, int best, int text_min, float *z, float ev[3], int idx[3]);
int  vb_decoder_persist_supported(VbEngine *e);
int  vb_decoder_persist_launch(VbEngine *e, const float *d_adapter, int adapter_row, int n_steps, int prev_token, int pos);

/* vb_decode_v2.cu: one weight pass for up to 8 columns (independent streams, or drafted positions of one stream) */
typedef struct { struct VbEngine *engine; const float *d_adapter; int adapter_row, n_steps, prev_token, pos; } VbV2Col;
int  vb_decoder_v2_supported(VbEngine *e);
int  vb_decoder_v2_launch(VbEngine *lead, const VbV2Col *cols, int nb, int n_steps, int verify, VbDecState *st_host);

/* vb_encoder.cu */
void vb_encoder_layers_dev(VbEngine *e, float *d_x, int new_len, int cache_len, int logical_start, int update_tail);
void vb_adapter_dev(VbEngine *e, const float *d_enc, int enc_len, float *d_out);
void vb_enc_layer_qkv_dev(VbEngine *e, int l, const float *x, int M, int pos0, float *kb, float *vb, int row_off);
void vb_enc_layer_rest_dev(VbEngine *e, int l, float *x, int M, const float *kb, const float *vb, int q_off);
void vb_conv_view_dev(VbEngine *e, const float *in, int cin, int stride, int n_out,
                      const uint16_t *w_kc, const float *bias, float *out, int cout);

/* vb_mel.cu */
vox_mel_ctx_t *vb_mel_ctx_init_on(VbEngine *e, int left_pad_samples);
int    vb_mel_feed_zeros(vox_mel_ctx_t *c, int n);
int    vb_mel_feed_device(vox_mel_ctx_t *c, const float *d_samples, int n);
float *vb_mel_dev_frames(vox_mel_ctx_t *c, int *n_frames, int *frame_offset);
int    vb_mel_recording_frames(int n_samples, int delay_tokens);
void   vb_mel_recording_range(VbEngine *e, const float *pcm_host, int n_samples, int f0, int f1, float *d_out);

/* vb_stream_dev.cu */
void vb_d2d(VbEngine *e, void *dst, const void *src, size_t bytes);
void vb_dzero(VbEngine *e, void *dst, size_t bytes);
void vb_h2d(VbEngine *e, void *dst, const void *src, size_t bytes);
void vb_d2h_sync(VbEngine *e, void *dst, const void *src, size_t bytes);
void vb_sync(VbEngine *e);
void vb_build_prompt_dev(VbEngine *e, float *d_out, const float *d_adapter, int n, int bos, int pad);
void vb_conv_stem_full_dev(VbEngine *e, const float *d_mel, int mel_frames, float *d_out, int *out_len);
void vb_conv_stem_range_dev(VbEngine *e, const float *d_mel, int mel_first, int mel_frames, int p0, int p1, float *d_out);
void vb_gemv_bf16_dev(VbEngine *e, float *y, const float *x, const uint16_t *W, const float *bias, int K, int N);
int  vb_gemv_cols_dev(VbEngine *e, const float *A, int lda, const uint16_t *W, const float *bias, float *C, int ldc, int M, int N, int K, int epi);

#ifdef __cplusplus
}
#endif
#endif
