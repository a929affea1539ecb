/*
 * vb_encoder.cu -- audio encoder (32 causal sliding-window layers), adapter and the
 * model-block entry points of /root/reference voxtral.h:309-320 on device memory.
 *
 *   vb_encoder_layers_dev  == the layer loop of vox_encoder_forward_incremental
 *                             (voxtral_encoder.c:519-636): RMSNorm -> q/k/v (+bias on q,v)
 *                             -> RoPE theta=1e6 -> causal window-750 MHA -> wo(+bias)+res ->
 *                             RMSNorm -> SwiGLU(w1,w3) -> w2(+bias)+res ; final RMSNorm.
 *   vb_adapter_dev         == vox_adapter_forward (voxtral_encoder.c:642-674)
 *
 * Two routes through a layer, same arithmetic per element (bit-identical, tests/test_gpu_blocks_parity.py):
 *   calls of >= 512 positions  every producer writes the next tensor-core kernel's operands as bf16 planes: RMSNorm -> planes ->
 *                              wq|wk|wv GEMM whose epilogue adds the bias, applies RoPE, appends K/V to the cache and writes the
 *                              attention's Q/K planes -> tcgen05 attention -> planes -> wo -> RMSNorm -> planes -> w1|w3 with
 *                              SiLU(g)*u -> planes -> w2 (vb_gemm_tc.cu, vb_attn_tc.cu)
 *   shorter calls              one kernel per step with f32 rows in between (vb_ops.cu), the GEMMs splitting their input themselves
 *
 * Encoder KV: the reference keeps a growing f32 cache that it compacts to the last 750
 * positions before every call (voxtral_encoder.c:388-406,463-466).  Only those 750 rows can
 * ever be attended again.  HBM holds a [32][750 + 2048][2048] cache per K and V:
 *   - small calls (<= 1024 new positions: every live-stream call) APPEND their K/V rows in place
 *     and attend over [last <=750 rows | new rows] of the cache; when the cache is full the last
 *     750 rows move to the front (once per ~400 live calls) -- no per-call copies;
 *   - large one-shot calls work on a scratch [cache_len + new_len][2048] buffer whose prefix is
 *     the cached tail; afterwards the last <=750 rows become the cache content.
 */
#include "vb_ops.cuh"
#include <string.h>

#define ENC_DIM VOX_ENC_DIM
#define ENC_HID VOX_ENC_HIDDEN
#define ENC_WIN VOX_ENC_WINDOW

#define ENC_APPEND_MAX 1024                                   /* calls up to this many positions append in place */
#define ENC_CACHE_ROWS (ENC_WIN + 2048)                       /* rows per layer in the K / V cache */

static void enc_alloc_tail(VbEngine *e) {
    if (e->d_enc_tail_k) return;
    size_t bytes = (size_t)VOX_ENC_LAYERS * ENC_CACHE_ROWS * VB_ENC_ATT * sizeof(float);
    const size_t wb = e->weight_bytes;
    e->d_enc_tail_k = (float *)vb_dev_alloc_owned(e, bytes);
    e->d_enc_tail_v = (float *)vb_dev_alloc_owned(e, bytes);
    e->weight_bytes = wb;
    e->kv_bytes += 2 * bytes;
}

/* One encoder layer in two halves, so that a sequence-sharded run can exchange K/V halos in between
 * (SURVEY.md section 8e): layer-l K/V of a position depend only on that position's layer-(l-1) state, so ranks that
 * each hold a contiguous position range proceed in lock-step and rank r only needs rank r-1's last 750 K/V rows.
 *   first half : RMSNorm -> [wq|wk|wv] GEMM (+bias) -> RoPE(pos0+i) -> K,V rows written at kb/vb[row_off + i]
 *   second half: window-750 attention over kb/vb rows [0, q_off + M) with q_offset = q_off -> wo(+bias)+res -> RMSNorm ->
 *                SwiGLU -> w2(+bias)+res.  x: [M,1280] updated in place. */
extern "C" void vb_enc_layer_qkv_dev(VbEngine *e, int l, const float *x, int M, int pos0, float *kb, float *vb, int row_off) {
    const VbEncLayerDev &w = e->enc[l];
    float *qkv = vb_ws(e, 2, (size_t)M * VB_ENC_QKV * 4);
    if (vb_gemm_tc_fused_ok(M)) {                                 /* long calls: RMSNorm writes the GEMM's bf16 planes directly */
        uint16_t *xnp = (uint16_t *)vb_ws(e, VB_WS_ENC_XNP, (size_t)3 * M * ENC_DIM * 2 + 256);
        vb_rmsnorm_rows_planes(e, xnp, x, w.attn_norm, M, ENC_DIM, VOX_ENC_NORM_EPS);
        if (vb_gemm_tc_qkv_ok(M) && vb_attn_tc_enabled()) {
            /* bias + RoPE + K/V append + the attention's Q/K planes as the GEMM's epilogue: no f32 qkv, no k_rope_split */
            vb_gemm_tc_qkv_rope(e, xnp, w.wqkv, w.bqkv, M, ENC_DIM, e->d_enc_inv_freq, pos0,
                                vb_attn_tc_qplanes(e, M, VOX_ENC_HEADS), vb_attn_tc_kplanes(e, row_off + M, VOX_ENC_HEADS),
                                kb, vb, row_off, row_off + M);
            return;
        }
        vb_gemm_tc_planes(e, xnp, w.wqkv, w.bqkv, qkv, VB_ENC_QKV, M, VB_ENC_QKV, ENC_DIM, VB_EPI_STORE, nullptr);
    } else {
        float *xn  = vb_ws(e, 1, (size_t)M * ENC_DIM * 4);
        vb_rmsnorm_rows(e, xn, x, w.attn_norm, nullptr, M, ENC_DIM, VOX_ENC_NORM_EPS);
        vb_gemm_bf16w(e, xn, ENC_DIM, w.wqkv, w.bqkv, qkv, VB_ENC_QKV, M, VB_ENC_QKV, ENC_DIM, VB_EPI_STORE);
    }
    vb_rope_split(e, qkv, VB_ENC_QKV, M, VOX_ENC_HEADS, VOX_ENC_KV_HEADS, VOX_ENC_HEAD_DIM,
                  e->d_enc_inv_freq, pos0, kb, vb, row_off, -1);
}

extern "C" void vb_enc_layer_rest_dev(VbEngine *e, int l, float *x, int M, const float *kb, const float *vb, int q_off) {
    const VbEncLayerDev &w = e->enc[l];
    float *qkv = vb_ws(e, 2, (size_t)M * VB_ENC_QKV * 4);           /* q part written by the first half */
    const float scale = 1.0f / sqrtf((float)VOX_ENC_HEAD_DIM);
    if (vb_gemm_tc_fused_ok(M) && vb_attn_tc_enabled() &&
        vb_attn_tc_usable(M, q_off + M, VOX_ENC_HEADS, VOX_ENC_KV_HEADS, VOX_ENC_HEAD_DIM, VB_ENC_QKV, VB_ENC_ATT, VB_ENC_ATT)) {
        /* long calls: every producer writes the next GEMM's A operand as bf16 planes (no f32 round trip, no k_split_planes):
         * attention -> wo, RMSNorm -> w1|w3, SiLU(g)*u -> w2.  Same values as the unfused path below, bit for bit. */
        uint16_t *attp = (uint16_t *)vb_ws(e, VB_WS_ENC_ATTP, (size_t)3 * M * VB_ENC_ATT * 2 + 256);
        uint16_t *xnp  = (uint16_t *)vb_ws(e, VB_WS_ENC_XNP, (size_t)3 * M * ENC_DIM * 2 + 256);
        uint16_t *gp   = (uint16_t *)vb_ws(e, VB_WS_ENC_GP, (size_t)3 * M * ENC_HID * 2 + 256);
        if (vb_gemm_tc_qkv_ok(M))                                /* Q planes and the new rows' K planes are already there */
            vb_attention_tc_pre(e, nullptr, 0, kb, vb, VB_ENC_ATT, M, q_off + M, VOX_ENC_HEADS, scale, ENC_WIN, q_off, attp);
        else
            vb_attention_tc(e, nullptr, 0, qkv, VB_ENC_QKV, kb, vb, VB_ENC_ATT, M, q_off + M, VOX_ENC_HEADS, scale, ENC_WIN, q_off, attp);
        vb_gemm_tc_planes(e, attp, w.wo, w.bo, x, ENC_DIM, M, ENC_DIM, VB_ENC_ATT, VB_EPI_RESIDUAL, nullptr);
        vb_rmsnorm_rows_planes(e, xnp, x, w.ffn_norm, M, ENC_DIM, VOX_ENC_NORM_EPS);
        vb_gemm_tc_planes(e, xnp, w.w13, nullptr, nullptr, 0, M, 2 * ENC_HID, ENC_DIM, VB_EPI_SWIGLU, gp);
        vb_gemm_tc_planes(e, gp, w.w2, w.b2, x, ENC_DIM, M, ENC_DIM, ENC_HID, VB_EPI_RESIDUAL, nullptr);
        return;
    }
    float *xn  = vb_ws(e, 1, (size_t)M * ENC_DIM * 4);
    float *att = vb_ws(e, 3, (size_t)M * VB_ENC_ATT * 4);
    float *g   = vb_ws(e, 4, (size_t)M * ENC_HID * 4);
    vb_attention_rows(e, att, VB_ENC_ATT, qkv, VB_ENC_QKV, kb, vb, VB_ENC_ATT, M, q_off + M,
                      VOX_ENC_HEADS, VOX_ENC_KV_HEADS, VOX_ENC_HEAD_DIM, scale, ENC_WIN, q_off);
    vb_gemm_bf16w(e, att, VB_ENC_ATT, w.wo, w.bo, x, ENC_DIM, M, ENC_DIM, VB_ENC_ATT, VB_EPI_RESIDUAL);
    vb_rmsnorm_rows(e, xn, x, w.ffn_norm, nullptr, M, ENC_DIM, VOX_ENC_NORM_EPS);
    vb_gemm_bf16w(e, xn, ENC_DIM, w.w13, nullptr, g, ENC_HID, M, 2 * ENC_HID, ENC_DIM, VB_EPI_SWIGLU);
    vb_gemm_bf16w(e, g, ENC_HID, w.w2, w.b2, x, ENC_DIM, M, ENC_DIM, ENC_HID, VB_EPI_RESIDUAL);
}

/* x: [new_len,1280] device, updated in place to the encoder output (final norm applied).
 * cache_len: rows of valid tail (<=750) BEFORE this call; logical_start: RoPE position of x[0]. */
extern "C" void vb_encoder_layers_dev(VbEngine *e, float *x, int M, int cache_len, int logical_start, int update_tail) {
    if (M <= 0) return;
    enc_alloc_tail(e);
    const size_t row = (size_t)VB_ENC_ATT * sizeof(float);
    const size_t layer_rows = (size_t)ENC_CACHE_ROWS * VB_ENC_ATT;
    if (cache_len == 0 && update_tail) e->enc_tail_len = 0;   /* a fresh stream (vox_stream_init / vox_cuda_reset_caches); the cache-less full forward leaves it alone */
    int phys = e->enc_tail_len;                               /* rows valid at the front of every layer's cache */
    const int p = cache_len < phys ? cache_len : phys;        /* prior rows the window can still see (<= 750) */

    if (update_tail && M <= ENC_APPEND_MAX) {
        /* ---- append in place ---- */
        if (phys + M > ENC_CACHE_ROWS) {                      /* full: keep the last p rows (source and destination cannot overlap) */
            for (int l = 0; l < VOX_ENC_LAYERS; l++) {
                float *ck = e->d_enc_tail_k + l * layer_rows, *cv = e->d_enc_tail_v + l * layer_rows;
                VB_CUDA_OK(cudaMemcpyAsync(ck, ck + (size_t)(phys - p) * VB_ENC_ATT, p * row, cudaMemcpyDeviceToDevice, e->stream));
                VB_CUDA_OK(cudaMemcpyAsync(cv, cv + (size_t)(phys - p) * VB_ENC_ATT, p * row, cudaMemcpyDeviceToDevice, e->stream));
            }
            phys = p;
        }
        for (int l = 0; l < VOX_ENC_LAYERS; l++) {
            float *kb = e->d_enc_tail_k + l * layer_rows + (size_t)(phys - p) * VB_ENC_ATT;
            float *vb = e->d_enc_tail_v + l * layer_rows + (size_t)(phys - p) * VB_ENC_ATT;
            vb_enc_layer_qkv_dev(e, l, x, M, logical_start, kb, vb, p);      /* new rows land at cache rows [phys, phys + M) */
            vb_enc_layer_rest_dev(e, l, x, M, kb, vb, p);
        }
        e->enc_tail_len = phys + M;
    } else {
        /* ---- one-shot: scratch [p + M] rows per layer, then the last <= 750 rows become the cache ---- */
        const int total = p + M;
        float *kb  = vb_ws(e, 5, (size_t)total * VB_ENC_ATT * 4);
        float *vb  = vb_ws(e, 6, (size_t)total * VB_ENC_ATT * 4);
        const int keep = total < ENC_WIN ? total : ENC_WIN;
        for (int l = 0; l < VOX_ENC_LAYERS; l++) {
            float *tk = e->d_enc_tail_k + l * layer_rows, *tv = e->d_enc_tail_v + l * layer_rows;
            if (p > 0) {
                VB_CUDA_OK(cudaMemcpyAsync(kb, tk + (size_t)(phys - p) * VB_ENC_ATT, p * row, cudaMemcpyDeviceToDevice, e->stream));
                VB_CUDA_OK(cudaMemcpyAsync(vb, tv + (size_t)(phys - p) * VB_ENC_ATT, p * row, cudaMemcpyDeviceToDevice, e->stream));
            }
            vb_enc_layer_qkv_dev(e, l, x, M, logical_start, kb, vb, p);
            vb_enc_layer_rest_dev(e, l, x, M, kb, vb, p);
            if (!update_tail) continue;
            VB_CUDA_OK(cudaMemcpyAsync(tk, kb + (size_t)(total - keep) * VB_ENC_ATT, keep * row, cudaMemcpyDeviceToDevice, e->stream));
            VB_CUDA_OK(cudaMemcpyAsync(tv, vb + (size_t)(total - keep) * VB_ENC_ATT, keep * row, cudaMemcpyDeviceToDevice, e->stream));
        }
        if (update_tail) e->enc_tail_len = keep;
    }
    vb_rmsnorm_rows(e, x, x, e->d_enc_norm, nullptr, M, ENC_DIM, VOX_ENC_NORM_EPS);
}

/* [enc_len,1280] -> [enc_len/4,3072]: the 4x "reshape" is free on row-major data
 * ([T,5120] is the same memory), Linear(5120->3072) -> GELU -> Linear(3072->3072). */
extern "C" void vb_adapter_dev(VbEngine *e, const float *d_enc, int enc_len, float *d_out) {
    int T = enc_len / VOX_DOWNSAMPLE;
    if (T <= 0) return;
    float *mid = vb_ws(e, 7, (size_t)T * VOX_DEC_DIM * 4);
    vb_gemm_bf16w(e, d_enc, ENC_DIM * VOX_DOWNSAMPLE, e->d_adapter0, nullptr, mid, VOX_DEC_DIM, T, VOX_DEC_DIM,
                  ENC_DIM * VOX_DOWNSAMPLE, VB_EPI_GELU);
    vb_gemm_bf16w(e, mid, VOX_DEC_DIM, e->d_adapter1, nullptr, d_out, VOX_DEC_DIM, T, VOX_DEC_DIM, VOX_DEC_DIM, VB_EPI_STORE);
}

/* Causal conv as a GEMM over a strided view (see DESIGN.md "conv stem"): `in` is position-major
 * [rows, cin] and already carries the left context rows (2 for stride 1, 1 for stride 2), so
 * output j reads rows j*stride .. j*stride+2, i.e. 3*cin contiguous floats; the weights were
 * re-ordered to [cout][k][cin] at load time.  out: [n_out, cout], bias + GELU fused. */
extern "C" void vb_conv_view_dev(VbEngine *e, const float *in, int cin, int stride, int n_out,
                                 const uint16_t *w_kc, const float *bias, float *out, int cout) {
    vb_gemm_bf16w(e, in, stride * cin, w_kc, bias, out, cout, n_out, cout, 3 * cin, VB_EPI_GELU);
}

/* ---------------------------------------------------------------------------------------------
 * Host-pointer entry points (reference voxtral.h:309-320); returned buffers are malloc'd.
 * ------------------------------------------------------------------------------------------- */
extern "C" {

/* Bookkeeping identical to voxtral_encoder.c:463-477: compact to the window before appending. */
static void enc_counters_before(vox_ctx_t *c, int new_len) {
    if (c->enc_kv_cache_len + new_len > ENC_WIN && c->enc_kv_cache_len > ENC_WIN) {
        c->enc_kv_pos_offset += c->enc_kv_cache_len - ENC_WIN;
        c->enc_kv_cache_len = ENC_WIN;
    }
}

int vox_cuda_encoder_step(vox_ctx_t *ctx, float *d_x, int new_len) {
    VbEngine *e = vb_engine(ctx);
    if (new_len <= 0) return 0;
    enc_counters_before(ctx, new_len);
    int cache_len = ctx->enc_kv_cache_len;
    vb_encoder_layers_dev(e, d_x, new_len, cache_len, ctx->enc_kv_pos_offset + cache_len, 1);
    ctx->enc_kv_cache_len = cache_len + new_len;
    if (ctx->enc_kv_cache_len > ctx->enc_kv_cache_max) ctx->enc_kv_cache_max = ctx->enc_kv_cache_len;
    return 0;
}

/* ---- building blocks for a sequence-sharded encoder run (all pointers are DEVICE pointers; see tools/sharded_encode.py) ---- */
int vox_cuda_mel_conv_stem(vox_ctx_t *ctx, const float *d_pcm_padded_mel /* [F,128] mel frames */, int mel_frames, float *d_out) {
    VbEngine *e = vb_engine(ctx);
    int n = 0;
    vb_conv_stem_full_dev(e, d_pcm_padded_mel, mel_frames, d_out, &n);
    return n;
}
int vox_cuda_encoder_layer_qkv(vox_ctx_t *ctx, int layer, const float *d_x, int M, int pos0, float *d_k, float *d_v, int row_off) {
    if (layer < 0 || layer >= VOX_ENC_LAYERS || M <= 0) return -1;
    vb_enc_layer_qkv_dev(vb_engine(ctx), layer, d_x, M, pos0, d_k, d_v, row_off);
    return 0;
}
int vox_cuda_encoder_layer_rest(vox_ctx_t *ctx, int layer, float *d_x, int M, const float *d_k, const float *d_v, int q_off) {
    if (layer < 0 || layer >= VOX_ENC_LAYERS || M <= 0) return -1;
    vb_enc_layer_rest_dev(vb_engine(ctx), layer, d_x, M, d_k, d_v, q_off);
    return 0;
}
int vox_cuda_encoder_final_norm(vox_ctx_t *ctx, float *d_x, int M) {
    VbEngine *e = vb_engine(ctx);
    vb_rmsnorm_rows(e, d_x, d_x, e->d_enc_norm, nullptr, M, ENC_DIM, VOX_ENC_NORM_EPS);
    return 0;
}
int vox_cuda_adapter(vox_ctx_t *ctx, const float *d_enc, int enc_len, float *d_out) {
    vb_adapter_dev(vb_engine(ctx), d_enc, enc_len, d_out);
    return enc_len / VOX_DOWNSAMPLE;
}
void vox_cuda_sync(vox_ctx_t *ctx) { VB_CUDA_OK(cudaStreamSynchronize(vb_engine(ctx)->stream)); }

float *vox_encoder_forward_incremental(vox_ctx_t *ctx, const float *x_new, int new_len, int *out_len) {
    if (new_len <= 0) { *out_len = 0; return NULL; }
    VbEngine *e = vb_engine(ctx);
    VB_CUDA_OK(cudaSetDevice(e->device));
    size_t bytes = (size_t)new_len * ENC_DIM * 4;
    float *dx = vb_ws(e, 0, bytes);
    VB_CUDA_OK(cudaMemcpyAsync(dx, x_new, bytes, cudaMemcpyHostToDevice, e->stream));
    vox_cuda_encoder_step(ctx, dx, new_len);
    float *out = (float *)malloc(bytes);
    VB_CUDA_OK(cudaMemcpyAsync(out, dx, bytes, cudaMemcpyDeviceToHost, e->stream));
    VB_CUDA_OK(cudaStreamSynchronize(e->stream));
    *out_len = new_len;
    return out;
}

float *vox_adapter_forward(vox_ctx_t *ctx, const float *enc_out, int enc_seq_len, int *out_seq_len) {
    VbEngine *e = vb_engine(ctx);
    VB_CUDA_OK(cudaSetDevice(e->device));
    int T = enc_seq_len / VOX_DOWNSAMPLE;
    *out_seq_len = T;
    if (T <= 0) return (float *)malloc(4);
    float *din = vb_ws(e, 0, (size_t)enc_seq_len * ENC_DIM * 4);
    float *dout = vb_ws(e, 8, (size_t)T * VOX_DEC_DIM * 4);
    VB_CUDA_OK(cudaMemcpyAsync(din, enc_out, (size_t)T * 4 * ENC_DIM * 4, cudaMemcpyHostToDevice, e->stream));
    vb_adapter_dev(e, din, enc_seq_len, dout);
    float *out = (float *)malloc((size_t)T * VOX_DEC_DIM * 4);
    VB_CUDA_OK(cudaMemcpyAsync(out, dout, (size_t)T * VOX_DEC_DIM * 4, cudaMemcpyDeviceToHost, e->stream));
    VB_CUDA_OK(cudaStreamSynchronize(e->stream));
    return out;
}

/* Full (non-incremental) encoder: conv stem over the whole mel then all layers with an empty
 * cache (voxtral_encoder.c:135-312).  Odd lengths: conv1 output count is ceil(F/2) there
 * (right zero tap); handled by the stem helper in vb_stream_dev.cu. */
float *vox_encoder_forward(vox_ctx_t *ctx, const float *mel, int mel_frames, int *out_seq_len) {
    VbEngine *e = vb_engine(ctx);
    VB_CUDA_OK(cudaSetDevice(e->device));
    int P = (mel_frames + 1) / 2;
    *out_seq_len = P;
    if (P <= 0) return NULL;
    float *dmel = vb_ws(e, 9, (size_t)(mel_frames + 4) * VOX_MEL_BINS * 4);
    VB_CUDA_OK(cudaMemcpyAsync(dmel, mel, (size_t)mel_frames * VOX_MEL_BINS * 4, cudaMemcpyHostToDevice, e->stream));
    float *dx = vb_ws(e, 0, (size_t)P * ENC_DIM * 4);
    int got = 0;
    vb_conv_stem_full_dev(e, dmel, mel_frames, dx, &got);
    vb_encoder_layers_dev(e, dx, P, 0, 0, /*update_tail=*/0);   /* the reference's full forward has no cache */
    float *out = (float *)malloc((size_t)P * ENC_DIM * 4);
    VB_CUDA_OK(cudaMemcpyAsync(out, dx, (size_t)P * ENC_DIM * 4, cudaMemcpyDeviceToHost, e->stream));
    VB_CUDA_OK(cudaStreamSynchronize(e->stream));
    return out;
}

int vox_encoder_kv_cache_preallocate(vox_ctx_t *ctx, int max_pos) {
    VbEngine *e = vb_engine(ctx);
    enc_alloc_tail(e);
    if (max_pos > ctx->enc_kv_cache_max) ctx->enc_kv_cache_max = max_pos;
    return 0;
}

}  /* extern "C" */
