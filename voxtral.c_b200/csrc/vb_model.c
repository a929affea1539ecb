/*
 * vb_model.c -- vox_load / vox_free / vox_set_delay: checkpoint -> HBM.
 *
 * Replaces /root/reference voxtral.c:31-80 (time conditioning), :93-256 (load), :262-349 (free),
 * voxtral_encoder.c:50-117 and voxtral_decoder.c:49-108 (tensor binding).  Host C.
 *
 * The public vox_ctx_t is filled exactly like the reference does (bf16 members point into the
 * mmap, small tensors are widened to malloc'd f32), then every tensor is copied ONCE to HBM in the
 * layouts the kernels want (vb_engine.h).  After vox_load returns, no kernel ever reads host memory.
 */
#include "vb_engine.h"

#include <math.h>
#include <string.h>
#include <sys/time.h>

#define ENC_PFX "mm_streams_embeddings.embedding_module.whisper_encoder"
#define EMB_PFX "mm_streams_embeddings.embedding_module"

static const safetensor_t *need(safetensors_file_t *sf, const char *who, const char *name) {
    const safetensor_t *t = safetensors_find(sf, name);
    if (!t) fprintf(stderr, "%s: weight not found: %s\n", who, name);
    return t;
}
static float *get_f32(safetensors_file_t *sf, const char *who, const char *name) {
    const safetensor_t *t = need(sf, who, name);
    return t ? safetensors_get_f32(sf, t) : NULL;
}
static uint16_t *get_bf16(safetensors_file_t *sf, const char *who, const char *name) {
    const safetensor_t *t = need(sf, who, name);
    return t ? safetensors_get_bf16_direct(sf, t) : NULL;
}

/* ---- device placement helpers ---- */
static void put(VbEngine *e, void *dev, const void *host, size_t bytes) {
    vb_load_copy(e, dev, host, bytes);
    vb_register_mirror(e, host, bytes, dev);
}

/* [a;b;c] stacked row-wise into one device matrix */
static uint16_t *stack3(VbEngine *e, const uint16_t *a, size_t na, const uint16_t *b, size_t nb,
                        const uint16_t *c, size_t nc) {
    uint16_t *d = vb_dev_alloc_owned(e, (na + nb + nc) * 2);
    put(e, d, a, na * 2); put(e, d + na, b, nb * 2); put(e, d + na + nb, c, nc * 2);
    return d;
}

/* rows of g and u interleaved: out[2i] = g[i], out[2i+1] = u[i] */
static uint16_t *interleave2(VbEngine *e, const uint16_t *g, const uint16_t *u, int rows, int cols) {
    size_t rb = (size_t)cols * 2;
    uint16_t *d = vb_dev_alloc_owned(e, 2 * (size_t)rows * rb);
    vb_load_copy_2d(e, d, 2 * rb, g, rb, rb, rows);                     /* the interleaved layout is built by the copy engine */
    vb_load_copy_2d(e, (uint8_t *)d + rb, 2 * rb, u, rb, rb, rows);
    return d;
}

static uint16_t f32_to_bf16_exact(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }

/* conv weight [cout][cin][3] (f32 values that came from bf16, so the narrowing is exact)
 * -> [cout][3][cin] bf16 for the strided-view GEMM */
static uint16_t *conv_reorder(VbEngine *e, const float *w, int cout, int cin) {
    size_t n = (size_t)cout * cin * 3;
    uint16_t *h = malloc(n * 2);
    for (int o = 0; o < cout; o++)
        for (int i = 0; i < cin; i++)
            for (int k = 0; k < 3; k++)
                h[((size_t)o * 3 + k) * cin + i] = f32_to_bf16_exact(w[((size_t)o * cin + i) * 3 + k]);
    uint16_t *d = vb_dev_alloc_owned(e, n * 2);
    VB_CUDA_OK(cudaMemcpy(d, h, n * 2, cudaMemcpyHostToDevice));
    free(h);
    return d;
}

/* ---- time conditioning (voxtral.c:31-80), same f32 expressions and loop order ---- */
static void time_embedding(float *out, float t) {
    const int half = VOX_DEC_DIM / 2;
    const float log_theta = logf(10000.0f);
    for (int i = 0; i < half; i++) {
        float inv_freq = expf(-log_theta * (float)i / (float)half);
        float emb = t * inv_freq;
        out[i] = cosf(emb);
        out[i + half] = sinf(emb);
    }
}
static float gelu_host(float v) {
    float inner = 0.7978845608028654f * (v + 0.044715f * v * v * v);
    return 0.5f * v * (1.0f + tanhf(inner));
}
static void update_time_conditioning(VbEngine *e) {
    vox_ctx_t *c = &e->pub;
    time_embedding(c->t_cond, (float)c->delay_tokens);
    size_t n = (size_t)VOX_DEC_LAYERS * VOX_DEC_DIM;
    if (!c->ada_scale) c->ada_scale = malloc(n * sizeof(float));
    for (int l = 0; l < VOX_DEC_LAYERS; l++) {
        const vox_dec_layer_t *L = &c->decoder.layers[l];
        float hidden[VOX_ADA_NORM_DIM];
        for (int i = 0; i < VOX_ADA_NORM_DIM; i++) {
            const float *row = L->ada_norm_down + (size_t)i * VOX_DEC_DIM;
            float s = 0.0f;
            for (int j = 0; j < VOX_DEC_DIM; j++) s += row[j] * c->t_cond[j];
            hidden[i] = gelu_host(s);
        }
        float *dst = c->ada_scale + (size_t)l * VOX_DEC_DIM;
        for (int i = 0; i < VOX_DEC_DIM; i++) {
            const float *row = L->ada_norm_up + (size_t)i * VOX_ADA_NORM_DIM;
            float s = 0.0f;
            for (int j = 0; j < VOX_ADA_NORM_DIM; j++) s += row[j] * hidden[j];
            dst[i] = s;
        }
    }
    if (!e->d_ada_scale) e->d_ada_scale = vb_dev_alloc_owned(e, n * sizeof(float));
    VB_CUDA_OK(cudaMemcpy(e->d_ada_scale, c->ada_scale, n * sizeof(float), cudaMemcpyHostToDevice));
}

static float *rope_inv_freq_dev(VbEngine *e, int head_dim) {
    int half = head_dim / 2;
    float h[128];
    for (int d = 0; d < half; d++) h[d] = 1.0f / powf(VOX_ROPE_THETA, (float)(2 * d) / (float)head_dim); /* voxtral_kernels.c:494 */
    float *dv = vb_dev_alloc_owned(e, (size_t)half * 4);
    VB_CUDA_OK(cudaMemcpy(dv, h, (size_t)half * 4, cudaMemcpyHostToDevice));
    return dv;
}

static int load_encoder(VbEngine *e, safetensors_file_t *sf) {
    vox_encoder_t *enc = &e->pub.encoder;
    char n[512];
    enc->conv0_weight = get_f32(sf, "encoder", ENC_PFX ".conv_layers.0.conv.weight");
    enc->conv0_bias   = get_f32(sf, "encoder", ENC_PFX ".conv_layers.0.conv.bias");
    enc->conv1_weight = get_f32(sf, "encoder", ENC_PFX ".conv_layers.1.conv.weight");
    enc->conv1_bias   = get_f32(sf, "encoder", ENC_PFX ".conv_layers.1.conv.bias");
    if (!enc->conv0_weight || !enc->conv1_weight || !enc->conv0_bias || !enc->conv1_bias) return -1;
    e->d_conv0_wk = conv_reorder(e, enc->conv0_weight, VOX_ENC_DIM, VOX_MEL_BINS);
    e->d_conv1_wk = conv_reorder(e, enc->conv1_weight, VOX_ENC_DIM, VOX_ENC_DIM);
    e->d_conv0_b = vb_dev_upload(e, enc->conv0_bias, VOX_ENC_DIM * 4);
    e->d_conv1_b = vb_dev_upload(e, enc->conv1_bias, VOX_ENC_DIM * 4);

    const size_t att = (size_t)VB_ENC_ATT * VOX_ENC_DIM, ffn = (size_t)VOX_ENC_HIDDEN * VOX_ENC_DIM;
    for (int i = 0; i < VOX_ENC_LAYERS; i++) {
        vox_enc_layer_t *L = &enc->layers[i];
        VbEncLayerDev *D = &e->enc[i];
#define TN(suffix) (snprintf(n, sizeof n, ENC_PFX ".transformer.layers.%d." suffix, i), n)
        L->wq_weight_bf16 = get_bf16(sf, "encoder", TN("attention.wq.weight"));
        L->wk_weight_bf16 = get_bf16(sf, "encoder", TN("attention.wk.weight"));
        L->wv_weight_bf16 = get_bf16(sf, "encoder", TN("attention.wv.weight"));
        L->wo_weight_bf16 = get_bf16(sf, "encoder", TN("attention.wo.weight"));
        L->w1_weight_bf16 = get_bf16(sf, "encoder", TN("feed_forward.w1.weight"));
        L->w2_weight_bf16 = get_bf16(sf, "encoder", TN("feed_forward.w2.weight"));
        L->w3_weight_bf16 = get_bf16(sf, "encoder", TN("feed_forward.w3.weight"));
        L->wq_bias = get_f32(sf, "encoder", TN("attention.wq.bias"));
        L->wv_bias = get_f32(sf, "encoder", TN("attention.wv.bias"));
        L->wo_bias = get_f32(sf, "encoder", TN("attention.wo.bias"));
        L->attention_norm = get_f32(sf, "encoder", TN("attention_norm.weight"));
        L->w2_bias = get_f32(sf, "encoder", TN("feed_forward.w2.bias"));
        L->ffn_norm = get_f32(sf, "encoder", TN("ffn_norm.weight"));
#undef TN
        if (!L->wq_weight_bf16 || !L->wk_weight_bf16 || !L->wv_weight_bf16 || !L->wo_weight_bf16 ||
            !L->w1_weight_bf16 || !L->w2_weight_bf16 || !L->w3_weight_bf16 || !L->wq_bias || !L->wv_bias ||
            !L->wo_bias || !L->attention_norm || !L->w2_bias || !L->ffn_norm) {
            fprintf(stderr, "encoder: failed to load layer %d weights\n", i);
            return -1;
        }
        D->wqkv = stack3(e, L->wq_weight_bf16, att, L->wk_weight_bf16, att, L->wv_weight_bf16, att);
        D->wo = vb_dev_upload(e, L->wo_weight_bf16, att * 2);
        D->w13 = interleave2(e, L->w1_weight_bf16, L->w3_weight_bf16, VOX_ENC_HIDDEN, VOX_ENC_DIM);
        D->w2 = vb_dev_upload(e, L->w2_weight_bf16, ffn * 2);
        float *bq = calloc(VB_ENC_QKV, sizeof(float));                 /* wk has no bias */
        memcpy(bq, L->wq_bias, VB_ENC_ATT * 4);
        memcpy(bq + 2 * VB_ENC_ATT, L->wv_bias, VB_ENC_ATT * 4);
        float *dbq = vb_dev_alloc_owned(e, VB_ENC_QKV * 4);
        VB_CUDA_OK(cudaMemcpy(dbq, bq, VB_ENC_QKV * 4, cudaMemcpyHostToDevice));
        free(bq);
        D->bqkv = dbq;
        D->bo = vb_dev_upload(e, L->wo_bias, VOX_ENC_DIM * 4);
        D->b2 = vb_dev_upload(e, L->w2_bias, VOX_ENC_DIM * 4);
        D->attn_norm = vb_dev_upload(e, L->attention_norm, VOX_ENC_DIM * 4);
        D->ffn_norm = vb_dev_upload(e, L->ffn_norm, VOX_ENC_DIM * 4);
        if (vox_verbose >= 2) fprintf(stderr, "  Encoder layer %d/%d loaded\n", i + 1, VOX_ENC_LAYERS);
    }
    enc->norm = get_f32(sf, "encoder", ENC_PFX ".transformer.norm.weight");
    if (!enc->norm) return -1;
    e->d_enc_norm = vb_dev_upload(e, enc->norm, VOX_ENC_DIM * 4);
    e->d_enc_inv_freq = rope_inv_freq_dev(e, VOX_ENC_HEAD_DIM);
    return 0;
}

static int load_adapter(VbEngine *e, safetensors_file_t *sf) {
    vox_adapter_t *a = &e->pub.adapter;
    a->linear0_weight_bf16 = get_bf16(sf, "adapter", EMB_PFX ".audio_language_projection.0.weight");
    a->linear1_weight_bf16 = get_bf16(sf, "adapter", EMB_PFX ".audio_language_projection.2.weight");
    if (!a->linear0_weight_bf16 || !a->linear1_weight_bf16) return -1;
    e->d_adapter0 = vb_dev_upload(e, a->linear0_weight_bf16, (size_t)VOX_DEC_DIM * VOX_ENC_DIM * VOX_DOWNSAMPLE * 2);
    e->d_adapter1 = vb_dev_upload(e, a->linear1_weight_bf16, (size_t)VOX_DEC_DIM * VOX_DEC_DIM * 2);
    return 0;
}

static int load_decoder(VbEngine *e, safetensors_file_t *sf) {
    vox_decoder_t *dec = &e->pub.decoder;
    char n[512];
    dec->tok_embeddings_bf16 = get_bf16(sf, "decoder", EMB_PFX ".tok_embeddings.weight");
    if (!dec->tok_embeddings_bf16) return -1;
    e->d_tok_emb = vb_dev_upload(e, dec->tok_embeddings_bf16, (size_t)VOX_VOCAB_SIZE * VOX_DEC_DIM * 2);
    const size_t nq = (size_t)VB_DEC_Q * VOX_DEC_DIM, nkv = (size_t)VB_DEC_KV * VOX_DEC_DIM;
    const size_t nff = (size_t)VOX_DEC_HIDDEN * VOX_DEC_DIM;
    for (int i = 0; i < VOX_DEC_LAYERS; i++) {
        vox_dec_layer_t *L = &dec->layers[i];
        VbDecLayerDev *D = &e->dec[i];
#define TN(suffix) (snprintf(n, sizeof n, "layers.%d." suffix, i), n)
        L->ada_norm_down = get_f32(sf, "decoder", TN("ada_rms_norm_t_cond.0.weight"));
        L->ada_norm_up = get_f32(sf, "decoder", TN("ada_rms_norm_t_cond.2.weight"));
        L->wq_weight_bf16 = get_bf16(sf, "decoder", TN("attention.wq.weight"));
        L->wk_weight_bf16 = get_bf16(sf, "decoder", TN("attention.wk.weight"));
        L->wv_weight_bf16 = get_bf16(sf, "decoder", TN("attention.wv.weight"));
        L->wo_weight_bf16 = get_bf16(sf, "decoder", TN("attention.wo.weight"));
        L->attention_norm = get_f32(sf, "decoder", TN("attention_norm.weight"));
        L->w1_weight_bf16 = get_bf16(sf, "decoder", TN("feed_forward.w1.weight"));
        L->w2_weight_bf16 = get_bf16(sf, "decoder", TN("feed_forward.w2.weight"));
        L->w3_weight_bf16 = get_bf16(sf, "decoder", TN("feed_forward.w3.weight"));
        L->ffn_norm = get_f32(sf, "decoder", TN("ffn_norm.weight"));
#undef TN
        if (!L->wq_weight_bf16 || !L->wk_weight_bf16 || !L->wv_weight_bf16 || !L->wo_weight_bf16 ||
            !L->w1_weight_bf16 || !L->w2_weight_bf16 || !L->w3_weight_bf16 || !L->attention_norm ||
            !L->ffn_norm || !L->ada_norm_down || !L->ada_norm_up) {
            fprintf(stderr, "decoder: failed to load layer %d\n", i);
            return -1;
        }
        D->wqkv = stack3(e, L->wq_weight_bf16, nq, L->wk_weight_bf16, nkv, L->wv_weight_bf16, nkv);
        D->wo = vb_dev_upload(e, L->wo_weight_bf16, nq * 2);
        D->w13 = interleave2(e, L->w1_weight_bf16, L->w3_weight_bf16, VOX_DEC_HIDDEN, VOX_DEC_DIM);
        D->w2 = vb_dev_upload(e, L->w2_weight_bf16, nff * 2);
        D->attn_norm = vb_dev_upload(e, L->attention_norm, VOX_DEC_DIM * 4);
        D->ffn_norm = vb_dev_upload(e, L->ffn_norm, VOX_DEC_DIM * 4);
        if (vox_verbose >= 2) fprintf(stderr, "  Decoder layer %d/%d loaded\n", i + 1, VOX_DEC_LAYERS);
    }
    dec->norm = get_f32(sf, "decoder", "norm.weight");
    if (!dec->norm) return -1;
    e->d_dec_norm = vb_dev_upload(e, dec->norm, VOX_DEC_DIM * 4);
    e->d_dec_inv_freq = rope_inv_freq_dev(e, VOX_DEC_HEAD_DIM);
    return 0;
}

vox_ctx_t *vox_load(const char *model_dir) {
    VbEngine *volatile e = calloc(1, sizeof *e);           /* volatile: read after a longjmp back into this frame */
    if (!e) return NULL;
    vox_ctx_t *volatile ctx = &e->pub;
    snprintf(ctx->model_dir, sizeof ctx->model_dir, "%s", model_dir);
    ctx->delay_tokens = 6;
    ctx->use_bf16 = 1;
    ctx->kv_cache_fp16 = 0;

    VB_API_GUARD({ fprintf(stderr, "vox_load: loading failed (device allocation or copy error)\n");
                   if (e->pin_base) { cudaStreamSynchronize(e->stream); cudaHostUnregister((void *)e->pin_base); e->pin_base = NULL; }
                   vox_free(ctx); return NULL; });
    if (vb_device_init(e) != 0) {
        fprintf(stderr, "vox_load: no usable CUDA device; refusing to load (no CPU fallback)\n");
        free(e);
        VB_API_END;
        return NULL;
    }
    char path[1024];
    snprintf(path, sizeof path, "%s/consolidated.safetensors", model_dir);
    if (vox_verbose >= 2) fprintf(stderr, "Loading model from %s\n", path);
    safetensors_file_t *sf = safetensors_open(path);
    if (!sf) {
        fprintf(stderr, "vox_load: cannot open %s\n", path);
        vb_device_shutdown(e);
        free(e);
        VB_API_END;
        return NULL;
    }
    ctx->safetensors = sf;
    if (vox_verbose >= 1) fprintf(stderr, "Loading weights...\n");
    /* Weight path: the reference "loads" by mmap and pays the page faults in the first matmuls
     * (voxtral_safetensors.c:225).  Here the mapping is registered read-only as pinned memory, so every tensor goes to HBM with
     * an asynchronous DMA at PCIe rate straight into its final layout (stacked q|k|v, row-interleaved w1|w3), one sync at the end. */
    struct timeval tv0; gettimeofday(&tv0, NULL);
    if (!getenv("VOX_CUDA_NO_HOSTREG")) {
        if (cudaHostRegister(sf->data, sf->file_size, cudaHostRegisterReadOnly) == cudaSuccess) { e->pin_base = (const uint8_t *)sf->data; e->pin_bytes = sf->file_size; }
        else cudaGetLastError();
    }
    if (load_encoder(e, sf) != 0) { fprintf(stderr, "vox_load: failed to load encoder\n"); VB_API_END; vox_free(ctx); return NULL; }
    if (load_adapter(e, sf) != 0) { fprintf(stderr, "vox_load: failed to load adapter\n"); VB_API_END; vox_free(ctx); return NULL; }
    if (load_decoder(e, sf) != 0) { fprintf(stderr, "vox_load: failed to load decoder\n"); VB_API_END; vox_free(ctx); return NULL; }
    update_time_conditioning(e);
    vb_decoder_alloc(e);
    VB_CUDA_OK(cudaStreamSynchronize(e->stream));
    if (e->pin_base) { cudaHostUnregister((void *)e->pin_base); e->pin_base = NULL; e->pin_bytes = 0; }
    { struct timeval tv1; gettimeofday(&tv1, NULL); e->load_ms = (tv1.tv_sec - tv0.tv_sec) * 1e3 + (tv1.tv_usec - tv0.tv_usec) / 1e3; }
    vb_set_default_engine(e);
    if (vox_verbose >= 1)
        fprintf(stderr, "Model loaded. (%.2f GB of weights resident in HBM on device %d, %.2f s)\n",
                (double)e->weight_bytes / 1e9, e->device, e->load_ms / 1e3);
    VB_API_END;
    return ctx;
}

/* A second context on the same weights: own decoder KV ring, encoder tail, scratch and counters; the bf16 matrices, the
 * small f32 tensors (incl. the time conditioning: call vox_set_delay on the parent BEFORE forking), the host-side public
 * tensors and the CUDA stream are the parent's.  One vox_stream_t per context, as in the reference (voxtral.c:1226-1228);
 * forks exist so that several streams can share one weight pass (vox_cuda_streams_decode).  Free forks before the parent. */
vox_ctx_t *vox_cuda_ctx_fork(vox_ctx_t *parent) {
    if (!parent) return NULL;
    VbEngine *pe = vb_engine(parent);
    VbEngine *e = malloc(sizeof *e);
    if (!e) return NULL;
    memcpy(e, pe, sizeof *e);
    e->parent = pe->parent ? pe->parent : pe;
    e->owned = NULL; e->n_owned = e->cap_owned = 0;
    e->d_kv_k = e->d_kv_v = NULL; e->kv_bytes = 0; e->d_state = NULL;
    e->d_x = e->d_q = e->d_attn_out = e->d_gate = e->d_logits = NULL;
    e->d_part_m = e->d_part_l = e->d_part_o = NULL; e->d_argmax = NULL;
    e->d_tokens = NULL; e->tokens_cap = 0; e->h_tokens_pinned = NULL; e->d_embed_in = NULL;
    e->d_mega_bar = NULL; e->step_graph_ready = 0;
    memset(&e->v2, 0, sizeof e->v2);
    memset(e->ws, 0, sizeof e->ws); memset(e->ws_bytes, 0, sizeof e->ws_bytes);
    e->d_enc_tail_k = e->d_enc_tail_v = NULL; e->enc_tail_len = 0;
    e->dist = NULL; e->d_dist_adapter = NULL; e->dist_adapter_cap = 0;
    e->launches = 0; e->last_decode_ms = e->last_encoder_ms = e->last_mel_ms = 0; e->last_decode_steps = e->last_encoder_positions = 0;
    e->total_decode_ms = e->total_encoder_ms = 0; e->total_decode_steps = e->total_encoder_positions = 0;
    vox_ctx_t *c = &e->pub;
    c->kv_cache_len = c->kv_cache_max = c->kv_pos_offset = 0;
    c->enc_kv_cache_len = c->enc_kv_cache_max = c->enc_kv_pos_offset = 0;
    if (cudaSetDevice(e->device) != cudaSuccess || cudaEventCreate(&e->ev0) != cudaSuccess || cudaEventCreate(&e->ev1) != cudaSuccess ||
        cudaEventCreate(&e->ev_user0) != cudaSuccess || cudaEventCreate(&e->ev_user1) != cudaSuccess) { free(e); return NULL; }
    const size_t wb = e->weight_bytes;
    vb_decoder_alloc(e);
    e->weight_bytes = wb;
    return c;
}

void vox_free(vox_ctx_t *ctx) {
    if (!ctx) return;
    VbEngine *e = vb_engine(ctx);
    if (e->parent) {                                       /* a fork owns only its KV / scratch / events */
        vb_decoder_free(e);
        vb_device_shutdown(e);
        free(e);
        return;
    }
#define FREE0(p) do { free(p); (p) = NULL; } while (0)
    FREE0(ctx->encoder.conv0_weight); FREE0(ctx->encoder.conv0_bias);
    FREE0(ctx->encoder.conv1_weight); FREE0(ctx->encoder.conv1_bias);
    for (int i = 0; i < VOX_ENC_LAYERS; i++) {
        vox_enc_layer_t *L = &ctx->encoder.layers[i];
        FREE0(L->wq_bias); FREE0(L->wv_bias); FREE0(L->wo_bias);
        FREE0(L->attention_norm); FREE0(L->w2_bias); FREE0(L->ffn_norm);
    }
    FREE0(ctx->encoder.norm);
    for (int i = 0; i < VOX_DEC_LAYERS; i++) {
        vox_dec_layer_t *L = &ctx->decoder.layers[i];
        FREE0(L->ada_norm_down); FREE0(L->ada_norm_up); FREE0(L->attention_norm); FREE0(L->ffn_norm);
    }
    FREE0(ctx->decoder.norm);
    FREE0(ctx->ada_scale);
#undef FREE0
    vb_decoder_free(e);
    vox_cuda_dist_shutdown(ctx);
    if (e->pin_base) { cudaStreamSynchronize(e->stream); cudaHostUnregister((void *)e->pin_base); e->pin_base = NULL; }   /* a failed vox_load */
    vb_device_shutdown(e);
    if (ctx->safetensors) safetensors_close((safetensors_file_t *)ctx->safetensors);
    free(e);
}

void vox_set_delay(vox_ctx_t *ctx, int delay_ms) {
    if (delay_ms < 80) delay_ms = 80;
    if (delay_ms > 2400) delay_ms = 2400;
    ctx->delay_tokens = delay_ms / 80;                                 /* 1 token = 80 ms */
    VbEngine *e = vb_engine(ctx);
    vb_sync(e);
    update_time_conditioning(e);
}
