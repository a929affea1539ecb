/*
 * voxtral_b200.h -- C ABI of libvoxtral_b200.so, the B200 (sm_100a) engine that
 * stands behind the antirez/voxtral.c API.
 *
 * Every declaration in sections 1-6 replaces, symbol for symbol, an interface
 * of the reference (cited as file:line into /root/reference).  The structs in
 * section 1 are laid out field-for-field like the reference's public structs so
 * that a program compiled against the reference's own voxtral.h (its unchanged
 * main.c in particular) links and runs against this library;
 * tests/test_abi_layout.py checks sizeof/offsetof against the reference headers
 * whenever /root/reference is present.
 *
 * Section 7 is new: the device-resident monolithic entry points (the shape the
 * reference's Metal backend ended up with, voxtral_metal.h:219,245,254) plus
 * introspection used by tests and bench.py.
 *
 * There is no CPU fallback anywhere behind this header: if no sm_100 device
 * (or no CUDA driver) is present, vox_load() fails loudly and returns NULL and
 * the host-pointer kernel wrappers abort with a diagnostic.
 */
#ifndef VOXTRAL_B200_H
#define VOXTRAL_B200_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * 0. Model constants                      (reference voxtral.h:19-50)
 * ---------------------------------------------------------------------- */
#define VOX_SAMPLE_RATE   16000
#define VOX_MEL_BINS      128
#define VOX_HOP_LENGTH    160
#define VOX_WINDOW_SIZE   400
#define VOX_FRAME_RATE    12.5f
#define VOX_LOG_MEL_MAX   1.5f

#define VOX_ENC_DIM       1280
#define VOX_ENC_LAYERS    32
#define VOX_ENC_HEADS     32
#define VOX_ENC_KV_HEADS  32
#define VOX_ENC_HEAD_DIM  64
#define VOX_ENC_HIDDEN    5120
#define VOX_ENC_WINDOW    750
#define VOX_ENC_NORM_EPS  1e-5f

#define VOX_DOWNSAMPLE    4

#define VOX_DEC_DIM       3072
#define VOX_DEC_LAYERS    26
#define VOX_DEC_HEADS     32
#define VOX_DEC_KV_HEADS  8
#define VOX_DEC_HEAD_DIM  128
#define VOX_DEC_HIDDEN    9216
#define VOX_DEC_WINDOW    8192
#define VOX_DEC_NORM_EPS  1e-5f
#define VOX_VOCAB_SIZE    131072
#define VOX_ADA_NORM_DIM  32
#define VOX_ROPE_THETA    1000000.0f

#define VOX_MAX_ALT       4

/* ------------------------------------------------------------------------
 * 1. Public context structs               (reference voxtral.h:56-204)
 *    Host-side mirrors.  bf16 members point into the mmap'd checkpoint, small
 *    f32 members are malloc'd; every one of them also has a device copy that
 *    the engine actually computes from (looked up by host address when a
 *    caller hands one to a section-4 kernel wrapper).
 * ---------------------------------------------------------------------- */
typedef struct {
    float *wq_weight;   uint16_t *wq_weight_bf16;   float *wq_bias;
    float *wk_weight;   uint16_t *wk_weight_bf16;
    float *wv_weight;   uint16_t *wv_weight_bf16;   float *wv_bias;
    float *wo_weight;   uint16_t *wo_weight_bf16;   float *wo_bias;
    float *attention_norm;
    float *w1_weight;   uint16_t *w1_weight_bf16;
    float *w2_weight;   uint16_t *w2_weight_bf16;   float *w2_bias;
    float *w3_weight;   uint16_t *w3_weight_bf16;
    float *ffn_norm;
} vox_enc_layer_t;                          /* voxtral.h:56-82 */

typedef struct {
    float *conv0_weight, *conv0_bias;       /* [1280,128,3], [1280] */
    float *conv1_weight, *conv1_bias;       /* [1280,1280,3], [1280] */
    vox_enc_layer_t layers[VOX_ENC_LAYERS];
    float *norm;
} vox_encoder_t;                            /* voxtral.h:84-96 */

typedef struct {
    float *ada_norm_down;                   /* [32,3072] */
    float *ada_norm_up;                     /* [3072,32] */
    float *wq_weight;   uint16_t *wq_weight_bf16;
    float *wk_weight;   uint16_t *wk_weight_bf16;
    float *wv_weight;   uint16_t *wv_weight_bf16;
    float *wo_weight;   uint16_t *wo_weight_bf16;
    float *attention_norm;
    float *w1_weight;   uint16_t *w1_weight_bf16;
    float *w2_weight;   uint16_t *w2_weight_bf16;
    float *w3_weight;   uint16_t *w3_weight_bf16;
    float *ffn_norm;
} vox_dec_layer_t;                          /* voxtral.h:102-126 */

typedef struct {
    float *tok_embeddings;  uint16_t *tok_embeddings_bf16;   /* [131072,3072] */
    vox_dec_layer_t layers[VOX_DEC_LAYERS];
    float *norm;
} vox_decoder_t;                            /* voxtral.h:128-138 */

typedef struct {
    float *linear0_weight;  uint16_t *linear0_weight_bf16;   /* [3072,5120] */
    float *linear1_weight;  uint16_t *linear1_weight_bf16;   /* [3072,3072] */
} vox_adapter_t;                            /* voxtral.h:144-149 */

typedef struct {
    vox_encoder_t encoder;
    vox_adapter_t adapter;
    vox_decoder_t decoder;

    void *safetensors;
    char model_dir[512];

    /* Decoder KV cache bookkeeping.  The four pointers stay NULL in this
     * engine (the cache lives in HBM as an 8192-slot ring per layer); the
     * three counters follow the reference's compaction arithmetic
     * (voxtral_decoder.c:317-347,615-623) so callers observe the same values. */
    float *kv_cache_k, *kv_cache_v;
    uint16_t *kv_cache_k_f16, *kv_cache_v_f16;
    int kv_cache_fp16;
    int kv_cache_len;
    int kv_cache_max;
    int kv_pos_offset;

    int delay_tokens;
    float t_cond[VOX_DEC_DIM];
    float *ada_scale;                       /* host copy [26*3072] */

    int use_bf16;

    /* Encoder KV cache bookkeeping (device resident; pointers stay NULL). */
    float *enc_kv_cache_k, *enc_kv_cache_v;
    int enc_kv_cache_len;
    int enc_kv_cache_max;
    int enc_kv_cache_is_shared;
    int enc_kv_pos_offset;

    /* Reference scratch members: unused here, kept for layout. */
    int enc_inc_cap;
    float *enc_inc_x_norm, *enc_inc_q, *enc_inc_k, *enc_inc_v;
    float *enc_inc_attn_out, *enc_inc_proj_out;
    float *enc_inc_gate, *enc_inc_up, *enc_inc_ffn_out;
    int *enc_inc_positions;
    float *enc_inc_rope_freqs;
    float *dec_x, *dec_x_norm, *dec_q, *dec_k, *dec_v;
    float *dec_attn_out, *dec_proj_out;
    float *dec_gate, *dec_up, *dec_ffn_out;
    float *dec_rope_freqs;
} vox_ctx_t;                                /* voxtral.h:155-204 */

/* ------------------------------------------------------------------------
 * 2. Model + streaming API                (reference voxtral.h:217-302)
 * ---------------------------------------------------------------------- */
typedef struct vox_stream vox_stream_t;

vox_ctx_t *vox_load(const char *model_dir);                 /* voxtral.c:116  */
void vox_free(vox_ctx_t *ctx);                              /* voxtral.c:262  */
void vox_set_delay(vox_ctx_t *ctx, int delay_ms);           /* voxtral.c:1629 */

vox_stream_t *vox_stream_init(vox_ctx_t *ctx);              /* voxtral.c:1190 */
int  vox_stream_feed(vox_stream_t *s, const float *samples, int n_samples); /* :1236 */
int  vox_stream_finish(vox_stream_t *s);                    /* voxtral.c:1247 */
int  vox_stream_flush(vox_stream_t *s);                     /* voxtral.c:1588 */
int  vox_stream_get(vox_stream_t *s, const char **out_tokens, int max);     /* :1267 */
int  vox_stream_get_alt(vox_stream_t *s, const char **out_tokens,
                        int max_tokens, int n_alt);         /* voxtral.c:1288 */
void vox_stream_set_alt(vox_stream_t *s, int n_alt, float cutoff);          /* :1277 */
void vox_set_processing_interval(vox_stream_t *s, float seconds);           /* :1617 */
void vox_stream_set_continuous(vox_stream_t *s, int enable);                /* :1625 */
void vox_stream_free(vox_stream_t *s);                      /* voxtral.c:1303 */

char *vox_transcribe(vox_ctx_t *ctx, const char *wav_path);                 /* :1573 */
char *vox_transcribe_audio(vox_ctx_t *ctx, const float *samples, int n);    /* :1338 */
char *vox_transcribe_stdin(vox_ctx_t *ctx);                                 /* :1371 */

/* ------------------------------------------------------------------------
 * 3. Model-block entry points             (reference voxtral.h:309-328)
 *    Host-pointer in / host-pointer out; returned buffers are malloc'd and
 *    owned by the caller exactly as in the reference.
 * ---------------------------------------------------------------------- */
float *vox_encoder_forward(vox_ctx_t *ctx, const float *mel, int mel_frames,
                           int *out_seq_len);               /* voxtral_encoder.c:135 */
float *vox_encoder_forward_incremental(vox_ctx_t *ctx, const float *x_new,
                                       int new_len, int *out_len);          /* :452 */
float *vox_adapter_forward(vox_ctx_t *ctx, const float *enc_out,
                           int enc_seq_len, int *out_seq_len);              /* :642 */
int   vox_decoder_forward(vox_ctx_t *ctx, const float *input_embeds,
                          float *logits);                   /* voxtral_decoder.c:586 */
void  vox_decoder_prefill(vox_ctx_t *ctx, const float *input_embeds,
                          int seq_len);                     /* voxtral_decoder.c:410 */
int   vox_decoder_kv_cache_preallocate(vox_ctx_t *ctx, int max_seq);        /* :206 */
int   vox_encoder_kv_cache_preallocate(vox_ctx_t *ctx, int max_pos);        /* enc :330 */

/* ------------------------------------------------------------------------
 * 4. Kernel dispatch surface              (reference voxtral_kernels.h:18-163)
 *    Same names, argument order and semantics; host f32 row-major tensors.
 *    Each call stages its operands to HBM, runs the CUDA kernel and copies
 *    the result back -- this is the per-op parity seam, not the fast path.
 * ---------------------------------------------------------------------- */
void vox_add_inplace(float *a, const float *b, int n);
void vox_mul_inplace(float *a, const float *b, int n);
void vox_axpy(float *a, float scale, const float *b, int n);
void vox_scale(float *x, float s, int n);
void vox_copy(float *dst, const float *src, int n);
void vox_matmul(float *C, const float *A, const float *B, int M, int K, int N);
void vox_matmul_t(float *C, const float *A, const float *B, int M, int K, int N);
void vox_linear(float *y, const float *x, const float *W, const float *b,
                int seq_len, int in_dim, int out_dim);
void vox_linear_nobias(float *y, const float *x, const float *W,
                       int seq_len, int in_dim, int out_dim);
void vox_linear_nobias_bf16(float *y, const float *x, const uint16_t *W_bf16,
                            int seq_len, int in_dim, int out_dim);
void vox_linear_bf16(float *y, const float *x, const uint16_t *W_bf16,
                     const float *b, int seq_len, int in_dim, int out_dim);
void vox_matmul_t_bf16(float *C, const float *A, const uint16_t *B_bf16,
                       int M, int K, int N);
void vox_conv1d(float *out, const float *in, const float *weight, const float *bias,
                int channels_in, int channels_out, int length,
                int kernel_size, int stride, int padding);
void vox_causal_conv1d(float *out, const float *in, const float *weight, const float *bias,
                       int channels_in, int channels_out, int length,
                       int kernel_size, int stride);
void vox_rms_norm(float *out, const float *x, const float *weight,
                  int seq_len, int hidden, float eps);
void vox_silu(float *x, int n);
void vox_gelu(float *x, int n);
void vox_softmax(float *x, int rows, int cols);
void vox_causal_attention(float *out, const float *Q, const float *K, const float *V,
                          int seq_q, int seq_k, int n_heads, int n_kv_heads,
                          int head_dim, float scale, int window_size, int q_offset);
void vox_compute_rope_freqs(float *freqs, const int *pos, int seq, int dim, float theta);
void vox_apply_rope(float *x, const float *freqs, int seq, int heads, int head_dim);

extern int vox_verbose;                                     /* voxtral.c:24 */
extern int vox_monitor;                                     /* voxtral.c:25 */

/* ------------------------------------------------------------------------
 * 5. Audio front end                      (reference voxtral_audio.h:12-69)
 * ---------------------------------------------------------------------- */
extern int vox_verbose_audio;
float *vox_load_wav(const char *path, int *out_n_samples);
float *vox_parse_wav_buffer(const uint8_t *data, size_t size, int *out_n_samples);
float *vox_read_pcm_stdin(int *out_n_samples);
float *vox_mel_spectrogram(const float *samples, int n_samples, int *out_frames);

typedef struct vox_mel_ctx vox_mel_ctx_t;
vox_mel_ctx_t *vox_mel_ctx_init(int left_pad_samples);
int    vox_mel_feed(vox_mel_ctx_t *ctx, const float *samples, int n_samples);
int    vox_mel_finish(vox_mel_ctx_t *ctx, int right_pad_samples);
float *vox_mel_data(vox_mel_ctx_t *ctx, int *out_n_frames);
int    vox_mel_frame_offset(vox_mel_ctx_t *ctx);
void   vox_mel_discard_before(vox_mel_ctx_t *ctx, int keep_from_frame);
void   vox_mel_free(vox_mel_ctx_t *ctx);

/* Microphone capture (reference voxtral_mic.h:13-23): Linux stubs, as in the
 * reference's own non-Apple build (voxtral_mic_macos.c:124-142). */
int  vox_mic_start(void);
int  vox_mic_read(float *out, int max_samples);
int  vox_mic_read_available(void);
void vox_mic_stop(void);

/* ------------------------------------------------------------------------
 * 6. Host I/O utilities kept in C
 *    tokenizer  (reference voxtral_tokenizer.h:16-34)
 *    safetensors (reference voxtral_safetensors.h:17-85)
 * ---------------------------------------------------------------------- */
typedef struct vox_tokenizer vox_tokenizer_t;
vox_tokenizer_t *vox_tokenizer_load(const char *path);
void vox_tokenizer_free(vox_tokenizer_t *tok);
const char *vox_tokenizer_decode(vox_tokenizer_t *tok, int token_id);
char *vox_tokenizer_decode_seq(vox_tokenizer_t *tok, const int *tokens, int n_tokens);
int vox_tokenizer_bos(vox_tokenizer_t *tok);
int vox_tokenizer_eos(vox_tokenizer_t *tok);
int vox_tokenizer_vocab_size(vox_tokenizer_t *tok);

#define SAFETENSORS_MAX_TENSORS 1024
typedef enum {
    DTYPE_F32 = 0, DTYPE_F16 = 1, DTYPE_BF16 = 2, DTYPE_I32 = 3,
    DTYPE_I64 = 4, DTYPE_BOOL = 5, DTYPE_UNKNOWN = -1
} safetensor_dtype_t;
typedef struct {
    char name[256];
    safetensor_dtype_t dtype;
    int ndim;
    int64_t shape[8];
    size_t data_offset;
    size_t data_size;
} safetensor_t;
typedef struct {
    char *path;
    void *data;
    size_t file_size;
    size_t header_size;
    char *header_json;
    int num_tensors;
    safetensor_t tensors[SAFETENSORS_MAX_TENSORS];
} safetensors_file_t;
safetensors_file_t *safetensors_open(const char *path);
void safetensors_close(safetensors_file_t *sf);
const safetensor_t *safetensors_find(const safetensors_file_t *sf, const char *name);
const void *safetensors_data(const safetensors_file_t *sf, const safetensor_t *t);
float *safetensors_get_f32(const safetensors_file_t *sf, const safetensor_t *t);
uint16_t *safetensors_get_bf16(const safetensors_file_t *sf, const safetensor_t *t);
uint16_t *safetensors_get_bf16_direct(const safetensors_file_t *sf, const safetensor_t *t);
int safetensor_is_bf16(const safetensor_t *t);
int64_t safetensor_numel(const safetensor_t *t);
void safetensor_print(const safetensor_t *t);
void safetensors_print_all(const safetensors_file_t *sf);

/* ------------------------------------------------------------------------
 * 7. B200 additions (no reference counterpart except the Metal precedent)
 * ---------------------------------------------------------------------- */

/* Device-resident monolithic steps, mirroring vox_metal_encoder_full_step /
 * vox_metal_decoder_prefill_step / vox_metal_decoder_full_step
 * (voxtral_metal.h:219,245,254).  All pointers below are DEVICE pointers. */

/* Run `new_len` post-conv-stem positions [new_len,1280] (f32, device) through
 * the 32 encoder layers against the device encoder KV ring; result overwrites
 * x in place (final RMSNorm applied).  Returns 0, or -1 on error. */
int vox_cuda_encoder_step(vox_ctx_t *ctx, float *d_x, int new_len);

/* Prefill `n` prompt embeddings [n,3072] (f32, device). */
int vox_cuda_decoder_prefill(vox_ctx_t *ctx, const float *d_embeds, int n);

/* Greedy-decode up to n_steps tokens entirely on the device.  Step i consumes
 * d_adapter[(first_pos+i)*3072 ..] + tok_embed(previous token) exactly like
 * voxtral.c:1057-1061; prev_token seeds the first step.  Token ids are written
 * to out_tokens (HOST).  Stops early after emitting EOS (token 2), like the
 * loop at voxtral.c:1056-1093.  Returns the number of tokens written. */
int vox_cuda_decoder_steps(vox_ctx_t *ctx, const float *d_adapter, int first_pos,
                           int n_steps, int prev_token, int *out_tokens);

/* Building blocks for the sequence-sharded encoder (BASELINE.json configs[4]).  A layer is split in two
 * so that ranks holding contiguous position ranges can exchange the 750-row K/V halo between the halves:
 *   vox_cuda_encoder_layer_qkv : RMSNorm -> q|k|v (+bias) -> RoPE(pos0+i); K,V rows go to d_k/d_v[row_off + i] ([rows,2048] f32)
 *   vox_cuda_encoder_layer_rest: attention over d_k/d_v rows [0, q_off+M) with query i at row q_off+i, then wo, FFN, residuals
 * x is [M,1280] f32 on the device and is updated in place.  All work is enqueued on the ctx's stream; vox_cuda_sync waits for it. */
int  vox_cuda_mel_conv_stem(vox_ctx_t *ctx, const float *d_mel, int mel_frames, float *d_out);   /* [F,128] -> [ceil(F/2),1280] */
int  vox_cuda_encoder_layer_qkv(vox_ctx_t *ctx, int layer, const float *d_x, int M, int pos0, float *d_k, float *d_v, int row_off);
int  vox_cuda_encoder_layer_rest(vox_ctx_t *ctx, int layer, float *d_x, int M, const float *d_k, const float *d_v, int q_off);
int  vox_cuda_encoder_final_norm(vox_ctx_t *ctx, float *d_x, int M);
int  vox_cuda_adapter(vox_ctx_t *ctx, const float *d_enc, int enc_len, float *d_out);           /* -> enc_len/4 rows of 3072 */
void vox_cuda_sync(vox_ctx_t *ctx);
/* device view of an incremental mel context's frames ([n_frames,128] f32) and the prompt-embedding builder
 * (out[i] = adapter[i] + tok_embed(i == 0 ? BOS : STREAMING_PAD), voxtral.c:990-999) */
float *vox_cuda_mel_device_frames(vox_mel_ctx_t *mel, int *n_frames);
int  vox_cuda_mel_feed_zeros(vox_mel_ctx_t *mel, int n);
int  vox_cuda_build_prompt(vox_ctx_t *ctx, float *d_out, const float *d_adapter, int n);

/* Introspection for tests and bench.py */
typedef struct {
    int    device;               /* CUDA ordinal this ctx lives on */
    int    sm_count;
    int    cc_major, cc_minor;
    size_t weight_bytes_hbm;     /* bytes of checkpoint resident in HBM */
    size_t kv_bytes_hbm;
    unsigned long long kernel_launches;  /* kernels launched by this library so far */
    double last_decode_kernel_ms;        /* device time of the last vox_cuda_decoder_steps */
    int    last_decode_steps;
    double last_encoder_kernel_ms;       /* device time of the last encoder+adapter pass */
    int    last_encoder_positions;
    double last_mel_kernel_ms;
    double total_decode_kernel_ms;       /* cumulative device time inside decode steps (CUDA events) */
    long long total_decode_steps;
    double total_encoder_ms;             /* cumulative host-observed time of encoder+adapter passes */
    long long total_encoder_positions;
    double load_ms;                      /* wall time of vox_load (checkpoint -> HBM) */
    long long verify_passes, verify_tokens;   /* verify mode: weight passes spent / tokens emitted (tokens / passes = tokens per pass) */
} vox_cuda_info_t;
int vox_cuda_get_info(vox_ctx_t *ctx, vox_cuda_info_t *out);
const char *vox_cuda_version(void);
/* Decode driver: 0 = auto (5 when the device supports it, else 3, else 1), 1 = one kernel per phase replayed as a CUDA
 * graph (validation, per-phase profiling), 3 = persistent cooperative kernel with direct streaming loads (round 1),
 * 5 = persistent kernel with a decoupled TMA weight stream, dynamic row chunks and up to 8 activation columns
 * (vb_decode_v2.cu).  All produce the same tokens. */
void vox_cuda_set_decode_mode(vox_ctx_t *ctx, int mode);

/* Exact multi-token decoding of ONE stream: depth 2..8 = that many consecutive positions per weight pass,
 * the first fed with the last emitted token, the others with drafts (a successor table learned from the stream itself, else
 * "repeat the last token"); the longest prefix whose drafts were right is accepted, so the ids are those of plain greedy
 * decoding and a pass yields between 1 and depth tokens.  Pays off when the output is repetitive (streaming pad tokens,
 * recurring word pieces); costs ~10-30 % per pass otherwise.  Default 1 (off); env VOX_CUDA_VERIFY sets the default. */
void vox_cuda_set_verify_depth(vox_ctx_t *ctx, int depth);

/* ---- several streams on one GPU sharing one weight pass ----
 * vox_cuda_ctx_fork: a second context on the parent's weights (own decoder KV ring, encoder tail, scratch; shared bf16
 * matrices, time conditioning and CUDA stream).  One vox_stream_t per context, as in the reference (voxtral.c:1226-1228).
 * Call vox_set_delay on the parent before forking; vox_free() the forks before the parent.
 * vox_cuda_stream_set_deferred(s, 1): vox_stream_feed / flush / finish run mel -> encoder -> adapter only.
 * vox_cuda_streams_decode(streams, n <= 8): prefill where needed, then ONE persistent kernel advances all decoders together
 * (a weight element is read once and multiplied into n activation columns) until every stream has consumed its adapter
 * rows; tokens are queued per stream exactly as vox_stream_feed would have queued them.  Returns the number of tokens
 * generated, -1 if a stream cannot be batched (continuous mode, alternatives, another device). */
vox_ctx_t *vox_cuda_ctx_fork(vox_ctx_t *parent);
void vox_cuda_stream_set_deferred(vox_stream_t *s, int on);
int  vox_cuda_streams_decode(vox_stream_t **streams, int n);

/* Forget both KV caches (decoder ring positions and encoder tail), like a freshly loaded ctx. */
void vox_cuda_reset_caches(vox_ctx_t *ctx);

/* Same as vox_stream_feed but the PCM already lives in HBM (d_samples is a DEVICE pointer on the
 * ctx's device): the path a GPU-resident audio front end would use, and what bench.py times for the
 * "inputs resident in HBM" figure. */
int vox_cuda_stream_feed_device(vox_stream_t *s, const float *d_samples, int n_samples);

/* Minimal device-memory helpers so a host program can stage buffers for the calls above
 * without linking the CUDA runtime itself. */
void *vox_cuda_malloc(vox_ctx_t *ctx, size_t bytes);
void  vox_cuda_free(vox_ctx_t *ctx, void *d_ptr);
int   vox_cuda_memcpy_h2d(vox_ctx_t *ctx, void *d_dst, const void *h_src, size_t bytes);
int   vox_cuda_memcpy_d2h(vox_ctx_t *ctx, void *h_dst, const void *d_src, size_t bytes);

/* Device-timeline stopwatch on the ctx's stream (CUDA events): start, run any API calls, stop. */
void   vox_cuda_timer_start(vox_ctx_t *ctx);
double vox_cuda_timer_stop_ms(vox_ctx_t *ctx);

/* Test hooks: copy one decoder layer's KV ring ([8192][1024] f32 each, slot = position & 8191) or the last step's logits
 * ([131072] f32) to host memory. */
int vox_cuda_debug_copy_kv(vox_ctx_t *ctx, int layer, float *h_k, float *h_v);
int vox_cuda_debug_copy_logits(vox_ctx_t *ctx, float *h_logits);

/* ---- ONE long recording over the GPUs of a node: sequence-sharded encoder (vb_dist.c; BASELINE.json configs[4]) ----
 * One process per GPU.  rank r owns a contiguous 4-aligned range of encoder positions (vox_cuda_shard_plan); per layer it
 * sends its last 750 K/V rows to rank r+1 (ncclSend/ncclRecv on the ctx's stream, no host sync in the layer loop); adapter
 * rows are all-gathered.  Exact: same per-row arithmetic as the unsharded encoder.  The decoder does not shard
 * (autoregressive): one rank calls vox_cuda_decode_adapter.  The 128-byte NCCL id is created on one rank and handed to the
 * others by the launcher.  world == 1 needs no NCCL and runs the same code unsharded. */
int  vox_cuda_shard_plan(int n_positions, int world, int rank, int *p0, int *p1, int *halo_rows);
int  vox_cuda_dist_unique_id(void *out128);
int  vox_cuda_dist_init(vox_ctx_t *ctx, int rank, int world, const void *id128);
void vox_cuda_dist_shutdown(vox_ctx_t *ctx);
int  vox_cuda_encode_sharded(vox_ctx_t *ctx, const float *pcm, int n_samples, float **d_adapter_out, int *n_tokens,
                             int *n_positions, double *encode_ms);
int  vox_cuda_decode_adapter(vox_ctx_t *ctx, const float *d_adapter, int n_tokens, int *out_ids, int max_ids);

/* Test hook for the error paths: the n-th device allocation from now on, and every later one, fails with an out-of-memory
 * error (n < 0: off).  With it vox_load returns NULL, vox_stream_init NULL, vox_stream_feed/flush/finish -1 -- the
 * reference's error returns (voxtral.c:132-158,1199-1200,1237) instead of a process abort. */
void vox_cuda_debug_fail_alloc_after(long long n);

/* Copy the device stream state a test wants to inspect back to the host. */
int vox_cuda_stream_token_ids(vox_stream_t *s, int *out, int max);   /* all ids generated so far */
int vox_cuda_stream_counts(vox_stream_t *s, int *mel_frames, int *adapter_tokens,
                           int *decoder_steps);

#ifdef __cplusplus
}
#endif
#endif /* VOXTRAL_B200_H */
