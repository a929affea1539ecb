/*
 * vox_oracle.h -- CPU restatement of the reference arithmetic (TEST INFRASTRUCTURE ONLY; see vox_oracle.c).
 */
#ifndef VOX_ORACLE_H
#define VOX_ORACLE_H
#include <stdint.h>

void orc_add(float *a, const float *b, int n);
void orc_mul(float *a, const float *b, int n);
void orc_silu(float *x, int n);
void orc_gelu(float *x, int n);
void orc_softmax(float *x, int rows, int cols);
void orc_linear_bf16(float *y, const float *x, const uint16_t *W, const float *b, int M, int K, int N);
void orc_rms_norm(float *out, const float *x, const float *w, int rows, int hidden, float eps);
void orc_rope_freqs(float *freqs, const int *pos, int seq, int dim, float theta);
void orc_apply_rope(float *x, const float *freqs, int seq, int heads, int head_dim);
void orc_causal_attention(float *out, const float *Q, const float *K, const float *V, int seq_q, int seq_k,
                          int n_heads, int n_kv_heads, int head_dim, float scale, int window, int q_offset);
int  orc_causal_conv1d_out_len(int length, int ks, int stride);
void orc_causal_conv1d(float *out, const float *in, const float *w, const float *bias, int cin, int cout,
                       int length, int ks, int stride);
void orc_mel_filters(float *filt);
void orc_mel_frames(float *mel, const float *padded, int n_frames);
int  orc_stream_mel(float *mel, const float *pcm, int n, int delay_tokens);
void orc_time_embedding(float *out, float t, int dim);
void orc_ada_scale(float *scale, const float *down, const float *up, const float *t_cond, int dim, int hid);
int  orc_argmax(const float *x, int n);

typedef struct {
    const uint16_t *wq, *wk, *wv, *wo, *w1, *w2, *w3;   /* bf16, row-major [out,in] */
    const float *attn_norm, *ffn_norm, *ada_scale;      /* ada_scale may be NULL */
} orc_dec_layer;
void orc_decoder_layer_step(float *x, const orc_dec_layer *L, float *kc, float *vc, int pos, int logical_pos,
                            int dim, int n_heads, int n_kv, int hd, int hidden, int window, float theta, float eps);
typedef struct {
    const uint16_t *wq, *wk, *wv, *wo, *w1, *w2, *w3;   /* bf16, row-major [out,in] */
    const float *bq, *bv, *bo, *b2;                     /* wk, w1, w3 have no bias */
    const float *attn_norm, *ffn_norm;
} orc_enc_layer;
void orc_encoder_layer(float *x, int m, const orc_enc_layer *L, float *kc, float *vc, int cache_len, int first_pos,
                       int dim, int n_heads, int hd, int hidden, int window, float theta, float eps);
void orc_conv_stem(float *out, const float *mel, int frames, const float *w0, const float *b0, const float *w1, const float *b1,
                   int mel_bins, int dim);
void orc_adapter(float *out, const float *enc, int rows, const uint16_t *w0, const uint16_t *w1, int enc_dim, int dec_dim);
void orc_stream_counts(int n_samples, int delay_tokens, int *mel_frames, int *enc_positions, int *adapter_tokens,
                       int *decoder_steps);
#endif
