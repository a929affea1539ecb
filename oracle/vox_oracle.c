/*
 * vox_oracle.c -- CPU restatement of the reference's arithmetic for the Voxtral hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the engine (voxtral.c_b200/) includes, links or calls this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load oracle/_build/liboracle.so.
 *
 * What it is: a plain, scalar C restatement (no BLAS, no SIMD, no -ffast-math) of every numeric function on
 * the path, each citing the reference lines it follows (paths are into /root/reference).  Dimensions are
 * arguments, so tests can run it at sizes that finish in milliseconds.
 *
 * How it is pinned (so that it can be trusted as a checker):
 *   - tests/test_cpu_oracle.py runs every function below against the SAME function of the unmodified reference
 *     compiled by oracle/Makefile into oracle/_ref/libvoxref.so, on seeded inputs (bit-exact where the
 *     reference is plain f32 arithmetic, 1e-6 relative where it is -ffast-math / OpenBLAS);
 *   - the decoder step (orc_decoder_layer_step + norm + logits + argmax) has no same-shaped reference function; it is pinned
 *     at the model's real dimensions against vox_decoder_forward on a hand-filled vox_ctx_t (tests/c/pin_decoder_step.c);
 *     likewise orc_encoder_layer / orc_adapter against vox_encoder_forward_incremental (two calls, cache carry) and
 *     vox_adapter_forward (tests/c/pin_encoder_adapter.c);
 *   - the streaming mel restatement is additionally checked against the mel checksums the survey recorded
 *     from the reference on samples/jfk.wav when that file is available.
 * The reference ships no numeric golden vectors of its own; model-level parity
 * (encoder / decoder / whole stream) is therefore pinned with the compiled reference itself and the traces
 * under tests/golden/ (oracle/ref_trace.c), not with this file.
 */
#include "vox_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static float bf16_to_f32(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

/* ---- K5: elementwise (voxtral_kernels.c:30-48, 369-406) ---- */
void orc_add(float *a, const float *b, int n) { for (int i = 0; i < n; i++) a[i] += b[i]; }
void orc_mul(float *a, const float *b, int n) { for (int i = 0; i < n; i++) a[i] *= b[i]; }
void orc_silu(float *x, int n) { for (int i = 0; i < n; i++) x[i] = x[i] / (1.0f + expf(-x[i])); }
void orc_gelu(float *x, int n) {
    for (int i = 0; i < n; i++) {
        float v = x[i], inner = 0.7978845608028654f * (v + 0.044715f * v * v * v);
        x[i] = 0.5f * v * (1.0f + tanhf(inner));
    }
}
void orc_softmax(float *x, int rows, int cols) {
    for (int r = 0; r < rows; r++) {
        float *row = x + (size_t)r * cols, mx = row[0], sum = 0.0f;
        for (int c = 1; c < cols; c++) if (row[c] > mx) mx = row[c];
        for (int c = 0; c < cols; c++) { row[c] = expf(row[c] - mx); sum += row[c]; }
        float inv = 1.0f / sum;
        for (int c = 0; c < cols; c++) row[c] *= inv;
    }
}

/* ---- K1: y[M,N] = x[M,K] W_bf16[N,K]^T (+b), bf16 -> f32 by <<16, f32 accumulate
 *      (voxtral_kernels.c:124-128 conversion, :154-195 matvec, :197-264 dispatch) ---- */
void orc_linear_bf16(float *y, const float *x, const uint16_t *W, const float *b, int M, int K, int N) {
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            const uint16_t *w = W + (size_t)n * K;
            const float *xr = x + (size_t)m * K;
            float s = b ? b[n] : 0.0f;
            for (int k = 0; k < K; k++) s += bf16_to_f32(w[k]) * xr[k];
            y[(size_t)m * N + n] = s;
        }
}

/* ---- K2: RMSNorm (voxtral_kernels.c:346-363) ---- */
void orc_rms_norm(float *out, const float *x, const float *w, int rows, int hidden, float eps) {
    for (int r = 0; r < rows; r++) {
        const float *xr = x + (size_t)r * hidden;
        float ss = 0.0f;
        for (int i = 0; i < hidden; i++) ss += xr[i] * xr[i];
        float inv = 1.0f / sqrtf(ss / hidden + eps);
        for (int i = 0; i < hidden; i++) out[(size_t)r * hidden + i] = xr[i] * inv * w[i];
    }
}

/* ---- K3: RoPE tables and interleaved-pair rotation (voxtral_kernels.c:488-526) ---- */
void orc_rope_freqs(float *freqs, const int *pos, int seq, int dim, float theta) {
    int half = dim / 2;
    for (int s = 0; s < seq; s++)
        for (int d = 0; d < half; d++) {
            float freq = 1.0f / powf(theta, (float)(2 * d) / (float)dim);
            float ang = (float)pos[s] * freq;
            freqs[((size_t)s * half + d) * 2] = cosf(ang);
            freqs[((size_t)s * half + d) * 2 + 1] = sinf(ang);
        }
}
void orc_apply_rope(float *x, const float *freqs, int seq, int heads, int head_dim) {
    int half = head_dim / 2;
    for (int s = 0; s < seq; s++)
        for (int h = 0; h < heads; h++) {
            float *v = x + ((size_t)s * heads + h) * head_dim;
            for (int d = 0; d < half; d++) {
                float c = freqs[((size_t)s * half + d) * 2], sn = freqs[((size_t)s * half + d) * 2 + 1];
                float x0 = v[2 * d], x1 = v[2 * d + 1];
                v[2 * d] = x0 * c - x1 * sn;
                v[2 * d + 1] = x0 * sn + x1 * c;
            }
        }
}

/* ---- K4: causal sliding-window GQA attention, online softmax in key order (voxtral_kernels.c:412-482) ---- */
void orc_causal_attention(float *out, const float *Q, const float *K, const float *V, int seq_q, int seq_k,
                          int n_heads, int n_kv_heads, int head_dim, float scale, int window, int q_offset) {
    int per = n_heads / n_kv_heads, qh = n_heads * head_dim, kvh = n_kv_heads * head_dim;
    for (int h = 0; h < n_heads; h++) {
        int kv = h / per;
        for (int i = 0; i < seq_q; i++) {
            const float *q = Q + (size_t)i * qh + h * head_dim;
            float *o = out + (size_t)i * qh + h * head_dim;
            int g = q_offset + i, k0 = 0, k1 = g + 1;
            if (window > 0 && g - window + 1 > 0) k0 = g - window + 1;
            if (k1 > seq_k) k1 = seq_k;
            float mx = -1e30f, sum = 0.0f;
            for (int d = 0; d < head_dim; d++) o[d] = 0.0f;
            for (int j = k0; j < k1; j++) {
                const float *k = K + (size_t)j * kvh + kv * head_dim, *v = V + (size_t)j * kvh + kv * head_dim;
                float s = 0.0f;
                for (int d = 0; d < head_dim; d++) s += q[d] * k[d];
                s *= scale;
                if (s > mx) {
                    float c = expf(mx - s);
                    sum = sum * c + 1.0f;
                    for (int d = 0; d < head_dim; d++) o[d] = o[d] * c + v[d];
                    mx = s;
                } else {
                    float p = expf(s - mx);
                    sum += p;
                    for (int d = 0; d < head_dim; d++) o[d] += p * v[d];
                }
            }
            if (sum > 0.0f) { float inv = 1.0f / sum; for (int d = 0; d < head_dim; d++) o[d] *= inv; }
        }
    }
}

/* ---- E1: causal conv1d on channel-major tensors: left pad = k - stride, OOB taps are zero
 *      (voxtral_kernels.c:293-340; out_len = ceil((L - k + pad)/stride + 1)) ---- */
int orc_causal_conv1d_out_len(int length, int ks, int stride) {
    float nf = ((float)length - ks + (ks - stride)) / (float)stride + 1.0f;
    int n = (int)ceilf(nf);
    return n < 0 ? 0 : n;
}
void orc_causal_conv1d(float *out, const float *in, const float *w, const float *bias, int cin, int cout,
                       int length, int ks, int stride) {
    int out_len = orc_causal_conv1d_out_len(length, ks, stride), left = ks - stride;
    for (int oc = 0; oc < cout; oc++)
        for (int ol = 0; ol < out_len; ol++) {
            float s = 0.0f;
            for (int ic = 0; ic < cin; ic++)
                for (int k = 0; k < ks; k++) {
                    int il = ol * stride - left + k;
                    if (il >= 0 && il < length) s += in[(size_t)ic * length + il] * w[((size_t)oc * cin + ic) * ks + k];
                }
            out[(size_t)oc * out_len + ol] = s + (bias ? bias[oc] : 0.0f);
        }
}

/* ---- M1/M3: mel filter bank (Slaney), DFT tables, periodic Hann (voxtral_audio.c:223-285, 528-542) ---- */
static float hz2mel(float f) {
    float m = 3.0f * f / 200.0f;
    if (f >= 1000.0f) m = 15.0f + logf(f / 1000.0f) * (27.0f / logf(6.4f));
    return m;
}
static float mel2hz(float m) {
    float f = 200.0f * m / 3.0f;
    if (m >= 15.0f) f = 1000.0f * expf((logf(6.4f) / 27.0f) * (m - 15.0f));
    return f;
}
void orc_mel_filters(float *filt /* [128][201] */) {
    enum { NM = 128, NF = 201 };
    float ff[NM + 2], fd[NM + 1];
    float lo = hz2mel(0.0f), hi = hz2mel(8000.0f);
    for (int i = 0; i < NM + 2; i++) ff[i] = mel2hz(lo + (hi - lo) * (float)i / (float)(NM + 1));
    for (int i = 0; i < NM + 1; i++) { fd[i] = ff[i + 1] - ff[i]; if (fd[i] == 0.0f) fd[i] = 1e-6f; }
    for (int m = 0; m < NM; m++) {
        float enorm = 2.0f / (ff[m + 2] - ff[m]);
        for (int k = 0; k < NF; k++) {
            float fk = (float)k * 8000.0f / (float)(NF - 1);
            float down = (fk - ff[m]) / fd[m], up = (ff[m + 2] - fk) / fd[m + 1];
            float v = fminf(down, up);
            filt[m * NF + k] = (v < 0.0f ? 0.0f : v) * enorm;
        }
    }
}

/* ---- M2: frames of an already padded signal (voxtral_audio.c:454-513): direct 201x400 DFT -> power -> mel ->
 *      max(log10(max(.,1e-10)), -6.5) -> (v+4)/4 ---- */
void orc_mel_frames(float *mel, const float *padded, int n_frames) {
    enum { NM = 128, NF = 201, NFFT = 400, HOP = 160 };
    static float filt[NM * NF], dcos[NF * NFFT], dsin[NF * NFFT], win[NFFT];
    static int ready = 0;
    if (!ready) {
        orc_mel_filters(filt);
        for (int k = 0; k < NF; k++)
            for (int n = 0; n < NFFT; n++) {
                float ang = 2.0f * (float)M_PI * (float)k * (float)n / (float)NFFT;
                dcos[k * NFFT + n] = cosf(ang); dsin[k * NFFT + n] = sinf(ang);
            }
        for (int i = 0; i < NFFT; i++) win[i] = 0.5f * (1.0f - cosf(2.0f * (float)M_PI * (float)i / (float)NFFT));
        ready = 1;
    }
    float w[NFFT], p[NF];
    for (int t = 0; t < n_frames; t++) {
        for (int i = 0; i < NFFT; i++) w[i] = padded[(size_t)t * HOP + i] * win[i];
        for (int k = 0; k < NF; k++) {
            float re = 0, im = 0;
            for (int n = 0; n < NFFT; n++) { re += w[n] * dcos[k * NFFT + n]; im += w[n] * dsin[k * NFFT + n]; }
            p[k] = re * re + im * im;
        }
        for (int m = 0; m < NM; m++) {
            float s = 0.0f;
            for (int k = 0; k < NF; k++) s += filt[m * NF + k] * p[k];
            if (s < 1e-10f) s = 1e-10f;
            float v = log10f(s);
            if (v < 1.5f - 8.0f) v = 1.5f - 8.0f;
            mel[(size_t)t * NM + m] = (v + 4.0f) / 4.0f;
        }
    }
}

/* ---- M2+M4+S1+S2: the mel the STREAM path sees for a complete recording of n samples:
 *      [200 + 32*1280 zeros | audio | align + 17*1280 zeros | reflect 200] with the last frame dropped
 *      (voxtral.c:1203,1593-1606; voxtral_audio.c:544-545,584-633).  Returns the frame count; mel may be NULL. */
int orc_stream_mel(float *mel, const float *pcm, int n, int delay_tokens) {
    int left = 200 + 32 * 1280;
    int align = (1280 - n % 1280) % 1280;
    int right = align + (delay_tokens + 1 + 10) * 1280;
    int total = left + n + right + 200;
    int frames = (total - 400) / 160 + 1 - 1;
    if (!mel) return frames;
    float *buf = calloc((size_t)total, sizeof(float));
    memcpy(buf + left, pcm, (size_t)n * sizeof(float));
    int end = left + n + right;
    for (int i = 0; i < 200; i++) buf[end + i] = buf[end - 2 - i];    /* reflection lands on the zero right pad */
    orc_mel_frames(mel, buf, frames);
    free(buf);
    return frames;
}

/* ---- D1: time conditioning (voxtral.c:31-80) ---- */
void orc_time_embedding(float *out, float t, int dim) {
    int half = dim / 2;
    float lt = logf(10000.0f);
    for (int i = 0; i < half; i++) {
        float e = t * expf(-lt * (float)i / (float)half);
        out[i] = cosf(e); out[i + half] = sinf(e);
    }
}
void orc_ada_scale(float *scale, const float *down /*[hid,dim]*/, const float *up /*[dim,hid]*/, const float *t_cond,
                   int dim, int hid) {
    float h[64];
    for (int i = 0; i < hid; i++) {
        float s = 0.0f;
        for (int j = 0; j < dim; j++) s += down[(size_t)i * dim + j] * t_cond[j];
        h[i] = s;
    }
    orc_gelu(h, hid);
    for (int i = 0; i < dim; i++) {
        float s = 0.0f;
        for (int j = 0; j < hid; j++) s += up[(size_t)i * hid + j] * h[j];
        scale[i] = s;
    }
}

/* ---- D3 tail: greedy argmax, first maximum wins (voxtral_decoder.c:697-704) ---- */
int orc_argmax(const float *x, int n) {
    int best = 0;
    for (int i = 1; i < n; i++) if (x[i] > x[best]) best = i;
    return best;
}

/* ---- D3 body at arbitrary dimensions: one token through one decoder layer against an explicit KV cache
 *      (voxtral_decoder.c:654-692).  kc/vc: [max_seq, n_kv*hd]; pos = physical index of the new row. ---- */
void orc_decoder_layer_step(float *x, const orc_dec_layer *L, float *kc, float *vc, int pos, int logical_pos,
                            int dim, int n_heads, int n_kv, int hd, int hidden, int window, float theta, float eps) {
    int qd = n_heads * hd, kvd = n_kv * hd;
    float *xn = malloc(sizeof(float) * dim), *q = malloc(sizeof(float) * qd), *att = malloc(sizeof(float) * qd);
    float *proj = malloc(sizeof(float) * dim), *g = malloc(sizeof(float) * hidden), *u = malloc(sizeof(float) * hidden);
    float *fr = malloc(sizeof(float) * hd);
    orc_rms_norm(xn, x, L->attn_norm, 1, dim, eps);
    orc_linear_bf16(q, xn, L->wq, NULL, 1, dim, qd);
    orc_linear_bf16(kc + (size_t)pos * kvd, xn, L->wk, NULL, 1, dim, kvd);
    orc_linear_bf16(vc + (size_t)pos * kvd, xn, L->wv, NULL, 1, dim, kvd);
    orc_rope_freqs(fr, &logical_pos, 1, hd, theta);
    orc_apply_rope(q, fr, 1, n_heads, hd);
    orc_apply_rope(kc + (size_t)pos * kvd, fr, 1, n_kv, hd);
    orc_causal_attention(att, q, kc, vc, 1, pos + 1, n_heads, n_kv, hd, 1.0f / sqrtf((float)hd), window, pos);
    orc_linear_bf16(proj, att, L->wo, NULL, 1, qd, dim);
    orc_add(x, proj, dim);
    orc_rms_norm(xn, x, L->ffn_norm, 1, dim, eps);
    if (L->ada_scale) for (int i = 0; i < dim; i++) xn[i] *= (1.0f + L->ada_scale[i]);
    orc_linear_bf16(g, xn, L->w1, NULL, 1, dim, hidden);
    orc_silu(g, hidden);
    orc_linear_bf16(u, xn, L->w3, NULL, 1, dim, hidden);
    orc_mul(g, u, hidden);
    orc_linear_bf16(proj, g, L->w2, NULL, 1, hidden, dim);
    orc_add(x, proj, dim);
    free(xn); free(q); free(att); free(proj); free(g); free(u); free(fr);
}

/* ---- E2 body at arbitrary dimensions: m new positions through one encoder layer against an explicit K/V cache
 *      (voxtral_encoder.c:519-619).  kc/vc: [max_rows, n_heads*hd] holding cache_len rows already; the new rows are appended.
 *      first_pos = logical position of x[0] (RoPE); attention uses physical indices, window-limited (K4).
 *      Biases: q, v, wo and w2 have one, k / w1 / w3 do not (voxtral.h:56-81). ---- */
void orc_encoder_layer(float *x, int m, const orc_enc_layer *L, float *kc, float *vc, int cache_len, int first_pos,
                       int dim, int n_heads, int hd, int hidden, int window, float theta, float eps) {
    const int qd = n_heads * hd;
    float *xn = malloc(sizeof(float) * (size_t)m * dim), *q = malloc(sizeof(float) * (size_t)m * qd);
    float *att = malloc(sizeof(float) * (size_t)m * qd), *proj = malloc(sizeof(float) * (size_t)m * dim);
    float *g = malloc(sizeof(float) * (size_t)m * hidden), *u = malloc(sizeof(float) * (size_t)m * hidden);
    float *fr = malloc(sizeof(float) * (size_t)m * hd);
    int *pos = malloc(sizeof(int) * (size_t)m);
    float *knew = kc + (size_t)cache_len * qd, *vnew = vc + (size_t)cache_len * qd;
    for (int i = 0; i < m; i++) pos[i] = first_pos + i;
    orc_rope_freqs(fr, pos, m, hd, theta);
    /* attention block */
    orc_rms_norm(xn, x, L->attn_norm, m, dim, eps);
    orc_linear_bf16(q, xn, L->wq, L->bq, m, dim, qd);
    orc_linear_bf16(knew, xn, L->wk, NULL, m, dim, qd);
    orc_linear_bf16(vnew, xn, L->wv, L->bv, m, dim, qd);
    orc_apply_rope(q, fr, m, n_heads, hd);
    orc_apply_rope(knew, fr, m, n_heads, hd);
    orc_causal_attention(att, q, kc, vc, m, cache_len + m, n_heads, n_heads, hd, 1.0f / sqrtf((float)hd), window, cache_len);
    orc_linear_bf16(proj, att, L->wo, L->bo, m, qd, dim);
    orc_add(x, proj, m * dim);
    /* gated feed-forward block */
    orc_rms_norm(xn, x, L->ffn_norm, m, dim, eps);
    orc_linear_bf16(g, xn, L->w1, NULL, m, dim, hidden);
    orc_silu(g, m * hidden);
    orc_linear_bf16(u, xn, L->w3, NULL, m, dim, hidden);
    orc_mul(g, u, m * hidden);
    orc_linear_bf16(proj, g, L->w2, L->b2, m, hidden, dim);
    orc_add(x, proj, m * dim);
    free(xn); free(q); free(att); free(proj); free(g); free(u); free(fr); free(pos);
}

/* ---- E1/E4, whole-sequence form (voxtral_encoder.c:146-185): mel [frames, bins] -> conv k3 s1 -> GELU -> conv k3 s2 -> GELU ->
 *      [ceil(frames/2), dim], both convs causal (K5 orc_causal_conv1d works channel-major, hence the two layout changes).
 *      The streaming form with its tails and odd-row residual (voxtral.c:537-715) is orchestration around the same arithmetic
 *      and is pinned at stream level by the traces under tests/golden/. ---- */
void orc_conv_stem(float *out, const float *mel, int frames, const float *w0, const float *b0, const float *w1, const float *b1,
                   int mel_bins, int dim) {
    const int l0 = orc_causal_conv1d_out_len(frames, 3, 1), l1 = orc_causal_conv1d_out_len(l0, 3, 2);
    float *cm = malloc(sizeof(float) * (size_t)mel_bins * frames);
    float *c0 = malloc(sizeof(float) * (size_t)dim * l0), *c1 = malloc(sizeof(float) * (size_t)dim * l1);
    for (int f = 0; f < frames; f++) for (int b = 0; b < mel_bins; b++) cm[(size_t)b * frames + f] = mel[(size_t)f * mel_bins + b];
    orc_causal_conv1d(c0, cm, w0, b0, mel_bins, dim, frames, 3, 1);
    orc_gelu(c0, dim * l0);
    orc_causal_conv1d(c1, c0, w1, b1, dim, dim, l0, 3, 2);
    orc_gelu(c1, dim * l1);
    for (int p = 0; p < l1; p++) for (int d = 0; d < dim; d++) out[(size_t)p * dim + d] = c1[(size_t)d * l1 + p];
    free(cm); free(c0); free(c1);
}

/* ---- E5: adapter (voxtral_encoder.c:642-674): four consecutive encoder rows are one input row (a free reshape of row-major
 *      data), Linear(4*enc_dim -> dec_dim), GELU, Linear(dec_dim -> dec_dim), no biases.  rows must be a multiple of 4. ---- */
void orc_adapter(float *out, const float *enc, int rows, const uint16_t *w0, const uint16_t *w1, int enc_dim, int dec_dim) {
    const int t = rows / 4;
    float *h = malloc(sizeof(float) * (size_t)t * dec_dim);
    orc_linear_bf16(h, enc, w0, NULL, t, 4 * enc_dim, dec_dim);
    orc_gelu(h, t * dec_dim);
    orc_linear_bf16(out, h, w1, NULL, t, dec_dim, dec_dim);
    free(h);
}

/* ---- stream bookkeeping restated as pure integer functions ---- */
void orc_stream_counts(int n_samples, int delay_tokens, int *mel_frames, int *enc_positions, int *adapter_tokens,
                       int *decoder_steps) {
    int f = orc_stream_mel(NULL, NULL, n_samples, delay_tokens);
    int p = f / 2, t = p / 4, prompt = 1 + 32 + delay_tokens;
    if (mel_frames) *mel_frames = f;
    if (enc_positions) *enc_positions = p;
    if (adapter_tokens) *adapter_tokens = t;
    if (decoder_steps) *decoder_steps = t >= prompt ? t - (prompt - 1) : 0;
}
