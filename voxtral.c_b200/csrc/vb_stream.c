/*
 * vb_stream.c -- the vox_stream_t state machine: feed -> [mel] -> [conv stem] -> [encoder] ->
 * [adapter] -> [decoder] -> token queue.  Host C; every tensor it moves around lives in HBM.
 *
 * Replaces /root/reference voxtral.c:352-1635.  The control flow (when each block runs, with how
 * many rows, and what is carried between calls) follows the reference decision by decision, because
 * any off-by-one shifts every later token (SURVEY.md section 7 hard part 5):
 *   S1 left pad 32*1280 zeros (+200)            voxtral.c:1203
 *   S2 right pad align + 17*1280 zeros          voxtral.c:1593-1606
 *   S3 prompt = BOS + 38 x STREAMING_PAD        voxtral.c:972,1005-1012
 *   S4 encoder gate 312 / interval*100 mel      voxtral.c:793-795
 *   conv stem tails / odd-frame residual         voxtral.c:537-715
 *   4x adapter alignment residual                voxtral.c:824-890
 *   continuous-mode restart policy               voxtral.c:1137-1187
 * What differs is where the data is: mel frames, conv tails, encoder residual, the adapter buffer,
 * the prompt embeddings and the autoregressive token feedback never leave the GPU; the host sees
 * only counts and, once per call, the token ids that were generated.
 */
#include "vb_engine.h"

#include <ctype.h>
#include <math.h>
#include <string.h>
#include <sys/time.h>

#define TOKEN_BOS            1
#define TOKEN_EOS            2
#define TOKEN_STREAMING_PAD  32
#define TOKEN_TEXT_MIN       1000
#define SAMPLES_PER_TOKEN    1280
#define OFFLINE_BUFFER_TOKENS 10
#define FIRST_CHUNK_MIN_MEL  312
#define DEFAULT_INTERVAL_S   2.0f
#define MAX_DECODE_KV        2000
#define MAX_NON_TEXT_STREAK  64
#define MAX_NO_DECODE_SAMPLES (VOX_SAMPLE_RATE * 20)
#define EMPTY_RESTARTS_FOR_FULL_RESET 2

/* workspace slots owned by the stream pipeline (vb_engine.h: 12-23) */
enum { WS_CONV0_IN = 12, WS_CONV0_OUT = 13, WS_CONV1_IN = 14, WS_CONV1_OUT = 15, WS_COMBINED = 16,
       WS_PROMPT = 17, WS_ADAPTER_CHUNK = 18, WS_TMP = 19 };

struct vox_stream {
    vox_ctx_t *ctx;
    VbEngine *e;
    vox_tokenizer_t *tokenizer;

    vox_mel_ctx_t *mel;
    int64_t real_samples_fed;
    int mel_cursor;

    /* conv stem carry (device) */
    float *d_mel_tail;        /* [2][128]  */
    float *d_conv0_tail;      /* [1280] last conv0 row handed to conv1 */
    float *d_conv0_resid;     /* [1280] odd conv0 row waiting for a partner */
    int conv0_residual_count, conv_stem_initialized;
    float *d_enc_resid;       /* [3][1280] encoder rows waiting for 4x alignment */
    int enc_residual_count;

    /* adapter rows (device, growing) */
    float *d_adapter;
    int total_adapter, adapter_cap, adapter_pos_offset;

    /* decoder bookkeeping */
    int decoder_started, gen_pos, prev_token, eos_seen;
    int nontext_streak, text_since_restart, empty_restarts, waiting_prompt;
    int64_t last_decode_sample;
    int finished, continuous;
    int failed;               /* a CUDA/allocation failure unwound out of this stream: every later call returns -1 */
    int defer_decode;         /* vox_cuda_stream_set_deferred: feed/flush/finish stop after the adapter; vox_cuda_streams_decode runs the decoder */

    const char **token_queue;
    int queue_head, queue_tail, queue_cap;
    int n_alt; float alt_cutoff;

    int min_new_mel;
    double encoder_ms, decoder_ms, prefill_ms;
    int n_generated, n_text_tokens;

    int *ids; int n_ids, cap_ids;       /* every generated id, for introspection */
    int *tok_buf; int tok_buf_cap;
};

/* realloc that reports through the error boundary instead of dereferencing NULL later */
static void *xrealloc(void *p, size_t bytes) {
    void *q = realloc(p, bytes);
    if (!q) { fprintf(stderr, "voxtral_b200: out of host memory (%zu bytes)\n", bytes); vb_cuda_fail(cudaErrorMemoryAllocation, __FILE__, __LINE__); }
    return q;
}

static double now_ms(void) {
    struct timeval tv; gettimeofday(&tv, NULL);
    return tv.tv_sec * 1000.0 + tv.tv_usec / 1000.0;
}

/* ---------------------------------------------------------------- token queue */
typedef enum { TOK_TEXT, TOK_CONTROL, TOK_INVALID, TOK_EOS } tok_class;

static tok_class classify(vox_stream_t *s, int id) {
    if (id == TOKEN_EOS) return TOK_EOS;
    if (id < TOKEN_TEXT_MIN) return TOK_CONTROL;
    const char *p = vox_tokenizer_decode(s->tokenizer, id);
    return (!p || !p[0]) ? TOK_INVALID : TOK_TEXT;
}

static void enqueue(vox_stream_t *s, const char *alts[VOX_MAX_ALT]) {
    int next = (s->queue_tail + 1) % s->queue_cap;
    if (next == s->queue_head) {
        int ncap = s->queue_cap * 2, n = 0;
        const char **nq = calloc((size_t)ncap * VOX_MAX_ALT, sizeof *nq);
        if (!nq) return;
        for (int i = s->queue_head; i != s->queue_tail; i = (i + 1) % s->queue_cap, n++)
            memcpy(&nq[n * VOX_MAX_ALT], &s->token_queue[i * VOX_MAX_ALT], VOX_MAX_ALT * sizeof *nq);
        free(s->token_queue);
        s->token_queue = nq; s->queue_head = 0; s->queue_tail = n; s->queue_cap = ncap;
        next = (s->queue_tail + 1) % s->queue_cap;
    }
    memcpy(&s->token_queue[s->queue_tail * VOX_MAX_ALT], alts, VOX_MAX_ALT * sizeof *alts);
    s->queue_tail = next;
}

/* Alternatives for a text position (voxtral.c:911-966): p_i = exp(l_i - l_max) / Z; the runner-up text tokens, best first,
 * are kept while 1 - p_i / p_best <= cutoff.  Z and the three candidates come from a device kernel over the step's logits
 * (vb_alt_candidates); l_max is the greedy token's logit, so its own exp() is exactly 1. */
static void fill_alts(vox_stream_t *s, int best, const char *alts[VOX_MAX_ALT]) {
    memset(alts, 0, VOX_MAX_ALT * sizeof *alts);
    alts[0] = vox_tokenizer_decode(s->tokenizer, best);
    if (s->n_alt <= 1) return;
    float z, ev[3]; int idx[3];
    vb_alt_candidates(s->e, best, TOKEN_TEXT_MIN, &z, ev, idx);
    const float inv = 1.0f / z, p_best = 1.0f * inv;
    if (!(p_best > 0)) return;
    for (int k = 0; k + 1 < s->n_alt && k < 3; k++) {
        if (idx[k] < 0) break;
        const float p = ev[k] * inv;
        if (1.0f - p / p_best > s->alt_cutoff) break;
        alts[k + 1] = vox_tokenizer_decode(s->tokenizer, idx[k]);
    }
}

static void remember_id(vox_stream_t *s, int id) {
    if (s->n_ids == s->cap_ids) {
        s->cap_ids = s->cap_ids ? s->cap_ids * 2 : 1024;
        s->ids = xrealloc(s->ids, sizeof(int) * (size_t)s->cap_ids);
    }
    s->ids[s->n_ids++] = id;
}

/* ---------------------------------------------------------------- conv stem (voxtral.c:537-715) */
/* d_mel_new: [n,128] device.  Returns the number of post-conv positions written to *out (device). */
static int conv_stem(vox_stream_t *s, const float *d_mel_new, int n, float **out) {
    VbEngine *e = s->e;
    const int C = VOX_ENC_DIM, B = VOX_MEL_BINS;
    *out = NULL;
    if (n <= 0) return 0;
    const int first = !s->conv_stem_initialized;

    /* conv0 over [tail(2) | new(n)] -> n rows; the first chunk sees true zero left padding */
    float *in0 = vb_ws(e, WS_CONV0_IN, (size_t)(n + 2) * B * 4);
    if (first) vb_dzero(e, in0, (size_t)2 * B * 4);
    else vb_d2d(e, in0, s->d_mel_tail, (size_t)2 * B * 4);
    vb_d2d(e, in0 + 2 * B, d_mel_new, (size_t)n * B * 4);
    float *c0 = vb_ws(e, WS_CONV0_OUT, (size_t)n * C * 4);
    vb_conv_view_dev(e, in0, B, 1, n, e->d_conv0_wk, e->d_conv0_b, c0, C);
    /* mel tail for the next call: last two new frames; with a single new frame the older slot
     * is zeroed, as the reference does (voxtral.c:573-579,609-615) */
    if (n >= 2) vb_d2d(e, s->d_mel_tail, d_mel_new + (size_t)(n - 2) * B, (size_t)2 * B * 4);
    else { vb_dzero(e, s->d_mel_tail, (size_t)B * 4); vb_d2d(e, s->d_mel_tail + B, d_mel_new, (size_t)B * 4); }
    s->conv_stem_initialized = 1;

    /* stride alignment: conv1 consumes conv0 rows two at a time */
    const int prev_res = s->conv0_residual_count;
    const int total = prev_res + n, new_res = total & 1;
    const int from_new = n - new_res, feed = prev_res + from_new;
    if (feed <= 0) {
        if (new_res) vb_d2d(e, s->d_conv0_resid, c0 + (size_t)(n - 1) * C, (size_t)C * 4);
        s->conv0_residual_count = new_res;
        return 0;
    }
    /* conv1 input = [previous conv0 row (zeros at stream start) | old residual | from_new rows] */
    float *in1 = vb_ws(e, WS_CONV1_IN, (size_t)(feed + 1) * C * 4);
    if (first) vb_dzero(e, in1, (size_t)C * 4);
    else vb_d2d(e, in1, s->d_conv0_tail, (size_t)C * 4);
    if (prev_res) vb_d2d(e, in1 + C, s->d_conv0_resid, (size_t)C * 4);
    vb_d2d(e, in1 + (size_t)(1 + prev_res) * C, c0, (size_t)from_new * C * 4);
    if (new_res) vb_d2d(e, s->d_conv0_resid, c0 + (size_t)(n - 1) * C, (size_t)C * 4);
    s->conv0_residual_count = new_res;
    vb_d2d(e, s->d_conv0_tail, in1 + (size_t)feed * C, (size_t)C * 4);

    int n_out = feed / 2;
    float *c1 = vb_ws(e, WS_CONV1_OUT, (size_t)n_out * C * 4);
    vb_conv_view_dev(e, in1, C, 2, n_out, e->d_conv1_wk, e->d_conv1_b, c1, C);
    *out = c1;
    return n_out;
}

/* ---------------------------------------------------------------- adapter rows */
static int adapter_reserve(vox_stream_t *s, int extra) {
    int phys = s->total_adapter - s->adapter_pos_offset;
    if (phys + extra <= s->adapter_cap) return 0;
    int ncap = s->adapter_cap ? s->adapter_cap * 2 : 256;
    while (ncap < phys + extra) ncap *= 2;
    float *nd = vb_dev_alloc((size_t)ncap * VOX_DEC_DIM * 4);
    if (phys > 0) vb_d2d(s->e, nd, s->d_adapter, (size_t)phys * VOX_DEC_DIM * 4);
    vb_sync(s->e);
    if (s->d_adapter) cudaFree(s->d_adapter);
    s->d_adapter = nd; s->adapter_cap = ncap;
    return 0;
}

static void adapter_compact(vox_stream_t *s) {            /* voxtral.c:717-731 */
    int consumed = s->gen_pos - s->adapter_pos_offset;
    if (consumed <= 0) return;
    int remaining = (s->total_adapter - s->adapter_pos_offset) - consumed;
    if (remaining > 0) {
        size_t bytes = (size_t)remaining * VOX_DEC_DIM * 4;
        float *tmp = vb_ws(s->e, WS_TMP, bytes);
        vb_d2d(s->e, tmp, s->d_adapter + (size_t)consumed * VOX_DEC_DIM, bytes);
        vb_d2d(s->e, s->d_adapter, tmp, bytes);
    }
    s->adapter_pos_offset += consumed;
}

static void reset_decoder_state(vox_stream_t *s) {        /* voxtral.c:733-749 */
    s->ctx->kv_cache_len = 0; s->ctx->kv_pos_offset = 0;
    s->total_adapter = 0; s->adapter_pos_offset = 0; s->gen_pos = 0;
    s->decoder_started = 0; s->prev_token = TOKEN_BOS; s->eos_seen = 0;
    s->n_generated = 0; s->nontext_streak = 0; s->text_since_restart = 0; s->waiting_prompt = 0;
}

static int reset_full_state(vox_stream_t *s) {            /* voxtral.c:752-780 */
    vox_mel_ctx_t *nm = vb_mel_ctx_init_on(s->e, 32 * SAMPLES_PER_TOKEN);
    if (!nm) return -1;
    vox_mel_free(s->mel);
    s->mel = nm; s->mel_cursor = 0;
    s->conv_stem_initialized = 0; s->conv0_residual_count = 0; s->enc_residual_count = 0;
    s->ctx->enc_kv_cache_len = 0; s->ctx->enc_kv_pos_offset = 0;
    reset_decoder_state(s);
    return 0;
}

/* ---------------------------------------------------------------- encoder side (voxtral.c:783-907) */
static void run_encoder(vox_stream_t *s) {
    VbEngine *e = s->e;
    int mel_frames = 0, mel_offset = 0;
    float *d_mel = vb_mel_dev_frames(s->mel, &mel_frames, &mel_offset);
    int total_mel = mel_offset + mel_frames;
    if (s->mel_cursor < mel_offset) s->mel_cursor = mel_offset;
    int mel_start = s->mel_cursor - mel_offset;
    int new_mel = total_mel - s->mel_cursor;
    int need = s->conv_stem_initialized ? s->min_new_mel : FIRST_CHUNK_MIN_MEL;
    if (new_mel < need && !s->finished) return;
    if (new_mel <= 0) return;

    double t0 = now_ms();
    float *conv_out = NULL;
    int conv_len = conv_stem(s, d_mel + (size_t)mel_start * VOX_MEL_BINS, new_mel, &conv_out);
    s->mel_cursor = total_mel;
    if (conv_len <= 0) { vox_mel_discard_before(s->mel, s->mel_cursor); return; }

    /* 32 encoder layers, in place on conv_out */
    vox_cuda_encoder_step(s->ctx, conv_out, conv_len);
    float *enc_out = conv_out;
    int enc_len = conv_len;

    /* 4x alignment for the adapter reshape */
    int total_enc = s->enc_residual_count + enc_len;
    int usable = (total_enc / VOX_DOWNSAMPLE) * VOX_DOWNSAMPLE, leftover = total_enc - usable;
    int from_res = s->enc_residual_count < usable ? s->enc_residual_count : usable;
    int from_enc = usable - from_res;
    if (usable > 0) {
        size_t row = (size_t)VOX_ENC_DIM * 4;
        float *comb = vb_ws(e, WS_COMBINED, (size_t)usable * row);
        if (from_res > 0) vb_d2d(e, comb, s->d_enc_resid, from_res * row);
        if (from_enc > 0) vb_d2d(e, comb + (size_t)from_res * VOX_ENC_DIM, enc_out, from_enc * row);
        int T = usable / VOX_DOWNSAMPLE;
        adapter_reserve(s, T);
        int phys = s->total_adapter - s->adapter_pos_offset;
        vb_adapter_dev(e, comb, usable, s->d_adapter + (size_t)phys * VOX_DEC_DIM);
        s->total_adapter += T;
    }
    if (leftover > 0) {
        /* rows not yet consumed: what is left of the old residual (only when usable == 0) followed
         * by the unused tail of enc_out.  The reference's copy assumes usable >= residual count
         * (voxtral.c:880-888); for the sub-4-row corner it reads out of bounds, here the rows are kept. */
        size_t row = (size_t)VOX_ENC_DIM * 4;
        int keep_old = s->enc_residual_count - from_res;
        float *tmp = vb_ws(e, WS_TMP, (size_t)3 * row);
        if (keep_old > 0) vb_d2d(e, tmp, s->d_enc_resid + (size_t)from_res * VOX_ENC_DIM, keep_old * row);
        vb_d2d(e, tmp + (size_t)keep_old * VOX_ENC_DIM, enc_out + (size_t)from_enc * VOX_ENC_DIM,
               (size_t)(leftover - keep_old) * row);
        vb_d2d(e, s->d_enc_resid, tmp, leftover * row);
    }
    s->enc_residual_count = leftover;

    vb_sync(e);
    double dt = now_ms() - t0;
    s->encoder_ms += dt;
    e->last_encoder_ms = dt; e->last_encoder_positions = conv_len;
    e->total_encoder_ms += dt; e->total_encoder_positions += conv_len;
    if (vox_monitor) { fprintf(stderr, "\xe2\x96\xb6"); fflush(stderr); }
    if (vox_verbose >= 2)
        fprintf(stderr, "  Encoder inc: %d mel -> %d conv -> %d usable (total adapter: %d, residual: %d)\n",
                new_mel, conv_len, usable, s->total_adapter, leftover);
    vox_mel_discard_before(s->mel, s->mel_cursor);
}

/* ---------------------------------------------------------------- decoder side (voxtral.c:969-1188) */
/* Reference KV bookkeeping for one appended position (voxtral_decoder.c:612-623,692). */
static void kv_counters_step(vox_ctx_t *c) {
    int pos = c->kv_cache_len;
    if (pos >= c->kv_cache_max) {
        if (c->kv_cache_len > VOX_DEC_WINDOW) {
            c->kv_pos_offset += c->kv_cache_len - VOX_DEC_WINDOW;
            c->kv_cache_len = VOX_DEC_WINDOW;
            pos = c->kv_cache_len;
        }
        if (pos >= c->kv_cache_max) {
            int m = c->kv_cache_max > 0 ? c->kv_cache_max : 1;
            while (m < pos + 1024) m *= 2;
            c->kv_cache_max = m;
        }
    }
    c->kv_cache_len = pos + 1;
}

typedef struct { int text, control, invalid, eos; } step_stats;

/* Account for one generated token exactly like the loop body at voxtral.c:1063-1092. */
static void on_token(vox_stream_t *s, int tok, step_stats *st) {
    s->prev_token = tok;
    s->n_generated++;
    s->last_decode_sample = s->real_samples_fed;
    remember_id(s, tok);
    tok_class cls = classify(s, tok);
    if (cls == TOK_TEXT) {
        const char *alts[VOX_MAX_ALT];
        fill_alts(s, tok, alts);
        if (alts[0]) {
            enqueue(s, alts);
            s->n_text_tokens++; s->text_since_restart = 1; s->empty_restarts = 0;
        }
        s->nontext_streak = 0;
        if (st) st->text++;
    } else if (cls == TOK_CONTROL) { s->nontext_streak++; if (st) st->control++; }
    else if (cls == TOK_INVALID) { s->nontext_streak++; if (st) st->invalid++; }
    if (tok == TOKEN_EOS) { s->eos_seen = 1; if (st) st->eos = 1; }
}

/* Generate up to n tokens from adapter rows [gen_pos, gen_pos+n); returns tokens produced. */
static int generate(vox_stream_t *s, int n, step_stats *st) {
    VbEngine *e = s->e;
    vox_ctx_t *c = s->ctx;
    if (n > s->tok_buf_cap) { s->tok_buf_cap = n + 256; s->tok_buf = xrealloc(s->tok_buf, sizeof(int) * (size_t)s->tok_buf_cap); }
    int produced = 0;
    while (produced < n && !s->eos_seen) {
        int want = s->n_alt > 1 ? 1 : n - produced;      /* alternatives are taken from each step's logits (still in HBM) */
        int row = s->gen_pos - s->adapter_pos_offset;
        int pos = c->kv_pos_offset + c->kv_cache_len;
        int got = vb_decoder_run_steps(e, s->d_adapter, row, want, s->prev_token, pos, s->tok_buf);
        if (got <= 0) break;
        for (int i = 0; i < got; i++) {
            kv_counters_step(c);
            on_token(s, s->tok_buf[i], st);
            s->gen_pos++;
            produced++;
            if (s->eos_seen) break;
        }
    }
    return produced;
}

/* Prompt prefill + the first decode step (voxtral.c:990-1012).  Prompt embeddings for positions 0..prompt_len-2 are
 * prefilled; the last prompt position is an ordinary decode step whose "previous token" is STREAMING_PAD. */
static void start_decoder(vox_stream_t *s) {
    VbEngine *e = s->e;
    vox_ctx_t *c = s->ctx;
    const int prompt_len = 1 + 32 + c->delay_tokens;
    s->waiting_prompt = 0;
    double t0 = now_ms();
    int pre = prompt_len - 1;
    float *prompt = vb_ws(e, WS_PROMPT, (size_t)pre * VOX_DEC_DIM * 4);
    vb_build_prompt_dev(e, prompt, s->d_adapter, pre, TOKEN_BOS, TOKEN_STREAMING_PAD);
    c->kv_cache_len = 0; c->kv_pos_offset = 0;
    if (c->kv_cache_max == 0) c->kv_cache_max = VOX_DEC_WINDOW + pre + 1024;      /* kv_cache_init, voxtral_decoder.c:423 */
    vb_decoder_prefill_dev(e, prompt, pre, 0);
    c->kv_cache_len = pre;
    s->gen_pos = s->adapter_pos_offset + pre;
    s->prev_token = TOKEN_STREAMING_PAD;
    generate(s, 1, NULL);
    s->decoder_started = 1;
    double dt = now_ms() - t0;
    s->decoder_ms += dt; s->prefill_ms += dt;
    if (vox_monitor) { fprintf(stderr, "\xc2\xb7"); fflush(stderr); }
}

static void run_decoder(vox_stream_t *s) {
    vox_ctx_t *c = s->ctx;
    const int prompt_len = 1 + 32 + c->delay_tokens;
    int cur = s->total_adapter - s->adapter_pos_offset;

    if (!s->decoder_started && cur < prompt_len) {
        if (vox_monitor && !s->waiting_prompt) { fprintf(stderr, "\xe2\x8c\x9b"); fflush(stderr); s->waiting_prompt = 1; }
        return;
    }
    if (!s->decoder_started) start_decoder(s);

    if (s->decoder_started && !s->eos_seen && s->gen_pos < s->total_adapter) {
        double t0 = now_ms();
        step_stats st = { 0, 0, 0, 0 };
        int steps = generate(s, s->total_adapter - s->gen_pos, &st);
        if (steps > 0) {
            double dt = now_ms() - t0;
            s->decoder_ms += dt;
            if (vox_monitor) {
                int slow = dt / steps > 40;
                const char *sym, *sev = "";
                if (st.text > 0) sym = slow ? "\xe2\x96\xb8" : "\xe2\x96\xaa";
                else if (st.invalid > 0) sym = slow ? "\xe2\x9c\x98" : "\xe2\x9c\x97";
                else if (st.control > 0) sym = slow ? "\xe2\x96\xb9" : "\xe2\x96\xab";
                else if (st.eos) sym = "\xe2\x97\xa6";
                else sym = "\xe2\x96\xaa";
                if (st.text == 0 && (st.control > 0 || st.invalid > 0)) {
                    if (s->nontext_streak >= MAX_NON_TEXT_STREAK - 8) sev = "\xe2\x98\xa0";
                    else if (s->nontext_streak >= MAX_NON_TEXT_STREAK / 2) sev = "\xe2\x9a\xa0";
                }
                fprintf(stderr, "%s%s", sym, sev); fflush(stderr);
            }
        }
    }

    adapter_compact(s);

    /* live-stream restart policy (voxtral.c:1137-1187) */
    int need_restart = 0, full_reset = 0;
    if (s->continuous) {
        if (s->eos_seen) need_restart = 1;
        else if (s->decoder_started && c->kv_cache_len > MAX_DECODE_KV) need_restart = 2;
        else if (s->decoder_started && s->nontext_streak >= MAX_NON_TEXT_STREAK) need_restart = 3;
        else if (!s->finished && s->real_samples_fed - s->last_decode_sample >= MAX_NO_DECODE_SAMPLES) need_restart = 4;
    }
    if (need_restart) {
        if (s->text_since_restart) s->empty_restarts = 0; else s->empty_restarts++;
        if (need_restart >= 2 || s->empty_restarts >= EMPTY_RESTARTS_FOR_FULL_RESET) full_reset = 1;
        if (vox_monitor) {
            const char *sym = need_restart == 1 ? "\xe2\x86\xba" : need_restart == 2 ? "\xe2\x9f\xb3"
                            : need_restart == 3 ? "\xe2\x86\xaf" : "\xe2\x8c\x9a";
            fprintf(stderr, "%s%s", sym, full_reset ? "\xe2\x99\xbb" : "\xe2\x9c\x82"); fflush(stderr);
        }
        if (full_reset) { if (reset_full_state(s) != 0) reset_decoder_state(s); s->empty_restarts = 0; }
        else reset_decoder_state(s);
        s->last_decode_sample = s->real_samples_fed;
    }
}

/* ---------------------------------------------------------------- public stream API */
vox_stream_t *vox_stream_init(vox_ctx_t *ctx) {
    if (!ctx) return NULL;
    vox_stream_t *volatile s = calloc(1, sizeof *s);
    if (!s) return NULL;
    VB_API_GUARD({ fprintf(stderr, "vox_stream_init: device allocation failed\n"); return NULL; });
    s->ctx = ctx; s->e = vb_engine(ctx);
    VB_CUDA_OK(cudaSetDevice(s->e->device));
    char path[1024];
    snprintf(path, sizeof path, "%s/tekken.json", ctx->model_dir);
    s->tokenizer = vox_tokenizer_load(path);
    if (!s->tokenizer) { free(s); VB_API_END; return NULL; }
    s->mel = vb_mel_ctx_init_on(s->e, 32 * SAMPLES_PER_TOKEN);
    s->queue_cap = 256;
    s->token_queue = calloc((size_t)s->queue_cap * VOX_MAX_ALT, sizeof *s->token_queue);
    s->n_alt = 1;
    s->d_mel_tail = vb_dev_alloc((size_t)2 * VOX_MEL_BINS * 4);
    s->d_conv0_tail = vb_dev_alloc((size_t)VOX_ENC_DIM * 4);
    s->d_conv0_resid = vb_dev_alloc((size_t)VOX_ENC_DIM * 4);
    s->d_enc_resid = vb_dev_alloc((size_t)3 * VOX_ENC_DIM * 4);
    s->prev_token = TOKEN_BOS;
    ctx->enc_kv_cache_len = 0; ctx->enc_kv_pos_offset = 0;
    s->min_new_mel = (int)(DEFAULT_INTERVAL_S * 100.0f);
    VB_API_END;
    return s;
}

int vox_stream_feed(vox_stream_t *s, const float *samples, int n_samples) {
    if (!s || s->finished || s->failed || n_samples <= 0) return -1;
    VB_API_GUARD({ s->failed = 1; return -1; });
    VB_CUDA_OK(cudaSetDevice(s->e->device));
    vox_mel_feed(s->mel, samples, n_samples);
    s->real_samples_fed += n_samples;
    run_encoder(s);
    if (!s->defer_decode) run_decoder(s);
    VB_API_END;
    return 0;
}

int vox_cuda_stream_feed_device(vox_stream_t *s, const float *d_samples, int n_samples) {
    if (!s || s->finished || s->failed || n_samples <= 0) return -1;
    VB_API_GUARD({ s->failed = 1; return -1; });
    VB_CUDA_OK(cudaSetDevice(s->e->device));
    vb_mel_feed_device(s->mel, d_samples, n_samples);
    s->real_samples_fed += n_samples;
    run_encoder(s);
    if (!s->defer_decode) run_decoder(s);
    VB_API_END;
    return 0;
}

int vox_stream_flush(vox_stream_t *s) {
    if (!s || s->finished || s->failed) return -1;
    VB_API_GUARD({ s->failed = 1; return -1; });
    VB_CUDA_OK(cudaSetDevice(s->e->device));
    int align = (int)((SAMPLES_PER_TOKEN - (s->real_samples_fed % SAMPLES_PER_TOKEN)) % SAMPLES_PER_TOKEN);
    int right_pad = align + (s->ctx->delay_tokens + 1 + OFFLINE_BUFFER_TOKENS) * SAMPLES_PER_TOKEN;
    vb_mel_feed_zeros(s->mel, right_pad);                 /* zeros go straight to the mel, not counted as audio */
    int saved = s->min_new_mel;
    s->min_new_mel = 1;
    run_encoder(s);
    if (!s->defer_decode) run_decoder(s);
    s->min_new_mel = saved;
    VB_API_END;
    return 0;
}

int vox_stream_finish(vox_stream_t *s) {
    if (!s || s->finished || s->failed) return -1;
    if (vox_stream_flush(s) != 0) return -1;
    VB_API_GUARD({ s->failed = 1; return -1; });
    s->finished = 1;
    vox_mel_finish(s->mel, 0);
    if (vox_verbose >= 2)
        fprintf(stderr, "Stream finished: %lld real samples (%.1f sec)\n", (long long)s->real_samples_fed,
                (double)s->real_samples_fed / VOX_SAMPLE_RATE);
    run_encoder(s);
    if (!s->defer_decode) run_decoder(s);
    VB_API_END;
    return 0;
}

int vox_stream_get(vox_stream_t *s, const char **out_tokens, int max) {
    if (!s || max <= 0) return 0;
    int n = 0;
    while (n < max && s->queue_head != s->queue_tail) {
        out_tokens[n++] = s->token_queue[s->queue_head * VOX_MAX_ALT];
        s->queue_head = (s->queue_head + 1) % s->queue_cap;
    }
    return n;
}

int vox_stream_get_alt(vox_stream_t *s, const char **out_tokens, int max_tokens, int n_alt) {
    if (!s || max_tokens <= 0 || n_alt <= 0) return 0;
    if (n_alt > VOX_MAX_ALT) n_alt = VOX_MAX_ALT;
    int n = 0;
    while (n < max_tokens && s->queue_head != s->queue_tail) {
        for (int a = 0; a < n_alt; a++) out_tokens[n * n_alt + a] = s->token_queue[s->queue_head * VOX_MAX_ALT + a];
        n++;
        s->queue_head = (s->queue_head + 1) % s->queue_cap;
    }
    return n;
}

void vox_stream_set_alt(vox_stream_t *s, int n_alt, float cutoff) {
    if (!s) return;
    s->n_alt = n_alt < 1 ? 1 : n_alt > VOX_MAX_ALT ? VOX_MAX_ALT : n_alt;
    s->alt_cutoff = cutoff < 0 ? 0 : cutoff > 1 ? 1 : cutoff;
}

void vox_set_processing_interval(vox_stream_t *s, float seconds) {
    if (!s) return;
    if (seconds <= 0) seconds = 0;
    s->min_new_mel = (int)(seconds * 100.0f);
    if (s->min_new_mel < 1) s->min_new_mel = 1;
}

void vox_stream_set_continuous(vox_stream_t *s, int enable) { if (s) s->continuous = enable; }

void vox_stream_free(vox_stream_t *s) {
    if (!s) return;
    if (vox_verbose >= 1) {   /* same two lines benchmark.py:25-30 parses */
        fprintf(stderr, "Encoder: %d mel -> %d tokens (%.0f ms)\n", s->mel_cursor, s->total_adapter, s->encoder_ms);
        if (s->n_text_tokens > 0) {
            double gen_ms = s->decoder_ms - s->prefill_ms;
            fprintf(stderr, "Decoder: %d text tokens (%d steps) in %.0f ms (prefill %.0f ms + %.1f ms/step)\n",
                    s->n_text_tokens, s->n_generated, s->decoder_ms, s->prefill_ms,
                    s->n_generated > 1 ? gen_ms / (s->n_generated - 1) : 0);
        }
    }
    cudaStreamSynchronize(s->e->stream);                /* teardown never aborts: a failed context still frees its host state */
    vox_mel_free(s->mel);
    if (s->tokenizer) vox_tokenizer_free(s->tokenizer);
    cudaFree(s->d_adapter); cudaFree(s->d_mel_tail); cudaFree(s->d_conv0_tail);
    cudaFree(s->d_conv0_resid); cudaFree(s->d_enc_resid);
    free(s->token_queue); free(s->ids); free(s->tok_buf);
    free(s);
}

/* ---------------------------------------------------------------- several streams, one weight pass */
/* A decode step reads 6.86 GB of weights whatever the number of activation columns, so up to 8 streams on one GPU share every
 * weight byte (vb_decode_v2.cu).  Each stream lives on its own context (vox_cuda_ctx_fork: own KV ring / encoder tail / scratch,
 * shared weights and CUDA stream) and is put in deferred mode, where feed/flush/finish run mel -> encoder -> adapter only;
 * vox_cuda_streams_decode() then advances the decoders of all of them together until every stream has consumed its adapter rows.
 * Per stream the result is what run_decoder() would have produced (same kernels, same per-column arithmetic); the live-stream
 * restart policy (vox_stream_set_continuous) and alternatives (n_alt > 1) need per-step host decisions and stay single-stream. */
void vox_cuda_stream_set_deferred(vox_stream_t *s, int on) { if (s) s->defer_decode = on ? 1 : 0; }

int vox_cuda_streams_decode(vox_stream_t **ss, int n) {
    if (!ss || n < 1 || n > 8) return -1;
    VbEngine *lead = ss[0]->e;
    for (int i = 0; i < n; i++) {
        vox_stream_t *s = ss[i];
        if (!s || s->continuous || s->n_alt > 1 || s->e->device != lead->device || s->e->stream != lead->stream) {
            fprintf(stderr, "vox_cuda_streams_decode: stream %d cannot be batched (continuous mode, alternatives, or another device/ctx group)\n", i);
            return -1;
        }
    }
    VB_API_GUARD({ for (int i = 0; i < n; i++) ss[i]->failed = 1; return -1; });
    VB_CUDA_OK(cudaSetDevice(lead->device));
    int total = 0;
    for (int i = 0; i < n; i++) {                         /* prompts: per-stream prefill (tcgen05 GEMMs) + first step */
        vox_stream_t *s = ss[i];
        const int prompt_len = 1 + 32 + s->ctx->delay_tokens;
        if (!s->decoder_started && s->total_adapter - s->adapter_pos_offset >= prompt_len) { start_decoder(s); total++; }
    }
    if (!vb_decoder_v2_supported(lead)) {                  /* no batched kernel on this device: one stream after the other */
        for (int i = 0; i < n; i++) { int before = ss[i]->n_generated; run_decoder(ss[i]); total += ss[i]->n_generated - before; }
        VB_API_END;
        return total;
    }
    double t0 = now_ms();
    for (;;) {
        VbV2Col cols[8]; int who[8]; int nb = 0, longest = 0;
        for (int i = 0; i < n; i++) {
            vox_stream_t *s = ss[i];
            int pending = s->total_adapter - s->gen_pos;
            if (!s->decoder_started || s->eos_seen || pending <= 0) continue;
            VbV2Col *c = &cols[nb];
            c->engine = s->e; c->d_adapter = s->d_adapter; c->adapter_row = s->gen_pos - s->adapter_pos_offset;
            c->n_steps = pending; c->prev_token = s->prev_token; c->pos = s->ctx->kv_pos_offset + s->ctx->kv_cache_len;
            if (pending > longest) longest = pending;
            who[nb++] = i;
        }
        if (nb == 0) break;
        VbDecState st[8];
        cudaEvent_t e0 = lead->ev0, e1 = lead->ev1;
        VB_CUDA_OK(cudaEventRecord(e0, lead->stream));
        if (vb_decoder_v2_launch(lead, cols, nb, longest, 0, st) != 0) VB_FAIL("batched decode launch failed");
        VB_CUDA_OK(cudaEventRecord(e1, lead->stream));
        cudaError_t serr = cudaStreamSynchronize(lead->stream);
        if (serr != cudaSuccess) fprintf(stderr, "vox_cuda_streams_decode: decode kernel failed (%d columns, wait-guard code %d)\n", nb, lead->v2.err_host ? *lead->v2.err_host : -1);
        VB_CUDA_OK(serr);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        int steps = 0;
        for (int b = 0; b < nb; b++) {
            vox_stream_t *s = ss[who[b]];
            int got = st[b].n_out;
            if (got > steps) steps = got;
            if (got <= 0) continue;
            if (got > s->tok_buf_cap) { s->tok_buf_cap = got + 256; s->tok_buf = xrealloc(s->tok_buf, sizeof(int) * (size_t)s->tok_buf_cap); }
            vb_d2h_sync(s->e, s->tok_buf, s->e->d_tokens, (size_t)got * 4);
            for (int i = 0; i < got; i++) {
                kv_counters_step(s->ctx);
                on_token(s, s->tok_buf[i], NULL);
                s->gen_pos++; total++;
                if (s->eos_seen) break;
            }
        }
        lead->last_decode_ms = ms; lead->last_decode_steps = steps;
        lead->total_decode_ms += ms; lead->total_decode_steps += steps;
        if (steps == 0) break;
    }
    double dt = now_ms() - t0;
    for (int i = 0; i < n; i++) { ss[i]->decoder_ms += dt / n; adapter_compact(ss[i]); }
    VB_API_END;
    return total;
}

/* ---------------------------------------------------------------- introspection (section 7 of the header) */
int vox_cuda_stream_token_ids(vox_stream_t *s, int *out, int max) {
    if (!s) return 0;
    int n = s->n_ids < max ? s->n_ids : max;
    if (out && n > 0) memcpy(out, s->ids, sizeof(int) * (size_t)n);
    return s->n_ids;
}
int vox_cuda_stream_counts(vox_stream_t *s, int *mel_frames, int *adapter_tokens, int *decoder_steps) {
    if (!s) return -1;
    if (mel_frames) *mel_frames = s->mel_cursor;
    if (adapter_tokens) *adapter_tokens = s->total_adapter;
    if (decoder_steps) *decoder_steps = s->n_ids;
    return 0;
}

/* ---------------------------------------------------------------- convenience API (voxtral.c:1338-1586) */
static void trim_ws(char *t) {
    size_t len = strlen(t), a = 0, b = len;
    while (a < len && isspace((unsigned char)t[a])) a++;
    while (b > a && isspace((unsigned char)t[b - 1])) b--;
    memmove(t, t + a, b - a);
    t[b - a] = 0;
}

typedef struct { char *p; size_t len, cap; } strbuf;
static void sb_drain(strbuf *sb, vox_stream_t *s) {
    const char *toks[64];
    int n;
    while ((n = vox_stream_get(s, toks, 64)) > 0)
        for (int i = 0; i < n; i++) {
            size_t l = strlen(toks[i]);
            if (sb->len + l + 1 > sb->cap) { while (sb->len + l + 1 > sb->cap) sb->cap *= 2; sb->p = xrealloc(sb->p, sb->cap); }
            memcpy(sb->p + sb->len, toks[i], l + 1);
            sb->len += l;
        }
}

char *vox_transcribe_audio(vox_ctx_t *ctx, const float *samples, int n_samples) {
    vox_stream_t *s = vox_stream_init(ctx);
    if (!s) return NULL;
    if (vox_stream_feed(s, samples, n_samples) != 0 || vox_stream_finish(s) != 0) {
        if (s->failed) { vox_stream_free(s); return NULL; }      /* device failure; n_samples <= 0 still yields "" like the reference */
        if (!s->finished) vox_stream_finish(s);
    }
    strbuf sb = { malloc(1024), 0, 1024 };
    sb.p[0] = 0;
    sb_drain(&sb, s);
    vox_stream_free(s);
    trim_ws(sb.p);
    return sb.p;
}

char *vox_transcribe(vox_ctx_t *ctx, const char *wav_path) {
    int n = 0;
    float *samples = vox_load_wav(wav_path, &n);
    if (!samples) { fprintf(stderr, "vox_transcribe: cannot load %s\n", wav_path); return NULL; }
    if (vox_verbose >= 1) fprintf(stderr, "Audio: %d samples (%.1f seconds)\n", n, (float)n / VOX_SAMPLE_RATE);
    char *text = vox_transcribe_audio(ctx, samples, n);
    free(samples);
    return text;
}

/* stdin: a RIFF header means "buffer everything, transcribe offline"; anything else is raw
 * s16le 16 kHz mono and is streamed in 4096-sample reads (voxtral.c:1371-1571). */
char *vox_transcribe_stdin(vox_ctx_t *ctx) {
    uint8_t head[4];
    if (fread(head, 1, 4, stdin) < 4) { fprintf(stderr, "vox_transcribe_stdin: not enough data on stdin\n"); return NULL; }
    if (!memcmp(head, "RIFF", 4)) {
        if (vox_verbose >= 2) fprintf(stderr, "Detected WAV format on stdin\n");
        size_t cap = 1u << 20, size = 4;
        uint8_t *buf = malloc(cap);
        memcpy(buf, head, 4);
        for (;;) {
            if (size == cap) { cap *= 2; buf = xrealloc(buf, cap); }
            size_t got = fread(buf + size, 1, cap - size, stdin);
            if (!got) break;
            size += got;
        }
        int n = 0;
        float *samples = vox_parse_wav_buffer(buf, size, &n);
        free(buf);
        if (!samples) { fprintf(stderr, "Invalid WAV data on stdin\n"); return NULL; }
        if (vox_verbose >= 1) fprintf(stderr, "Audio: %d samples (%.1f seconds)\n", n, (float)n / VOX_SAMPLE_RATE);
        char *text = vox_transcribe_audio(ctx, samples, n);
        free(samples);
        return text;
    }
    if (vox_verbose >= 2) fprintf(stderr, "Streaming raw s16le 16kHz mono from stdin\n");
    vox_stream_t *s = vox_stream_init(ctx);
    if (!s) return NULL;
    {
        int16_t sv[2]; memcpy(sv, head, 4);
        float f[2] = { sv[0] / 32768.0f, sv[1] / 32768.0f };
        vox_stream_feed(s, f, 2);
    }
    strbuf sb = { malloc(1024), 0, 1024 };
    sb.p[0] = 0;
    int16_t raw[4096]; float fb[4096];
    for (;;) {
        size_t got = fread(raw, sizeof(int16_t), 4096, stdin);
        if (!got) { vox_stream_finish(s); sb_drain(&sb, s); break; }
        for (size_t i = 0; i < got; i++) fb[i] = raw[i] / 32768.0f;
        vox_stream_feed(s, fb, (int)got);
        sb_drain(&sb, s);
    }
    vox_stream_free(s);
    trim_ws(sb.p);
    return sb.p;
}

/* ---------------------------------------------------------------- decoder host-pointer API */
int vox_decoder_kv_cache_preallocate(vox_ctx_t *ctx, int max_seq) {
    if (ctx->kv_cache_max == 0) ctx->kv_cache_max = max_seq;
    return 0;
}

void vox_decoder_prefill(vox_ctx_t *ctx, const float *input_embeds, int seq_len) {   /* voxtral_decoder.c:410 */
    if (seq_len <= 0) return;
    VbEngine *e = vb_engine(ctx);
    VB_CUDA_OK(cudaSetDevice(e->device));
    if (ctx->kv_cache_max == 0) ctx->kv_cache_max = VOX_DEC_WINDOW + seq_len + 1024;
    else if (ctx->kv_cache_len + seq_len > ctx->kv_cache_max) {
        int m = ctx->kv_cache_max;
        while (m < ctx->kv_cache_len + seq_len + 1024) m *= 2;
        ctx->kv_cache_max = m;
    }
    size_t bytes = (size_t)seq_len * VOX_DEC_DIM * 4;
    float *d = vb_ws(e, WS_PROMPT, bytes);
    vb_h2d(e, d, input_embeds, bytes);
    vb_decoder_prefill_dev(e, d, seq_len, ctx->kv_pos_offset + ctx->kv_cache_len);
    vb_sync(e);
    ctx->kv_cache_len += seq_len;
}

int vox_decoder_forward(vox_ctx_t *ctx, const float *input_embeds, float *logits) {  /* voxtral_decoder.c:586 */
    VbEngine *e = vb_engine(ctx);
    VB_API_GUARD({ return TOKEN_EOS; });               /* out of memory -> EOS, voxtral_decoder.c:621,649 */
    VB_CUDA_OK(cudaSetDevice(e->device));
    if (ctx->kv_cache_max == 0) ctx->kv_cache_max = VOX_DEC_WINDOW + 1024;
    kv_counters_step(ctx);
    int pos = ctx->kv_pos_offset + ctx->kv_cache_len - 1;
    vb_h2d(e, e->d_embed_in, input_embeds, (size_t)VOX_DEC_DIM * 4);
    int tok = vb_decoder_step_from_embed(e, e->d_embed_in, pos, logits);
    VB_API_END;
    return tok;
}

int vox_cuda_decoder_prefill(vox_ctx_t *ctx, const float *d_embeds, int n) {
    VbEngine *e = vb_engine(ctx);
    if (ctx->kv_cache_max == 0) ctx->kv_cache_max = VOX_DEC_WINDOW + n + 1024;
    vb_decoder_prefill_dev(e, d_embeds, n, ctx->kv_pos_offset + ctx->kv_cache_len);
    ctx->kv_cache_len += n;
    return 0;
}

int vox_cuda_decoder_steps(vox_ctx_t *ctx, const float *d_adapter, int first_pos, int n_steps,
                           int prev_token, int *out_tokens) {
    VbEngine *e = vb_engine(ctx);
    int pos = ctx->kv_pos_offset + ctx->kv_cache_len;
    int got = vb_decoder_run_steps(e, d_adapter, first_pos, n_steps, prev_token, pos, out_tokens);
    for (int i = 0; i < got; i++) kv_counters_step(ctx);
    return got;
}

int vox_cuda_get_info(vox_ctx_t *ctx, vox_cuda_info_t *out) {
    if (!ctx || !out) return -1;
    VbEngine *e = vb_engine(ctx);
    memset(out, 0, sizeof *out);
    out->device = e->device; out->sm_count = e->sm_count; out->cc_major = e->cc_major; out->cc_minor = e->cc_minor;
    out->weight_bytes_hbm = e->weight_bytes; out->kv_bytes_hbm = e->kv_bytes;
    out->kernel_launches = e->launches;
    out->last_decode_kernel_ms = e->last_decode_ms; out->last_decode_steps = e->last_decode_steps;
    out->last_encoder_kernel_ms = e->last_encoder_ms; out->last_encoder_positions = e->last_encoder_positions;
    out->last_mel_kernel_ms = e->last_mel_ms;
    out->total_decode_kernel_ms = e->total_decode_ms; out->total_decode_steps = e->total_decode_steps;
    out->total_encoder_ms = e->total_encoder_ms; out->total_encoder_positions = e->total_encoder_positions;
    out->load_ms = e->load_ms;
    out->verify_passes = e->verify_passes; out->verify_tokens = e->verify_tokens;
    return 0;
}

float *vox_cuda_mel_device_frames(vox_mel_ctx_t *mel, int *n_frames) {
    int off = 0;
    return vb_mel_dev_frames(mel, n_frames, &off);
}
int vox_cuda_mel_feed_zeros(vox_mel_ctx_t *mel, int n) { return vb_mel_feed_zeros(mel, n); }
int vox_cuda_build_prompt(vox_ctx_t *ctx, float *d_out, const float *d_adapter, int n) {
    vb_build_prompt_dev(vb_engine(ctx), d_out, d_adapter, n, TOKEN_BOS, TOKEN_STREAMING_PAD);
    return 0;
}

int vox_cuda_debug_copy_kv(vox_ctx_t *ctx, int layer, float *h_k, float *h_v) {
    VbEngine *e = vb_engine(ctx);
    if (!e->d_kv_k || layer < 0 || layer >= VOX_DEC_LAYERS) return -1;
    size_t n = (size_t)VB_KV_SLOTS * VB_DEC_KV;
    vb_d2h_sync(e, h_k, e->d_kv_k + (size_t)layer * n, n * 4);
    vb_d2h_sync(e, h_v, e->d_kv_v + (size_t)layer * n, n * 4);
    return 0;
}
int vox_cuda_debug_copy_logits(vox_ctx_t *ctx, float *h_logits) {
    VbEngine *e = vb_engine(ctx);
    if (!e->d_logits) return -1;
    vb_d2h_sync(e, h_logits, e->d_logits, (size_t)VOX_VOCAB_SIZE * 4);
    return 0;
}

void vox_cuda_set_decode_mode(vox_ctx_t *ctx, int mode) {
    if (ctx) vb_engine(ctx)->decode_mode = mode;
}

void vox_cuda_set_verify_depth(vox_ctx_t *ctx, int depth) {
    if (!ctx) return;
    vb_engine(ctx)->verify_depth = depth < 2 ? 1 : depth > 8 ? 8 : depth;
}

void vox_cuda_reset_caches(vox_ctx_t *ctx) {
    if (!ctx) return;
    vb_sync(vb_engine(ctx));
    ctx->kv_cache_len = 0; ctx->kv_pos_offset = 0;
    ctx->enc_kv_cache_len = 0; ctx->enc_kv_pos_offset = 0;
}

const char *vox_cuda_version(void) { return "voxtral_b200 0.1 (sm_100a)"; }
