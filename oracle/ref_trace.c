/*
 * ref_trace.c -- TEST INFRASTRUCTURE ONLY: runs the UNMODIFIED reference pipeline
 * (oracle/_ref/libvoxref.so, built from /root/reference by oracle/Makefile) on a WAV and records
 * what crosses its model-block boundaries, for tests/golden.
 *
 * How: this executable defines vox_encoder_forward_incremental / vox_adapter_forward /
 * vox_decoder_prefill / vox_decoder_forward itself.  The reference's stream code (voxtral.c) calls
 * those through the PLT of the shared library, so the dynamic linker binds the calls to the
 * definitions below (symbol interposition; this binary is linked with -rdynamic), and each wrapper
 * forwards to the real function found with dlsym(RTLD_NEXT).  No reference source is modified or
 * copied.
 *
 * Output (little-endian, in <out_dir>):
 *   trace.json        manifest with counts
 *   tokens.i32        every token id returned by vox_decoder_forward, in order
 *   logits_top.f32/.i32   per step: top-8 values / ids
 *   logits_probe.f32  per step: logits at 64 fixed probe ids
 *   enc_in_*.f32 / enc_out_*.f32   conv-stem output fed to / returned by each encoder call
 *   adapter_*.f32     each adapter output
 *   step_embed.f32    every decoder input embedding [steps,3072]
 *   prefill_embed.f32 the prefill embeddings
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "voxtral.h"
#include "voxtral_audio.h"
#include "voxtral_kernels.h"

#define N_PROBE 64
static const char *g_dir = ".";
static FILE *f_tokens, *f_topv, *f_topi, *f_probe, *f_step_embed;
static int n_steps, n_enc_calls, n_adapter_calls, n_prefill;
static int probe_ids[N_PROBE];

static FILE *open_out(const char *name) {
    char p[1024];
    snprintf(p, sizeof p, "%s/%s", g_dir, name);
    FILE *f = fopen(p, "wb");
    if (!f) { perror(p); exit(1); }
    return f;
}
static void dump(const char *name, const void *data, size_t bytes) {
    FILE *f = open_out(name);
    fwrite(data, 1, bytes, f);
    fclose(f);
}

float *vox_encoder_forward_incremental(vox_ctx_t *ctx, const float *x_new, int new_len, int *out_len) {
    static float *(*real)(vox_ctx_t *, const float *, int, int *);
    if (!real) real = dlsym(RTLD_NEXT, "vox_encoder_forward_incremental");
    char name[64];
    snprintf(name, sizeof name, "enc_in_%d.f32", n_enc_calls);
    dump(name, x_new, (size_t)new_len * VOX_ENC_DIM * 4);
    float *out = real(ctx, x_new, new_len, out_len);
    snprintf(name, sizeof name, "enc_out_%d.f32", n_enc_calls);
    if (out) dump(name, out, (size_t)*out_len * VOX_ENC_DIM * 4);
    n_enc_calls++;
    return out;
}

float *vox_adapter_forward(vox_ctx_t *ctx, const float *enc_out, int enc_seq_len, int *out_seq_len) {
    static float *(*real)(vox_ctx_t *, const float *, int, int *);
    if (!real) real = dlsym(RTLD_NEXT, "vox_adapter_forward");
    float *out = real(ctx, enc_out, enc_seq_len, out_seq_len);
    char name[64];
    snprintf(name, sizeof name, "adapter_%d.f32", n_adapter_calls++);
    if (out) dump(name, out, (size_t)*out_seq_len * VOX_DEC_DIM * 4);
    return out;
}

void vox_decoder_prefill(vox_ctx_t *ctx, const float *input_embeds, int seq_len) {
    static void (*real)(vox_ctx_t *, const float *, int);
    if (!real) real = dlsym(RTLD_NEXT, "vox_decoder_prefill");
    char name[64];
    snprintf(name, sizeof name, n_prefill ? "prefill_embed_%d.f32" : "prefill_embed.f32", n_prefill);
    dump(name, input_embeds, (size_t)seq_len * VOX_DEC_DIM * 4);
    n_prefill++;
    real(ctx, input_embeds, seq_len);
}

static int cmp_desc(const void *a, const void *b) {
    float x = ((const float *)a)[0], y = ((const float *)b)[0];
    return x < y ? 1 : x > y ? -1 : 0;
}

int vox_decoder_forward(vox_ctx_t *ctx, const float *input_embeds, float *logits) {
    static int (*real)(vox_ctx_t *, const float *, float *);
    if (!real) real = dlsym(RTLD_NEXT, "vox_decoder_forward");
    fwrite(input_embeds, 4, VOX_DEC_DIM, f_step_embed);
    int tok = real(ctx, input_embeds, logits);
    int32_t t32 = tok;
    fwrite(&t32, 4, 1, f_tokens);
    /* top-8 by value (ties: lower id first, as the reference argmax) */
    float topv[8]; int32_t topi[8];
    for (int k = 0; k < 8; k++) { topv[k] = -1e30f; topi[k] = -1; }
    for (int i = 0; i < VOX_VOCAB_SIZE; i++) {
        float v = logits[i];
        if (v > topv[7]) {
            int k = 7;
            while (k > 0 && v > topv[k - 1]) { topv[k] = topv[k - 1]; topi[k] = topi[k - 1]; k--; }
            topv[k] = v; topi[k] = i;
        }
    }
    (void)cmp_desc;
    fwrite(topv, 4, 8, f_topv);
    fwrite(topi, 4, 8, f_topi);
    float pv[N_PROBE];
    for (int k = 0; k < N_PROBE; k++) pv[k] = logits[probe_ids[k]];
    fwrite(pv, 4, N_PROBE, f_probe);
    n_steps++;
    return tok;
}

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s <model_dir> <wav> <out_dir> [feed_chunk_samples]\n"
                        "  scenario knobs (environment): TRACE_DELAY_MS, TRACE_FLUSH_AFTER_CHUNK, TRACE_ALT_N + TRACE_ALT_CUTOFF,\n"
                        "  TRACE_CONTINUOUS, TRACE_INTERVAL\n", argv[0]);
        return 2;
    }
    g_dir = argv[3];
    int chunk = argc > 4 ? atoi(argv[4]) : 0;
    const char *e_delay = getenv("TRACE_DELAY_MS"), *e_flush = getenv("TRACE_FLUSH_AFTER_CHUNK"), *e_alt = getenv("TRACE_ALT_N");
    const char *e_cut = getenv("TRACE_ALT_CUTOFF"), *e_cont = getenv("TRACE_CONTINUOUS"), *e_int = getenv("TRACE_INTERVAL");
    const int n_alt = e_alt ? atoi(e_alt) : 1;
    uint32_t r = 0xC0FFEEu;
    for (int k = 0; k < N_PROBE; k++) { r = r * 1664525u + 1013904223u; probe_ids[k] = (int)((r >> 8) % VOX_VOCAB_SIZE); }
    f_tokens = open_out("tokens.i32"); f_topv = open_out("logits_top.f32"); f_topi = open_out("logits_top.i32");
    f_probe = open_out("logits_probe.f32"); f_step_embed = open_out("step_embed.f32");

    vox_verbose = 1;
    vox_ctx_t *ctx = vox_load(argv[1]);
    if (!ctx) return 1;
    int n = 0;
    float *pcm = vox_load_wav(argv[2], &n);
    if (!pcm) return 1;
    if (e_delay) vox_set_delay(ctx, atoi(e_delay));
    vox_stream_t *s = vox_stream_init(ctx);
    if (!s) return 1;
    if (e_int) vox_set_processing_interval(s, (float)atof(e_int));
    if (e_cont) vox_stream_set_continuous(s, atoi(e_cont));
    if (n_alt > 1) vox_stream_set_alt(s, n_alt, e_cut ? (float)atof(e_cut) : 0.5f);
    /* text pieces, in order, for the text-level golden; with alternatives: one line per position, fields separated by TAB,
     * and text.txt keeps the best piece only.  drain.txt records how many positions each drain returned. */
    FILE *ftext = open_out("text.txt");
    FILE *falt = open_out("alt.txt");
    FILE *fdrain = open_out("drain.txt");
    const char *toks[64 * 8];
    int got;
#define DRAIN(tag) do { \
        int total = 0; \
        if (n_alt > 1) { \
            while ((got = vox_stream_get_alt(s, toks, 64, n_alt)) > 0) { total += got; \
                for (int i = 0; i < got; i++) { \
                    fputs(toks[i * n_alt], ftext); \
                    for (int k = 0; k < n_alt; k++) { if (k) fputc('\t', falt); fputs(toks[i * n_alt + k] ? toks[i * n_alt + k] : "<null>", falt); } \
                    fputc('\n', falt); } } \
        } else while ((got = vox_stream_get(s, toks, 64)) > 0) { total += got; for (int i = 0; i < got; i++) fputs(toks[i], ftext); } \
        fprintf(fdrain, "%s %d\n", tag, total); } while (0)
    if (chunk <= 0) vox_stream_feed(s, pcm, n);
    else {
        int ci = 0;
        for (int off = 0; off < n; off += chunk, ci++) {
            vox_stream_feed(s, pcm + off, n - off < chunk ? n - off : chunk);
            DRAIN("feed");
            if (e_flush && ci == atoi(e_flush)) { vox_stream_flush(s); DRAIN("flush"); }
        }
    }
    vox_stream_finish(s);
    DRAIN("finish");
    fclose(ftext); fclose(falt); fclose(fdrain);
    vox_stream_free(s);

    FILE *fj = open_out("trace.json");
    fprintf(fj, "{\"samples\": %d, \"feed_chunk\": %d, \"decoder_steps\": %d, \"encoder_calls\": %d, "
                "\"adapter_calls\": %d, \"prefills\": %d, \"probe_ids\": [", n, chunk, n_steps, n_enc_calls,
            n_adapter_calls, n_prefill);
    for (int k = 0; k < N_PROBE; k++) fprintf(fj, "%s%d", k ? "," : "", probe_ids[k]);
    fprintf(fj, "]}\n");
    fclose(fj);
    fclose(f_tokens); fclose(f_topv); fclose(f_topi); fclose(f_probe); fclose(f_step_embed);
    vox_free(ctx);
    free(pcm);
    return 0;
}
