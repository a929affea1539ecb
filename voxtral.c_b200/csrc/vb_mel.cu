/*
 * vb_mel.cu -- log-mel front end on the GPU (HOT LOOP A of SURVEY.md section 3.1).
 *
 * Replaces the numeric half of /root/reference voxtral_audio.c:219-672:
 *   tables   build_mel_filters / hertz_to_mel / mel_to_hertz (:223-285), periodic Hann (:540-542),
 *            DFT cos/sin (:528-538) -- computed ON THE HOST in f32 with the reference's exact
 *            expressions (they are one-off and their f32 rounding is part of the model input),
 *            then uploaded once per process (transposed so that threads index the contiguous dim).
 *   frames   mel_compute_available (:454-513): w = x[160t..160t+399]*hann; re/im[k] = sum_n w[n]cos/sin;
 *            p = re^2+im^2; mel[m] = sum_k filt[m,k] p[k]; max(log10(max(mel,1e-10)), -6.5); (v+4)/4.
 *   stream   vox_mel_ctx_init/feed/finish/data/frame_offset/discard_before (:515-662), including the
 *            200+left_pad zero prefix, the 200-sample reflect over the tail at finish and the dropped
 *            last frame.
 *
 * N_FFT = 400 is not a power of two; the reference does a direct 201x400 DFT and so does this kernel
 * (37 MFLOP per audio second -- the front end is <0.1% of the pipeline).
 *
 * Device residency: samples are staged to HBM on feed; frames are produced in HBM and stay there for
 * the conv stem.  vox_mel_data() (host-pointer API) lazily mirrors the frames to host memory.
 */
#include "vb_engine.h"
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define N_MEL   128
#define N_FFT   400
#define N_FREQ  201
#define KPAD    208        /* N_FREQ padded for the transposed tables */
#define HOP     160
#define FRAMES_PER_CTA 8
#define KEEP_TAIL 512      /* trailing samples always retained (reflect padding needs 202) */

struct MelTables { float *cosT, *sinT, *filtT, *window; int device; };
static MelTables g_tab[16];
static int g_tab_n = 0;

/* ---- host-side table construction (reference voxtral_audio.c:223-285, 528-542) ---- */
static float hz_to_mel(float f) {
    const float min_log_hz = 1000.0f, min_log_mel = 15.0f, logstep = 27.0f / logf(6.4f);
    float m = 3.0f * f / 200.0f;
    if (f >= min_log_hz) m = min_log_mel + logf(f / min_log_hz) * logstep;
    return m;
}
static float mel_to_hz(float m) {
    const float min_log_hz = 1000.0f, min_log_mel = 15.0f, logstep = logf(6.4f) / 27.0f;
    float f = 200.0f * m / 3.0f;
    if (m >= min_log_mel) f = min_log_hz * expf(logstep * (m - min_log_mel));
    return f;
}
static void build_filters(float *filt /* [N_MEL][N_FREQ] */) {
    float fft_f[N_FREQ], ff[N_MEL + 2], fd[N_MEL + 1];
    for (int i = 0; i < N_FREQ; i++) fft_f[i] = (float)i * ((float)VOX_SAMPLE_RATE / 2.0f) / (float)(N_FREQ - 1);
    float lo = hz_to_mel(0.0f), hi = hz_to_mel((float)VOX_SAMPLE_RATE / 2.0f);
    for (int i = 0; i < N_MEL + 2; i++) ff[i] = mel_to_hz(lo + (hi - lo) * (float)i / (float)(N_MEL + 1));
    for (int i = 0; i < N_MEL + 1; i++) { fd[i] = ff[i + 1] - ff[i]; if (fd[i] == 0.0f) fd[i] = 1e-6f; }
    for (int m = 0; m < N_MEL; m++) {
        float enorm = 2.0f / (ff[m + 2] - ff[m]);
        for (int k = 0; k < N_FREQ; k++) {
            float down = (fft_f[k] - ff[m]) / fd[m], up = (ff[m + 2] - fft_f[k]) / fd[m + 1];
            float v = fminf(down, up);
            filt[m * N_FREQ + k] = (v < 0.0f ? 0.0f : v) * enorm;
        }
    }
}

static const MelTables *mel_tables(int device) {
    for (int i = 0; i < g_tab_n; i++) if (g_tab[i].device == device) return &g_tab[i];
    if (g_tab_n >= (int)(sizeof g_tab / sizeof g_tab[0])) VB_FAIL("mel tables: more than 16 devices in one process");
    float *cosT = (float *)calloc((size_t)N_FFT * KPAD, 4), *sinT = (float *)calloc((size_t)N_FFT * KPAD, 4);
    float *filt = (float *)malloc((size_t)N_MEL * N_FREQ * 4), *filtT = (float *)calloc((size_t)N_FREQ * N_MEL, 4);
    float win[N_FFT];
    for (int k = 0; k < N_FREQ; k++)
        for (int n = 0; n < N_FFT; n++) {
            float ang = 2.0f * (float)M_PI * (float)k * (float)n / (float)N_FFT;
            cosT[n * KPAD + k] = cosf(ang);
            sinT[n * KPAD + k] = sinf(ang);
        }
    build_filters(filt);
    for (int m = 0; m < N_MEL; m++) for (int k = 0; k < N_FREQ; k++) filtT[k * N_MEL + m] = filt[m * N_FREQ + k];
    for (int i = 0; i < N_FFT; i++) win[i] = 0.5f * (1.0f - cosf(2.0f * (float)M_PI * (float)i / (float)N_FFT));
    MelTables *t = &g_tab[g_tab_n++];
    t->device = device;
    t->cosT = (float *)vb_dev_alloc((size_t)N_FFT * KPAD * 4);
    t->sinT = (float *)vb_dev_alloc((size_t)N_FFT * KPAD * 4);
    t->filtT = (float *)vb_dev_alloc((size_t)N_FREQ * N_MEL * 4);
    t->window = (float *)vb_dev_alloc(N_FFT * 4);
    VB_CUDA_OK(cudaMemcpy(t->cosT, cosT, (size_t)N_FFT * KPAD * 4, cudaMemcpyHostToDevice));
    VB_CUDA_OK(cudaMemcpy(t->sinT, sinT, (size_t)N_FFT * KPAD * 4, cudaMemcpyHostToDevice));
    VB_CUDA_OK(cudaMemcpy(t->filtT, filtT, (size_t)N_FREQ * N_MEL * 4, cudaMemcpyHostToDevice));
    VB_CUDA_OK(cudaMemcpy(t->window, win, N_FFT * 4, cudaMemcpyHostToDevice));
    free(cosT); free(sinT); free(filt); free(filtT);
    return t;
}

/* ---- the kernel: FRAMES_PER_CTA frames per CTA, thread = DFT bin (then mel bin) ---- */
__global__ void __launch_bounds__(256)
k_mel_frames(const float *__restrict__ samples, long long first_sample /* index of frame0's first sample in `samples` */,
             int n_frames, const float *__restrict__ cosT, const float *__restrict__ sinT,
             const float *__restrict__ filtT, const float *__restrict__ window, float *__restrict__ mel_out) {
    __shared__ float wf[FRAMES_PER_CTA][N_FFT];
    __shared__ float pw[FRAMES_PER_CTA][KPAD];
    const int f0 = blockIdx.x * FRAMES_PER_CTA;
    const int nf = min(FRAMES_PER_CTA, n_frames - f0);
    for (int i = threadIdx.x; i < FRAMES_PER_CTA * N_FFT; i += 256) {
        int f = i / N_FFT, n = i % N_FFT;
        wf[f][n] = f < nf ? samples[first_sample + (long long)(f0 + f) * HOP + n] * window[n] : 0.f;
    }
    __syncthreads();
    const int k = threadIdx.x;
    if (k < N_FREQ) {
        float re[FRAMES_PER_CTA], im[FRAMES_PER_CTA];
#pragma unroll
        for (int f = 0; f < FRAMES_PER_CTA; f++) { re[f] = 0.f; im[f] = 0.f; }
        for (int n = 0; n < N_FFT; n++) {
            float c = cosT[n * KPAD + k], s = sinT[n * KPAD + k];
#pragma unroll
            for (int f = 0; f < FRAMES_PER_CTA; f++) {
                float w = wf[f][n];
                re[f] = fmaf(w, c, re[f]);
                im[f] = fmaf(w, s, im[f]);
            }
        }
#pragma unroll
        for (int f = 0; f < FRAMES_PER_CTA; f++) pw[f][k] = re[f] * re[f] + im[f] * im[f];
    }
    __syncthreads();
    const int m = threadIdx.x;
    if (m < N_MEL) {
        float acc[FRAMES_PER_CTA];
#pragma unroll
        for (int f = 0; f < FRAMES_PER_CTA; f++) acc[f] = 0.f;
        for (int kk = 0; kk < N_FREQ; kk++) {
            float fl = filtT[kk * N_MEL + m];
#pragma unroll
            for (int f = 0; f < FRAMES_PER_CTA; f++) acc[f] = fmaf(fl, pw[f][kk], acc[f]);
        }
#pragma unroll
        for (int f = 0; f < FRAMES_PER_CTA; f++) {
            if (f < nf) {
                float v = acc[f] < 1e-10f ? 1e-10f : acc[f];
                v = log10f(v);
                const float lo = VOX_LOG_MEL_MAX - 8.0f;
                if (v < lo) v = lo;
                mel_out[(size_t)(f0 + f) * N_MEL + m] = (v + 4.0f) / 4.0f;
            }
        }
    }
}

/* ---- incremental context ---- */
struct vox_mel_ctx {
    VbEngine *e;
    const MelTables *tab;
    float *d_samples; long long sample_base; int n_local, cap_local;   /* device window of the padded signal */
    long long n_total;                                                 /* padded samples seen so far */
    float *d_mel; int mel_cap;                                         /* frames [mel_offset, mel_offset+n_mel) */
    int n_mel, mel_offset;
    int finished;
    float *h_mel; int h_mel_cap; int h_mel_valid_from, h_mel_valid_n;  /* lazy host mirror for vox_mel_data */
    float *h_stage; int h_stage_cap;                                   /* pinned staging for feeds */
};

static void mel_reserve_samples(vox_mel_ctx *c, int extra) {
    if (c->n_local + extra <= c->cap_local) return;
    int ncap = c->cap_local ? c->cap_local : 65536;
    while (ncap < c->n_local + extra) ncap *= 2;
    float *nd = (float *)vb_dev_alloc((size_t)ncap * 4);
    VB_CUDA_OK(cudaMemcpyAsync(nd, c->d_samples, (size_t)c->n_local * 4, cudaMemcpyDeviceToDevice, c->e->stream));
    VB_CUDA_OK(cudaStreamSynchronize(c->e->stream));
    cudaFree(c->d_samples);
    c->d_samples = nd; c->cap_local = ncap;
}

static void mel_reserve_frames(vox_mel_ctx *c, int extra) {
    if (c->n_mel + extra <= c->mel_cap) return;
    int ncap = c->mel_cap ? c->mel_cap : 1024;
    while (ncap < c->n_mel + extra) ncap *= 2;
    float *nd = (float *)vb_dev_alloc((size_t)ncap * N_MEL * 4);
    if (c->n_mel) VB_CUDA_OK(cudaMemcpyAsync(nd, c->d_mel, (size_t)c->n_mel * N_MEL * 4, cudaMemcpyDeviceToDevice, c->e->stream));
    VB_CUDA_OK(cudaStreamSynchronize(c->e->stream));
    cudaFree(c->d_mel);
    c->d_mel = nd; c->mel_cap = ncap;
}

/* Compute every frame whose 400-sample window is available; then drop dead samples. */
static int mel_compute_ready(vox_mel_ctx *c) {
    long long next = (long long)c->mel_offset + c->n_mel;            /* next global frame */
    long long avail = (c->n_total - N_FFT) / HOP + 1;                 /* frames that fit */
    if (c->n_total < N_FFT) avail = 0;
    int todo = (int)(avail - next);
    if (todo > 0) {
        mel_reserve_frames(c, todo);
        long long first = next * HOP - c->sample_base;
        int blocks = (todo + FRAMES_PER_CTA - 1) / FRAMES_PER_CTA;
        k_mel_frames<<<blocks, 256, 0, c->e->stream>>>(c->d_samples, first, todo, c->tab->cosT, c->tab->sinT,
                                                       c->tab->filtT, c->tab->window, c->d_mel + (size_t)c->n_mel * N_MEL);
        VB_CUDA_OK(cudaGetLastError());
        c->e->launches++;
        c->n_mel += todo;
        next += todo;
    } else todo = 0;
    /* retire samples no future frame (or the finish-time reflection) can touch */
    long long keep_from = next * HOP;
    if (keep_from > c->n_total - KEEP_TAIL) keep_from = c->n_total - KEEP_TAIL;
    if (keep_from > c->sample_base + 65536) {                        /* amortise: at most once per ~4 s of audio */
        int drop = (int)(keep_from - c->sample_base), remain = c->n_local - drop;
        float *tmp = vb_ws(c->e, 11, (size_t)remain * 4);
        VB_CUDA_OK(cudaMemcpyAsync(tmp, c->d_samples + drop, (size_t)remain * 4, cudaMemcpyDeviceToDevice, c->e->stream));
        VB_CUDA_OK(cudaMemcpyAsync(c->d_samples, tmp, (size_t)remain * 4, cudaMemcpyDeviceToDevice, c->e->stream));
        c->n_local = remain; c->sample_base = keep_from;
    }
    return todo;
}

extern "C" {

vox_mel_ctx_t *vb_mel_ctx_init_on(VbEngine *e, int left_pad_samples);

vox_mel_ctx_t *vox_mel_ctx_init(int left_pad_samples) {
    vb_require_gpu("vox_mel_ctx_init");
    return vb_mel_ctx_init_on(vb_default_engine(), left_pad_samples);
}

vox_mel_ctx_t *vb_mel_ctx_init_on(VbEngine *e, int left_pad_samples) {
    vox_mel_ctx *c = (vox_mel_ctx *)calloc(1, sizeof *c);
    c->e = e;
    VB_CUDA_OK(cudaSetDevice(c->e->device));
    c->tab = mel_tables(c->e->device);
    int left = 200 + left_pad_samples;                               /* voxtral_audio.c:544-545 */
    mel_reserve_samples(c, left + 16000);
    VB_CUDA_OK(cudaMemsetAsync(c->d_samples, 0, (size_t)left * 4, c->e->stream));
    c->n_local = left; c->n_total = left; c->sample_base = 0;
    return c;
}

static void mel_append(vox_mel_ctx *c, const float *host, int n, int zeros) {
    mel_reserve_samples(c, n);
    if (zeros == 2) {                                                /* `host` is really a device pointer */
        VB_CUDA_OK(cudaMemcpyAsync(c->d_samples + c->n_local, host, (size_t)n * 4, cudaMemcpyDeviceToDevice, c->e->stream));
    } else if (zeros) {
        VB_CUDA_OK(cudaMemsetAsync(c->d_samples + c->n_local, 0, (size_t)n * 4, c->e->stream));
    } else {
        if (n > c->h_stage_cap) {
            if (c->h_stage) cudaFreeHost(c->h_stage);
            c->h_stage_cap = n + n / 4 + 4096;
            VB_CUDA_OK(cudaMallocHost((void **)&c->h_stage, (size_t)c->h_stage_cap * 4));
        }
        VB_CUDA_OK(cudaStreamSynchronize(c->e->stream));             /* staging buffer may still be in flight */
        memcpy(c->h_stage, host, (size_t)n * 4);
        VB_CUDA_OK(cudaMemcpyAsync(c->d_samples + c->n_local, c->h_stage, (size_t)n * 4, cudaMemcpyHostToDevice, c->e->stream));
    }
    c->n_local += n; c->n_total += n;
}

int vox_mel_feed(vox_mel_ctx_t *c, const float *samples, int n_samples) {
    if (!c || n_samples <= 0) return 0;
    mel_append(c, samples, n_samples, 0);
    return mel_compute_ready(c);
}

int vb_mel_feed_device(vox_mel_ctx_t *c, const float *d_samples, int n) {
    if (!c || n <= 0) return 0;
    mel_append(c, d_samples, n, 2);
    return mel_compute_ready(c);
}

/* device-to-device variant used by the stream path for zero padding */
int vb_mel_feed_zeros(vox_mel_ctx_t *c, int n) {
    if (!c || n <= 0) return 0;
    mel_append(c, NULL, n, 1);
    return mel_compute_ready(c);
}

__global__ void k_reflect_tail(float *s, int n_local, int real_end_local, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int src = real_end_local - 2 - i;
    s[n_local + i] = src >= 0 ? s[src] : 0.f;
}

int vox_mel_finish(vox_mel_ctx_t *c, int right_pad_samples) {
    if (!c) return 0;
    if (c->finished) return c->n_mel;
    if (right_pad_samples > 0) mel_append(c, NULL, right_pad_samples, 1);
    /* 200-sample reflection about the end of the signal before the explicit right pad
     * (voxtral_audio.c:602-618): s[n+i] = s[real_end-2-i] */
    mel_reserve_samples(c, 200);
    int real_end_local = c->n_local - right_pad_samples;
    k_reflect_tail<<<1, 256, 0, c->e->stream>>>(c->d_samples, c->n_local, real_end_local, 200);
    c->e->launches++;
    c->n_local += 200; c->n_total += 200;
    mel_compute_ready(c);
    if (c->n_mel > 0) c->n_mel--;                                    /* drop last frame (:627-628) */
    c->finished = 1;
    return c->n_mel;
}

float *vox_mel_data(vox_mel_ctx_t *c, int *out_n_frames) {
    if (!c) { if (out_n_frames) *out_n_frames = 0; return NULL; }
    if (out_n_frames) *out_n_frames = c->n_mel;
    if (c->n_mel > c->h_mel_cap) {
        c->h_mel_cap = c->n_mel * 2 + 256;
        c->h_mel = (float *)realloc(c->h_mel, (size_t)c->h_mel_cap * N_MEL * 4);
    }
    if (c->n_mel > 0) {
        VB_CUDA_OK(cudaMemcpyAsync(c->h_mel, c->d_mel, (size_t)c->n_mel * N_MEL * 4, cudaMemcpyDeviceToHost, c->e->stream));
        VB_CUDA_OK(cudaStreamSynchronize(c->e->stream));
    }
    return c->h_mel;
}

int vox_mel_frame_offset(vox_mel_ctx_t *c) { return c ? c->mel_offset : 0; }

void vox_mel_discard_before(vox_mel_ctx_t *c, int keep_from_frame) {
    if (!c || keep_from_frame <= c->mel_offset) return;
    int drop = keep_from_frame - c->mel_offset;
    if (drop > c->n_mel) drop = c->n_mel;
    if (drop <= 0) return;
    int remain = c->n_mel - drop;
    if (remain > 0) {
        float *tmp = vb_ws(c->e, 11, (size_t)remain * N_MEL * 4);
        VB_CUDA_OK(cudaMemcpyAsync(tmp, c->d_mel + (size_t)drop * N_MEL, (size_t)remain * N_MEL * 4, cudaMemcpyDeviceToDevice, c->e->stream));
        VB_CUDA_OK(cudaMemcpyAsync(c->d_mel, tmp, (size_t)remain * N_MEL * 4, cudaMemcpyDeviceToDevice, c->e->stream));
    }
    c->n_mel = remain; c->mel_offset += drop;
}

void vox_mel_free(vox_mel_ctx_t *c) {
    if (!c) return;
    cudaStreamSynchronize(c->e->stream);
    cudaFree(c->d_samples); cudaFree(c->d_mel);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    free(c->h_mel);
    free(c);
}

/* device accessors for the stream path */
float *vb_mel_dev_frames(vox_mel_ctx_t *c, int *n_frames, int *frame_offset) {
    if (n_frames) *n_frames = c->n_mel;
    if (frame_offset) *frame_offset = c->mel_offset;
    return c->d_mel;
}

/* Batch spectrogram with reflect padding on both sides and the last STFT frame dropped
 * (vox_mel_spectrogram, voxtral_audio.c:294-399; not used by the stream path). */
float *vox_mel_spectrogram(const float *samples, int n_samples, int *out_frames) {
    vb_require_gpu("vox_mel_spectrogram");
    VbEngine *e = vb_default_engine();
    VB_CUDA_OK(cudaSetDevice(e->device));
    const int pad = N_FFT / 2, padded = n_samples + 2 * pad;
    int total = (padded - N_FFT) / HOP + 1, frames = total - 1;
    if (frames <= 0) {
        fprintf(stderr, "vox_mel_spectrogram: audio too short (%d samples)\n", n_samples);
        return NULL;
    }
    float *hp = (float *)malloc((size_t)padded * 4);
    for (int i = 0; i < pad; i++) { int s = pad - i; hp[i] = s < n_samples ? samples[s] : 0.0f; }
    memcpy(hp + pad, samples, (size_t)n_samples * 4);
    for (int i = 0; i < pad; i++) { int s = n_samples - 2 - i; hp[pad + n_samples + i] = s >= 0 ? samples[s] : 0.0f; }
    const MelTables *t = mel_tables(e->device);
    float *ds = (float *)vb_dev_alloc((size_t)padded * 4), *dm = (float *)vb_dev_alloc((size_t)frames * N_MEL * 4);
    VB_CUDA_OK(cudaMemcpyAsync(ds, hp, (size_t)padded * 4, cudaMemcpyHostToDevice, e->stream));
    k_mel_frames<<<(frames + FRAMES_PER_CTA - 1) / FRAMES_PER_CTA, 256, 0, e->stream>>>(ds, 0, frames, t->cosT, t->sinT,
                                                                                      t->filtT, t->window, dm);
    e->launches++;
    float *out = (float *)malloc((size_t)frames * N_MEL * 4);
    VB_CUDA_OK(cudaMemcpyAsync(out, dm, (size_t)frames * N_MEL * 4, cudaMemcpyDeviceToHost, e->stream));
    VB_CUDA_OK(cudaStreamSynchronize(e->stream));
    cudaFree(ds); cudaFree(dm); free(hp);
    *out_frames = frames;
    return out;
}

}  /* extern "C" */

/* ---- frames [f0, f1) of a COMPLETE recording, as the stream path would produce them (vb_dist.c) ----
 * The padded signal the stream path builds for a whole recording is [200 + 32*1280 zeros | pcm | align + (delay+1+10)*1280 zeros |
 * 200 reflected samples (= zeros: they mirror the zero padding)] (voxtral.c:1203,1593-1606; voxtral_audio.c:544-545,602-629) and
 * frame t reads padded samples [160 t, 160 t + 400).  A rank of a sequence-sharded run needs only its own frames, so only its
 * slice of the PCM crosses PCIe. */
extern "C" int vb_mel_recording_frames(int n_samples, int delay_tokens) {
    const long long align = (1280 - n_samples % 1280) % 1280;
    const long long total = 200 + 32 * 1280 + (long long)n_samples + align + (long long)(delay_tokens + 1 + 10) * 1280 + 200;
    return (int)((total - N_FFT) / HOP + 1 - 1);                      /* the last frame is dropped (voxtral_audio.c:627-628) */
}

extern "C" void vb_mel_recording_range(VbEngine *e, const float *pcm_host, int n_samples, int f0, int f1, float *d_out) {
    if (f1 <= f0) return;
    const MelTables *tab = mel_tables(e->device);
    const long long left = 200 + 32 * 1280;
    const long long s0 = (long long)f0 * HOP, s1 = (long long)(f1 - 1) * HOP + N_FFT;   /* padded samples [s0, s1) */
    const size_t len = (size_t)(s1 - s0);
    float *ds = vb_ws(e, 9, len * 4);
    VB_CUDA_OK(cudaMemsetAsync(ds, 0, len * 4, e->stream));
    long long a = s0 > left ? s0 : left, b = s1 < left + n_samples ? s1 : left + n_samples;   /* the part that is real audio */
    if (b > a) VB_CUDA_OK(cudaMemcpyAsync(ds + (a - s0), pcm_host + (a - left), (size_t)(b - a) * 4, cudaMemcpyHostToDevice, e->stream));
    const int n = f1 - f0, blocks = (n + FRAMES_PER_CTA - 1) / FRAMES_PER_CTA;
    k_mel_frames<<<blocks, 256, 0, e->stream>>>(ds, 0, n, tab->cosT, tab->sinT, tab->filtT, tab->window, d_out);
    VB_CUDA_OK(cudaGetLastError());
    e->launches++;
}

