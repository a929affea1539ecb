/*
 * vb_wav.c -- host-side audio I/O (RIFF/WAVE parsing, stdin PCM, mic stubs).
 *
 * Replaces the I/O half of /root/reference voxtral_audio.c:43-217 and the
 * non-Apple stubs of voxtral_mic_macos.c:124-142.  SURVEY.md section 2 rows
 * 2 and 12 mark these OUT OF SCOPE as kernels (microseconds of host work) but
 * the symbols are part of the link surface main.c needs, so they stay in C.
 *
 * Semantics kept from the reference: 16-bit PCM only; multi-channel input is
 * averaged to mono; other sample rates are converted with the reference's
 * 2-tap linear interpolation (position computed in float, voxtral_audio.c:110-137);
 * a `data` chunk whose size is 0/-1/overlong means "rest of file" (piped ffmpeg).
 */
#include "voxtral_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int vox_verbose_audio = 0;

static unsigned rd16(const uint8_t *p) { return (unsigned)p[0] | ((unsigned)p[1] << 8); }
static uint32_t rd32(const uint8_t *p) { return (uint32_t)rd16(p) | ((uint32_t)rd16(p + 2) << 16); }

typedef struct { int fmt, channels, rate, bits; const uint8_t *pcm; long pcm_bytes; } wav_info;

static int wav_scan(const uint8_t *buf, size_t size, wav_info *w) {
    memset(w, 0, sizeof *w);
    if (size < 44 || memcmp(buf, "RIFF", 4) || memcmp(buf + 8, "WAVE", 4)) return -1;
    size_t off = 12;
    while (off + 8 <= size) {
        const uint8_t *ck = buf + off;
        uint32_t len = rd32(ck + 4);
        if (!memcmp(ck, "data", 4)) {
            w->pcm = ck + 8;
            long avail = (long)(size - off - 8);
            long want = (long)(int32_t)len;
            w->pcm_bytes = (want <= 0 || want > avail) ? avail : want;
            break;
        }
        if (!memcmp(ck, "fmt ", 4) && len >= 16 && off + 8 + len <= size) {
            w->fmt = (int)rd16(ck + 8);
            w->channels = (int)rd16(ck + 10);
            w->rate = (int)rd32(ck + 12);
            w->bits = (int)rd16(ck + 22);
        }
        if (off + 8 + (size_t)len > size) break;
        off += 8 + (size_t)len + (len & 1u);
    }
    return 0;
}

static float *pcm16_to_mono(const uint8_t *pcm, int frames, int channels) {
    float *out = malloc(sizeof(float) * (size_t)(frames > 0 ? frames : 1));
    if (!out) return NULL;
    for (int i = 0; i < frames; i++) {
        if (channels == 1) {
            out[i] = (float)(int16_t)rd16(pcm + 2 * (size_t)i) / 32768.0f;
        } else {
            float acc = 0;
            for (int c = 0; c < channels; c++)
                acc += (float)(int16_t)rd16(pcm + 2 * ((size_t)i * channels + c));
            out[i] = (acc / (float)channels) / 32768.0f;
        }
    }
    return out;
}

static float *to_16k(float *in, int n_in, int rate, int *n_out) {
    int n = (int)((long long)n_in * VOX_SAMPLE_RATE / rate);
    float *out = malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    if (!out) { free(in); return NULL; }
    for (int i = 0; i < n; i++) {
        float pos = (float)i * rate / VOX_SAMPLE_RATE;
        int k = (int)pos;
        float a = pos - k;
        if (k + 1 < n_in) out[i] = in[k] * (1.0f - a) + in[k + 1] * a;
        else out[i] = k < n_in ? in[k] : 0.0f;
    }
    free(in);
    *n_out = n;
    if (vox_verbose_audio)
        fprintf(stderr, "  Resampled %d -> %d Hz (%d samples)\n", rate, VOX_SAMPLE_RATE, n);
    return out;
}

float *vox_parse_wav_buffer(const uint8_t *data, size_t size, int *out_n_samples) {
    wav_info w;
    if (wav_scan(data, size, &w) != 0) {
        fprintf(stderr, "parse_wav_buffer: not a valid WAV file\n");
        return NULL;
    }
    if (w.fmt != 1 || w.bits != 16 || !w.pcm || w.channels < 1) {
        fprintf(stderr, "parse_wav_buffer: unsupported format (need 16-bit PCM, got fmt=%d bits=%d)\n",
                w.fmt, w.bits);
        return NULL;
    }
    int frames = (int)(w.pcm_bytes / (w.channels * 2));
    float *mono = pcm16_to_mono(w.pcm, frames, w.channels);
    if (!mono) return NULL;
    if (w.rate != VOX_SAMPLE_RATE) mono = to_16k(mono, frames, w.rate, &frames);
    if (mono) *out_n_samples = frames;
    return mono;
}

static uint8_t *slurp(FILE *f, size_t *size) {
    size_t cap = 1u << 20, n = 0;
    uint8_t *buf = malloc(cap);
    while (buf) {
        size_t got = fread(buf + n, 1, cap - n, f);
        n += got;
        if (got == 0) break;
        if (n == cap) {
            uint8_t *nb = realloc(buf, cap *= 2);
            if (!nb) { free(buf); return NULL; }
            buf = nb;
        }
    }
    *size = n;
    return buf;
}

float *vox_load_wav(const char *path, int *out_n_samples) {
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "vox_load_wav: cannot open %s\n", path); return NULL; }
    size_t size = 0;
    uint8_t *buf = slurp(f, &size);
    fclose(f);
    if (!buf || size == 0) { free(buf); return NULL; }
    float *s = vox_parse_wav_buffer(buf, size, out_n_samples);
    free(buf);
    return s;
}

float *vox_read_pcm_stdin(int *out_n_samples) {
    size_t size = 0;
    uint8_t *buf = slurp(stdin, &size);
    if (!buf) return NULL;
    if (size < 4) { fprintf(stderr, "vox_read_pcm_stdin: no data on stdin\n"); free(buf); return NULL; }
    fprintf(stderr, "Read %zu bytes from stdin\n", size);
    float *s;
    if (!memcmp(buf, "RIFF", 4)) {
        fprintf(stderr, "Detected WAV format on stdin\n");
        s = vox_parse_wav_buffer(buf, size, out_n_samples);
    } else {
        fprintf(stderr, "Treating stdin as raw s16le 16kHz mono\n");
        int frames = (int)(size / 2);
        s = pcm16_to_mono(buf, frames, 1);
        if (s) *out_n_samples = frames;
    }
    free(buf);
    return s;
}

/* Microphone capture exists only on macOS in the reference; everywhere else it
 * ships these exact failure semantics (voxtral_mic_macos.c:124-142). */
int vox_mic_start(void) {
    fprintf(stderr, "Microphone capture is not supported on this platform\n");
    return -1;
}
int vox_mic_read(float *out, int max_samples) { (void)out; (void)max_samples; return 0; }
int vox_mic_read_available(void) { return 0; }
void vox_mic_stop(void) {}
