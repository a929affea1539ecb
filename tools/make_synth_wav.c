/*
 * make_synth_wav -- seeded synthetic 16 kHz mono s16 PCM, written as a WAV so it
 * goes through vox_load_wav like any other input.
 *
 * Base signal is the one SURVEY.md section 8d specifies
 *   0.25*sin(2*pi*220*t)*(0.5+0.5*sin(2*pi*3*t)) + 0.05*u      (u ~ U(-1,1), xorshift32)
 * That alone is stationary (every 80 ms token sees the same spectrum, and the
 * reference then emits one repeated token), so a "syllable" layer is added on
 * top: consecutive segments of 40..240 ms, each either silent or a 3-harmonic
 * tone with its own pitch (90..420 Hz), a linear pitch glide and a raised-cosine
 * envelope.  All parameters come from the same xorshift32 stream, so the file is
 * a pure function of (seconds, seed).
 *
 * Usage: make_synth_wav <out.wav> <seconds> [seed_hex]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint32_t rng;
static uint32_t xs32(void) { rng ^= rng << 13; rng ^= rng >> 17; rng ^= rng << 5; return rng; }
static double u01(void) { return (double)xs32() / 4294967296.0; }

static void le32(uint8_t *p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }
static void le16(uint8_t *p, uint16_t v) { p[0] = v; p[1] = v >> 8; }

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <out.wav> <seconds> [seed_hex]\n", argv[0]); return 2; }
    double secs = atof(argv[2]);
    rng = argc > 3 ? (uint32_t)strtoul(argv[3], NULL, 16) : 0x5EED1234u;
    long n = (long)llround(secs * 16000.0);
    int16_t *pcm = malloc((size_t)n * 2 + 2);
    const double w1 = 2.0 * M_PI * 220.0 / 16000.0, w2 = 2.0 * M_PI * 3.0 / 16000.0;

    long seg_left = 0, seg_len = 1;
    double f0 = 0, glide = 0, amp = 0, phase = 0, h2 = 0, h3 = 0;
    for (long i = 0; i < n; i++) {
        if (seg_left == 0) {
            seg_len = seg_left = 640 + (long)(u01() * 3200.0);
            int voiced = u01() < 0.8;
            f0 = 90.0 + 330.0 * u01();
            glide = (u01() - 0.5) * 200.0;               /* Hz over the segment */
            amp = voiced ? 0.05 + 0.3 * u01() : 0.0;
            h2 = u01(); h3 = u01();
            phase = 0;
        }
        double pos = (double)(seg_len - seg_left) / (double)seg_len;
        double f = f0 + glide * pos;
        phase += 2.0 * M_PI * f / 16000.0;
        double env = 0.5 - 0.5 * cos(2.0 * M_PI * pos);
        double syl = amp * env * (sin(phase) + 0.5 * h2 * sin(2 * phase) + 0.33 * h3 * sin(3 * phase)) / 1.6;
        seg_left--;

        double u = u01() * 2.0 - 1.0;
        double x = 0.1 * sin(w1 * (double)i) * (0.5 + 0.5 * sin(w2 * (double)i)) + 0.02 * u + syl;
        long q = lround(x * 32767.0);
        if (q > 32767) q = 32767;
        if (q < -32768) q = -32768;
        pcm[i] = (int16_t)q;
    }
    uint8_t h[44];
    memcpy(h, "RIFF", 4); le32(h + 4, (uint32_t)(36 + n * 2)); memcpy(h + 8, "WAVEfmt ", 8);
    le32(h + 16, 16); le16(h + 20, 1); le16(h + 22, 1); le32(h + 24, 16000); le32(h + 28, 32000);
    le16(h + 32, 2); le16(h + 34, 16); memcpy(h + 36, "data", 4); le32(h + 40, (uint32_t)(n * 2));
    FILE *f = fopen(argv[1], "wb");
    if (!f) { perror(argv[1]); return 1; }
    fwrite(h, 1, 44, f); fwrite(pcm, 2, (size_t)n, f); fclose(f);
    return 0;
}
