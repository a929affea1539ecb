/*
 * vb_tokenizer.c -- Tekken id -> text table (decode only).
 *
 * Replaces /root/reference voxtral_tokenizer.c (API voxtral_tokenizer.h:16-34).
 * Behaviour kept: ids 0..999 come from special_tokens[].token_str, ids >= 1000
 * from base64(vocab[id-1000].token_bytes); pieces are C strings, so a piece
 * whose first byte is 0x00 (real rank 0) decodes to "" and is treated as a
 * non-text token by the stream code (voxtral.c:489-497); unknown ids -> NULL.
 *
 * Storage differs from the reference: all pieces live in one arena and the
 * lookup tables hold offsets, so load does 3 allocations instead of 131k.
 */
#include "voxtral_b200.h"
#include "vb_json.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define TK_SPECIAL 1000
#define TK_VOCAB   130072
#define TK_PIECE_MAX 255     /* decoded bytes kept per piece, as in the reference */

struct vox_tokenizer {
    char *arena; size_t used, cap;
    int32_t *vocab_off;      /* [TK_VOCAB]   -1 = absent */
    int32_t *special_off;    /* [TK_SPECIAL] -1 = absent */
    int n_vocab, n_special;
};

static int32_t arena_put(vox_tokenizer_t *t, const char *s, size_t n) {
    if (t->used + n + 1 > t->cap) {
        size_t nc = t->cap ? t->cap * 2 : (1u << 20);
        while (nc < t->used + n + 1) nc *= 2;
        char *na = realloc(t->arena, nc);
        if (!na) return -1;
        t->arena = na; t->cap = nc;
    }
    memcpy(t->arena + t->used, s, n);
    t->arena[t->used + n] = 0;
    int32_t off = (int32_t)t->used;
    t->used += n + 1;
    return off;
}

static int b64val(unsigned char c) {
    if (c >= 'A' && c <= 'Z') return c - 'A';
    if (c >= 'a' && c <= 'z') return c - 'a' + 26;
    if (c >= '0' && c <= '9') return c - '0' + 52;
    if (c == '+') return 62;
    if (c == '/') return 63;
    return -1;
}

static size_t b64decode(const char *in, char *out, size_t cap) {
    unsigned acc = 0; int bits = 0; size_t n = 0;
    for (; *in && *in != '='; in++) {
        int v = b64val((unsigned char)*in);
        if (v < 0) continue;
        acc = (acc << 6) | (unsigned)v; bits += 6;
        if (bits >= 8) { bits -= 8; if (n < cap) out[n++] = (char)((acc >> bits) & 0xFF); }
    }
    return n;
}

/* Walk an array of objects, handing (rank, string field) pairs to the table. */
static void load_table(jcur *c, vox_tokenizer_t *t, const char *field, int is_vocab) {
    if (!j_eat(c, '[')) { c->ok = 0; return; }
    while (c->ok && !j_eat(c, ']')) {
        if (j_eat(c, ',')) continue;
        if (!j_eat(c, '{')) { c->ok = 0; return; }
        long long rank = -1;
        char val[512]; int have = 0;
        while (c->ok && !j_eat(c, '}')) {
            if (j_eat(c, ',')) continue;
            char key[40];
            if (j_string(c, key, sizeof key) < 0 || !j_eat(c, ':')) { c->ok = 0; return; }
            if (!strcmp(key, "rank")) rank = j_int(c);
            else if (!strcmp(key, field) && j_peek(c) == '"') { j_string(c, val, sizeof val); have = 1; }
            else j_skip(c);
        }
        if (!have || !val[0] || rank < 0) continue;
        if (is_vocab && rank < TK_VOCAB) {
            char raw[TK_PIECE_MAX + 1];
            size_t n = b64decode(val, raw, TK_PIECE_MAX);
            t->vocab_off[rank] = arena_put(t, raw, n);
            if (rank + 1 > t->n_vocab) t->n_vocab = (int)rank + 1;
        } else if (!is_vocab && rank < TK_SPECIAL) {
            t->special_off[rank] = arena_put(t, val, strlen(val));
            if (rank + 1 > t->n_special) t->n_special = (int)rank + 1;
        }
    }
}

vox_tokenizer_t *vox_tokenizer_load(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "tokenizer: cannot open %s\n", path); return NULL; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *doc = sz > 0 ? malloc((size_t)sz + 1) : NULL;
    if (!doc || fread(doc, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(doc); return NULL; }
    fclose(f);
    doc[sz] = 0;

    vox_tokenizer_t *t = calloc(1, sizeof *t);
    t->vocab_off = malloc(sizeof(int32_t) * TK_VOCAB);
    t->special_off = malloc(sizeof(int32_t) * TK_SPECIAL);
    memset(t->vocab_off, 0xFF, sizeof(int32_t) * TK_VOCAB);
    memset(t->special_off, 0xFF, sizeof(int32_t) * TK_SPECIAL);

    jcur c = { doc, doc + sz, 1 };
    if (!j_eat(&c, '{')) c.ok = 0;
    while (c.ok && !j_eat(&c, '}')) {
        if (j_eat(&c, ',')) continue;
        char key[64];
        if (j_string(&c, key, sizeof key) < 0 || !j_eat(&c, ':')) { c.ok = 0; break; }
        if (!strcmp(key, "vocab")) load_table(&c, t, "token_bytes", 1);
        else if (!strcmp(key, "special_tokens")) load_table(&c, t, "token_str", 0);
        else j_skip(&c);
    }
    free(doc);
    if (!c.ok) { vox_tokenizer_free(t); return NULL; }
    if (vox_verbose >= 2)
        fprintf(stderr, "Tokenizer: %d vocab + %d special tokens\n", t->n_vocab, t->n_special);
    return t;
}

void vox_tokenizer_free(vox_tokenizer_t *t) {
    if (!t) return;
    free(t->arena); free(t->vocab_off); free(t->special_off); free(t);
}

const char *vox_tokenizer_decode(vox_tokenizer_t *t, int id) {
    if (!t) return NULL;
    if (id >= TK_SPECIAL && id < TK_SPECIAL + t->n_vocab) {
        int32_t o = t->vocab_off[id - TK_SPECIAL];
        return o >= 0 ? t->arena + o : NULL;
    }
    if (id >= 0 && id < t->n_special) {
        int32_t o = t->special_off[id];
        return o >= 0 ? t->arena + o : NULL;
    }
    return NULL;
}

char *vox_tokenizer_decode_seq(vox_tokenizer_t *t, const int *ids, int n) {
    size_t cap = 64, len = 0;
    char *out = malloc(cap);
    out[0] = 0;
    for (int i = 0; i < n; i++) {
        if (ids[i] >= 0 && ids[i] < TK_SPECIAL) continue;     /* control tokens are not text */
        const char *p = vox_tokenizer_decode(t, ids[i]);
        if (!p) continue;
        size_t l = strlen(p);
        if (len + l + 1 > cap) { while (len + l + 1 > cap) cap *= 2; out = realloc(out, cap); }
        memcpy(out + len, p, l + 1);
        len += l;
    }
    return out;
}

int vox_tokenizer_bos(vox_tokenizer_t *t) { (void)t; return 1; }
int vox_tokenizer_eos(vox_tokenizer_t *t) { (void)t; return 2; }
int vox_tokenizer_vocab_size(vox_tokenizer_t *t) { (void)t; return TK_SPECIAL + TK_VOCAB; }
