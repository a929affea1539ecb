/*
 * vb_json.h -- a forward-only JSON cursor, just enough for the two JSON
 * documents the engine reads (the safetensors header and tekken.json).
 * String values: \" \\ \/ \n \t \r \b \f and \uXXXX (BMP, emitted as UTF-8) are decoded.
 */
#ifndef VB_JSON_H
#define VB_JSON_H
#include <stdlib.h>
#include <string.h>

typedef struct { const char *p, *end; int ok; } jcur;

static inline void j_ws(jcur *c) {
    while (c->p < c->end && (*c->p == ' ' || *c->p == '\n' || *c->p == '\t' || *c->p == '\r')) c->p++;
}
static inline int j_eat(jcur *c, char ch) {
    j_ws(c);
    if (c->p < c->end && *c->p == ch) { c->p++; return 1; }
    return 0;
}
static inline int j_peek(jcur *c) { j_ws(c); return c->p < c->end ? (unsigned char)*c->p : -1; }

/* Decode a JSON string into dst (NUL-terminated, truncating at cap-1). Returns decoded length or -1. */
static inline int j_string(jcur *c, char *dst, size_t cap) {
    j_ws(c);
    if (c->p >= c->end || *c->p != '"') { c->ok = 0; return -1; }
    c->p++;
    size_t n = 0;
#define J_PUT(b) do { if (dst && n + 1 < cap) dst[n] = (char)(b); n++; } while (0)
    while (c->p < c->end && *c->p != '"') {
        unsigned char ch = (unsigned char)*c->p++;
        if (ch != '\\') { J_PUT(ch); continue; }
        if (c->p >= c->end) break;
        ch = (unsigned char)*c->p++;
        switch (ch) {
        case 'n': J_PUT('\n'); break;
        case 't': J_PUT('\t'); break;
        case 'r': J_PUT('\r'); break;
        case 'b': J_PUT('\b'); break;
        case 'f': J_PUT('\f'); break;
        case 'u': {
            unsigned cp = 0;
            for (int i = 0; i < 4 && c->p < c->end; i++, c->p++) {
                char h = *c->p;
                cp = cp * 16 + (unsigned)(h >= '0' && h <= '9' ? h - '0' : (h | 32) - 'a' + 10);
            }
            if (cp < 0x80) J_PUT(cp);
            else if (cp < 0x800) { J_PUT(0xC0 | (cp >> 6)); J_PUT(0x80 | (cp & 63)); }
            else { J_PUT(0xE0 | (cp >> 12)); J_PUT(0x80 | ((cp >> 6) & 63)); J_PUT(0x80 | (cp & 63)); }
            break;
        }
        default: J_PUT(ch);
        }
    }
#undef J_PUT
    if (dst && cap) dst[n < cap ? n : cap - 1] = 0;
    if (c->p >= c->end) { c->ok = 0; return -1; }
    c->p++;
    return (int)(n < cap || !dst ? n : cap - 1);
}
static inline long long j_int(jcur *c) {
    j_ws(c);
    char *e = NULL;
    long long v = strtoll(c->p, &e, 10);
    if (e == c->p) c->ok = 0;
    c->p = e;
    return v;
}
static inline void j_skip(jcur *c) {   /* skip any value */
    j_ws(c);
    if (c->p >= c->end) { c->ok = 0; return; }
    char ch = *c->p;
    if (ch == '"') { j_string(c, NULL, 0); return; }
    if (ch == '{' || ch == '[') {
        char close = ch == '{' ? '}' : ']';
        c->p++;
        for (;;) {
            j_ws(c);
            if (c->p >= c->end) { c->ok = 0; return; }
            if (*c->p == close) { c->p++; return; }
            if (*c->p == ',' || *c->p == ':') { c->p++; continue; }
            j_skip(c);
            if (!c->ok) return;
        }
    }
    while (c->p < c->end && *c->p != ',' && *c->p != '}' && *c->p != ']') c->p++;
}
#endif
