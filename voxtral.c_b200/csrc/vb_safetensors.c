/*
 * vb_safetensors.c -- mmap reader for `consolidated.safetensors`.
 *
 * Replaces /root/reference voxtral_safetensors.c (API: voxtral_safetensors.h:52-85).
 * Same observable behaviour (mmap PROT_READ/MAP_PRIVATE of the whole file,
 * <=1024 tensors, linear name lookup, F16/BF16 -> F32 widening), different
 * implementation: the header is scanned by a small recursive-descent JSON
 * walker that only understands what the safetensors header can contain.
 */
#define _GNU_SOURCE
#include "voxtral_b200.h"

#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "vb_json.h"

static safetensor_dtype_t dtype_of(const char *s) {
    if (!strcmp(s, "F32")) return DTYPE_F32;
    if (!strcmp(s, "F16")) return DTYPE_F16;
    if (!strcmp(s, "BF16")) return DTYPE_BF16;
    if (!strcmp(s, "I32")) return DTYPE_I32;
    if (!strcmp(s, "I64")) return DTYPE_I64;
    if (!strcmp(s, "BOOL")) return DTYPE_BOOL;
    return DTYPE_UNKNOWN;
}

/* One `"name": {"dtype":..,"shape":[..],"data_offsets":[a,b]}` entry. */
static int parse_entry(jcur *c, safetensor_t *t) {
    memset(t->shape, 0, sizeof t->shape);
    t->ndim = 0; t->dtype = DTYPE_UNKNOWN; t->data_offset = t->data_size = 0;
    if (!j_eat(c, '{')) return 0;
    while (c->ok && !j_eat(c, '}')) {
        char key[32];
        if (j_eat(c, ',')) continue;
        if (j_string(c, key, sizeof key) < 0 || !j_eat(c, ':')) return 0;
        if (!strcmp(key, "dtype")) {
            char v[16];
            if (j_string(c, v, sizeof v) < 0) return 0;
            t->dtype = dtype_of(v);
        } else if (!strcmp(key, "shape")) {
            if (!j_eat(c, '[')) return 0;
            while (c->ok && !j_eat(c, ']')) {
                if (j_eat(c, ',')) continue;
                long long d = j_int(c);
                if (t->ndim < 8) t->shape[t->ndim++] = d;
            }
        } else if (!strcmp(key, "data_offsets")) {
            if (!j_eat(c, '[')) return 0;
            long long a = j_int(c); j_eat(c, ',');
            long long b = j_int(c);
            if (!j_eat(c, ']')) return 0;
            t->data_offset = (size_t)a;
            t->data_size = (size_t)(b - a);
        } else {
            j_skip(c);
        }
    }
    return c->ok;
}

safetensors_file_t *safetensors_open(const char *path) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return NULL;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 8) { close(fd); return NULL; }
    void *map = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return NULL;

    uint64_t hlen;
    memcpy(&hlen, map, 8);
    if (hlen == 0 || hlen > (uint64_t)st.st_size - 8) { munmap(map, (size_t)st.st_size); return NULL; }

    safetensors_file_t *sf = calloc(1, sizeof *sf);
    if (!sf) { munmap(map, (size_t)st.st_size); return NULL; }
    sf->path = strdup(path);
    sf->data = map;
    sf->file_size = (size_t)st.st_size;
    sf->header_size = (size_t)hlen;
    sf->header_json = malloc(hlen + 1);
    memcpy(sf->header_json, (const char *)map + 8, hlen);
    sf->header_json[hlen] = 0;

    jcur c = { sf->header_json, sf->header_json + hlen, 1 };
    if (!j_eat(&c, '{')) goto bad;
    while (c.ok && !j_eat(&c, '}')) {
        if (j_eat(&c, ',')) continue;
        char name[256];
        if (j_string(&c, name, sizeof name) < 0 || !j_eat(&c, ':')) goto bad;
        if (!strcmp(name, "__metadata__")) { j_skip(&c); continue; }
        if (sf->num_tensors >= SAFETENSORS_MAX_TENSORS) { j_skip(&c); continue; }
        safetensor_t *t = &sf->tensors[sf->num_tensors];
        snprintf(t->name, sizeof t->name, "%s", name);
        if (!parse_entry(&c, t)) goto bad;
        if (8 + hlen + t->data_offset + t->data_size > sf->file_size) goto bad;
        sf->num_tensors++;
    }
    if (!c.ok) goto bad;
    return sf;
bad:
    safetensors_close(sf);
    return NULL;
}

void safetensors_close(safetensors_file_t *sf) {
    if (!sf) return;
    if (sf->data) munmap(sf->data, sf->file_size);
    free(sf->header_json);
    free(sf->path);
    free(sf);
}

const safetensor_t *safetensors_find(const safetensors_file_t *sf, const char *name) {
    for (int i = 0; i < sf->num_tensors; i++)
        if (!strcmp(sf->tensors[i].name, name)) return &sf->tensors[i];
    return NULL;
}

const void *safetensors_data(const safetensors_file_t *sf, const safetensor_t *t) {
    if (!sf || !t) return NULL;
    return (const uint8_t *)sf->data + 8 + sf->header_size + t->data_offset;
}

int64_t safetensor_numel(const safetensor_t *t) {
    int64_t n = 1;
    for (int i = 0; i < t->ndim; i++) n *= t->shape[i];
    return n;
}

int safetensor_is_bf16(const safetensor_t *t) { return t && t->dtype == DTYPE_BF16; }

static float half_to_float(uint16_t h) {
    uint32_t s = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 0x3FFu, bits;
    if (e == 31) bits = s | 0x7F800000u | (m << 13);
    else if (e) bits = s | ((e + 112u) << 23) | (m << 13);
    else if (!m) bits = s;
    else {
        int sh = 0;
        while (!(m & 0x400u)) { m <<= 1; sh++; }
        bits = s | ((113u - (uint32_t)sh) << 23) | ((m & 0x3FFu) << 13);
    }
    float f; memcpy(&f, &bits, 4); return f;
}

float *safetensors_get_f32(const safetensors_file_t *sf, const safetensor_t *t) {
    if (!sf || !t) return NULL;
    int64_t n = safetensor_numel(t);
    if (n <= 0) return NULL;
    float *out = malloc((size_t)n * sizeof(float));
    if (!out) return NULL;
    const void *src = safetensors_data(sf, t);
    if (t->dtype == DTYPE_F32) {
        memcpy(out, src, (size_t)n * 4);
    } else if (t->dtype == DTYPE_BF16) {
        const uint16_t *b = src;
        for (int64_t i = 0; i < n; i++) { uint32_t u = (uint32_t)b[i] << 16; memcpy(&out[i], &u, 4); }
    } else if (t->dtype == DTYPE_F16) {
        const uint16_t *h = src;
        for (int64_t i = 0; i < n; i++) out[i] = half_to_float(h[i]);
    } else {
        free(out);
        return NULL;
    }
    return out;
}

uint16_t *safetensors_get_bf16(const safetensors_file_t *sf, const safetensor_t *t) {
    if (!sf || !t || t->dtype != DTYPE_BF16) return NULL;
    size_t bytes = (size_t)safetensor_numel(t) * 2;
    uint16_t *out = malloc(bytes);
    if (out) memcpy(out, safetensors_data(sf, t), bytes);
    return out;
}

uint16_t *safetensors_get_bf16_direct(const safetensors_file_t *sf, const safetensor_t *t) {
    if (!sf || !t || t->dtype != DTYPE_BF16) return NULL;
    return (uint16_t *)safetensors_data(sf, t);
}

void safetensor_print(const safetensor_t *t) {
    static const char *names[] = { "F32", "F16", "BF16", "I32", "I64", "BOOL" };
    printf("%-90s %-5s [", t->name, t->dtype >= 0 && t->dtype <= 5 ? names[t->dtype] : "?");
    for (int i = 0; i < t->ndim; i++) printf("%s%lld", i ? ", " : "", (long long)t->shape[i]);
    printf("]  %zu bytes\n", t->data_size);
}

void safetensors_print_all(const safetensors_file_t *sf) {
    printf("%s: %d tensors\n", sf->path ? sf->path : "?", sf->num_tensors);
    for (int i = 0; i < sf->num_tensors; i++) safetensor_print(&sf->tensors[i]);
}
