/*
 * vb_decode_persist_common.cuh -- building blocks shared by the persistent decode kernels
 * (vb_decode_v2.cu: TMA weight ring, dynamic row chunks, up to 8 activation columns -- the default;
 * vb_decode_persist.cu: direct streaming loads, 512 compute threads synchronising on named barrier 1):
 * PTX wrappers (mbarrier, cp.async.bulk, release/acquire), the grid barrier, the static slab schedule.
 */
#ifndef VB_DECODE_PERSIST_COMMON_CUH
#define VB_DECODE_PERSIST_COMMON_CUH
#include "vb_decode_common.cuh"

#define MK_CONS      512                       /* compute threads per CTA (16 warps) */
#define MK_GROUP     16                        /* rows per reduction group */
#define MK_SPIN_LIMIT (4000000000ll)           /* ~2 s of SM clocks: trap instead of hanging the GPU */
#define MK_ATT_FLOATS (16 * 4 * 132)           /* intra-CTA attention merge scratch: 16 warps x 4 heads x (m,l,pad,pad,o[128]) */
#define MK_PROF_SLOTS 1024                     /* timestamps per CTA */

struct MegaArgs {
    DecParams p;
    int n_steps, pos0, token0, adapter_row0;
    unsigned int *bar;                          /* [0] grid barrier counter, [16..23] attention tickets, [32] error word */
    int *err;
    long long *prof;                            /* optional: per-CTA phase timestamps of step `prof_step` (else NULL) */
    int prof_step;
    int l2_ahead;                               /* bytes per CTA kept prefetched into L2 ahead of consumption */
    const uint16_t *emb_img;                    /* tc kernel only: decode-tiled image of the tied embedding (p.tok_emb stays row-major for lookups) */
};

#define PROF(tag) do { if (a.prof && step == a.prof_step && tid == 0 && prof_n < MK_PROF_SLOTS) \
    a.prof[(size_t)blockIdx.x * MK_PROF_SLOTS + prof_n++] = clock64(); } while (0)

/* ------------------------------------------------------------------ PTX wrappers */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void *src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void red_release_add(unsigned int *p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cons_bar() { asm volatile("bar.sync 1, %0;" :: "n"(MK_CONS) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void spin_guard(long long &t0, int *err, int code) {
    long long now = clock64();
    if (t0 == 0) t0 = now;
    else if (now - t0 > MK_SPIN_LIMIT) { if (err) atomicExch(err, code); __trap(); }
}

/* ------------------------------------------------------------------ the static slab schedule */
struct Phase { const uint16_t *W; int row_bytes; int row0; int nrows; int rc; };

__device__ __forceinline__ void rows_of(int total_units, int unit_rows, int &row0, int &nrows) {
    long long a = (long long)total_units * blockIdx.x / gridDim.x;
    long long b = (long long)total_units * (blockIdx.x + 1) / gridDim.x;
    row0 = (int)a * unit_rows; nrows = (int)(b - a) * unit_rows;
}
/* ph: 0 QKV, 1 WO, 2 W13, 3 W2 (per layer), 4 LOGITS */
__device__ __forceinline__ Phase phase_of(const DecParams &p, int layer, int ph) {
    Phase f;
    switch (ph) {
    case 0:  f.W = p.wqkv[layer]; f.row_bytes = VOX_DEC_DIM * 2; rows_of(VB_DEC_QKV / 2, 2, f.row0, f.nrows); f.rc = 4; break;
    case 1:  f.W = p.wo[layer];   f.row_bytes = VB_DEC_Q * 2;    rows_of(VOX_DEC_DIM, 1, f.row0, f.nrows);    f.rc = 3; break;
    case 2:  f.W = p.w13[layer];  f.row_bytes = VOX_DEC_DIM * 2; rows_of(VOX_DEC_HIDDEN, 2, f.row0, f.nrows); f.rc = 4; break;
    case 3:  f.W = p.w2[layer];   f.row_bytes = VOX_DEC_HIDDEN * 2; rows_of(VOX_DEC_DIM, 1, f.row0, f.nrows); f.rc = 1; break;
    default: f.W = p.tok_emb;     f.row_bytes = VOX_DEC_DIM * 2; rows_of(VOX_VOCAB_SIZE, 1, f.row0, f.nrows); f.rc = 4; break;
    }
    return f;
}

/* ------------------------------------------------------------------ consumer building blocks */
__device__ __forceinline__ float cons_block_sum(float v, float *sred) {
    v = vb_warp_sum(v);
    cons_bar();
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
    cons_bar();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) t += sred[i];
    return t;
}

template <int CPT>
__device__ __forceinline__ void load_x_cols_cg(float (&xr)[CPT * 8], const float *x, int NT) {
    const int t = threadIdx.x;
#pragma unroll
    for (int c = 0; c < CPT; c++) {
        if (t < NT) {
            const float4 *p = reinterpret_cast<const float4 *>(x + (size_t)(c * NT + t) * 8);
            float4 a = __ldcg(p), b = __ldcg(p + 1);
            xr[c * 8 + 0] = a.x; xr[c * 8 + 1] = a.y; xr[c * 8 + 2] = a.z; xr[c * 8 + 3] = a.w;
            xr[c * 8 + 4] = b.x; xr[c * 8 + 5] = b.y; xr[c * 8 + 6] = b.z; xr[c * 8 + 7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) xr[c * 8 + j] = 0.f;
        }
    }
}

template <int CPT>
__device__ __forceinline__ void rmsnorm_cols_cons(float (&xr)[CPT * 8], const float *__restrict__ w,
                                                  const float *__restrict__ ada, int NT, int hidden, float *sred) {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < CPT * 8; j++) ss = fmaf(xr[j], xr[j], ss);
    float tot = cons_block_sum(ss, sred);
    float rinv = 1.0f / sqrtf(tot / (float)hidden + VOX_DEC_NORM_EPS);
    const int t = threadIdx.x;
    if (t < NT) {
#pragma unroll
        for (int c = 0; c < CPT; c++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                int k = (c * NT + t) * 8 + j;
                float v = xr[c * 8 + j] * rinv * w[k];
                if (ada) v *= (1.0f + ada[k]);
                xr[c * 8 + j] = v;
            }
    }
}

/* Grid-wide barrier among the consumer threads of all CTAs.  bar.sync makes the CTA's prior writes
 * happen-before thread 0's release-RED; the acquire poll + bar.sync orders everyone's later (L2, ld.cg) reads. */
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int &gen, int *err) {
    gen++;
    cons_bar();
    if (threadIdx.x == 0) {
        red_release_add(bar, 1u);
        const unsigned int target = gen * gridDim.x;
        long long t0 = 0;
        while (ld_acquire_u32(bar) < target) spin_guard(t0, err, 1);
    }
    cons_bar();
}

/* ------------------------------------------------------------------ attention inside the megakernel */
/* kv head h is served by the CTAs with (cta & 7) == h; they split the valid ring slots between them.  Inside a
 * CTA the 16 warps take interleaved slots (all 4 query heads of the kv head share each K/V row read), merge
 * through shared memory to ONE partial per query head, publish it, and the last CTA to arrive for a kv head
 * (atomic ticket) combines that head group's partials into attn_out -- so the whole attention is one phase. */
__device__ __forceinline__ void mega_attention(const DecParams &p, int layer, int pos, volatile int *is_last_flag, float *att_scr,
                                               unsigned int *tickets) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kvh = blockIdx.x & 7, si = blockIdx.x >> 3;
    const int nsplit = (gridDim.x >> 3) + ((int)(gridDim.x & 7) > kvh ? 1 : 0);
    const int n_valid = min(pos + 1, VB_KV_SLOTS);
    const int s0 = (int)((long long)n_valid * si / nsplit), s1 = (int)((long long)n_valid * (si + 1) / nsplit);
    const float scale = 1.0f / sqrtf((float)HD);
    float4 qv[4];
#pragma unroll
    for (int hq = 0; hq < 4; hq++) qv[hq] = __ldcg(reinterpret_cast<const float4 *>(p.q + (kvh * 4 + hq) * HD + lane * 4));
    float m[4], l[4]; float4 o[4];
#pragma unroll
    for (int hq = 0; hq < 4; hq++) { m[hq] = -1e30f; l[hq] = 0.f; o[hq] = make_float4(0.f, 0.f, 0.f, 0.f); }
    const float *kb = p.kv_k + (size_t)layer * VB_KV_SLOTS * VB_DEC_KV + kvh * HD + lane * 4;
    const float *vb = p.kv_v + (size_t)layer * VB_KV_SLOTS * VB_DEC_KV + kvh * HD + lane * 4;
    for (int sb = s0 + warp; sb < s1; sb += 64) {          /* this warp's slots: sb, sb+16, sb+32, sb+48 */
        float4 k4[4], v4[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int s = min(sb + 16 * u, s1 - 1);
            k4[u] = __ldcg(reinterpret_cast<const float4 *>(kb + (size_t)s * VB_DEC_KV));
            v4[u] = __ldcg(reinterpret_cast<const float4 *>(vb + (size_t)s * VB_DEC_KV));
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (sb + 16 * u < s1) {
                float sc[4];
#pragma unroll
                for (int hq = 0; hq < 4; hq++)
                    sc[hq] = qv[hq].x * k4[u].x + qv[hq].y * k4[u].y + qv[hq].z * k4[u].z + qv[hq].w * k4[u].w;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                    for (int hq = 0; hq < 4; hq++) sc[hq] += __shfl_xor_sync(0xffffffffu, sc[hq], off);
#pragma unroll
                for (int hq = 0; hq < 4; hq++) {
                    float sv = sc[hq] * scale;
                    float mn = fmaxf(m[hq], sv);
                    float c = expf(m[hq] - mn), pw = expf(sv - mn);
                    l[hq] = l[hq] * c + pw;
                    o[hq].x = o[hq].x * c + pw * v4[u].x; o[hq].y = o[hq].y * c + pw * v4[u].y;
                    o[hq].z = o[hq].z * c + pw * v4[u].z; o[hq].w = o[hq].w * c + pw * v4[u].w;
                    m[hq] = mn;
                }
            }
        }
    }
    /* ---- merge the 16 warps: scratch[warp][hq] = {m, l, -, -, o[128]} ---- */
#pragma unroll
    for (int hq = 0; hq < 4; hq++) {
        float *dst = att_scr + (size_t)(warp * 4 + hq) * 132;
        if (lane == 0) { dst[0] = m[hq]; dst[1] = l[hq]; }
        *reinterpret_cast<float4 *>(dst + 4 + lane * 4) = o[hq];
    }
    cons_bar();
    if (warp < 4) {                                         /* warp hq merges query head kvh*4+hq */
        const int hq = warp;
        float M = -1e30f;
#pragma unroll
        for (int w = 0; w < 16; w++) M = fmaxf(M, att_scr[(size_t)(w * 4 + hq) * 132]);
        float L = 0.f; float4 O = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < 16; w++) {
            const float *src = att_scr + (size_t)(w * 4 + hq) * 132;
            float lw = src[1];
            if (lw > 0.f) {
                float c = expf(src[0] - M);
                float4 ow = *reinterpret_cast<const float4 *>(src + 4 + lane * 4);
                L = fmaf(c, lw, L);
                O.x = fmaf(c, ow.x, O.x); O.y = fmaf(c, ow.y, O.y); O.z = fmaf(c, ow.z, O.z); O.w = fmaf(c, ow.w, O.w);
            }
        }
        const size_t pi = (size_t)si * VOX_DEC_HEADS + (kvh * 4 + hq);
        if (lane == 0) { p.part_m[pi] = M; p.part_l[pi] = L; }
        *reinterpret_cast<float4 *>(p.part_o + pi * HD + lane * 4) = O;
    }
    cons_bar();
    /* ---- ticket: the last CTA of this kv head combines ---- */
    if (tid == 0) {
        __threadfence();
        unsigned int old = atomicAdd(&tickets[kvh], 1u);
        int last = (old == (unsigned int)(nsplit - 1));
        if (last) { tickets[kvh] = 0u; __threadfence(); }
        *is_last_flag = last;
    }
    cons_bar();
    if (*is_last_flag && warp < 4) {
        /* all loads are issued before any is consumed (no data-dependent control flow): one L2 round trip */
        const int h = kvh * 4 + warp;
        float mi = -1e30f, li = 0.f;
        if (lane < nsplit) {
            mi = __ldcg(p.part_m + (size_t)lane * VOX_DEC_HEADS + h);
            li = __ldcg(p.part_l + (size_t)lane * VOX_DEC_HEADS + h);
        }
        float M = mi;
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o2));
        float wi = li > 0.f ? expf(mi - M) : 0.f;
        float L = vb_warp_sum(wi * li);
        float4 O = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 ov[20];
#pragma unroll
        for (int i = 0; i < 20; i++)
            ov[i] = i < nsplit ? __ldcg(reinterpret_cast<const float4 *>(p.part_o + ((size_t)i * VOX_DEC_HEADS + h) * HD + lane * 4))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 20; i++) {
            float c = __shfl_sync(0xffffffffu, wi, i);
            O.x = fmaf(c, ov[i].x, O.x); O.y = fmaf(c, ov[i].y, O.y); O.z = fmaf(c, ov[i].z, O.z); O.w = fmaf(c, ov[i].w, O.w);
        }
        float inv = L > 0.f ? 1.0f / L : 0.f;
        *reinterpret_cast<float4 *>(p.attn_out + h * HD + lane * 4) = make_float4(O.x * inv, O.y * inv, O.z * inv, O.w * inv);
    }
}


void vb_mega_prof_begin(VbEngine *e, MegaArgs &a, int n_steps);
void vb_mega_prof_report(VbEngine *e, const MegaArgs &a, const char *label);
#endif
