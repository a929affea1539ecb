/*
 * vb_decode_mega.cu -- the decode loop as ONE persistent cooperative kernel.
 *
 * Why: a decode step is 6.86 GB of weights through ~130 dependent GEMV phases of 4-17 us each
 * (profiles/r01_launches_graph.md).  Launched as separate kernels, every phase pays a launch ramp, a
 * cold prologue and a drain: the GEMV core runs at 99% of HBM peak inside the 125 us logits kernel but
 * the step as a whole reaches only ~37%.  Here the whole multi-step greedy loop is one kernel:
 *
 *   grid   = one CTA per SM (cooperative launch => co-resident), 16 consumer warps + 1 producer warp
 *   weights: the producer warp walks the CTA's static slab schedule (layer, phase, row chunk) and streams
 *            it with cp.async.bulk (TMA 1-D bulk copies, mbarrier complete_tx) into a 6 x 36 KB shared-
 *            memory ring.  Weights do not depend on activations, so the producer NEVER waits for a phase
 *            boundary: while the consumers sit in a grid barrier or in the attention phase, the ring fills
 *            with the next phase's rows, and HBM keeps streaming.
 *   compute: consumers own fixed k-columns (8 consecutive k per 16-byte LDS.128, conflict free), keep the
 *            activation vector in registers, accumulate 16 rows, then one warp transpose-reduce + one
 *            named barrier per 16 rows.  Epilogues (RoPE + KV ring write, residual add, SiLU*up,
 *            logits + argmax) are the same as in vb_decode.cu.
 *   phases are separated by a grid-wide barrier (one atomic + one acquire spin per CTA); activations that
 *            cross CTAs live in L2 and are read with ld.global.cg.
 *   the autoregressive feedback (argmax -> next embedding) stays on the device; the kernel runs n_steps
 *            steps or until EOS and writes the token ids.
 *
 * Reference semantics: voxtral_decoder.c:586-706 per step, voxtral.c:1056-1093 for the loop.
 */
#include "vb_decode_common.cuh"
#include <cooperative_groups.h>
#include <string.h>

#define MK_CONS      512                       /* consumer threads (16 warps) */
#define MK_THREADS   (MK_CONS + 32)            /* + 1 producer warp */
#define MK_SLOTS     6
#define MK_SLOT_BYTES 36864                    /* 2 rows of K=9216 / 4 rows of K=4096 / 4 rows of K=3072 (24 KB) */
#define MK_GROUP     16                        /* rows per reduction group */
#define MK_SPIN_LIMIT (4000000000ll)           /* ~2 s of SM clocks: trap instead of hanging the GPU */

struct MegaArgs {
    DecParams p;
    int n_steps, pos0, token0, adapter_row0;
    unsigned int *bar;                          /* grid barrier counter (zeroed before launch) */
    int *err;
};

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
__device__ __forceinline__ void cons_bar() { asm volatile("bar.sync 1, %0;" :: "n"(MK_CONS) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

/* ------------------------------------------------------------------ shared state */
struct MegaSmem {
    uint64_t full[MK_SLOTS], empty[MK_SLOTS];
    float red[2][16][MK_GROUP];
    float sred[16];
    unsigned long long cand[16];
    volatile int abort_flag;                    /* consumers -> producer: stop issuing */
};

struct Ring {
    uint8_t *slots; MegaSmem *sm; uint32_t it;  /* it: chunks handled so far by this thread's role */
    __device__ __forceinline__ int slot() const { return (int)(it % MK_SLOTS); }
    __device__ __forceinline__ uint32_t parity() const { return (it / MK_SLOTS) & 1u; }
};

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
    case 1:  f.W = p.wo[layer];   f.row_bytes = VB_DEC_Q * 2;    rows_of(VOX_DEC_DIM, 1, f.row0, f.nrows);    f.rc = 4; break;
    case 2:  f.W = p.w13[layer];  f.row_bytes = VOX_DEC_DIM * 2; rows_of(VOX_DEC_HIDDEN, 2, f.row0, f.nrows); f.rc = 4; break;
    case 3:  f.W = p.w2[layer];   f.row_bytes = VOX_DEC_HIDDEN * 2; rows_of(VOX_DEC_DIM, 1, f.row0, f.nrows); f.rc = 2; break;
    default: f.W = p.tok_emb;     f.row_bytes = VOX_DEC_DIM * 2; rows_of(VOX_VOCAB_SIZE, 1, f.row0, f.nrows); f.rc = 4; break;
    }
    return f;
}

/* ------------------------------------------------------------------ producer */
__device__ void producer_issue_phase(Ring &r, const Phase &f, int *err) {
    /* chunk boundaries restart at every 16-row reduction group so that consumer groups own whole chunks */
    for (int g0 = 0; g0 < f.nrows; g0 += MK_GROUP) {
        const int gr = min(MK_GROUP, f.nrows - g0);
        for (int c0 = 0; c0 < gr; c0 += f.rc) {
            const int rows = min(f.rc, gr - c0);
            const int s = r.slot();
            long long t0 = 0;
            while (!mbar_try_wait(&r.sm->empty[s], r.parity() ^ 1u)) {
                if (r.sm->abort_flag) return;
                spin_guard(t0, err, 2);
            }
            if (r.sm->abort_flag) return;
            const uint32_t bytes = (uint32_t)rows * (uint32_t)f.row_bytes;
            mbar_expect_tx(&r.sm->full[s], bytes);
            bulk_g2s(r.slots + (size_t)s * MK_SLOT_BYTES,
                     reinterpret_cast<const uint8_t *>(f.W) + (size_t)(f.row0 + g0 + c0) * f.row_bytes, bytes, &r.sm->full[s]);
            r.it++;
        }
    }
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

/* Consume one GEMV phase from the ring.  RC = rows per chunk (must match phase_of().rc). */
template <int CPT, int RC, typename Epi>
__device__ __forceinline__ void consume_phase(Ring &r, const Phase &f, int NT, const float (&xr)[CPT * 8],
                                              int &redbuf, int *err, Epi epi) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool active = t < NT;
    for (int g0 = 0; g0 < f.nrows; g0 += MK_GROUP) {
        const int gr = min(MK_GROUP, f.nrows - g0);
        float acc[MK_GROUP];
#pragma unroll
        for (int i = 0; i < MK_GROUP; i++) acc[i] = 0.f;
#pragma unroll
        for (int cb = 0; cb < MK_GROUP / RC; cb++) {
            if (cb * RC < gr) {
                const int s = r.slot();
                long long t0 = 0;
                while (!mbar_try_wait(&r.sm->full[s], r.parity())) spin_guard(t0, err, 3);
                const uint8_t *base = r.slots + (size_t)s * MK_SLOT_BYTES + (size_t)t * 16;
#pragma unroll
                for (int rr = 0; rr < RC; rr++) {
                    if (cb * RC + rr < gr && active) {
#pragma unroll
                        for (int c = 0; c < CPT; c++) {
                            uint4 w = *reinterpret_cast<const uint4 *>(base + (size_t)rr * f.row_bytes + (size_t)c * NT * 16);
                            acc[cb * RC + rr] = dot8(w, &xr[c * 8], acc[cb * RC + rr]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&r.sm->empty[s]);
                r.it++;
            }
        }
        float tot = warp_transpose_reduce<MK_GROUP>(acc, lane);
        if ((lane & 1) == 0) r.sm->red[redbuf][warp][lane >> 1] = tot;
        cons_bar();
        if (warp == 0) {
            float sum = 0.f;
            if (lane < MK_GROUP) {
#pragma unroll
                for (int wv = 0; wv < 16; wv++) sum += r.sm->red[redbuf][wv][lane];
            }
            epi(f.row0 + g0 + lane, sum, lane, lane < gr);
        }
        redbuf ^= 1;          /* double-buffered: the next group's barrier orders reuse two groups later */
    }
}

__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int &gen, int *err) {
    gen++;
    cons_bar();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        const unsigned int target = gen * gridDim.x;
        long long t0 = 0;
        while (ld_acquire_u32(bar) < target) spin_guard(t0, err, 1);
        __threadfence();
    }
    cons_bar();
}

/* ------------------------------------------------------------------ attention inside the megakernel */
__device__ __forceinline__ void mega_attn_partial(const DecParams &p, int layer, int pos) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kvh = warp & 7, sub = warp >> 3;
    const int NS = gridDim.x * NS_PER_CTA, sg = blockIdx.x * NS_PER_CTA + sub;
    const int n_valid = min(pos + 1, VB_KV_SLOTS);
    const int s0 = (int)((long long)n_valid * sg / NS), s1 = (int)((long long)n_valid * (sg + 1) / NS);
    const float scale = 1.0f / sqrtf((float)HD);
    float4 qv[4];
#pragma unroll
    for (int hq = 0; hq < 4; hq++) qv[hq] = __ldcg(reinterpret_cast<const float4 *>(p.q + (kvh * 4 + hq) * HD + lane * 4));
    float m[4], l[4]; float4 o[4];
#pragma unroll
    for (int hq = 0; hq < 4; hq++) { m[hq] = -1e30f; l[hq] = 0.f; o[hq] = make_float4(0.f, 0.f, 0.f, 0.f); }
    const float *kb = p.kv_k + (size_t)layer * VB_KV_SLOTS * VB_DEC_KV + kvh * HD + lane * 4;
    const float *vb = p.kv_v + (size_t)layer * VB_KV_SLOTS * VB_DEC_KV + kvh * HD + lane * 4;
    for (int sb = s0; sb < s1; sb += 4) {
        float4 k4[4], v4[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int s = min(sb + u, s1 - 1);
            k4[u] = __ldcg(reinterpret_cast<const float4 *>(kb + (size_t)s * VB_DEC_KV));
            v4[u] = __ldcg(reinterpret_cast<const float4 *>(vb + (size_t)s * VB_DEC_KV));
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (sb + u < s1) {
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
#pragma unroll
    for (int hq = 0; hq < 4; hq++) {
        int h = kvh * 4 + hq;
        size_t pi = (size_t)sg * VOX_DEC_HEADS + h;
        if (lane == 0) { p.part_m[pi] = m[hq]; p.part_l[pi] = l[hq]; }
        *reinterpret_cast<float4 *>(p.part_o + pi * HD + lane * 4) = o[hq];
    }
}

/* CTA c < 128 combines head (c & 31), dims [(c>>5)*32, +32); 16 warps stride over the partials. */
__device__ __forceinline__ void mega_attn_combine(const DecParams &p, MegaSmem *sm) {
    if (blockIdx.x >= 128) return;
    const int h = blockIdx.x & 31, dq = blockIdx.x >> 5;
    const int NS = gridDim.x * NS_PER_CTA;
    const int dl = threadIdx.x & 31, pl = threadIdx.x >> 5, d = dq * 32 + dl;
    float M = -1e30f;
    for (int i = threadIdx.x; i < NS; i += MK_CONS) M = fmaxf(M, __ldcg(p.part_m + (size_t)i * VOX_DEC_HEADS + h));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
    cons_bar();
    if (dl == 0) sm->sred[pl] = M;
    cons_bar();
#pragma unroll
    for (int i = 0; i < 16; i++) M = fmaxf(M, sm->sred[i]);
    float num = 0.f, den = 0.f;
    for (int i = pl; i < NS; i += 16) {
        size_t pi = (size_t)i * VOX_DEC_HEADS + h;
        float li = __ldcg(p.part_l + pi);
        if (li > 0.f) {
            float w = expf(__ldcg(p.part_m + pi) - M);
            den = fmaf(w, li, den);
            num = fmaf(w, __ldcg(p.part_o + pi * HD + d), num);
        }
    }
    float *nb = &sm->red[0][0][0], *db = &sm->red[1][0][0];          /* reuse as [16][16] x2: index pl*32+dl needs 512 */
    /* red is 2*16*16 = 512 floats per half: exactly [16 warps][32 dims] when viewed flat over both halves */
    (void)db;
    float *flat = nb;                                                  /* 1024 floats total: num then den */
    flat[pl * 32 + dl] = num;
    cons_bar();
    float n2 = 0.f;
    if (pl == 0) {
#pragma unroll
        for (int i = 0; i < 16; i++) n2 += flat[i * 32 + dl];
    }
    cons_bar();
    flat[pl * 32 + dl] = den;
    cons_bar();
    if (pl == 0) {
        float d2 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i++) d2 += flat[i * 32 + dl];
        p.attn_out[h * HD + d] = d2 > 0.f ? n2 / d2 : 0.f;
    }
    cons_bar();
}

/* ------------------------------------------------------------------ the kernel */
extern __shared__ __align__(1024) uint8_t mk_smem_raw[];

__global__ void __launch_bounds__(MK_THREADS, 1) k_dec_mega(MegaArgs a) {
    uint8_t *slots = mk_smem_raw;
    MegaSmem *sm = reinterpret_cast<MegaSmem *>(mk_smem_raw + (size_t)MK_SLOTS * MK_SLOT_BYTES);
    const DecParams &p = a.p;
    const int tid = threadIdx.x;

    if (tid == 0) {
        for (int i = 0; i < MK_SLOTS; i++) { mbar_init(&sm->full[i], 1); mbar_init(&sm->empty[i], 16); }
        sm->abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= MK_CONS) {
        /* ===================== producer warp: one elected lane streams the slab schedule ===================== */
        if (tid == MK_CONS) {
            Ring r{ slots, sm, 0u };
            for (int step = 0; step < a.n_steps && !sm->abort_flag; step++) {
                for (int layer = 0; layer < VOX_DEC_LAYERS && !sm->abort_flag; layer++)
                    for (int ph = 0; ph < 4 && !sm->abort_flag; ph++) {
                        Phase f = phase_of(p, layer, ph);
                        producer_issue_phase(r, f, a.err);
                    }
                if (!sm->abort_flag) { Phase f = phase_of(p, 0, 4); producer_issue_phase(r, f, a.err); }
            }
            /* every bulk copy that was issued must land before the CTA may exit (smem is its target):
             * chunks it-1 .. it-MK_SLOTS are the only ones that can still be in flight */
            for (int back = 1; back <= MK_SLOTS; back++) {
                if (r.it < (uint32_t)back) break;
                uint32_t j = r.it - back;
                long long t0 = 0;
                while (!mbar_try_wait(&sm->full[j % MK_SLOTS], (j / MK_SLOTS) & 1u)) spin_guard(t0, a.err, 4);
            }
        }
        __syncthreads();          /* matches the consumers' final barrier */
        return;
    }

    /* ===================== consumers ===================== */
    Ring r{ slots, sm, 0u };
    unsigned int gen = 0;
    int redbuf = 0;
    int pos = a.pos0, token = a.token0, arow = a.adapter_row0;
    const float *adapter = *p.adapter_pp;
    const int lane = tid & 31;
    int n_done = 0, eos = 0;

    for (int step = 0; step < a.n_steps; step++) {
        /* x = adapter[arow] + tok_emb[token] (voxtral.c:1057-1061).  Every CTA needs all of x as GEMV input;
         * the residual stream itself lives in global memory and each row is owned by the CTA that owns that
         * output row of wo/w2, so that CTA (re)initialises its rows here. */
        const float *arow_p = adapter + (size_t)arow * VOX_DEC_DIM;
        const uint16_t *erow_p = p.tok_emb + (size_t)token * VOX_DEC_DIM;
        {
            int r0, n; rows_of(VOX_DEC_DIM, 1, r0, n);
            if (tid < n) p.x[r0 + tid] = arow_p[r0 + tid] + __uint_as_float((uint32_t)erow_p[r0 + tid] << 16);
        }
        const int slot = pos & (VB_KV_SLOTS - 1);

        for (int layer = 0; layer < VOX_DEC_LAYERS; layer++) {
            /* ---- QKV ---- */
            {
                const int NT = VOX_DEC_DIM / 8;
                float xr[8];
                if (layer == 0) {
                    if (tid < NT) {
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            int k = tid * 8 + j;
                            xr[j] = arow_p[k] + __uint_as_float((uint32_t)erow_p[k] << 16);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; j++) xr[j] = 0.f;
                    }
                } else {
                    load_x_cols_cg<1>(xr, p.x, NT);
                }
                rmsnorm_cols_cons<1>(xr, p.attn_norm[layer], nullptr, NT, VOX_DEC_DIM, sm->sred);
                Phase f = phase_of(p, layer, 0);
                float *kdst = p.kv_k + ((size_t)layer * VB_KV_SLOTS + slot) * VB_DEC_KV;
                float *vdst = p.kv_v + ((size_t)layer * VB_KV_SLOTS + slot) * VB_DEC_KV;
                const float *inv_freq = p.inv_freq;
                float *q = p.q;
                consume_phase<1, 4>(r, f, NT, xr, redbuf, a.err, [&](int row, float v, int, bool valid) {
                    float other = __shfl_xor_sync(0xffffffffu, v, 1);
                    if (!valid) return;
                    if (row < VB_DEC_Q + VB_DEC_KV) {
                        int d = (row & (HD - 1)) >> 1;
                        float sn, cs;
                        sincosf((float)pos * inv_freq[d], &sn, &cs);
                        float y = (row & 1) ? (other * sn + v * cs) : (v * cs - other * sn);
                        if (row < VB_DEC_Q) q[row] = y; else kdst[row - VB_DEC_Q] = y;
                    } else {
                        vdst[row - VB_DEC_Q - VB_DEC_KV] = v;
                    }
                });
            }
            grid_barrier(a.bar, gen, a.err);
            /* ---- attention ---- */
            mega_attn_partial(p, layer, pos);
            grid_barrier(a.bar, gen, a.err);
            mega_attn_combine(p, sm);
            grid_barrier(a.bar, gen, a.err);
            /* ---- wo + residual ---- */
            {
                const int NT = VB_DEC_Q / 8;
                float xr[8];
                load_x_cols_cg<1>(xr, p.attn_out, NT);
                Phase f = phase_of(p, layer, 1);
                float *x = p.x;
                consume_phase<1, 4>(r, f, NT, xr, redbuf, a.err, [&](int row, float v, int, bool valid) {
                    if (valid) x[row] = __ldcg(x + row) + v;
                });
            }
            grid_barrier(a.bar, gen, a.err);
            /* ---- w1|w3 + SiLU*up ---- */
            {
                const int NT = VOX_DEC_DIM / 8;
                float xr[8];
                load_x_cols_cg<1>(xr, p.x, NT);
                rmsnorm_cols_cons<1>(xr, p.ffn_norm[layer], p.ada + (size_t)layer * VOX_DEC_DIM, NT, VOX_DEC_DIM, sm->sred);
                Phase f = phase_of(p, layer, 2);
                float *gate = p.gate;
                consume_phase<1, 4>(r, f, NT, xr, redbuf, a.err, [&](int row, float v, int, bool valid) {
                    float other = __shfl_xor_sync(0xffffffffu, v, 1);
                    if (valid && !(row & 1)) gate[row >> 1] = vb_silu(v) * other;
                });
            }
            grid_barrier(a.bar, gen, a.err);
            /* ---- w2 + residual ---- */
            {
                const int NT = VOX_DEC_HIDDEN / 8 / 3;
                float xr[24];
                load_x_cols_cg<3>(xr, p.gate, NT);
                Phase f = phase_of(p, layer, 3);
                float *x = p.x;
                consume_phase<3, 2>(r, f, NT, xr, redbuf, a.err, [&](int row, float v, int, bool valid) {
                    if (valid) x[row] = __ldcg(x + row) + v;
                });
            }
            grid_barrier(a.bar, gen, a.err);
        }
        /* ---- final norm, logits, argmax ---- */
        {
            const int NT = VOX_DEC_DIM / 8;
            float xr[8];
            load_x_cols_cg<1>(xr, p.x, NT);
            rmsnorm_cols_cons<1>(xr, p.final_norm, nullptr, NT, VOX_DEC_DIM, sm->sred);
            Phase f = phase_of(p, 0, 4);
            float *logits = p.logits;
            unsigned long long best = 0ull;
            consume_phase<1, 4>(r, f, NT, xr, redbuf, a.err, [&](int row, float v, int, bool valid) {
                if (!valid) return;
                logits[row] = v;
                unsigned long long c = pack_cand(v, row);
                if (c > best) best = c;
            });
            if (tid < 32) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                    if (other > best) best = other;
                }
                if (tid == 0) p.argmax[blockIdx.x] = best;
            }
        }
        grid_barrier(a.bar, gen, a.err);
        {
            unsigned long long best = 0ull;
            for (int i = tid; i < (int)gridDim.x; i += MK_CONS) {
                unsigned long long c = __ldcg(p.argmax + i);
                if (c > best) best = c;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                if (other > best) best = other;
            }
            cons_bar();
            if (lane == 0) sm->cand[tid >> 5] = best;
            cons_bar();
#pragma unroll
            for (int i = 0; i < 16; i++) if (sm->cand[i] > best) best = sm->cand[i];
            token = cand_index(best);
        }
        if (blockIdx.x == 0 && tid == 0) p.tokens[n_done] = token;
        n_done++; pos++; arow++;
        if (token == VB_TOKEN_EOS) { eos = 1; break; }
    }

    /* stop the producer (it may be several chunks into a step that will never be consumed) */
    if (tid == 0) sm->abort_flag = 1;
    if (blockIdx.x == 0 && tid == 0) {
        VbDecState st;
        st.pos = pos; st.token = token; st.eos = eos; st.n_out = n_done; st.adapter_row = arow;
        st.pad[0] = st.pad[1] = st.pad[2] = 0;
        *p.st = st;
    }
    __syncthreads();              /* with the producer warp: all bulk copies have landed */
}

/* ------------------------------------------------------------------ host */
static size_t mega_smem_bytes() { return (size_t)MK_SLOTS * MK_SLOT_BYTES + sizeof(MegaSmem) + 64; }

extern "C" int vb_decoder_mega_supported(VbEngine *e) {
    static int cached = -1;
    if (cached >= 0) return cached;
    int coop = 0, max_smem = 0, blocks = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device);
    if (!coop || (size_t)max_smem < mega_smem_bytes()) { cached = 0; return 0; }
    if (cudaFuncSetAttribute(k_dec_mega, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mega_smem_bytes()) != cudaSuccess) { cached = 0; return 0; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_dec_mega, MK_THREADS, mega_smem_bytes()) != cudaSuccess || blocks < 1) { cached = 0; return 0; }
    cached = 1;
    return 1;
}

/* Runs up to n_steps greedy steps; state (pos/token/adapter_row) comes from the arguments; returns the number
 * of tokens produced (stops after EOS).  Tokens are left in e->d_tokens / the final state in e->d_state. */
extern "C" int vb_decoder_mega_launch(VbEngine *e, const float *d_adapter, int adapter_row, int n_steps,
                                      int prev_token, int pos) {
    if (!e->d_mega_bar) {
        const size_t wb = e->weight_bytes;
        e->d_mega_bar = (unsigned int *)vb_dev_alloc_owned(e, 256);
        e->weight_bytes = wb;
    }
    struct { VbDecState st; const float *adapter; } h;
    memset(&h, 0, sizeof h);
    h.st.pos = pos; h.st.token = prev_token; h.st.adapter_row = adapter_row; h.adapter = d_adapter;
    VB_CUDA_OK(cudaMemcpyAsync(e->d_state, &h, sizeof h, cudaMemcpyHostToDevice, e->stream));
    VB_CUDA_OK(cudaMemsetAsync(e->d_mega_bar, 0, 256, e->stream));
    MegaArgs a;
    a.p = vb_make_dec_params(e, 1);
    a.n_steps = n_steps; a.pos0 = pos; a.token0 = prev_token; a.adapter_row0 = adapter_row;
    a.bar = e->d_mega_bar; a.err = (int *)(e->d_mega_bar + 32);
    void *args[] = { &a };
    VB_CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_dec_mega, dim3(e->sm_count), dim3(MK_THREADS), args,
                                           mega_smem_bytes(), e->stream));
    e->launches += 1;
    return 0;
}
