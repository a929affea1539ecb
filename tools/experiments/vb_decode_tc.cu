/*
 * vb_decode_tc.cu -- persistent cooperative decode kernel: TMA weight ring + tensor-core consumer.
 *
 * Why a third persistent variant (profiles/r01_decode.md): a decode step is HBM bound, but ~130 phase boundaries per step
 * each idle HBM for ~5 us, and nothing downstream of HBM can make the time up: L2->SM delivers only ~1.1x the HBM rate
 * chip-wide, and the CUDA-core GEMV (8 shifts + 8 FMAs per 16 bytes) is issue-bound at <2x the HBM rate even out of shared
 * memory.  The only store big and fast enough to absorb a boundary is shared memory (148 x ~165 KB = 24 MB, one boundary's
 * worth of HBM time) -- provided the consumer drains it several times faster than HBM fills it.  So:
 *
 *   weight image  : the decoder matrices are re-tiled once (k_build_tc_image) into the order this kernel consumes them: per
 *                   CTA slab, per 16-row group, per 1024-k chunk one contiguous <=32 KB block [rows][1024 k], 16-byte pieces
 *                   XOR-swizzled by row parity so the fragment loads below are bank-conflict free without padding.  (Issuing a
 *                   chunk as 16 separate 2 KB row copies from the row-major matrix costs ~70 cycles per cp.async.bulk --
 *                   1100 of the 1424 cycles a chunk may take at the HBM rate -- and the producer could not stay ahead.)
 *   producer warp : walks the CTA's static slab schedule and streams one chunk per cp.async.bulk (mbarrier complete_tx)
 *                   into a 5 x 32 KB ring, never waiting for a phase boundary -- weights do not depend on activations;
 *   consumers     : 16 warps; the activation vector is split once per phase into three bf16 planes (hi + mid + lo = all
 *                   24 mantissa bits of the f32 value) kept in shared memory; each warp owns 64 of a chunk's 1024 k and
 *                   issues mma.sync.m16n8k16 (A = 16 weight rows straight from the ring with LDS.128 -- the k order inside
 *                   a fragment is permuted consistently for A and B, so no ldmatrix/transposes --, B = the three planes in
 *                   columns 0..2, f32 accumulate): 6 LDS + 4 MMA per warp per 32 KB chunk instead of ~1000 instructions;
 *   everything else (grid barriers, ticketed split-S attention, epilogues, on-device token feedback) is shared with the
 *   other persistent kernels.
 *
 * Products bf16 x bf16 are exact in f32; accumulation is f32 in the tensor core, in a different order than the CUDA-core
 * kernels -- logits agree to ~1e-5, token ids are checked identical (tests/test_gpu_stream_parity.py).
 *
 * Reference semantics: voxtral_decoder.c:586-706 per step, voxtral.c:1056-1093 for the loop.
 */
#include "vb_decode_persist_common.cuh"
#include <string.h>

/* with VOX_CUDA_MEGA_PROF: also record how many chunks the producer is ahead at every stamp (upper half of the buffer) */
#undef PROF
#define PROF(tag) do { if (a.prof && step == a.prof_step && tid == 0 && prof_n < MK_PROF_SLOTS / 2) { \
    a.prof[(size_t)blockIdx.x * MK_PROF_SLOTS + MK_PROF_SLOTS / 2 + prof_n] = (long long)(sm->prod_it - r.it); \
    a.prof[(size_t)blockIdx.x * MK_PROF_SLOTS + prof_n++] = clock64(); } } while (0)

#define TK_THREADS    (MK_CONS + 32)           /* + 1 producer warp */
#define TK_SLOTS      5
#define TK_KC         1024                     /* k columns per chunk */
#define TK_ROW_BYTES  (TK_KC * 2)
#define TK_SLOT_BYTES (16 * TK_ROW_BYTES)
/* 16-byte piece c of row r of a chunk lives at piece c ^ ((r & 1) << 2): the 8 lanes of a quarter warp (two rows x four
 * consecutive pieces) then cover all eight 16-byte bank groups of a 128-byte line */
#define TK_SWZ(r, c)  ((c) ^ (((r) & 1) << 2))
#define TK_PLANE_STRIDE(K) ((K) * 2 + 64)      /* bytes per activation plane, same bank argument */
#define TK_PLANES_BYTES ((3 * TK_PLANE_STRIDE(VOX_DEC_HIDDEN) + 127) / 128 * 128)

struct TcSmem {
    uint64_t full[TK_SLOTS], empty[TK_SLOTS];
    float red[2][16][MK_GROUP];
    float sred[16];
    unsigned long long cand[16];
    volatile int abort_flag;
    volatile int is_last;
    volatile unsigned int prod_it;               /* chunks issued so far (diagnostics only) */
};

/* phase_of() with the logits phase reading the tiled embedding image (p.tok_emb stays row-major for the embedding lookup) */
__device__ __forceinline__ Phase tc_phase_of(const DecParams &p, const uint16_t *emb_img, int layer, int ph) {
    Phase f = phase_of(p, layer, ph);
    if (ph == 4) f.W = emb_img;
    return f;
}

/* ------------------------------------------------------------------ chunk schedule (producer and consumers agree on it) */
/* step -> layer -> {qkv, wo, w13, w2} -> logits; inside a phase: 16-row groups, inside a group: K/1024 chunks */
struct ChunkCursor {
    int step, layer, ph, g0, kc, nkc;
    Phase f;
    const uint16_t *emb_img;
    __device__ void set_phase(const DecParams &p) {
        f = tc_phase_of(p, emb_img, layer, ph);
        nkc = f.row_bytes / (TK_KC * 2); g0 = 0; kc = 0;
    }
    __device__ void start(const DecParams &p, const uint16_t *emb) { emb_img = emb; step = 0; layer = 0; ph = 0; set_phase(p); }
    __device__ void advance(const DecParams &p, int n_steps) {
        if (++kc < nkc) return;
        kc = 0; g0 += MK_GROUP;
        if (g0 < f.nrows) return;
        if (ph == 4) { ph = 0; layer = 0; step++; }
        else if (ph == 3) { if (layer == VOX_DEC_LAYERS - 1) ph = 4; else { layer++; ph = 0; } }
        else ph++;
        if (step < n_steps) set_phase(p);
    }
};

/* all 32 lanes of the producer warp run this; lane r copies row r of the chunk */
__device__ void tc_producer(uint8_t *slots, TcSmem *sm, const DecParams &p, const uint16_t *emb_img, int n_steps, int *err,
                            uint32_t &it_out, long long *prof) {
    const int lane = threadIdx.x & 31;
    ChunkCursor cur;
    cur.start(p, emb_img);
    uint32_t it = 0;
    long long t_wait = 0, t_begin = clock64();
    while (cur.step < n_steps) {
        const int s = (int)(it % TK_SLOTS);
        const uint32_t par = (it / TK_SLOTS) & 1u;
        long long t0 = 0;
        bool aborted = false;
        const long long tw0 = clock64();
        while (!mbar_try_wait(&sm->empty[s], par ^ 1u)) {
            if (sm->abort_flag) { aborted = true; break; }
            spin_guard(t0, err, 2);
        }
        aborted = __any_sync(0xffffffffu, aborted || sm->abort_flag);
        t_wait += clock64() - tw0;
        if (aborted) break;
        const int gr = min(MK_GROUP, cur.f.nrows - cur.g0);
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)gr * TK_ROW_BYTES;
            mbar_expect_tx(&sm->full[s], bytes);
            bulk_g2s(slots + (size_t)s * TK_SLOT_BYTES,
                     reinterpret_cast<const uint8_t *>(cur.f.W) + (size_t)(cur.f.row0 + cur.g0) * cur.f.row_bytes + (size_t)cur.kc * bytes,
                     bytes, &sm->full[s]);
        }
        it++;
        if (lane == 0) sm->prod_it = it;
        cur.advance(p, n_steps);
    }
    it_out = it;
    if (prof && lane == 0) {
        prof[(size_t)blockIdx.x * MK_PROF_SLOTS + MK_PROF_SLOTS - 1] = t_wait;
        prof[(size_t)blockIdx.x * MK_PROF_SLOTS + MK_PROF_SLOTS - 2] = clock64() - t_begin;
        prof[(size_t)blockIdx.x * MK_PROF_SLOTS + MK_PROF_SLOTS - 3] = it;
        unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        prof[(size_t)blockIdx.x * MK_PROF_SLOTS + MK_PROF_SLOTS - 4] = smid;
    }
}

/* ------------------------------------------------------------------ consumer */
struct TcRing {
    uint8_t *slots; TcSmem *sm; uint32_t it;
    __device__ __forceinline__ int slot() const { return (int)(it % TK_SLOTS); }
    __device__ __forceinline__ uint32_t parity() const { return (it / TK_SLOTS) & 1u; }
};

__device__ __forceinline__ uint32_t bf16_rn_bits(float f) {          /* finite inputs only */
    uint32_t u = __float_as_uint(f);
    return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
/* 8 consecutive activations -> their three bf16 planes (hi + mid + lo reproduces all 24 mantissa bits) */
__device__ __forceinline__ void store_planes8(uint8_t *planes, int K, int k, const float *v) {
    uint32_t w[3][4];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t h = bf16_rn_bits(v[j]);
        float r1 = v[j] - __uint_as_float(h << 16);
        uint32_t m = bf16_rn_bits(r1);
        float r2 = r1 - __uint_as_float(m << 16);
        uint32_t l = bf16_rn_bits(r2);
        if (j & 1) { w[0][j >> 1] |= h << 16; w[1][j >> 1] |= m << 16; w[2][j >> 1] |= l << 16; }
        else       { w[0][j >> 1] = h;        w[1][j >> 1] = m;        w[2][j >> 1] = l; }
    }
#pragma unroll
    for (int n = 0; n < 3; n++)
        *reinterpret_cast<uint4 *>(planes + (size_t)n * TK_PLANE_STRIDE(K) + (size_t)k * 2) = make_uint4(w[n][0], w[n][1], w[n][2], w[n][3]);
}
template <int CPT>
__device__ __forceinline__ void publish_planes(uint8_t *planes, int K, int NT, const float (&xr)[CPT * 8]) {
    const int t = threadIdx.x;
    if (t < NT) {
#pragma unroll
        for (int c = 0; c < CPT; c++) store_planes8(planes, K, (c * NT + t) * 8, &xr[c * 8]);
    }
    cons_bar();
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

/* Consume one GEMV phase from the ring.  K = row length of this phase's matrix (a multiple of 1024). */
template <typename Epi>
__device__ __forceinline__ void tc_phase(TcRing &r, const Phase &f, int K, const uint8_t *planes, int &redbuf, int *err, Epi epi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nkc = K / TK_KC;
    const uint8_t *bsrc = planes + (size_t)g * TK_PLANE_STRIDE(K);          /* only lanes with g < 3 load B */
    for (int g0 = 0; g0 < f.nrows; g0 += MK_GROUP) {
        const int gr = min(MK_GROUP, f.nrows - g0);
        float c[4] = { 0.f, 0.f, 0.f, 0.f };
        for (int kc = 0; kc < nkc; kc++) {
            const int s = r.slot();
            long long t0 = 0;
            while (!mbar_try_wait(&r.sm->full[s], r.parity())) spin_guard(t0, err, 3);
            const uint8_t *sb = r.slots + (size_t)s * TK_SLOT_BYTES;
#pragma unroll
            for (int b2 = 0; b2 < 2; b2++) {
                const int kb = (warp * 2 + b2) * 32 + t * 8;                  /* first of this lane's 8 k inside the chunk */
                const int pc = TK_SWZ(g, kb >> 3);                            /* rows g and g+8 have the same parity */
                const uint4 alo = *reinterpret_cast<const uint4 *>(sb + (size_t)g * TK_ROW_BYTES + (size_t)pc * 16);
                const uint4 ahi = *reinterpret_cast<const uint4 *>(sb + (size_t)(g + 8) * TK_ROW_BYTES + (size_t)pc * 16);
                uint4 bv = make_uint4(0u, 0u, 0u, 0u);
                if (g < 3) bv = *reinterpret_cast<const uint4 *>(bsrc + (size_t)(kc * TK_KC + kb) * 2);
                mma_bf16_16816(c, alo.x, ahi.x, alo.y, ahi.y, bv.x, bv.y);
                mma_bf16_16816(c, alo.z, ahi.z, alo.w, ahi.w, bv.z, bv.w);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&r.sm->empty[s]);
            r.it++;
        }
        /* columns 0..2 hold the hi/mid/lo partial dot products (3..7 are zero); sum them over the 4 lanes of a row */
        float rlo = c[0] + c[1], rhi = c[2] + c[3];
        rlo += __shfl_xor_sync(0xffffffffu, rlo, 1); rhi += __shfl_xor_sync(0xffffffffu, rhi, 1);
        rlo += __shfl_xor_sync(0xffffffffu, rlo, 2); rhi += __shfl_xor_sync(0xffffffffu, rhi, 2);
        if (t == 0) { r.sm->red[redbuf][warp][g] = rlo; r.sm->red[redbuf][warp][g + 8] = rhi; }
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

/* ------------------------------------------------------------------ the kernel */
extern __shared__ __align__(1024) uint8_t tk_smem_raw[];

__global__ void __launch_bounds__(TK_THREADS, 1) k_dec_tc(MegaArgs a) {
    uint8_t *slots = tk_smem_raw;
    uint8_t *planes = tk_smem_raw + (size_t)TK_SLOTS * TK_SLOT_BYTES;
    float *att_scr = reinterpret_cast<float *>(planes);             /* attention never overlaps a GEMV phase */
    TcSmem *sm = reinterpret_cast<TcSmem *>(planes + TK_PLANES_BYTES);
    const DecParams &p = a.p;
    const int tid = threadIdx.x;

    if (tid == 0) {
        for (int i = 0; i < TK_SLOTS; i++) { mbar_init(&sm->full[i], 1); mbar_init(&sm->empty[i], 16); }
        sm->abort_flag = 0; sm->prod_it = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= MK_CONS) {
        /* ===================== producer warp ===================== */
        uint32_t it = 0;
        tc_producer(slots, sm, p, a.emb_img, a.n_steps, a.err, it, a.prof);
        /* every bulk copy that was issued must land before the CTA may exit (smem is its target) */
        for (int back = 1; back <= TK_SLOTS; back++) {
            if (it < (uint32_t)back) break;
            uint32_t j = it - back;
            long long t0 = 0;
            while (!mbar_try_wait(&sm->full[j % TK_SLOTS], (j / TK_SLOTS) & 1u)) spin_guard(t0, a.err, 4);
        }
        __syncthreads();          /* matches the consumers' final barrier */
        return;
    }

    /* ===================== consumers ===================== */
    TcRing r{ slots, sm, 0u };
    unsigned int gen = 0;
    int redbuf = 0;
    int pos = a.pos0, token = a.token0, arow = a.adapter_row0;
    const float *adapter = *p.adapter_pp;
    const int lane = tid & 31;
    int n_done = 0, eos = 0, prof_n = 0;

    for (int step = 0; step < a.n_steps; step++) {
        const float *arow_p = adapter + (size_t)arow * VOX_DEC_DIM;
        const uint16_t *erow_p = p.tok_emb + (size_t)token * VOX_DEC_DIM;
        {   /* residual stream rows owned by this CTA: x = adapter[arow] + tok_emb[token] (voxtral.c:1057-1061) */
            int r0, n; rows_of(VOX_DEC_DIM, 1, r0, n);
            if (tid < n) p.x[r0 + tid] = arow_p[r0 + tid] + __uint_as_float((uint32_t)erow_p[r0 + tid] << 16);
        }
        const int slot = pos & (VB_KV_SLOTS - 1);

        for (int layer = 0; layer < VOX_DEC_LAYERS; layer++) {
            PROF(0);
            {   /* ---- RMSNorm -> [wq|wk|wv] -> RoPE -> KV ring ---- */
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
                } else load_x_cols_cg<1>(xr, p.x, NT);
                rmsnorm_cols_cons<1>(xr, p.attn_norm[layer], nullptr, NT, VOX_DEC_DIM, sm->sred);
                publish_planes<1>(planes, VOX_DEC_DIM, NT, xr);
                Phase f = tc_phase_of(p, a.emb_img, layer, 0);
                float *kdst = p.kv_k + ((size_t)layer * VB_KV_SLOTS + slot) * VB_DEC_KV;
                float *vdst = p.kv_v + ((size_t)layer * VB_KV_SLOTS + slot) * VB_DEC_KV;
                const float *inv_freq = p.inv_freq;
                float *q = p.q;
                tc_phase(r, f, VOX_DEC_DIM, planes, redbuf, a.err, [&](int row, float v, int, bool valid) {
                    float other = __shfl_xor_sync(0xffffffffu, v, 1);
                    if (!valid) return;
                    if (row < VB_DEC_Q + VB_DEC_KV) {
                        int d = (row & (HD - 1)) >> 1;
                        float sn, cs;
                        sincosf((float)pos * inv_freq[d], &sn, &cs);
                        float y = (row & 1) ? (other * sn + v * cs) : (v * cs - other * sn);
                        if (row < VB_DEC_Q) q[row] = y; else kdst[row - VB_DEC_Q] = y;
                    } else vdst[row - VB_DEC_Q - VB_DEC_KV] = v;
                });
            }
            PROF(1);
            grid_barrier(a.bar, gen, a.err);
            PROF(2);
            mega_attention(p, layer, pos, &sm->is_last, att_scr, a.bar + 16);
            PROF(3);
            grid_barrier(a.bar, gen, a.err);
            PROF(4);
            {   /* ---- wo + residual ---- */
                const int NT = VB_DEC_Q / 8;
                float xr[8];
                load_x_cols_cg<1>(xr, p.attn_out, NT);
                publish_planes<1>(planes, VB_DEC_Q, NT, xr);
                Phase f = tc_phase_of(p, a.emb_img, layer, 1);
                float *x = p.x;
                tc_phase(r, f, VB_DEC_Q, planes, redbuf, a.err, [&](int row, float v, int, bool valid) {
                    if (valid) x[row] = __ldcg(x + row) + v;
                });
            }
            PROF(5);
            grid_barrier(a.bar, gen, a.err);
            PROF(6);
            {   /* ---- RMSNorm*(1+ada) -> [w1|w3] -> SiLU(g)*u ---- */
                const int NT = VOX_DEC_DIM / 8;
                float xr[8];
                load_x_cols_cg<1>(xr, p.x, NT);
                rmsnorm_cols_cons<1>(xr, p.ffn_norm[layer], p.ada + (size_t)layer * VOX_DEC_DIM, NT, VOX_DEC_DIM, sm->sred);
                publish_planes<1>(planes, VOX_DEC_DIM, NT, xr);
                Phase f = tc_phase_of(p, a.emb_img, layer, 2);
                float *gate = p.gate;
                tc_phase(r, f, VOX_DEC_DIM, planes, redbuf, a.err, [&](int row, float v, int, bool valid) {
                    float other = __shfl_xor_sync(0xffffffffu, v, 1);
                    if (valid && !(row & 1)) gate[row >> 1] = vb_silu(v) * other;
                });
            }
            PROF(7);
            grid_barrier(a.bar, gen, a.err);
            PROF(8);
            {   /* ---- w2 + residual ---- */
                const int NT = VOX_DEC_HIDDEN / 8 / 3;
                float xr[24];
                load_x_cols_cg<3>(xr, p.gate, NT);
                publish_planes<3>(planes, VOX_DEC_HIDDEN, NT, xr);
                Phase f = tc_phase_of(p, a.emb_img, layer, 3);
                float *x = p.x;
                tc_phase(r, f, VOX_DEC_HIDDEN, planes, redbuf, a.err, [&](int row, float v, int, bool valid) {
                    if (valid) x[row] = __ldcg(x + row) + v;
                });
            }
            PROF(9);
            grid_barrier(a.bar, gen, a.err);
        }
        PROF(10);
        {   /* ---- final RMSNorm -> tied-embedding logits -> per-CTA argmax ---- */
            const int NT = VOX_DEC_DIM / 8;
            float xr[8];
            load_x_cols_cg<1>(xr, p.x, NT);
            rmsnorm_cols_cons<1>(xr, p.final_norm, nullptr, NT, VOX_DEC_DIM, sm->sred);
            publish_planes<1>(planes, VOX_DEC_DIM, NT, xr);
            Phase f = tc_phase_of(p, a.emb_img, 0, 4);
            float *logits = p.logits;
            unsigned long long best = 0ull;
            tc_phase(r, f, VOX_DEC_DIM, planes, redbuf, a.err, [&](int row, float v, int, bool valid) {
                if (!valid) return;
                logits[row] = v;
                unsigned long long cd = pack_cand(v, row);
                if (cd > best) best = cd;
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
        PROF(11);
        grid_barrier(a.bar, gen, a.err);
        PROF(12);
        {   /* global argmax: every CTA reduces the per-CTA candidates, so every CTA knows the token */
            unsigned long long best = 0ull;
            for (int i = tid; i < (int)gridDim.x; i += MK_CONS) {
                unsigned long long cd = __ldcg(p.argmax + i);
                if (cd > best) best = cd;
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

/* ------------------------------------------------------------------ weight image */
/* Launched with the decode grid: CTA b re-tiles exactly the rows rows_of() gives it in the decode kernel, so the image is a
 * permutation inside each CTA's slab and slab offsets are those of the row-major matrix. */
__global__ void __launch_bounds__(512) k_build_tc_image(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int total_units,
                                                         int unit_rows, int K) {
    int row0, nrows;
    rows_of(total_units, unit_rows, row0, nrows);
    const int ppr = K / 8;                                   /* 16-byte pieces per row */
    const int ppc = TK_KC / 8;                               /* ... per chunk row */
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    for (long long i = threadIdx.x; i < (long long)nrows * ppr; i += blockDim.x) {
        const int r = (int)(i / ppr), piece = (int)(i % ppr);
        const int g0 = r & ~(MK_GROUP - 1), rin = r & (MK_GROUP - 1);
        const int gr = min(MK_GROUP, nrows - g0);
        const int kc = piece / ppc, c = piece % ppc;
        const size_t dpiece = (size_t)(row0 + g0) * ppr + (size_t)kc * gr * ppc + (size_t)rin * ppc + TK_SWZ(rin, c);
        d4[dpiece] = s4[(size_t)(row0 + r) * ppr + piece];
    }
}

struct TcImageTab { const uint16_t *wqkv[VOX_DEC_LAYERS], *wo[VOX_DEC_LAYERS], *w13[VOX_DEC_LAYERS], *w2[VOX_DEC_LAYERS], *emb; };
static TcImageTab g_tc_tab;               /* one engine per process (one process per GPU) */
static VbEngine *g_tc_tab_owner = NULL;

static void tc_prepare_images(VbEngine *e) {
    if (e->d_tc_img && g_tc_tab_owner == e) return;
    const size_t n_qkv = (size_t)VB_DEC_QKV * VOX_DEC_DIM, n_wo = (size_t)VOX_DEC_DIM * VB_DEC_Q;
    const size_t n_w13 = (size_t)2 * VOX_DEC_HIDDEN * VOX_DEC_DIM, n_w2 = (size_t)VOX_DEC_DIM * VOX_DEC_HIDDEN;
    const size_t n_emb = (size_t)VOX_VOCAB_SIZE * VOX_DEC_DIM;
    const size_t total = (size_t)VOX_DEC_LAYERS * (n_qkv + n_wo + n_w13 + n_w2) + n_emb;
    const size_t wb = e->weight_bytes;
    e->d_tc_img = (uint16_t *)vb_dev_alloc_owned(e, total * 2);
    e->weight_bytes = wb;                                    /* a derived copy, not checkpoint bytes */
    DecParams p = vb_make_dec_params(e, 1);
    uint16_t *cur = e->d_tc_img;
    auto build = [&](const uint16_t *src, size_t n, int units, int unit_rows, int K) {
        k_build_tc_image<<<e->sm_count, 512, 0, e->stream>>>(src, cur, units, unit_rows, K);
        e->launches += 1;
        const uint16_t *at = cur; cur += n; return at;
    };
    for (int l = 0; l < VOX_DEC_LAYERS; l++) {               /* units / unit_rows must match phase_of() */
        g_tc_tab.wqkv[l] = build(p.wqkv[l], n_qkv, VB_DEC_QKV / 2, 2, VOX_DEC_DIM);
        g_tc_tab.wo[l]   = build(p.wo[l],   n_wo,  VOX_DEC_DIM, 1, VB_DEC_Q);
        g_tc_tab.w13[l]  = build(p.w13[l],  n_w13, VOX_DEC_HIDDEN, 2, VOX_DEC_DIM);
        g_tc_tab.w2[l]   = build(p.w2[l],   n_w2,  VOX_DEC_DIM, 1, VOX_DEC_HIDDEN);
    }
    g_tc_tab.emb = build(p.tok_emb, n_emb, VOX_VOCAB_SIZE, 1, VOX_DEC_DIM);
    VB_CUDA_OK(cudaGetLastError());
    g_tc_tab_owner = e;
}

/* ------------------------------------------------------------------ host */
static size_t tc_smem_bytes() { return (size_t)TK_SLOTS * TK_SLOT_BYTES + TK_PLANES_BYTES + sizeof(TcSmem) + 64; }

extern "C" int vb_decoder_tc_supported(VbEngine *e) {
    static int cached = -1;
    if (cached >= 0) return cached;
    int coop = 0, max_smem = 0, blocks = 0;
    static_assert(TK_PLANES_BYTES >= MK_ATT_FLOATS * 4, "attention scratch aliases the activation planes");
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device);
    if (!coop || e->sm_count > 160 || (size_t)max_smem < tc_smem_bytes()) { cached = 0; return 0; }
    if (cudaFuncSetAttribute(k_dec_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes()) != cudaSuccess) { cached = 0; return 0; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_dec_tc, TK_THREADS, tc_smem_bytes()) != cudaSuccess || blocks < 1) { cached = 0; return 0; }
    cached = 1;
    return 1;
}

extern "C" int vb_decoder_tc_launch(VbEngine *e, const float *d_adapter, int adapter_row, int n_steps, int prev_token, int pos) {
    if (!e->d_mega_bar) {
        const size_t wb = e->weight_bytes;
        e->d_mega_bar = (unsigned int *)vb_dev_alloc_owned(e, 256);
        e->weight_bytes = wb;
    }
    tc_prepare_images(e);
    struct { VbDecState st; const float *adapter; } h;
    memset(&h, 0, sizeof h);
    h.st.pos = pos; h.st.token = prev_token; h.st.adapter_row = adapter_row; h.adapter = d_adapter;
    VB_CUDA_OK(cudaMemcpyAsync(e->d_state, &h, sizeof h, cudaMemcpyHostToDevice, e->stream));
    VB_CUDA_OK(cudaMemsetAsync(e->d_mega_bar, 0, 256, e->stream));
    MegaArgs a;
    a.p = vb_make_dec_params(e, 1);
    for (int l = 0; l < VOX_DEC_LAYERS; l++) {
        a.p.wqkv[l] = g_tc_tab.wqkv[l]; a.p.wo[l] = g_tc_tab.wo[l]; a.p.w13[l] = g_tc_tab.w13[l]; a.p.w2[l] = g_tc_tab.w2[l];
    }
    a.emb_img = g_tc_tab.emb;
    a.n_steps = n_steps; a.pos0 = pos; a.token0 = prev_token; a.adapter_row0 = adapter_row;
    a.bar = e->d_mega_bar; a.err = (int *)(e->d_mega_bar + 32);
    a.l2_ahead = 0;
    vb_mega_prof_begin(e, a, n_steps);
    void *args[] = { &a };
    VB_CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_dec_tc, dim3(e->sm_count), dim3(TK_THREADS), args, tc_smem_bytes(), e->stream));
    e->launches += 1;
    vb_mega_prof_report(e, a, "tc-ring");
    return 0;
}
