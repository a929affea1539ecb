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
 *            it with cp.async.bulk (TMA 1-D bulk copies, mbarrier complete_tx) into a 7 x 24 KB shared-
 *            memory ring, and keeps a further 320 KB per CTA prefetched into L2 (cp.async.bulk.prefetch.L2).  Weights do not depend on activations, so the producer NEVER waits for a phase
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
#include "vb_decode_persist_common.cuh"
#include <string.h>

#define MK_THREADS   (MK_CONS + 32)            /* + 1 producer warp */
#define MK_SLOTS     7
#define MK_SLOT_BYTES 24576                    /* 4 rows of K=3072 / 3 rows of K=4096 / 1 row of K=9216 (18 KB) */
#define MK_L2_AHEAD  (320 * 1024)              /* bytes per CTA the producer keeps prefetched into L2 beyond the ring */

/* ------------------------------------------------------------------ shared state */
struct MegaSmem {
    uint64_t full[MK_SLOTS], empty[MK_SLOTS];
    float red[2][16][MK_GROUP];
    float sred[16];
    unsigned long long cand[16];
    volatile int abort_flag;                    /* consumers -> producer: stop issuing */
    volatile int is_last;                       /* attention ticket result for this CTA */
};

struct Ring {
    uint8_t *slots; MegaSmem *sm; uint32_t it;  /* it: chunks handled so far by this thread's role */
    __device__ __forceinline__ int slot() const { return (int)(it % MK_SLOTS); }
    __device__ __forceinline__ uint32_t parity() const { return (it / MK_SLOTS) & 1u; }
};

/* ------------------------------------------------------------------ producer */
/* Enumerates the CTA's weight chunks in consumption order: step -> layer -> {qkv, wo, w13, w2} -> logits.
 * Chunk boundaries restart at every 16-row reduction group so that consumer groups own whole chunks. */
struct SlabCursor {
    int step, layer, ph, g0, c0;
    Phase f;
    __device__ void start(const DecParams &p) { step = 0; layer = 0; ph = 0; g0 = 0; c0 = 0; f = phase_of(p, 0, 0); }
    __device__ bool next(const DecParams &p, int n_steps, const uint8_t *&ptr, uint32_t &bytes) {
        if (step >= n_steps) return false;
        const int gr = min(MK_GROUP, f.nrows - g0);
        const int rows = min(f.rc, gr - c0);
        ptr = reinterpret_cast<const uint8_t *>(f.W) + (size_t)(f.row0 + g0 + c0) * f.row_bytes;
        bytes = (uint32_t)rows * (uint32_t)f.row_bytes;
        c0 += f.rc;
        if (c0 >= gr) {
            c0 = 0; g0 += MK_GROUP;
            if (g0 >= f.nrows) {
                g0 = 0;
                if (ph == 4) { ph = 0; layer = 0; step++; }
                else if (ph == 3) { if (layer == VOX_DEC_LAYERS - 1) ph = 4; else { layer++; ph = 0; } }
                else ph++;
                if (step < n_steps) f = phase_of(p, layer, ph);
            }
        }
        return true;
    }
};

__device__ void producer_run(uint8_t *slots, MegaSmem *sm, const DecParams &p, int n_steps, int l2_ahead, int *err, uint32_t &it_out) {
    SlabCursor cur, pre;
    cur.start(p); pre.start(p);
    long long issued = 0, prefetched = 0;
    bool pre_live = true;
    uint32_t it = 0;
    const uint8_t *ptr; uint32_t bytes;
    while (!sm->abort_flag && cur.next(p, n_steps, ptr, bytes)) {
        while (pre_live && prefetched < issued + l2_ahead) {
            const uint8_t *pp; uint32_t pb;
            pre_live = pre.next(p, n_steps, pp, pb);
            if (pre_live) { l2_prefetch(pp, pb); prefetched += pb; }
        }
        const int s = (int)(it % MK_SLOTS);
        const uint32_t par = (it / MK_SLOTS) & 1u;
        long long t0 = 0;
        bool aborted = false;
        while (!mbar_try_wait(&sm->empty[s], par ^ 1u)) {
            if (sm->abort_flag) { aborted = true; break; }
            spin_guard(t0, err, 2);
        }
        if (aborted || sm->abort_flag) break;
        mbar_expect_tx(&sm->full[s], bytes);
        bulk_g2s(slots + (size_t)s * MK_SLOT_BYTES, ptr, bytes, &sm->full[s]);
        it++; issued += bytes;
    }
    it_out = it;
}

/* ------------------------------------------------------------------ ring consumer */
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
        for (int cb = 0; cb < (MK_GROUP + RC - 1) / RC; cb++) {
            if (cb * RC < gr) {
                const int s = r.slot();
                long long t0 = 0;
                while (!mbar_try_wait(&r.sm->full[s], r.parity())) spin_guard(t0, err, 3);
                const uint8_t *base = r.slots + (size_t)s * MK_SLOT_BYTES + (size_t)t * 16;
#pragma unroll
                for (int rr = 0; rr < RC; rr++) {
                    if (cb * RC + rr < MK_GROUP && cb * RC + rr < gr && active) {
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

/* ------------------------------------------------------------------ the kernel */
extern __shared__ __align__(1024) uint8_t mk_smem_raw[];

__global__ void __launch_bounds__(MK_THREADS, 1) k_dec_mega(MegaArgs a) {
    uint8_t *slots = mk_smem_raw;
    float *att_scr = reinterpret_cast<float *>(mk_smem_raw + (size_t)MK_SLOTS * MK_SLOT_BYTES);
    MegaSmem *sm = reinterpret_cast<MegaSmem *>(mk_smem_raw + (size_t)MK_SLOTS * MK_SLOT_BYTES + (size_t)MK_ATT_FLOATS * 4);
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
            uint32_t it = 0;
            producer_run(slots, sm, p, a.n_steps, a.l2_ahead, a.err, it);
            /* every bulk copy that was issued must land before the CTA may exit (smem is its target):
             * chunks it-1 .. it-MK_SLOTS are the only ones that can still be in flight */
            for (int back = 1; back <= MK_SLOTS; back++) {
                if (it < (uint32_t)back) break;
                uint32_t j = it - back;
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
    int n_done = 0, eos = 0, prof_n = 0;

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
            PROF(0);
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
            PROF(1);
            grid_barrier(a.bar, gen, a.err);
            PROF(2);
            /* ---- attention ---- */
            mega_attention(p, layer, pos, &sm->is_last, att_scr, a.bar + 16);
            PROF(3);
            grid_barrier(a.bar, gen, a.err);
            PROF(4);
            /* ---- wo + residual ---- */
            {
                const int NT = VB_DEC_Q / 8;
                float xr[8];
                load_x_cols_cg<1>(xr, p.attn_out, NT);
                Phase f = phase_of(p, layer, 1);
                float *x = p.x;
                consume_phase<1, 3>(r, f, NT, xr, redbuf, a.err, [&](int row, float v, int, bool valid) {
                    if (valid) x[row] = __ldcg(x + row) + v;
                });
            }
            PROF(5);
            grid_barrier(a.bar, gen, a.err);
            PROF(6);
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
            PROF(7);
            grid_barrier(a.bar, gen, a.err);
            PROF(8);
            /* ---- w2 + residual ---- */
            {
                const int NT = VOX_DEC_HIDDEN / 8 / 3;
                float xr[24];
                load_x_cols_cg<3>(xr, p.gate, NT);
                Phase f = phase_of(p, layer, 3);
                float *x = p.x;
                consume_phase<3, 1>(r, f, NT, xr, redbuf, a.err, [&](int row, float v, int, bool valid) {
                    if (valid) x[row] = __ldcg(x + row) + v;
                });
            }
            PROF(9);
            grid_barrier(a.bar, gen, a.err);
        }
        PROF(10);
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
        PROF(11);
        grid_barrier(a.bar, gen, a.err);
        PROF(12);
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
static size_t mega_smem_bytes() { return (size_t)MK_SLOTS * MK_SLOT_BYTES + (size_t)MK_ATT_FLOATS * 4 + sizeof(MegaSmem) + 64; }

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
    const char *la = getenv("VOX_CUDA_L2_AHEAD");
    a.l2_ahead = la ? atoi(la) : MK_L2_AHEAD;
    vb_mega_prof_begin(e, a, n_steps);
    void *args[] = { &a };
    VB_CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_dec_mega, dim3(e->sm_count), dim3(MK_THREADS), args,
                                           mega_smem_bytes(), e->stream));
    e->launches += 1;
    vb_mega_prof_report(e, a, "tma-ring");
    return 0;
}

/* ---- optional in-kernel phase profile (VOX_CUDA_MEGA_PROF=<step>) shared by both persistent kernels ---- */
static long long *g_d_prof = NULL;
void vb_mega_prof_begin(VbEngine *e, MegaArgs &a, int n_steps) {
    a.prof = NULL; a.prof_step = -1;
    const char *pe = getenv("VOX_CUDA_MEGA_PROF");
    if (pe && n_steps > atoi(pe)) {
        if (!g_d_prof) g_d_prof = (long long *)vb_dev_alloc((size_t)e->sm_count * MK_PROF_SLOTS * 8);
        VB_CUDA_OK(cudaMemsetAsync(g_d_prof, 0, (size_t)e->sm_count * MK_PROF_SLOTS * 8, e->stream));
        a.prof = g_d_prof; a.prof_step = atoi(pe);
    }
}
void vb_mega_prof_report(VbEngine *e, const MegaArgs &a, const char *label) {
    if (!a.prof) return;
    size_t n = (size_t)e->sm_count * MK_PROF_SLOTS;
    long long *h = (long long *)malloc(n * 8);
    VB_CUDA_OK(cudaMemcpyAsync(h, g_d_prof, n * 8, cudaMemcpyDeviceToHost, e->stream));
    VB_CUDA_OK(cudaStreamSynchronize(e->stream));
    /* per-layer stamps: 0 start, 1 qkv done, 2 bar, 3 attention done, 4 bar, 5 wo, 6 bar, 7 w13, 8 bar, 9 w2, (next 0 after bar) */
    static const char *names[10] = { "qkv", "bar", "attn", "bar", "wo", "bar", "w13", "bar", "w2", "bar" };
    const int ctas[3] = { 0, e->sm_count / 2, e->sm_count - 1 };
    for (int ci = 0; ci < 3; ci++) {
        long long *t = h + (size_t)ctas[ci] * MK_PROF_SLOTS;
        double sum[10] = { 0 };
        for (int l = 1; l < VOX_DEC_LAYERS - 1; l++)
            for (int k = 0; k < 10; k++) sum[k] += (double)(t[l * 10 + k + 1] - t[l * 10 + k]);
        fprintf(stderr, "[%s prof] cta %3d cycles/layer:", label, ctas[ci]);
        double tot = 0;
        for (int k = 0; k < 10; k++) { fprintf(stderr, " %s=%.0f", names[k], sum[k] / (VOX_DEC_LAYERS - 2)); tot += sum[k]; }
        fprintf(stderr, " | layer=%.0f | logits=%lld bar=%lld step=%lld\n", tot / (VOX_DEC_LAYERS - 2),
                t[26 * 10 + 1] - t[26 * 10], t[26 * 10 + 2] - t[26 * 10 + 1], t[26 * 10 + 2] - t[0]);
        if (!strcmp(label, "tc-ring")) {               /* producer lead (chunks) at each stamp, mean over layers */
            fprintf(stderr, "[%s prof] cta %3d ring lead at stamp:", label, ctas[ci]);
            for (int k = 0; k < 10; k++) {
                double s = 0;
                for (int l = 1; l < VOX_DEC_LAYERS - 1; l++) s += (double)t[MK_PROF_SLOTS / 2 + l * 10 + k];
                fprintf(stderr, " %d:%.1f", k, s / (VOX_DEC_LAYERS - 2));
            }
            fprintf(stderr, " | producer: %lld chunks, %.0f cycles/chunk of which %.0f waiting for a free slot\n", t[MK_PROF_SLOTS - 3],
                    (double)t[MK_PROF_SLOTS - 2] / (double)t[MK_PROF_SLOTS - 3], (double)t[MK_PROF_SLOTS - 1] / (double)t[MK_PROF_SLOTS - 3]);
        }
    }
    if (!strcmp(label, "tc-ring") && getenv("VOX_CUDA_MEGA_PROF_ALL")) {
        /* per CTA: SM id, cycles in the logits phase, summed cycles of the four GEMV phases of layers 1..24 */
        for (int c = 0; c < e->sm_count; c++) {
            long long *t = h + (size_t)c * MK_PROF_SLOTS;
            double ph = 0;
            for (int l = 1; l < VOX_DEC_LAYERS - 1; l++)
                ph += (double)((t[l * 10 + 1] - t[l * 10]) + (t[l * 10 + 5] - t[l * 10 + 4]) + (t[l * 10 + 7] - t[l * 10 + 6]) + (t[l * 10 + 9] - t[l * 10 + 8]));
            fprintf(stderr, "[tc-ring sm] cta %3d smid %3lld logits %lld phases/layer %.0f\n", c, t[MK_PROF_SLOTS - 4], t[26 * 10 + 1] - t[26 * 10], ph / (VOX_DEC_LAYERS - 2));
        }
    }
    free(h);
}
