/*
 * vb_decode_v2.cu -- persistent decode kernel, second generation: a decoupled weight stream and B activation columns.
 *
 * What round 1 measured (profiles/r01_decode.md): the GEMV core streams at 99 % of the HBM peak inside the long logits
 * phase, but a decode step is 131 dependent phases of 4-20 us and every phase boundary (drain, grid barrier, activation
 * reload, RMSNorm, ramp) idles HBM for ~5 us: 0.59 of the roofline.  And one stream can never beat one weight pass
 * (6.86 GB) per 80 ms of audio.  This kernel attacks both:
 *
 *   weight stream   one producer warp per CTA walks the WHOLE schedule (step -> layer -> phase) independently of the
 *                   activations and copies weight rows with cp.async.bulk (TMA, mbarrier complete_tx) into an 8 x 24 KB
 *                   shared-memory ring.  It never waits for a phase boundary: while the 12 consumer warps sit in a grid
 *                   barrier, in attention or in an RMSNorm, the next phase's rows keep arriving, and the consumers (which
 *                   read shared memory several times faster than HBM delivers) catch up afterwards.  The number of bulk
 *                   copies in flight per SM is capped (a.inflight_max) so that barrier polls and activation reloads
 *                   do not queue behind megabytes of prefetch -- the effect that cancelled round 1's gains.
 *   work balance    SMs do not get equal shares of HBM bandwidth (10-15 % spread).  The three large phases (QKV, w1|w3,
 *                   logits) are therefore handed out dynamically: a chunk = 4 consecutive output rows, taken from a
 *                   per-phase atomic counter by the producer, three grabs ahead.  Every output row is still computed by
 *                   exactly one CTA in a fixed order, so results do not depend on who took which chunk.
 *   B columns       every weight element read from shared memory is multiplied into NB activation columns (template
 *                   parameter; FFMA2 on column pairs).  A column is an independent stream (own KV ring, position,
 *                   adapter rows, token feedback) -- several vox_stream_t share one weight pass -- or, in verify mode,
 *                   consecutive positions of ONE stream (drafted tokens, longest matching prefix accepted).
 *                   The arithmetic of a column (FMA order, reduction tree) is the same for every NB, so a stream decodes
 *                   to the same ids alone or batched.
 *   uniform rows    wo (K=4096) and w2 (K=9216) are walked in 2 / 3 column blocks of 2048 / 3072 with the partial sums
 *                   kept per row in shared memory, so every phase has 8 activations x NB per thread in registers and the
 *                   kernel has one GEMV loop.  These two phases keep a static row partition (a row's blocks must meet in
 *                   one CTA).
 *
 * Per-step math and epilogues are those of vb_decode_persist.cu / the reference: voxtral_decoder.c:586-706, loop
 * voxtral.c:1056-1093.
 */
#include "vb_decode_persist_common.cuh"
#include <string.h>

#define V2_CONS        384                     /* consumer threads: 12 warps, 8 k-columns each for a 3072-wide row */
#define V2_CW          (V2_CONS / 32)
#define V2_THREADS     (V2_CONS + 32)          /* + producer warp */
#define V2_SLOTS       8
#define V2_SLOT_BYTES  24576                   /* 4 rows of 3072 bf16 */
#define V2_RC          4                       /* rows per chunk */
#define V2_MAXB        8
#define V2_SUBPHASES   (VOX_DEC_LAYERS * 4 + 1)
#define V2_ATT_FLOATS  (V2_CW * 4 * 132)
#define V2_PROF_SLOTS  320
#define V2_MAX_STEPS   2048
#define V2_NSPLIT_MAX  20
#define V2_DRAFT_TAB   512                     /* entries of the per-CTA successor table (verify mode) */

struct V2Col {
    const float *adapter;                      /* adapter rows of this stream */
    float *kv_k, *kv_v;                        /* [26][8192][1024] ring of this stream */
    float *logits;                             /* [131072] */
    int *tokens;                               /* out: generated ids */
    int pos0, token0, arow0, n_steps;
};

struct V2Args {
    DecParams p;                               /* weights (activation pointers in there are not used) */
    V2Col col[V2_MAXB];
    float *x, *q, *attn_out, *gate;            /* [nb][3072], [nb][4096], [nb][4096], [nb][9216] */
    float *part_m, *part_l, *part_o;           /* [nb][20][32], [nb][20][32], [nb][20][32][128] */
    unsigned long long *argmax;                /* [grid][V2_MAXB] */
    unsigned int *bar;                         /* [0] grid barrier, [32] error word, [64 + column * 8 + kv head] attention tickets */
    unsigned int *ctr;                         /* [n_steps][V2_SUBPHASES] chunk counters of the dynamic phases */
    int *err;                                  /* host-mapped word: which wait gave up (spin_guard code) -- survives the trap */
    VbDecState *st_out;                        /* [nb] */
    int nb, n_steps, inflight_max, dynamic, verify, dbg;
    long long *prof; int prof_step;
};

struct V2Smem {
    uint64_t full[V2_SLOTS], empty[V2_SLOTS];
    int meta_row0[V2_SLOTS], meta_nrows[V2_SLOTS];
    float red[2][V2_CW][32];
    float sred[V2_CW][V2_MAXB];
    float part[24][V2_MAXB];                    /* per-row partial sums of the column-block phases (wo, w2) */
    unsigned long long cand[V2_CW][V2_MAXB];
    /* column state, identical in every CTA */
    int c_pos[V2_MAXB], c_token[V2_MAXB], c_arow[V2_MAXB], c_done[V2_MAXB], c_nout[V2_MAXB], c_left[V2_MAXB];
    const float *c_adapter[V2_MAXB]; float *c_kv_k[V2_MAXB], *c_kv_v[V2_MAXB], *c_logits[V2_MAXB]; int *c_tokens[V2_MAXB];
    /* verify mode (one stream, columns = drafted positions): stream state, a small successor table as the drafter, counters */
    int v_pos, v_token, v_arow, v_left, v_nout, v_eos, v_passes;
    int dr_key[V2_DRAFT_TAB], dr_succ[V2_DRAFT_TAB];
    volatile int abort_flag, is_last;
    long long prof_t, prof_acc[16]; int prof_cnt[8];   /* profiled step: cycles before / inside v2_consume and chunks per sub-phase kind */
};

__device__ __forceinline__ bool mbar_test_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}

__device__ __forceinline__ void v2_bar() { asm volatile("bar.sync 1, %0;" :: "n"(V2_CONS) : "memory"); }

__device__ __forceinline__ void v2_grid_barrier(unsigned int *bar, unsigned int &gen, int *err) {
    gen++;
    v2_bar();
    if (threadIdx.x == 0) {
        red_release_add(bar, 1u);
        const unsigned int target = gen * gridDim.x;
        long long t0 = 0;
        while (ld_acquire_u32(bar) < target) spin_guard(t0, err, 1);
    }
    v2_bar();
}

/* ------------------------------------------------------------------ the weight schedule */
struct V2Phase { const uint8_t *W; int row_stride, seg_bytes, col_off, total_rows, dyn; };

/* ph: 0 QKV, 1 WO (2 column blocks), 2 W13, 3 W2 (3 column blocks), 4 LOGITS */
__device__ __forceinline__ V2Phase v2_phase(const DecParams &p, int layer, int ph, int part) {
    V2Phase f;
    switch (ph) {
    case 0:  f.W = (const uint8_t *)p.wqkv[layer]; f.row_stride = VOX_DEC_DIM * 2; f.seg_bytes = VOX_DEC_DIM * 2; f.col_off = 0;
             f.total_rows = VB_DEC_QKV; f.dyn = 1; break;
    case 1:  f.W = (const uint8_t *)p.wo[layer]; f.row_stride = VB_DEC_Q * 2; f.seg_bytes = 4096; f.col_off = part * 4096;
             f.total_rows = VOX_DEC_DIM; f.dyn = 0; break;
    case 2:  f.W = (const uint8_t *)p.w13[layer]; f.row_stride = VOX_DEC_DIM * 2; f.seg_bytes = VOX_DEC_DIM * 2; f.col_off = 0;
             f.total_rows = 2 * VOX_DEC_HIDDEN; f.dyn = 1; break;
    case 3:  f.W = (const uint8_t *)p.w2[layer]; f.row_stride = VOX_DEC_HIDDEN * 2; f.seg_bytes = VOX_DEC_DIM * 2; f.col_off = part * VOX_DEC_DIM * 2;
             f.total_rows = VOX_DEC_DIM; f.dyn = 0; break;
    default: f.W = (const uint8_t *)p.tok_emb; f.row_stride = VOX_DEC_DIM * 2; f.seg_bytes = VOX_DEC_DIM * 2; f.col_off = 0;
             f.total_rows = VOX_VOCAB_SIZE; f.dyn = 1; break;
    }
    return f;
}

__device__ __forceinline__ void v2_static_rows(int total, int unit, int &r0, int &r1) {
    const int units = total / unit;
    r0 = (int)((long long)units * blockIdx.x / gridDim.x) * unit;
    r1 = (int)((long long)units * (blockIdx.x + 1) / gridDim.x) * unit;
}

/* ------------------------------------------------------------------ producer */
/* ONE thread per CTA.  It is latency bound by construction (every chunk is a chain of try_wait -> meta -> expect_tx -> bulk copy),
 * so all of its state lives in registers and there is a single copy of the per-chunk code: a flat loop over the (sub)phases.
 * (A first version kept the state in a struct handed to a noinline function: ~30 local-memory accesses per chunk made the
 * producer, not HBM, the bottleneck at 17 B/cycle/SM.) */
__device__ void v2_producer(V2Smem *sm, uint8_t *slots, const V2Args &a) {
    int *err = a.err;
    uint32_t it = 0, landed = 0;
    const uint32_t inflight_max = (uint32_t)a.inflight_max;
    bool dead = false;
    long long n_chunks = 0, t_wait = 0, t_cap = 0;
    const bool prof = a.prof != nullptr;
    const long long t_begin = clock64();

    /* a free slot, and room under the in-flight cap; false = the consumers left (EOS): stop */
    auto acquire = [&]() -> bool {
        const int s = (int)(it % V2_SLOTS);
        const uint32_t par = (it / V2_SLOTS) & 1u;
        long long t0 = 0;
        long long ta = 0;
        if (prof) ta = clock64();
        while (!mbar_try_wait(&sm->empty[s], par ^ 1u)) {
            if (sm->abort_flag) { dead = true; return false; }
            spin_guard(t0, err, 2);
        }
        long long tb = 0;
        if (prof) { tb = clock64(); t_wait += tb - ta; }
        while (it - landed >= inflight_max) {
            if (mbar_try_wait(&sm->full[landed % V2_SLOTS], (landed / V2_SLOTS) & 1u)) landed++;
            else { if (sm->abort_flag) { dead = true; return false; } spin_guard(t0, err, 5); }
        }
        if (prof) t_cap += clock64() - tb;
        return true;
    };

    const int n_sub_total = a.n_steps * (VOX_DEC_LAYERS * 7 + 1);
#pragma unroll 1
    for (int flat = 0; flat < n_sub_total && !dead; flat++) {
        const int step = flat / (VOX_DEC_LAYERS * 7 + 1), idx = flat - step * (VOX_DEC_LAYERS * 7 + 1);
        const int layer = idx / 7;
        const int sub = layer == VOX_DEC_LAYERS ? 7 : idx - layer * 7;
        /* sub 0 QKV | 1,2 wo blocks | 3 w1|w3 | 4,5,6 w2 blocks | 7 logits */
        const int ph = sub == 0 ? 0 : sub <= 2 ? 1 : sub == 3 ? 2 : sub <= 6 ? 3 : 4;
        const int part = ph == 1 ? sub - 1 : ph == 3 ? sub - 4 : 0;
        const V2Phase f = v2_phase(a.p, layer == VOX_DEC_LAYERS ? 0 : layer, ph, part);
        const bool dyn = f.dyn && a.dynamic;
        unsigned int *ctr = a.ctr + (size_t)step * V2_SUBPHASES + (ph == 4 ? VOX_DEC_LAYERS * 4 : layer * 4 + ph);
        /* one chunk: slot, meta, expect_tx, bulk copy (one copy when the rows are contiguous, else one per row segment) */
        auto emit = [&](int row0, int nrows) -> bool {
            if (!acquire()) return false;
            const int s = (int)(it % V2_SLOTS);
            sm->meta_row0[s] = row0; sm->meta_nrows[s] = nrows;
            uint8_t *dst = slots + (size_t)s * V2_SLOT_BYTES;
            const uint8_t *src = f.W + (size_t)row0 * f.row_stride + f.col_off;
            mbar_expect_tx(&sm->full[s], (uint32_t)(nrows * f.seg_bytes));
            if (f.seg_bytes == f.row_stride) bulk_g2s(dst, src, (uint32_t)(nrows * f.seg_bytes), &sm->full[s]);
            else for (int k = 0; k < nrows; k++) bulk_g2s(dst + (size_t)k * f.seg_bytes, src + (size_t)k * f.row_stride, (uint32_t)f.seg_bytes, &sm->full[s]);
            it++; n_chunks++;
            return true;
        };
        if (dyn) {
            /* chunk ids from the phase's global counter.  An atomic takes ~2000 cycles when the memory system is busy, twice a
             * chunk's time: four grabs are kept in flight in four distinct registers (a rotating "g0 = g1" window would read the
             * newest result every iteration and serialise on it -- measured: 900 cycles per chunk). */
            const unsigned int n_dyn = (unsigned int)(f.total_rows / V2_RC);
            unsigned int g[4];
#pragma unroll
            for (int u = 0; u < 4; u++) g[u] = atomicAdd(ctr, 1u);
            bool more = true;
            while (more) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const unsigned int id = g[u];
                    if (id >= n_dyn) { more = false; break; }             /* ids only grow: the other three are past the end too */
                    g[u] = atomicAdd(ctr, 1u);
                    if (!emit((int)id * V2_RC, V2_RC)) { more = false; break; }
                }
            }
        } else {
            int r, r1;
            v2_static_rows(f.total_rows, f.dyn ? V2_RC : 1, r, r1);
            for (; r < r1; r += V2_RC) if (!emit(r, min(V2_RC, r1 - r))) break;
        }
        if (dead || !acquire()) break;
        {   /* end marker: "this CTA has no more rows in this phase" */
            const int s = (int)(it % V2_SLOTS);
            sm->meta_row0[s] = 0; sm->meta_nrows[s] = 0;
            mbar_arrive(&sm->full[s]);
            it++;
        }
    }
    if (a.prof) {
        long long *pp = a.prof + (size_t)blockIdx.x * V2_PROF_SLOTS + V2_PROF_SLOTS - 8;
        pp[0] = n_chunks; pp[1] = t_wait; pp[2] = t_cap; pp[3] = clock64() - t_begin;
    }
    /* every bulk copy that was issued must land before the CTA may exit (shared memory is its target) */
    while (landed < it) {
        long long t0 = 0;
        while (!mbar_try_wait(&sm->full[landed % V2_SLOTS], (landed / V2_SLOTS) & 1u)) spin_guard(t0, err, 4);
        landed++;
    }
}

/* ------------------------------------------------------------------ consumer: the one GEMV loop */
template <int R> struct V2Log2;
template <> struct V2Log2<1>  { static const int v = 0; };
template <> struct V2Log2<2>  { static const int v = 1; };
template <> struct V2Log2<4>  { static const int v = 2; };
template <> struct V2Log2<8>  { static const int v = 3; };
template <> struct V2Log2<16> { static const int v = 4; };
template <> struct V2Log2<32> { static const int v = 5; };

/* lane L ends up with the total (over the warp) of v[L >> (5 - log2 R)] */
template <int R>
__device__ __forceinline__ float v2_transpose_reduce(float (&v)[R], int lane) {
    int off = 16;
#pragma unroll
    for (int n = R; n > 1; n >>= 1, off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; i++) {
            float send = upper ? v[i] : v[i + n / 2];
            float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    float r = v[0];
#pragma unroll
    for (int o = (16 >> V2Log2<R>::v); o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;
}

/* activation registers: NB columns x 8 k-values.  NB >= 2 keeps column pairs packed for FFMA2. */
__device__ __forceinline__ unsigned long long v2_pack(float lo, float hi) {
    unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void v2_unpack(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
template <int NB> struct V2X {
    static constexpr int NP = NB / 2;
    unsigned long long v[NP][8];               /* {col 2j, col 2j+1} at k */
    /* f(b, k) -> value of column b at this thread's k-th element */
    template <typename F> __device__ __forceinline__ void fill(F f) {
#pragma unroll
        for (int j = 0; j < NP; j++)
#pragma unroll
            for (int k = 0; k < 8; k++) v[j][k] = v2_pack(f(2 * j, k), f(2 * j + 1, k));
    }
    __device__ __forceinline__ float get(int b, int k) const {
        float lo, hi; v2_unpack(v[b >> 1][k], lo, hi);
        return (b & 1) ? hi : lo;
    }
};
template <> struct V2X<1> {
    float v[8];
    template <typename F> __device__ __forceinline__ void fill(F f) {
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = f(0, k);
    }
    __device__ __forceinline__ float get(int, int k) const { return v[k]; }
};

/* acc[b] += w[k] * x[b][k] for the 8 weights of one 16-byte load, k ascending: the reference's order
 * (voxtral_kernels.c:154-195) in every column, whatever NB is. */
template <int NB>
__device__ __forceinline__ void v2_dot8(const uint4 w, const V2X<NB> &x, float (&acc)[NB]) {
    const float wf[8] = { vb_bf16_lo(w.x), vb_bf16_hi(w.x), vb_bf16_lo(w.y), vb_bf16_hi(w.y),
                          vb_bf16_lo(w.z), vb_bf16_hi(w.z), vb_bf16_lo(w.w), vb_bf16_hi(w.w) };
    if constexpr (NB == 1) {
#pragma unroll
        for (int k = 0; k < 8; k++) acc[0] = fmaf(wf[k], x.v[k], acc[0]);
    } else {
        unsigned long long a2[NB / 2];
#pragma unroll
        for (int j = 0; j < NB / 2; j++) asm("mov.b64 %0, {%1,%2};" : "=l"(a2[j]) : "f"(acc[2 * j]), "f"(acc[2 * j + 1]));
#pragma unroll
        for (int k = 0; k < 8; k++) {
            unsigned long long ww;
            asm("mov.b64 %0, {%1,%1};" : "=l"(ww) : "f"(wf[k]));
#pragma unroll
            for (int j = 0; j < NB / 2; j++) asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a2[j]) : "l"(ww), "l"(x.v[j][k]));
        }
#pragma unroll
        for (int j = 0; j < NB / 2; j++) asm("mov.b64 {%0,%1}, %2;" : "=f"(acc[2 * j]), "=f"(acc[2 * j + 1]) : "l"(a2[j]));
    }
}

/* Consume the chunks of one (sub)phase from the ring until the producer's end marker.
 * NT = threads that own a 16-byte column of the row segment (384 for 3072-wide, 256 for 2048-wide segments).
 * Partial sums of 16 (row, column) pairs are reduced together -- 4 chunks x 4 rows for one column, 2 chunks for two columns,
 * one chunk for four, half a chunk for eight -- so the cost of a reduction (15 shuffles, one shared-memory pass, one named
 * barrier) is the same per 16 outputs whatever NB is, and a slot is released as soon as its rows are in the accumulators.
 * epi(row, b, value, lane, valid) runs in ONE warp per reduction, the warps taking turns (grp % 12): the epilogue is a latency
 * chain (12 shared-memory reads, the sum, sincosf / expf, global stores) that would otherwise put the same warp on the critical
 * path of every reduction (measured: 284 of 876 cycles per chunk).  lane = (chunk * 4 + row_in_chunk) * NB + b. */
template <int NB, typename Epi>
__device__ __forceinline__ void v2_consume(V2Smem *sm, const uint8_t *slots, uint32_t &it, int seg_bytes, int NT,
                                           const V2X<NB> &x, int &redbuf, uint32_t &grp, int *err, long long (&tacc)[5], bool timing, int dbg, Epi epi) {
    constexpr int CPR = NB == 1 ? 4 : NB == 2 ? 2 : 1;     /* chunks per reduction */
    constexpr int HALVES = NB == 8 ? 2 : 1;                /* NB = 8: a chunk's 32 values are reduced as two halves (registers) */
    constexpr int RH = V2_RC / HALVES;                     /* rows per half */
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool active = t < NT;
    bool end = false, ready = false;
    while (!end) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = 0.f;
        int row0s[CPR], nrs[CPR], got = 0;
#pragma unroll
        for (int c = 0; c < CPR; c++) {
            row0s[c] = 0; nrs[c] = 0;
            if (end) continue;
            const int s = (int)(it % V2_SLOTS);
            const uint32_t par = (it / V2_SLOTS) & 1u;
            long long tq0 = 0;
            if (timing) tq0 = clock64();
            if (!ready) {                                   /* not already seen complete by the probe of the previous chunk */
                long long t0 = 0;
                while (!mbar_try_wait(&sm->full[s], par)) spin_guard(t0, err, 3);
            }
            long long tq1 = 0;
            if (timing) { tq1 = clock64(); tacc[0] += tq1 - tq0; }
            const int nrows = sm->meta_nrows[s];
            row0s[c] = sm->meta_row0[s]; nrs[c] = nrows;
            if (nrows == 0) { end = true; ready = false; }
            else {
                got++;
                const uint8_t *base = slots + (size_t)s * V2_SLOT_BYTES + (size_t)t * 16;
                /* probe the NEXT slot now (non-blocking): its ~100-cycle mbarrier round trip overlaps this chunk's math */
                ready = mbar_test_wait(&sm->full[(it + 1) % V2_SLOTS], ((it + 1) / V2_SLOTS) & 1u);
#pragma unroll
                for (int h = 0; h < HALVES; h++) {
                    if (active && !(dbg & 1)) {
                        uint4 w[RH];
#pragma unroll
                        for (int r = 0; r < RH; r++)
                            w[r] = (h * RH + r < nrows) ? *reinterpret_cast<const uint4 *>(base + (size_t)(h * RH + r) * seg_bytes) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                        for (int r = 0; r < RH; r++) {
                            float a[NB];
#pragma unroll
                            for (int b = 0; b < NB; b++) a[b] = 0.f;
                            v2_dot8<NB>(w[r], x, a);
#pragma unroll
                            for (int b = 0; b < NB; b++) acc[(HALVES == 2 ? 0 : c * V2_RC * NB) + r * NB + b] = a[b];
                        }
                    }
                    if (HALVES == 2 && !(dbg & 2)) {        /* NB = 8: reduce this half now, lanes 16h..16h+15 of the result row */
                        const float tot = v2_transpose_reduce<16>(acc, lane);
                        if (!(lane & 1)) sm->red[redbuf][warp][h * 16 + (lane >> 1)] = tot;
#pragma unroll
                        for (int i = 0; i < 16; i++) acc[i] = 0.f;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm->empty[s]);     /* the rows are in registers: the slot can be refilled */
            it++;
            if (timing) tacc[1] += clock64() - tq1;
        }
        if (got == 0) break;
        if (dbg & 2) continue;
        long long tr0 = 0;
        if (timing) tr0 = clock64();
        if (HALVES == 1) {
            const float tot = v2_transpose_reduce<16>(acc, lane);
            if (!(lane & 1)) sm->red[redbuf][warp][lane >> 1] = tot;
        }
        long long tr1 = 0;
        if (timing) { tr1 = clock64(); tacc[2] += tr1 - tr0; }
        v2_bar();
        long long tr2 = 0;
        if (timing) { tr2 = clock64(); tacc[3] += tr2 - tr1; }
        if (warp == (int)(grp % V2_CW)) {
            constexpr int NV = 16 * HALVES;                 /* values in this reduction: lane < NV */
            float sum = 0.f;
            if (lane < NV) {
#pragma unroll
                for (int wv = 0; wv < V2_CW; wv++) sum += sm->red[redbuf][wv][lane];
            }
            const int c = lane / (V2_RC * NB), r = (lane / NB) % V2_RC;
            int row0 = row0s[0], nr = nrs[0];
#pragma unroll
            for (int k = 1; k < CPR; k++) if (c == k) { row0 = row0s[k]; nr = nrs[k]; }
            epi(row0 + r, lane % NB, sum, lane, lane < NV && r < nr);
            if (timing) tacc[4] += clock64() - tr2;
        }
        grp++;
        redbuf ^= 1;
    }
}

/* ------------------------------------------------------------------ activation loads */
template <int NB>
__device__ __forceinline__ void v2_load_x(V2X<NB> &x, const float *src, int col_stride, int NT) {
    const int t = threadIdx.x;
    float4 lo[NB], hi[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        lo[b] = make_float4(0.f, 0.f, 0.f, 0.f); hi[b] = lo[b];
        if (t < NT) {
            const float4 *p = reinterpret_cast<const float4 *>(src + (size_t)b * col_stride + (size_t)t * 8);
            lo[b] = __ldcg(p); hi[b] = __ldcg(p + 1);
        }
    }
    x.fill([&](int b, int k) {
        const float4 q = k < 4 ? lo[b] : hi[b];
        const int kk = k & 3;
        return kk == 0 ? q.x : kk == 1 ? q.y : kk == 2 ? q.z : q.w;
    });
}

/* RMSNorm of each column over the 3072 values spread across the 384 threads (voxtral_kernels.c:346-363), optional (1+ada) */
template <int NB>
__device__ __forceinline__ void v2_rmsnorm(V2X<NB> &x, const float *__restrict__ w, const float *__restrict__ ada, V2Smem *sm) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    float ss[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) { float v = x.get(b, j); s = fmaf(v, v, s); }
        ss[b] = vb_warp_sum(s);
    }
    v2_bar();
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NB; b++) sm->sred[warp][b] = ss[b];
    }
    v2_bar();
    float wk[8], ak[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { wk[j] = w[t * 8 + j]; ak[j] = ada ? ada[t * 8 + j] : 0.f; }
    float rinv[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < V2_CW; i++) tot += sm->sred[i][b];
        rinv[b] = 1.0f / sqrtf(tot / (float)VOX_DEC_DIM + VOX_DEC_NORM_EPS);
    }
    V2X<NB> y;
    y.fill([&](int b, int j) {
        float v = x.get(b, j) * rinv[b] * wk[j];
        if (ada) v *= (1.0f + ak[j]);
        return v;
    });
    x = y;
}

/* ------------------------------------------------------------------ attention */
/* Work unit = (column, kv head): the CTAs are dealt round-robin to the 8 * n_active pairs and those of a pair split the valid
 * ring slots of that column between them (one column: 18-19 CTAs per kv head; eight columns: 2-3 CTAs per pair), so a CTA
 * handles exactly one pair per layer whatever NB is.  Inside the CTA the 12 warps take interleaved slots (the 4 query heads of
 * the kv head share each K/V row read), merge through shared memory to one partial per query head, publish it; the last CTA
 * to arrive for the pair (atomic ticket) combines the pair's partials into attn_out.  voxtral_kernels.c:412-482. */
template <int NB>
__device__ __forceinline__ void v2_attention(const V2Args &a, V2Smem *sm, float *att_scr, int layer) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int nact = 0, my_b = -1;
    {
        int idx[NB];
#pragma unroll
        for (int b = 0; b < NB; b++) if (!sm->c_done[b]) idx[nact++] = b;
        if (nact == 0) return;
        const int pair_of_cta = (int)(blockIdx.x % (unsigned)(8 * nact));
#pragma unroll
        for (int k = 0; k < NB; k++) if (k < nact && k == pair_of_cta / 8) my_b = idx[k];
    }
    const int npairs = 8 * nact;
    const int pair = (int)(blockIdx.x % (unsigned)npairs), si = (int)(blockIdx.x / (unsigned)npairs);
    const int nsplit = (int)(gridDim.x / (unsigned)npairs) + ((int)(gridDim.x % (unsigned)npairs) > pair ? 1 : 0);
    const int kvh = pair & 7, b = my_b;
    const float scale = 1.0f / sqrtf((float)HD);
    {
        const int pos = sm->c_pos[b];
        const int n_valid = min(pos + 1, VB_KV_SLOTS);
        const int s0 = (int)((long long)n_valid * si / nsplit), s1 = (int)((long long)n_valid * (si + 1) / nsplit);
        float4 qv[4];
#pragma unroll
        for (int hq = 0; hq < 4; hq++) qv[hq] = __ldcg(reinterpret_cast<const float4 *>(a.q + (size_t)b * VB_DEC_Q + (kvh * 4 + hq) * HD + lane * 4));
        /* Online softmax in blocks of 4 slots: the 16 (slot, head) partial dots of a block are reduced over the lanes together
         * (one transposed reduction: lane L ends up with the score of slot (L >> 3), head ((L >> 1) & 3)), every lane keeps the
         * running max / sum of ITS head only, and the 16 weights are broadcast back for the 4 x 4-dim accumulators each lane
         * owns: 40 shuffles and 2 expf per block instead of 80 shuffles and 32 expf with a per-slot update. */
        float m_my = -1e30f, l_my = 0.f; float4 o[4];
#pragma unroll
        for (int hq = 0; hq < 4; hq++) o[hq] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int my_u = lane >> 3;
        const float *kb = sm->c_kv_k[b] + (size_t)layer * VB_KV_SLOTS * VB_DEC_KV + kvh * HD + lane * 4;
        const float *vb = sm->c_kv_v[b] + (size_t)layer * VB_KV_SLOTS * VB_DEC_KV + kvh * HD + lane * 4;
        for (int sb = s0 + warp; sb < s1; sb += 4 * V2_CW) {       /* this warp's slots: sb, sb+12, sb+24, sb+36 */
            float4 k4[4], v4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int s = min(sb + V2_CW * u, s1 - 1);
                k4[u] = __ldcg(reinterpret_cast<const float4 *>(kb + (size_t)s * VB_DEC_KV));
                v4[u] = __ldcg(reinterpret_cast<const float4 *>(vb + (size_t)s * VB_DEC_KV));
            }
            float sc[16];
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int hq = 0; hq < 4; hq++)
                    sc[u * 4 + hq] = qv[hq].x * k4[u].x + qv[hq].y * k4[u].y + qv[hq].z * k4[u].z + qv[hq].w * k4[u].w;
            const float tot = v2_transpose_reduce<16>(sc, lane);
            const bool valid = sb + V2_CW * my_u < s1;
            const float sv = valid ? tot * scale : -1e30f;
            float bm = fmaxf(sv, __shfl_xor_sync(0xffffffffu, sv, 8));
            bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
            const float mn = fmaxf(m_my, bm);
            const float c = expf(m_my - mn);
            const float pw = valid ? expf(sv - mn) : 0.f;
            float ps = pw + __shfl_xor_sync(0xffffffffu, pw, 8);
            ps += __shfl_xor_sync(0xffffffffu, ps, 16);
            l_my = l_my * c + ps;
            m_my = mn;
#pragma unroll
            for (int hq = 0; hq < 4; hq++) {
                const float ch = __shfl_sync(0xffffffffu, c, hq * 2);
                o[hq].x *= ch; o[hq].y *= ch; o[hq].z *= ch; o[hq].w *= ch;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float pu = __shfl_sync(0xffffffffu, pw, (u * 4 + hq) * 2);
                    o[hq].x = fmaf(pu, v4[u].x, o[hq].x); o[hq].y = fmaf(pu, v4[u].y, o[hq].y);
                    o[hq].z = fmaf(pu, v4[u].z, o[hq].z); o[hq].w = fmaf(pu, v4[u].w, o[hq].w);
                }
            }
        }
        /* merge the 12 warps: scratch[warp][hq] = {m, l, -, -, o[128]}; lane 2 hq holds head hq's m and l */
#pragma unroll
        for (int hq = 0; hq < 4; hq++) {
            float *dst = att_scr + (size_t)(warp * 4 + hq) * 132;
            if (lane == hq * 2) { dst[0] = m_my; dst[1] = l_my; }
            *reinterpret_cast<float4 *>(dst + 4 + lane * 4) = o[hq];
        }
        v2_bar();
        if (warp < 4) {
            const int hq = warp;
            float M = -1e30f;
#pragma unroll
            for (int w = 0; w < V2_CW; w++) M = fmaxf(M, att_scr[(size_t)(w * 4 + hq) * 132]);
            float L = 0.f; float4 O = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < V2_CW; w++) {
                const float *src = att_scr + (size_t)(w * 4 + hq) * 132;
                float lw = src[1];
                if (lw > 0.f) {
                    float c = expf(src[0] - M);
                    float4 ow = *reinterpret_cast<const float4 *>(src + 4 + lane * 4);
                    L = fmaf(c, lw, L);
                    O.x = fmaf(c, ow.x, O.x); O.y = fmaf(c, ow.y, O.y); O.z = fmaf(c, ow.z, O.z); O.w = fmaf(c, ow.w, O.w);
                }
            }
            const size_t pi = ((size_t)b * V2_NSPLIT_MAX + si) * VOX_DEC_HEADS + (kvh * 4 + hq);
            if (lane == 0) { a.part_m[pi] = M; a.part_l[pi] = L; }
            *reinterpret_cast<float4 *>(a.part_o + pi * HD + lane * 4) = O;
        }
        v2_bar();
    }
    /* ticket: the last CTA of this (column, kv head) combines */
    unsigned int *ticket = a.bar + 64 + b * 8 + kvh;
    if (tid == 0) {
        __threadfence();
        unsigned int old = atomicAdd(ticket, 1u);
        int last = (old == (unsigned int)(nsplit - 1));
        if (last) { *ticket = 0u; __threadfence(); }
        sm->is_last = last;
    }
    v2_bar();
    if (sm->is_last && warp < 4) {
        const int h = kvh * 4 + warp;
        float mi = -1e30f, li = 0.f;
        if (lane < nsplit) {
            mi = __ldcg(a.part_m + ((size_t)b * V2_NSPLIT_MAX + lane) * VOX_DEC_HEADS + h);
            li = __ldcg(a.part_l + ((size_t)b * V2_NSPLIT_MAX + lane) * VOX_DEC_HEADS + h);
        }
        float M = mi;
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o2));
        float wi = li > 0.f ? expf(mi - M) : 0.f;
        float L = vb_warp_sum(wi * li);
        float4 O = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 ov[V2_NSPLIT_MAX];
#pragma unroll
        for (int i = 0; i < V2_NSPLIT_MAX; i++)
            ov[i] = i < nsplit ? __ldcg(reinterpret_cast<const float4 *>(a.part_o + (((size_t)b * V2_NSPLIT_MAX + i) * VOX_DEC_HEADS + h) * HD + lane * 4))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < V2_NSPLIT_MAX; i++) {
            float c = __shfl_sync(0xffffffffu, wi, i);
            O.x = fmaf(c, ov[i].x, O.x); O.y = fmaf(c, ov[i].y, O.y); O.z = fmaf(c, ov[i].z, O.z); O.w = fmaf(c, ov[i].w, O.w);
        }
        float inv = L > 0.f ? 1.0f / L : 0.f;
        *reinterpret_cast<float4 *>(a.attn_out + (size_t)b * VB_DEC_Q + h * HD + lane * 4) = make_float4(O.x * inv, O.y * inv, O.z * inv, O.w * inv);
    }
}

/* ------------------------------------------------------------------ verify mode: exact multi-token decoding */
/* The greedy loop of voxtral.c:1056-1093 emits one token per weight pass.  Here the columns of a pass are CONSECUTIVE positions
 * of one stream: column 0 gets the last emitted token (always valid), column j a DRAFT of the j-th next token.  Every column's
 * argmax o_j is exact given its inputs, so o_0 is always the next token, and o_j is the (j+1)-th next token iff drafts 1..j were
 * right (draft_j == o_{j-1}).  The longest such prefix is accepted: identical ids, between 1 and nb tokens per weight pass.
 * K/V rows written by rejected columns lie at positions the next pass rewrites before anything reads them (no ring wrap: the
 * host only uses this mode while pos + nb <= 8192).  Drafter: a 512-entry table of "token -> the token that followed it last
 * time" kept identically in every CTA's shared memory, falling back to repeating the last token (streaming ASR output is
 * dominated by repeated [STREAMING_PAD] and by re-occurring word pieces). */
__device__ __forceinline__ int v2_draft_next(const V2Smem *sm, int tok) {
    const int h = tok & (V2_DRAFT_TAB - 1);
    return sm->dr_key[h] == tok ? sm->dr_succ[h] : tok;
}
__device__ __forceinline__ void v2_verify_columns(V2Smem *sm, int nb) {
    int tok = sm->v_token;
    for (int j = 0; j < V2_MAXB; j++) {
        const bool on = j < nb && j < sm->v_left && !sm->v_eos;
        sm->c_pos[j] = sm->v_pos + j; sm->c_arow[j] = on ? sm->v_arow + j : 0; sm->c_token[j] = tok;
        sm->c_done[j] = on ? 0 : 1;
        tok = v2_draft_next(sm, tok);
    }
}

/* ------------------------------------------------------------------ the kernel */
extern __shared__ __align__(1024) uint8_t v2_smem_raw[];

#define V2PROF() do { if (a.prof && step == a.prof_step && tid == 0 && prof_n < V2_PROF_SLOTS) \
    a.prof[(size_t)blockIdx.x * V2_PROF_SLOTS + prof_n++] = clock64(); } while (0)

template <int NB>
__global__ void __launch_bounds__(V2_THREADS, 1) k_dec_v2(const __grid_constant__ V2Args a) {
    uint8_t *slots = v2_smem_raw;
    float *att_scr = reinterpret_cast<float *>(v2_smem_raw + (size_t)V2_SLOTS * V2_SLOT_BYTES);
    V2Smem *sm = reinterpret_cast<V2Smem *>(v2_smem_raw + (size_t)V2_SLOTS * V2_SLOT_BYTES + (size_t)V2_ATT_FLOATS * 4);
    const DecParams &p = a.p;
    const int tid = threadIdx.x, lane = tid & 31;
    int *err = a.err;

    if (tid == 0) {
        for (int i = 0; i < V2_SLOTS; i++) { mbar_init(&sm->full[i], 1); mbar_init(&sm->empty[i], V2_CW); }
        sm->abort_flag = 0; sm->is_last = 0;
        for (int i = 0; i < 16; i++) sm->prof_acc[i] = 0;
        for (int i = 0; i < 8; i++) sm->prof_cnt[i] = 0;
        for (int b = 0; b < V2_MAXB; b++) {
            const bool on = b < a.nb && a.col[b].n_steps > 0;
            sm->c_pos[b] = a.col[b].pos0; sm->c_token[b] = a.col[b].token0; sm->c_arow[b] = a.col[b].arow0;
            sm->c_done[b] = on ? 0 : 1; sm->c_nout[b] = 0; sm->c_left[b] = on ? a.col[b].n_steps : 0;
            sm->c_adapter[b] = a.col[b].adapter; sm->c_kv_k[b] = a.col[b].kv_k; sm->c_kv_v[b] = a.col[b].kv_v;
            sm->c_logits[b] = a.col[b].logits; sm->c_tokens[b] = a.col[b].tokens;
        }
        if (a.verify) {
            /* one stream: column j processes position pos+j with input token (j == 0: the last emitted token, else draft j);
             * all columns share the stream's KV ring / adapter rows; columns beyond the remaining adapter rows are off */
            sm->v_pos = a.col[0].pos0; sm->v_token = a.col[0].token0; sm->v_arow = a.col[0].arow0; sm->v_left = a.col[0].n_steps;
            sm->v_nout = 0; sm->v_eos = 0; sm->v_passes = 0;
            for (int i = 0; i < V2_DRAFT_TAB; i++) { sm->dr_key[i] = -1; sm->dr_succ[i] = -1; }
            v2_verify_columns(sm, a.nb);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= V2_CONS) {
        if (tid == V2_CONS) v2_producer(sm, slots, a);
        return;
    }

    /* ===================== consumers ===================== */
    uint32_t it = 0;
    unsigned int gen = 0;
    int redbuf = 0, prof_n = 0;
    uint32_t grp = 0;                                          /* reductions so far: the epilogue warp of a reduction is grp % 12 */
    long long tacc[5] = { 0, 0, 0, 0, 0 };                    /* profiled launches: cycles of this thread in wait-for-data / math / reduce / CTA barrier / epilogue */
    const bool timing = a.prof != nullptr && (lane == 0);
    int my_r0, my_r1;                                          /* residual-stream rows this CTA owns (static wo / w2 partition) */
    v2_static_rows(VOX_DEC_DIM, 1, my_r0, my_r1);

    for (int step = 0; step < a.n_steps; step++) {
        {   /* all columns finished? (identical decision in every CTA) */
            int live = 0;
#pragma unroll
            for (int b = 0; b < NB; b++) live |= !sm->c_done[b];
            if (!live) break;
        }
        /* residual stream rows owned by this CTA: x = adapter[arow] + tok_emb[token] (voxtral.c:1057-1061) */
        for (int i = tid; i < (my_r1 - my_r0) * NB; i += V2_CONS) {
            const int b = i % NB, r = my_r0 + i / NB;
            const float *ar = sm->c_adapter[b] + (size_t)(sm->c_done[b] ? 0 : sm->c_arow[b]) * VOX_DEC_DIM;   /* finished column: any valid row */
            const uint16_t *er = p.tok_emb + (size_t)sm->c_token[b] * VOX_DEC_DIM;
            a.x[(size_t)b * VOX_DEC_DIM + r] = ar[r] + __uint_as_float((uint32_t)er[r] << 16);
        }

        /* One GEMV loop for all 26 x 7 + 1 weight (sub)phases of the step (a single copy of the hot code in the I-cache):
         * sub 0 QKV | 1,2 wo column blocks | 3 w1|w3 | 4,5,6 w2 column blocks | 7 logits (after the last layer). */
        unsigned long long best = 0ull;                                /* logits: the epilogue lane's column is lane % NB throughout */
#pragma unroll 1
        for (int idx = 0; idx <= VOX_DEC_LAYERS * 7; idx++) {
            const int layer = idx / 7;
            const int sub = layer == VOX_DEC_LAYERS ? 7 : idx - layer * 7;
            if (sub == 0 || sub == 7) V2PROF();
            const bool pstep = a.prof && step == a.prof_step && tid == 0 && layer >= 1 && layer < VOX_DEC_LAYERS - 1;
            if (pstep) sm->prof_t = clock64();
            /* ---- this phase's activation columns: 8 values x NB per thread ---- */
            V2X<NB> x;
            int seg_bytes = VOX_DEC_DIM * 2, NT = V2_CONS;
            if (sub == 0 && layer == 0) {
                x.fill([&](int b, int j) {
                    const float *ar = sm->c_adapter[b] + (size_t)(sm->c_done[b] ? 0 : sm->c_arow[b]) * VOX_DEC_DIM + tid * 8;
                    const uint16_t *er = p.tok_emb + (size_t)sm->c_token[b] * VOX_DEC_DIM + tid * 8;
                    return ar[j] + __uint_as_float((uint32_t)er[j] << 16);
                });
            } else {
                const float *src = a.x; int stride = VOX_DEC_DIM;
                if (sub == 1 || sub == 2) { src = a.attn_out + (sub - 1) * 2048; stride = VB_DEC_Q; seg_bytes = 4096; NT = 256; }
                else if (sub >= 4 && sub <= 6) { src = a.gate + (sub - 4) * VOX_DEC_DIM; stride = VOX_DEC_HIDDEN; }
                v2_load_x<NB>(x, src, stride, NT);
            }
            if (sub == 0 || sub == 3 || sub == 7) {
                const float *nw = sub == 0 ? p.attn_norm[layer] : sub == 3 ? p.ffn_norm[layer] : p.final_norm;
                v2_rmsnorm<NB>(x, nw, sub == 3 ? p.ada + (size_t)layer * VOX_DEC_DIM : nullptr, sm);
            }
            const float *inv_freq = p.inv_freq;
            if (pstep) { const long long tn = clock64(); sm->prof_acc[sub] += tn - sm->prof_t; sm->prof_t = tn; sm->prof_cnt[sub] -= (int)it + 1; }
            v2_consume<NB>(sm, slots, it, seg_bytes, NT, x, redbuf, grp, err, tacc, timing, a.dbg, [&](int row, int b, float v, int, bool valid) {
                const float other = __shfl_xor_sync(0xffffffffu, v, NB);    /* row ^ 1 of the same column: RoPE pair / (gate, up) pair */
                if (!valid) return;
                switch (sub) {
                case 0: {   /* RoPE -> q, KV ring (voxtral_decoder.c:626-660) */
                    if (sm->c_done[b]) return;
                    const int pos = sm->c_pos[b], slot = pos & (VB_KV_SLOTS - 1);
                    if (row < VB_DEC_Q + VB_DEC_KV) {
                        const int d = (row & (HD - 1)) >> 1;
                        float sn, cs;
                        sincosf((float)pos * inv_freq[d], &sn, &cs);
                        const float y = (row & 1) ? (other * sn + v * cs) : (v * cs - other * sn);
                        if (row < VB_DEC_Q) a.q[(size_t)b * VB_DEC_Q + row] = y;
                        else sm->c_kv_k[b][((size_t)layer * VB_KV_SLOTS + slot) * VB_DEC_KV + row - VB_DEC_Q] = y;
                    } else sm->c_kv_v[b][((size_t)layer * VB_KV_SLOTS + slot) * VB_DEC_KV + row - VB_DEC_Q - VB_DEC_KV] = v;
                    break; }
                case 1: case 4: sm->part[row - my_r0][b] = v; break;
                case 5: sm->part[row - my_r0][b] += v; break;
                case 2: case 6: {   /* last column block: residual add */
                    float *xp = a.x + (size_t)b * VOX_DEC_DIM + row;
                    *xp = __ldcg(xp) + (sm->part[row - my_r0][b] + v);
                    break; }
                case 3: if (!(row & 1)) a.gate[(size_t)b * VOX_DEC_HIDDEN + (row >> 1)] = vb_silu(v) * other; break;   /* voxtral_decoder.c:682-686 */
                default: {
                    if (b < a.nb) sm->c_logits[b][row] = v;            /* (padding columns alias column 0's buffers) */
                    const unsigned long long c = pack_cand(v, row);
                    if (c > best) best = c;
                    break; }
                }
            });
            if (pstep) { sm->prof_acc[8 + sub] += clock64() - sm->prof_t; sm->prof_cnt[sub] += (int)it; }
            if (sub == 0) {
                V2PROF();
                v2_grid_barrier(a.bar, gen, err);
                V2PROF();
                v2_attention<NB>(a, sm, att_scr, layer);
                V2PROF();
                v2_grid_barrier(a.bar, gen, err);
                V2PROF();
            } else if (sub == 2 || sub == 3 || sub == 6) {
                V2PROF();
                v2_grid_barrier(a.bar, gen, err);
                if (sub != 6) V2PROF();
            }
        }
        {   /* per-CTA argmax per column: every warp ran some of the logits epilogues (its lanes' column is lane % NB) */
#pragma unroll
            for (int o = 16; o >= NB; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                if (other > best) best = other;
            }
            v2_bar();
            if (lane < NB) sm->cand[tid >> 5][lane] = best;
            v2_bar();
            if (tid < NB) {
                unsigned long long m = 0ull;
#pragma unroll
                for (int w = 0; w < V2_CW; w++) if (sm->cand[w][tid] > m) m = sm->cand[w][tid];
                a.argmax[(size_t)blockIdx.x * V2_MAXB + tid] = m;
            }
        }
        V2PROF();
        v2_grid_barrier(a.bar, gen, err);
        V2PROF();
        {   /* global argmax per column: every CTA reduces the per-CTA candidates, so every CTA knows the tokens */
            unsigned long long best[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) {
                best[b] = tid < (int)gridDim.x ? __ldcg(a.argmax + (size_t)tid * V2_MAXB + b) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    unsigned long long other = __shfl_xor_sync(0xffffffffu, best[b], o);
                    if (other > best[b]) best[b] = other;
                }
            }
            v2_bar();
            if (lane == 0) {
#pragma unroll
                for (int b = 0; b < NB; b++) sm->cand[tid >> 5][b] = best[b];
            }
            v2_bar();
            if (tid == 0 && !a.verify) {
                for (int b = 0; b < NB; b++) {
                    if (sm->c_done[b]) continue;
                    unsigned long long m = 0ull;
                    for (int w = 0; w < V2_CW; w++) if (sm->cand[w][b] > m) m = sm->cand[w][b];
                    const int tok = cand_index(m);
                    if (blockIdx.x == 0) sm->c_tokens[b][sm->c_nout[b]] = tok;
                    sm->c_nout[b]++; sm->c_token[b] = tok; sm->c_pos[b]++; sm->c_arow[b]++; sm->c_left[b]--;
                    if (tok == VB_TOKEN_EOS) sm->c_done[b] = 2;
                    else if (sm->c_left[b] <= 0) sm->c_done[b] = 1;
                }
            }
            if (tid == 0 && a.verify) {
                /* accept the longest prefix whose drafts were right; every CTA takes the same decision */
                int prev = sm->v_token;
                sm->v_passes++;
                for (int b = 0; b < NB; b++) {
                    if (sm->c_done[b]) break;
                    if (b > 0 && sm->c_token[b] != prev) break;            /* draft b was wrong: o_b was computed from a wrong input */
                    unsigned long long m = 0ull;
                    for (int w = 0; w < V2_CW; w++) if (sm->cand[w][b] > m) m = sm->cand[w][b];
                    const int tok = cand_index(m);
                    if (blockIdx.x == 0) sm->c_tokens[0][sm->v_nout] = tok;
                    const int h = prev & (V2_DRAFT_TAB - 1);
                    sm->dr_key[h] = prev; sm->dr_succ[h] = tok;             /* learn: prev was followed by tok */
                    sm->v_nout++; sm->v_pos++; sm->v_arow++; sm->v_left--;
                    prev = tok;
                    if (tok == VB_TOKEN_EOS) { sm->v_eos = 1; break; }
                }
                sm->v_token = prev;
                v2_verify_columns(sm, a.nb);
            }
            v2_bar();
        }
        V2PROF();
    }
    if (a.prof && lane == 0 && (tid == 0 || tid == V2_CONS - 32)) {            /* warp 0 and the epilogue warp */
        long long *pp = a.prof + (size_t)blockIdx.x * V2_PROF_SLOTS + V2_PROF_SLOTS - (tid == 0 ? 20 : 14);
        for (int i = 0; i < 5; i++) pp[i] = tacc[i];
        if (tid == 0) for (int i = 0; i < 16; i++) a.prof[(size_t)blockIdx.x * V2_PROF_SLOTS + 270 + i] = sm->prof_acc[i];
        if (tid == 0) for (int i = 0; i < 8; i++) a.prof[(size_t)blockIdx.x * V2_PROF_SLOTS + 286 + i] = sm->prof_cnt[i];
    }
    if (tid == 0) {
        sm->abort_flag = 1;                                            /* the producer may be ahead of an early exit (EOS) */
        if (blockIdx.x == 0) {
            for (int b = 0; b < a.nb; b++) {
                VbDecState st;
                st.pos = sm->c_pos[b]; st.token = sm->c_token[b]; st.eos = sm->c_done[b] == 2; st.n_out = sm->c_nout[b];
                st.adapter_row = sm->c_arow[b]; st.pad[0] = st.pad[1] = st.pad[2] = 0;
                if (a.verify) {
                    st.pos = sm->v_pos; st.token = sm->v_token; st.eos = sm->v_eos; st.n_out = b == 0 ? sm->v_nout : 0;
                    st.adapter_row = sm->v_arow; st.pad[0] = sm->v_passes;       /* weight passes spent on n_out tokens */
                }
                a.st_out[b] = st;
            }
        }
    }
}

/* ------------------------------------------------------------------ host side */
static size_t v2_smem_bytes() { return (size_t)V2_SLOTS * V2_SLOT_BYTES + (size_t)V2_ATT_FLOATS * 4 + sizeof(V2Smem); }

template <int NB> static cudaError_t v2_prepare() {
    return cudaFuncSetAttribute((const void *)k_dec_v2<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v2_smem_bytes());
}

extern "C" int vb_decoder_v2_supported(VbEngine *e) {
    if (e->v2_checked) return e->v2_ok;
    e->v2_checked = 1; e->v2_ok = 0;
    int coop = 0, blocks = 0, smem_optin = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device);
    const char *why = NULL; cudaError_t ce = cudaSuccess;
    if (!coop) why = "no cooperative launch";
    else if (e->sm_count > 8 * V2_NSPLIT_MAX) why = "more SMs than the attention split supports";
    else if ((size_t)smem_optin < v2_smem_bytes()) why = "not enough shared memory per block";
    else if ((ce = v2_prepare<1>()) != cudaSuccess || (ce = v2_prepare<2>()) != cudaSuccess || (ce = v2_prepare<4>()) != cudaSuccess ||
             (ce = v2_prepare<8>()) != cudaSuccess) why = "cudaFuncSetAttribute failed";
    else if ((ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_dec_v2<8>, V2_THREADS, v2_smem_bytes())) != cudaSuccess || blocks < 1)
        why = "kernel does not fit one CTA per SM";
    if (why) {
        if (vox_verbose >= 1 || getenv("VOX_CUDA_DECODE"))
            fprintf(stderr, "voxtral_b200: v2 decode kernel unavailable: %s (%s; smem need %zu of %d, blocks/SM %d)\n", why,
                    cudaGetErrorString(ce), v2_smem_bytes(), smem_optin, blocks);
        cudaGetLastError();
        return 0;
    }
    e->v2_ok = 1;
    return 1;
}

static int v2_alloc(VbEngine *e) {
    if (e->v2.x) return 0;
    const size_t wb = e->weight_bytes;
    VbV2Scratch *s = &e->v2;
    s->x = (float *)vb_dev_alloc_owned(e, (size_t)V2_MAXB * VOX_DEC_DIM * 4);
    s->q = (float *)vb_dev_alloc_owned(e, (size_t)V2_MAXB * VB_DEC_Q * 4);
    s->attn_out = (float *)vb_dev_alloc_owned(e, (size_t)V2_MAXB * VB_DEC_Q * 4);
    s->gate = (float *)vb_dev_alloc_owned(e, (size_t)V2_MAXB * VOX_DEC_HIDDEN * 4);
    s->part_m = (float *)vb_dev_alloc_owned(e, (size_t)V2_MAXB * V2_NSPLIT_MAX * VOX_DEC_HEADS * 4);
    s->part_l = (float *)vb_dev_alloc_owned(e, (size_t)V2_MAXB * V2_NSPLIT_MAX * VOX_DEC_HEADS * 4);
    s->part_o = (float *)vb_dev_alloc_owned(e, (size_t)V2_MAXB * V2_NSPLIT_MAX * VOX_DEC_HEADS * HD * 4);
    s->argmax = (unsigned long long *)vb_dev_alloc_owned(e, (size_t)e->sm_count * V2_MAXB * 8);
    s->bar = (unsigned int *)vb_dev_alloc_owned(e, 1024);    /* [0] grid barrier, [32] error word, [64..127] attention tickets */
    s->ctr = (unsigned int *)vb_dev_alloc_owned(e, (size_t)V2_MAX_STEPS * V2_SUBPHASES * 4);
    s->st = (VbDecState *)vb_dev_alloc_owned(e, sizeof(VbDecState) * V2_MAXB);
    if (cudaHostAlloc((void **)&s->err_host, 64, cudaHostAllocMapped) != cudaSuccess || cudaHostGetDevicePointer((void **)&s->err_dev, s->err_host, 0) != cudaSuccess) { s->err_host = NULL; s->err_dev = (int *)(s->bar + 32); cudaGetLastError(); }
    s->logits_extra = (float *)vb_dev_alloc_owned(e, (size_t)(V2_MAXB - 1) * VOX_VOCAB_SIZE * 4);
    s->prof = NULL;
    e->weight_bytes = wb;
    return 0;
}

static void v2_prof_report(VbEngine *e, const V2Args &a) {
    size_t n = (size_t)e->sm_count * V2_PROF_SLOTS;
    long long *h = (long long *)malloc(n * 8);
    if (cudaMemcpyAsync(h, a.prof, n * 8, cudaMemcpyDeviceToHost, e->stream) != cudaSuccess || cudaStreamSynchronize(e->stream) != cudaSuccess) { free(h); return; }
    /* stamps per layer: 0 start, 1 qkv, 2 bar, 3 attn, 4 bar, 5 wo, 6 bar, 7 w13, 8 bar, 9 w2, then (bar) next layer's 0 */
    static const char *names[10] = { "qkv", "bar", "attn", "bar", "wo", "bar", "w13", "bar", "w2", "bar" };
    const int ctas[3] = { 0, e->sm_count / 2, e->sm_count - 1 };
    for (int ci = 0; ci < 3; ci++) {
        long long *t = h + (size_t)ctas[ci] * V2_PROF_SLOTS;
        double sum[10] = { 0 };
        for (int l = 1; l < VOX_DEC_LAYERS - 1; l++)
            for (int k = 0; k < 10; k++) sum[k] += (double)(t[l * 10 + k + 1] - t[l * 10 + k]);
        fprintf(stderr, "[v2 prof nb=%d] cta %3d cycles/layer:", a.nb, ctas[ci]);
        double tot = 0;
        for (int k = 0; k < 10; k++) { fprintf(stderr, " %s=%.0f", names[k], sum[k] / (VOX_DEC_LAYERS - 2)); tot += sum[k]; }
        fprintf(stderr, " | layer=%.0f | logits=%lld bar=%lld feedback=%lld step=%lld\n", tot / (VOX_DEC_LAYERS - 2),
                t[26 * 10 + 1] - t[26 * 10], t[26 * 10 + 2] - t[26 * 10 + 1], t[26 * 10 + 3] - t[26 * 10 + 2], t[26 * 10 + 3] - t[0]);
        const long long *q = t + V2_PROF_SLOTS - 8;
        const double nc = (double)(q[0] ? q[0] : 1);
        fprintf(stderr, "[v2 prof nb=%d] cta %3d whole launch: producer %lld chunks in %lld cycles (%.0f/chunk; waiting for a free slot %.0f, for the in-flight cap %.0f per chunk)\n",
                a.nb, ctas[ci], q[0], q[3], (double)q[3] / nc, (double)q[1] / nc, (double)q[2] / nc);
        const long long *w0 = t + V2_PROF_SLOTS - 20, *w11 = t + V2_PROF_SLOTS - 14;
        fprintf(stderr, "[v2 prof nb=%d] cta %3d consumer cycles per chunk  warp 0: wait-data %.0f math %.0f reduce %.0f cta-barrier %.0f | epilogue warp: wait-data %.0f math %.0f reduce %.0f cta-barrier %.0f epilogue %.0f\n",
                a.nb, ctas[ci], w0[0] / nc, w0[1] / nc, w0[2] / nc, w0[3] / nc, w11[0] / nc, w11[1] / nc, w11[2] / nc, w11[3] / nc, w11[4] / nc);
        /* sub-phases 0 QKV | 1,2 wo blocks | 3 w1|w3 | 4,5,6 w2 blocks: cycles of thread 0 from the top of the sub-phase to the first
         * chunk (activation load, RMSNorm; the load's latency itself overlaps the first chunk) | inside the weight loop / chunks taken */
        const long long *sa = t + 270;
        fprintf(stderr, "[v2 prof nb=%d] cta %3d sub-phase cycles per layer (activation load + norm | weight loop / chunks):", a.nb, ctas[ci]);
        for (int k = 0; k < 7; k++) fprintf(stderr, " s%d=%.0f|%.0f/%.1f", k, sa[k] / (double)(VOX_DEC_LAYERS - 2), sa[8 + k] / (double)(VOX_DEC_LAYERS - 2), sa[16 + k] / (double)(VOX_DEC_LAYERS - 2));
        fprintf(stderr, "\n");
    }
    free(h);
}

/* cols[i].engine's KV ring / logits / token buffers are used for column i; scratch and the launch stream are the leader's. */
extern "C" int vb_decoder_v2_launch(VbEngine *lead, const VbV2Col *cols, int nb, int n_steps, int verify, VbDecState *st_host) {
    if (nb < 1 || nb > V2_MAXB || n_steps < 1) return -1;
    if (n_steps > V2_MAX_STEPS) n_steps = V2_MAX_STEPS;
    if (!vb_decoder_v2_supported(lead)) return -1;
    v2_alloc(lead);
    V2Args a;
    memset(&a, 0, sizeof a);
    a.p = vb_make_dec_params(lead, 1);
    VbV2Scratch *s = &lead->v2;
    for (int b = 0; b < nb; b++) {
        const VbV2Col *c = verify ? &cols[0] : &cols[b];               /* verify: every column is a position of the one stream */
        VbEngine *e = c->engine;
        vb_decoder_alloc(e);
        a.col[b].adapter = c->d_adapter; a.col[b].kv_k = e->d_kv_k; a.col[b].kv_v = e->d_kv_v;
        a.col[b].logits = (verify && b > 0) ? s->logits_extra + (size_t)(b - 1) * VOX_VOCAB_SIZE : e->d_logits;
        a.col[b].tokens = e->d_tokens;
        a.col[b].pos0 = c->pos; a.col[b].token0 = c->prev_token; a.col[b].arow0 = c->adapter_row;
        a.col[b].n_steps = c->n_steps < n_steps ? c->n_steps : n_steps;
    }
    for (int b = nb; b < V2_MAXB; b++) { a.col[b] = a.col[0]; a.col[b].n_steps = 0; }
    a.x = s->x; a.q = s->q; a.attn_out = s->attn_out; a.gate = s->gate;
    a.part_m = s->part_m; a.part_l = s->part_l; a.part_o = s->part_o; a.argmax = s->argmax;
    a.bar = s->bar; a.ctr = s->ctr; a.st_out = s->st; a.err = s->err_dev;
    if (s->err_host) *s->err_host = 0;
    a.nb = nb; a.n_steps = n_steps; a.verify = verify;
    const char *ev;
    a.inflight_max = (ev = getenv("VOX_CUDA_V2_INFLIGHT")) ? atoi(ev) : 3;
    if (a.inflight_max < 1) a.inflight_max = 1;
    if (a.inflight_max > V2_SLOTS) a.inflight_max = V2_SLOTS;
    a.dynamic = (ev = getenv("VOX_CUDA_V2_DYNAMIC")) ? atoi(ev) : 1;
    a.dbg = (ev = getenv("VOX_CUDA_V2_DBG")) ? atoi(ev) : 0;     /* diagnostics only (results are wrong): 1 = no FMAs, 2 = no reductions/epilogues */
    a.prof = NULL; a.prof_step = -1;
    if ((ev = getenv("VOX_CUDA_V2_PROF")) && n_steps > atoi(ev)) {
        if (!s->prof) { const size_t wb = lead->weight_bytes; s->prof = (long long *)vb_dev_alloc_owned(lead, (size_t)lead->sm_count * V2_PROF_SLOTS * 8); lead->weight_bytes = wb; }
        if (cudaMemsetAsync(s->prof, 0, (size_t)lead->sm_count * V2_PROF_SLOTS * 8, lead->stream) != cudaSuccess) return -1;
        a.prof = s->prof; a.prof_step = atoi(ev);
    }
    if (cudaMemsetAsync(s->bar, 0, 1024, lead->stream) != cudaSuccess) return -1;
    if (cudaMemsetAsync(s->ctr, 0, (size_t)n_steps * V2_SUBPHASES * 4, lead->stream) != cudaSuccess) return -1;
    void *args[] = { &a };
    const void *fn = nb == 1 ? (const void *)k_dec_v2<1> : nb == 2 ? (const void *)k_dec_v2<2> : nb <= 4 ? (const void *)k_dec_v2<4> : (const void *)k_dec_v2<8>;
    cudaError_t le = cudaLaunchCooperativeKernel(fn, dim3(lead->sm_count), dim3(V2_THREADS), args, v2_smem_bytes(), lead->stream);
    if (le != cudaSuccess) { fprintf(stderr, "voxtral_b200: v2 decode launch failed: %s\n", cudaGetErrorString(le)); return -1; }
    lead->launches += 1;
    if (a.prof) v2_prof_report(lead, a);
    if (st_host) {
        if (cudaMemcpyAsync(st_host, s->st, sizeof(VbDecState) * nb, cudaMemcpyDeviceToHost, lead->stream) != cudaSuccess) return -1;
    }
    return 0;
}
