/*
 * vb_decode_persist.cu -- persistent cooperative decode kernel, direct-streaming variant (round 1's default; now the
 * A/B reference for vb_decode_v2.cu and the fallback when that kernel does not fit a device).
 *
 * One CTA per SM runs the whole multi-step greedy loop; the 131 phases of a step are separated by grid barriers; the token
 * feedback stays on the device.  Weights are NOT staged through shared memory: every thread streams its own k-columns with
 * 128-bit ld.global.nc.L1::no_allocate loads, 16 rows in flight per thread (the GEMV core reaches 99% of the measured HBM
 * peak inside the logits phase).  What a phase boundary costs -- drain, barrier, activation reload, RMSNorm, ramp, ~5.5 us --
 * is NOT hidden here; the step sits at 0.59 of the HBM roofline (profiles/r01_decode.md, which also records the L2-prefetch,
 * TMA-ring and mma.sync variants that were tried: they live on under tools/experiments/).
 *
 * Reference semantics: voxtral_decoder.c:586-706 per step, voxtral.c:1056-1093 for the loop.
 */
#include "vb_decode_persist_common.cuh"
#include <string.h>


/* y[row] = W[row,:] . x for the CTA's rows, 16 rows of 128-bit loads in flight per thread. */
template <int CPT, typename Epi>
__device__ __forceinline__ void gemv_stream(const Phase &f, int K, int NT, const float (&xr)[CPT * 8],
                                            float (*red)[16][MK_GROUP], int &redbuf, Epi epi) {
    constexpr int R = (CPT == 1) ? 16 : 4;             /* rows per batch of loads (register budget) */
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool active = t < NT;
    const uint16_t *wt = f.W + (size_t)t * 8;
    for (int g0 = 0; g0 < f.nrows; g0 += MK_GROUP) {
        const int gr = min(MK_GROUP, f.nrows - g0);
        float acc[MK_GROUP];
#pragma unroll
        for (int rb = 0; rb < MK_GROUP; rb += R) {
            uint4 w[R][CPT];
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int c = 0; c < CPT; c++) {
                    if (active && rb + r < gr)
                        w[r][c] = ldg_stream16(wt + (size_t)(f.row0 + g0 + rb + r) * K + (size_t)c * NT * 8);
                    else w[r][c] = make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
            for (int r = 0; r < R; r++) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < CPT; c++) s = dot8(w[r][c], &xr[c * 8], s);
                acc[rb + r] = s;
            }
        }
        float tot = warp_transpose_reduce<MK_GROUP>(acc, lane);
        if ((lane & 1) == 0) red[redbuf][warp][lane >> 1] = tot;
        cons_bar();
        if (warp == 0) {
            float sum = 0.f;
            if (lane < MK_GROUP) {
#pragma unroll
                for (int wv = 0; wv < 16; wv++) sum += red[redbuf][wv][lane];
            }
            epi(f.row0 + g0 + lane, sum, lane, lane < gr);
        }
        redbuf ^= 1;
    }
}

__global__ void __launch_bounds__(MK_CONS, 1) k_dec_persist(MegaArgs a) {
    __shared__ float red[2][16][MK_GROUP];
    __shared__ float sred[16];
    __shared__ unsigned long long cand[16];
    __shared__ int is_last;
    __shared__ __align__(16) float att_scr[MK_ATT_FLOATS];
    const DecParams &p = a.p;
    const int tid = threadIdx.x, lane = tid & 31;

    unsigned int gen = 0;
    int redbuf = 0;
    int pos = a.pos0, token = a.token0, arow = a.adapter_row0;
    const float *adapter = *p.adapter_pp;
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
                rmsnorm_cols_cons<1>(xr, p.attn_norm[layer], nullptr, NT, VOX_DEC_DIM, sred);
                Phase f = phase_of(p, layer, 0);
                float *kdst = p.kv_k + ((size_t)layer * VB_KV_SLOTS + slot) * VB_DEC_KV;
                float *vdst = p.kv_v + ((size_t)layer * VB_KV_SLOTS + slot) * VB_DEC_KV;
                const float *inv_freq = p.inv_freq;
                float *q = p.q;
                gemv_stream<1>(f, VOX_DEC_DIM, NT, xr, red, redbuf, [&](int row, float v, int, bool valid) {
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
            mega_attention(p, layer, pos, &is_last, att_scr, a.bar + 16);
            PROF(3);
            grid_barrier(a.bar, gen, a.err);
            PROF(4);
            {   /* ---- wo + residual ---- */
                const int NT = VB_DEC_Q / 8;
                float xr[8];
                load_x_cols_cg<1>(xr, p.attn_out, NT);
                Phase f = phase_of(p, layer, 1);
                float *x = p.x;
                gemv_stream<1>(f, VB_DEC_Q, NT, xr, red, redbuf, [&](int row, float v, int, bool valid) {
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
                rmsnorm_cols_cons<1>(xr, p.ffn_norm[layer], p.ada + (size_t)layer * VOX_DEC_DIM, NT, VOX_DEC_DIM, sred);
                Phase f = phase_of(p, layer, 2);
                float *gate = p.gate;
                gemv_stream<1>(f, VOX_DEC_DIM, NT, xr, red, redbuf, [&](int row, float v, int, bool valid) {
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
                Phase f = phase_of(p, layer, 3);
                float *x = p.x;
                gemv_stream<3>(f, VOX_DEC_HIDDEN, NT, xr, red, redbuf, [&](int row, float v, int, bool valid) {
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
            rmsnorm_cols_cons<1>(xr, p.final_norm, nullptr, NT, VOX_DEC_DIM, sred);
            Phase f = phase_of(p, 0, 4);
            float *logits = p.logits;
            unsigned long long best = 0ull;
            gemv_stream<1>(f, VOX_DEC_DIM, NT, xr, red, redbuf, [&](int row, float v, int, bool valid) {
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
        {   /* global argmax: every CTA reduces the per-CTA candidates, so every CTA knows the token */
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
            if (lane == 0) cand[tid >> 5] = best;
            cons_bar();
#pragma unroll
            for (int i = 0; i < 16; i++) if (cand[i] > best) best = cand[i];
            token = cand_index(best);
        }
        if (blockIdx.x == 0 && tid == 0) p.tokens[n_done] = token;
        n_done++; pos++; arow++;
        if (token == VB_TOKEN_EOS) { eos = 1; break; }
    }
    if (blockIdx.x == 0 && tid == 0) {
        VbDecState st;
        st.pos = pos; st.token = token; st.eos = eos; st.n_out = n_done; st.adapter_row = arow;
        st.pad[0] = st.pad[1] = st.pad[2] = 0;
        *p.st = st;
    }
}

extern "C" int vb_decoder_persist_supported(VbEngine *e) {
    if (e->persist_checked) return e->persist_ok;                   /* per engine = per device */
    e->persist_checked = 1; e->persist_ok = 0;
    int coop = 0, blocks = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device);
    if (!coop || e->sm_count > 160) return 0;                       /* attention merge assumes <= 20 CTAs per kv head */
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_dec_persist, MK_CONS, 0) != cudaSuccess || blocks < 1) return 0;
    e->persist_ok = 1;
    return 1;
}

extern "C" int vb_decoder_persist_launch(VbEngine *e, const float *d_adapter, int adapter_row, int n_steps,
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
    a.l2_ahead = 0;
    vb_mega_prof_begin(e, a, n_steps);
    void *args[] = { &a };
    VB_CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_dec_persist, dim3(e->sm_count), dim3(MK_CONS), args, 0, e->stream));
    e->launches += 1;
    vb_mega_prof_report(e, a, "persist");
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
    }
    free(h);
}
