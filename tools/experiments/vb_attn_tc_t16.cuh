/*
 * EXPERIMENT (not built into the library): k_attn_tc_t16 -- the tensor-memory-P attention (vb_attn_tc.cu:k_attn_tc_t) with 16 softmax
 * warps on column quarters (4 warps per scheduler) and one mbarrier arrival per warp.  Built and run on a B200 once
 * (gpurun_out/r02i_variants.json): the attention parity tests pass, greedy ids identical, encoder pass of the 60 s clip 33.15 ms
 * vs 33.03 ms with 8 warps and 33.81 ms for the shared-memory default -- doubling the softmax warps buys nothing, so the softmax
 * side is not issue- or dependent-chain-bound either (profiles/r02_encoder.md).  To build it again, paste it after k_attn_tc_t in
 * vb_attn_tc.cu (it uses that file's helpers and vb_tc.cuh's tc_tmem_ld16 / tc_tmem_st8) and launch it with FA_THREADS16 threads.
 */
/* k_attn_tc_t with 16 softmax warps on column quarters (4 warps per scheduler instead of 2: the softmax side is latency bound on
 * dependent chains) and one mbarrier arrival per warp.  Selected by VOX_CUDA_ATTN_P=tmem16. */
#define FA_THREADS16 576
__global__ void __launch_bounds__(FA_THREADS16, 1)
k_attn_tc_t16(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
          float *__restrict__ out, int ldo, int seq_q, int seq_k, int cols, float scale, int window, int q_offset,
          uint16_t *__restrict__ oplanes /* if set: the result as [3][seq_q][cols] bf16 planes (A operand of the wo GEMM) instead of out */) {
    extern __shared__ uint8_t fa_smem_raw[];
    uint8_t *sm = reinterpret_cast<uint8_t *>(((uintptr_t)fa_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + FA_OFF_BAR);
    uint64_t *q_full = bars;                    /* 1 */
    uint64_t *k_full = bars + 1;                /* 3 */
    uint64_t *k_empty = bars + 4;               /* 3 */
    uint64_t *v_full = bars + 7;                /* 2 */
    uint64_t *v_empty = bars + 9;               /* 2 */
    uint64_t *s_full = bars + 11;               /* 2 */
    uint64_t *s_empty = bars + 13;              /* 2 */
    uint64_t *o_full = bars + 15;               /* 2 */
    uint64_t *o_empty = bars + 17;              /* 2 */
    uint64_t *p_full = bars + 19;               /* 2 */
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 21);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, q0 = blockIdx.x * FA_BQ;
    const int hoff = h * FA_HD;

    /* key blocks this query tile can see */
    const int g_first = q_offset + q0;
    const int g_last = q_offset + min(q0 + FA_BQ, seq_q) - 1;
    int k_lo = 0;
    if (window > 0 && g_first - window + 1 > 0) k_lo = g_first - window + 1;
    k_lo = (k_lo / FA_BK) * FA_BK;
    const int k_hi = min(g_last + 1, seq_k);                           /* exclusive */
    const int nb = k_hi > k_lo ? (k_hi - k_lo + FA_BK - 1) / FA_BK : 0;

    if (threadIdx.x == 0) {
        tc_mbar_init(q_full, 1);
        for (int i = 0; i < FA_K_SLOTS; i++) { tc_mbar_init(&k_full[i], 1); tc_mbar_init(&k_empty[i], 1); }
        for (int i = 0; i < 2; i++) {
            tc_mbar_init(&v_full[i], 1); tc_mbar_init(&v_empty[i], 1);
            tc_mbar_init(&s_full[i], 1); tc_mbar_init(&s_empty[i], 16);
            tc_mbar_init(&o_full[i], 1); tc_mbar_init(&o_empty[i], 16);
        }
        tc_mbar_init(&p_full[0], 16); tc_mbar_init(&p_full[1], 16);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tm_s = tmem_base, tm_o = tmem_base + 128;           /* + b * 64 */
    const uint32_t tm_p = tmem_base + 256;                              /* + b * 96 + plane * 32: P(j) as packed bf16 pairs, 32 columns per plane */

    if (warp == 0) {
        if (lane == 0 && nb > 0) {
            tc_mbar_expect(q_full, 3 * FA_QPLANE);
            for (int p = 0; p < 3; p++) tc_tma_load_2d(sm + FA_OFF_Q + p * FA_QPLANE, &tmQ, hoff, p * seq_q + q0, q_full);
            for (int t = 0; t <= nb; t++) {
                if (t < nb) {                                           /* K(t) */
                    const int s = t % FA_K_SLOTS, u = t / FA_K_SLOTS;
                    tc_mbar_wait(&k_empty[s], (u & 1) ^ 1);
                    tc_mbar_expect(&k_full[s], 3 * FA_KTILE);
                    for (int p = 0; p < 3; p++)
                        tc_tma_load_2d(sm + FA_OFF_K + (s * 3 + p) * FA_KTILE, &tmK, hoff, p * seq_k + k_lo + t * FA_BK, &k_full[s]);
                }
                if (t >= 1) {                                           /* V(t-1) */
                    const int j = t - 1, s = j & 1, u = j >> 1;
                    tc_mbar_wait(&v_empty[s], (u & 1) ^ 1);
                    tc_mbar_expect(&v_full[s], 3 * FA_VTILE);
                    for (int p = 0; p < 3; p++)
                        tc_tma_load_2d(sm + FA_OFF_V + (s * 3 + p) * FA_VTILE, &tmV, k_lo + j * FA_BK, p * cols + hoff, &v_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (nb > 0) {                                                   /* whole warp converged; one elected lane issues */
            const uint32_t idesc = tc_idesc(FA_BQ, FA_BK);              /* M128 N64 for both products */
            /* descriptors differ only in the start-address field (bits [0,14) = address >> 4): one base per operand tile,
             * + 2 per k step of 16 elements (32 B), + the tile pitch >> 4 per plane */
            const uint64_t dq = tc_smem_desc(s32(sm + FA_OFF_Q));
            const uint64_t dk0 = tc_smem_desc(s32(sm + FA_OFF_K)), dv0 = tc_smem_desc(s32(sm + FA_OFF_V));
            tc_mbar_wait(q_full, 0);
            for (int j = -1; j < nb; j++) {
                if (j + 1 < nb) {                                       /* S(j+1) = Q K(j+1)^T */
                    const int jj = j + 1, s = jj % FA_K_SLOTS, b = jj & 1;
                    tc_mbar_wait(&k_full[s], (jj / FA_K_SLOTS) & 1);
                    tc_mbar_wait(&s_empty[b], ((jj >> 1) & 1) ^ 1);
                    tc_fence_after();
                    if (tc_elect_one()) {
                        const uint64_t dk = dk0 + (uint64_t)(s * ((3 * FA_KTILE) >> 4));
#pragma unroll
                        for (int pr = 0; pr < 6; pr++) {
#pragma unroll
                            for (int k = 0; k < FA_HD / 16; k++)
                                tc_umma_bf16(tm_s + b * 64, dq + (uint64_t)(FA_PA(pr) * (FA_QPLANE >> 4) + 2 * k),
                                             dk + (uint64_t)(FA_PB(pr) * (FA_KTILE >> 4) + 2 * k), idesc, (pr | k) ? 1u : 0u);
                        }
                        tc_umma_commit(&k_empty[s]);
                        tc_umma_commit(&s_full[b]);
                    }
                    __syncwarp();
                }
                if (j >= 0) {                                           /* (P V)(j) */
                    const int s = j & 1, u = j >> 1;
                    tc_mbar_wait(&v_full[s], u & 1);
                    tc_mbar_wait(&p_full[s], u & 1);
                    tc_mbar_wait(&o_empty[s], (u & 1) ^ 1);
                    tc_fence_after();
                    if (tc_elect_one()) {
                        const uint64_t dv = dv0 + (uint64_t)(s * ((3 * FA_VTILE) >> 4));
#pragma unroll
                        for (int pr = 0; pr < 6; pr++) {
#pragma unroll
                            for (int k = 0; k < FA_BK / 16; k++)
                                tc_umma_bf16_ts(tm_o + s * 64, tm_p + s * 96 + FA_PA(pr) * 32 + 8 * k,
                                                dv + (uint64_t)(FA_PB(pr) * (FA_VTILE >> 4) + 2 * k), idesc, (pr | k) ? 1u : 0u);
                        }
                        tc_umma_commit(&v_empty[s]);
                        tc_umma_commit(&o_full[s]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        /* softmax warps 2..17: warps w, w+4, w+8, w+12 own the same TMEM lane quarter (warp % 4) and split the 64 columns (keys
         * of S, head dims of O) in quarters of 16.  Per key block the quarters exchange their partial row maximum; the partial
         * row sums meet at the end.  One mbarrier arrival per warp (lane 0 after __syncwarp). */
        const int qd = warp & 3, qt = (warp - 2) >> 2;
        const int r = qd * 32 + lane;                                   /* row of the query tile */
        const uint32_t lane_off = ((uint32_t)(qd * 32) << 16) + (uint32_t)(qt * 16);
        const int g = q_offset + q0 + r;                                /* index of this query in the key buffer */
        int lo = 0;
        if (window > 0 && g - window + 1 > 0) lo = g - window + 1;
        const int hi = min(g, seq_k - 1);
        int lo_max = 0;
        if (window > 0 && g_last - window + 1 > 0) lo_max = g_last - window + 1;
        const int hi_min = min(g_first, seq_k - 1);
        float o[16];
#pragma unroll
        for (int d = 0; d < 16; d++) o[d] = 0.f;
        float m = -1e30f, l = 0.f, alpha_prev = 1.f;
        float *xch = reinterpret_cast<float *>(sm + FA_OFF_P);          /* the P region of shared memory is free here: [2][4][128] maxima, then [4][128] sums */

        for (int j = 0; j < nb; j++) {
            const int b = j & 1;
            const int k0 = k_lo + j * FA_BK;
            uint32_t sv[16];
            tc_mbar_wait(&s_full[b], (j >> 1) & 1);
            tc_fence_after();
            tc_tmem_ld16(tm_s + b * 64 + lane_off, sv);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&s_empty[b]);

            float mx = -1e30f;
            if (k0 >= lo_max && k0 + FA_BK - 1 <= hi_min) {             /* interior block (uniform over the CTA) */
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    const float a = __uint_as_float(sv[c]) * scale;
                    sv[c] = __float_as_uint(a);
                    mx = fmaxf(mx, a);
                }
            } else {
                const int c_lo = lo - k0 - qt * 16, c_hi = hi - k0 - qt * 16;
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    float a = __uint_as_float(sv[c]) * scale;
                    a = (c >= c_lo && c <= c_hi) ? a : -1e30f;
                    sv[c] = __float_as_uint(a);
                    mx = fmaxf(mx, a);
                }
            }
            xch[(b * 4 + qt) * FA_BQ + r] = mx;
            asm volatile("bar.sync 1, 512;" ::: "memory");
            mx = fmaxf(fmaxf(xch[(b * 4 + 0) * FA_BQ + r], xch[(b * 4 + 1) * FA_BQ + r]),
                       fmaxf(xch[(b * 4 + 2) * FA_BQ + r], xch[(b * 4 + 3) * FA_BQ + r]));
            const float mn = fmaxf(m, mx);
            const float alpha = fa_exp(m - mn);
            float rs = 0.f;
#pragma unroll
            for (int c = 0; c < 16; c++) {
                const float a = __uint_as_float(sv[c]);
                const float p = a > -1e29f ? fa_exp(a - mn) : 0.f;
                sv[c] = __float_as_uint(p);
                rs += p;
            }
            l = l * alpha + rs;
            m = mn;

            {   /* P(j): this thread's 16 keys = 8 packed words per plane */
                uint32_t w0[8], w1[8], w2[8];
#pragma unroll
                for (int e2 = 0; e2 < 8; e2++) {
                    float x = __uint_as_float(sv[2 * e2]), y = __uint_as_float(sv[2 * e2 + 1]);
                    w0[e2] = tc_pack_bf16x2(x, y);
                    x -= vb_bf16_lo(w0[e2]); y -= vb_bf16_hi(w0[e2]);   /* exact */
                    w1[e2] = tc_pack_bf16x2(x, y);
                    x -= vb_bf16_lo(w1[e2]); y -= vb_bf16_hi(w1[e2]);
                    w2[e2] = tc_pack_bf16x2(x, y);
                }
                const uint32_t pa = tm_p + b * 96 + ((uint32_t)(qd * 32) << 16) + (uint32_t)(qt * 8);
                tc_tmem_st8(pa, w0);
                tc_tmem_st8(pa + 32, w1);
                tc_tmem_st8(pa + 64, w2);
                tc_tmem_wait_st();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&p_full[b]);

            if (j >= 1) {                                               /* O = O * alpha(j-1) + (P V)(j-1) */
                tc_mbar_wait(&o_full[b ^ 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
                tc_tmem_ld16(tm_o + (b ^ 1) * 64 + lane_off, sv);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(&o_empty[b ^ 1]);
#pragma unroll
                for (int d = 0; d < 16; d++) o[d] = fmaf(o[d], alpha_prev, __uint_as_float(sv[d]));
            }
            alpha_prev = alpha;
        }
        if (nb > 0) {
            const int b = (nb - 1) & 1;
            uint32_t sv[16];
            tc_mbar_wait(&o_full[b], ((nb - 1) >> 1) & 1);
            tc_fence_after();
            tc_tmem_ld16(tm_o + b * 64 + lane_off, sv);
#pragma unroll
            for (int d = 0; d < 16; d++) o[d] = fmaf(o[d], alpha_prev, __uint_as_float(sv[d]));
        }
        /* row sum = the four quarters' partial sums (same running maximum in all), added in a fixed order */
        float *lx = xch + 8 * FA_BQ;
        asm volatile("bar.sync 1, 512;" ::: "memory");                  /* everybody is past the last read of the maxima */
        lx[qt * FA_BQ + r] = l;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        l = (lx[0 * FA_BQ + r] + lx[1 * FA_BQ + r]) + (lx[2 * FA_BQ + r] + lx[3 * FA_BQ + r]);
        if (q0 + r < seq_q) {
            const float inv = l > 0.f ? 1.0f / l : 0.f;
            if (oplanes) {
                uint16_t *dst = oplanes + (size_t)(q0 + r) * cols + hoff + qt * 16;
                const size_t plane = (size_t)seq_q * cols;
#pragma unroll
                for (int d = 0; d < 16; d += 8) {
                    uint32_t w0[4], w1[4], w2[4];
#pragma unroll
                    for (int e2 = 0; e2 < 4; e2++) {
                        float x = o[d + 2 * e2] * inv, y = o[d + 2 * e2 + 1] * inv;
                        w0[e2] = tc_pack_bf16x2(x, y);
                        x -= vb_bf16_lo(w0[e2]); y -= vb_bf16_hi(w0[e2]);
                        w1[e2] = tc_pack_bf16x2(x, y);
                        x -= vb_bf16_lo(w1[e2]); y -= vb_bf16_hi(w1[e2]);
                        w2[e2] = tc_pack_bf16x2(x, y);
                    }
                    *reinterpret_cast<uint4 *>(dst + d) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
                    *reinterpret_cast<uint4 *>(dst + d + plane) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
                    *reinterpret_cast<uint4 *>(dst + d + 2 * plane) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
                }
            } else {
                float *dst = out + (size_t)(q0 + r) * ldo + hoff + qt * 16;
#pragma unroll
                for (int d = 0; d < 16; d += 4)
                    *reinterpret_cast<float4 *>(dst + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(512));
    }
}

