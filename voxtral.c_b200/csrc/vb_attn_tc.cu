/*
 * vb_attn_tc.cu -- banded causal attention of the encoder (head_dim 64, MHA, window 750) on the 5th-generation tensor
 * cores: S = Q K^T and O += P V as tcgen05.mma with TMEM accumulators, operands staged by TMA, online softmax in registers.
 * Semantics: voxtral_kernels.c:412-482 (keys [max(0,g-W+1), min(g,seq_k-1)], g = q_offset+i), as used by the encoder layer
 * loop voxtral_encoder.c:555-571.
 *
 * Numerics.  The reference computes in f32.  Every f32 operand is split into three bf16 planes x = x0 + x1 + x2 (each step of
 * the split is exact), and a product a*b is evaluated as the six plane products whose weight is >= 2^-18:
 * a0b0, a0b1, a1b0, a1b1, a0b2, a2b0 -- each exact in the f32 accumulator; what is dropped is <= 2^-26 |a||b|, below f32
 * rounding of the sum.  P (the softmax numerators, in [0,1]) is split the same way before P V.  Max / exp / sum are f32 in
 * registers (ex2.approx on (s - m) log2e: relative error 2^-22), so the result agrees with the CUDA-core kernels
 * (k_attn_tile64 / k_attn_warp) to a few f32 ulps -- below what the tensor core's accumulation order moves.
 *
 * One CTA = 128 queries of one head; keys in blocks of 64.  320 threads:
 *   warp 0     TMA: Q planes once (3 x [128 x 64] bf16), then per key block K planes (3 x [64 keys x 64 d]) into a 3-slot ring
 *              and V^T planes (3 x [64 d x 64 keys]) into a 2-slot ring, all SWIZZLE_128B boxes
 *   warp 1     MMA issue (whole warp converged, one elected lane issues, vb_tc.cuh:tc_elect_one): S(j+1) = Q K(j+1)^T is issued
 *              BEFORE P(j) V(j), so the tensor pipe works on the next scores while the softmax warps turn S(j) into P(j);
 *              24 + 24 tcgen05.mma (M128 N64 K16) per key block
 *   warps 2-9  softmax, thread = (query row, column half): warps w and w+4 own the same TMEM lane quarter and split the 64
 *              columns (keys of S, head dims of O).  tcgen05.ld S(j) (32 columns), mask (edge blocks only), running max -- the
 *              two halves exchange their partial maximum through shared memory + one named barrier -- and sum, P(j) -> three
 *              bf16 planes written to shared memory in the SWIZZLE_128B K-major layout the MMA reads, then the deferred
 *              O = O*alpha + (P V)(j-1) from TMEM into 32 registers.  Output: f32 rows, or bf16 planes for the wo GEMM.
 * TMEM: S double-buffered (2 x 64 columns), P V block result double-buffered (2 x 64 columns).
 * Measured: profiles/r02_encoder.md (188 us per layer at 3196 positions; the CUDA-core kernel k_attn_tile64 took 701 us).
 *
 * V^T: the B operand of P V has to be K-major, i.e. [d][key] with keys contiguous; k_vt_planes transposes V while splitting
 * it.  Q and K planes come from the GEMM's splitter (vb_tc_split_planes).
 */
#include "vb_tc.cuh"
#include <stdlib.h>

#define FA_BQ 128
#define FA_BK 64
#define FA_HD 64
#define FA_THREADS 320                          /* TMA warp, MMA warp, 8 softmax warps */
#define FA_QPLANE (FA_BQ * 128)                 /* 16 KB: 128 rows x 64 bf16 */
#define FA_KTILE  (FA_BK * 128)                 /*  8 KB: 64 keys x 64 bf16 */
#define FA_VTILE  (FA_HD * 128)                 /*  8 KB: 64 d rows x 64 keys */
#define FA_K_SLOTS 3
#define FA_V_SLOTS 2
#define FA_OFF_Q 0
#define FA_OFF_K (3 * FA_QPLANE)
#define FA_OFF_V (FA_OFF_K + FA_K_SLOTS * 3 * FA_KTILE)
#define FA_OFF_P (FA_OFF_V + FA_V_SLOTS * 3 * FA_VTILE)
#define FA_OFF_BAR (FA_OFF_P + 3 * FA_QPLANE)
#define FA_OFF_XCH (FA_OFF_BAR + 256)          /* 4 x 128 floats: partial row maxima, then 2 x 128: partial row sums */
#define FA_SMEM_BYTES (FA_OFF_XCH + 6 * FA_BQ * 4 + 1024 /*align*/)
#define FA_TMEM_COLS 256

/* the six plane pairs (a plane, b plane), largest first: (0,0) (0,1) (1,0) (1,1) (0,2) (2,0) */
#define FA_PA(pr) ((0x201100u >> (4 * (pr))) & 0xFu)
#define FA_PB(pr) ((0x021010u >> (4 * (pr))) & 0xFu)

/* e^x for x <= 0 through ex2.approx (relative error 2^-22; the argument product rounds at 2^-24 |x|) */
__device__ __forceinline__ float fa_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
__device__ __forceinline__ void fa_sts128(uint32_t saddr, const uint32_t (&w)[4]) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(saddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}

/* V f32 [seq_k, ldkv] -> V^T planes [3][cols][nk_pad] bf16 (keys contiguous), zero for keys >= seq_k.
 * CTA = 64 keys x 32 columns through a padded shared tile; a thread writes two keys (one 32-bit word) per plane. */
__global__ void __launch_bounds__(256)
k_vt_planes(const float *__restrict__ V, int ldkv, int seq_k, int cols, int nk_pad, uint16_t *__restrict__ vt) {
    __shared__ float t[64][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           /* 32 x 8 */
    const int k0 = blockIdx.x * 64, c0 = blockIdx.y * 32;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int key = k0 + ty + 8 * i;
        t[ty + 8 * i][tx] = key < seq_k ? V[(size_t)key * ldkv + c0 + tx] : 0.f;
    }
    __syncthreads();
    const size_t plane = (size_t)cols * nk_pad;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int col = c0 + ty + 8 * i;
        const int key = k0 + 2 * tx;
        if (key >= nk_pad) continue;
        uint32_t a0, a1, a2, b0, b1, b2;
        tc_split3(t[2 * tx][ty + 8 * i], a0, a1, a2);
        tc_split3(t[2 * tx + 1][ty + 8 * i], b0, b1, b2);
        uint32_t *dst = reinterpret_cast<uint32_t *>(vt + (size_t)col * nk_pad + key);
        dst[0] = a0 | (b0 << 16);
        *reinterpret_cast<uint32_t *>(reinterpret_cast<uint16_t *>(dst) + plane) = a1 | (b1 << 16);
        *reinterpret_cast<uint32_t *>(reinterpret_cast<uint16_t *>(dst) + 2 * plane) = a2 | (b2 << 16);
    }
}

__global__ void __launch_bounds__(FA_THREADS, 1)
k_attn_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
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
    uint64_t *p_full = bars + 19;               /* 1 */
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 20);

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
            tc_mbar_init(&s_full[i], 1); tc_mbar_init(&s_empty[i], 256);
            tc_mbar_init(&o_full[i], 1); tc_mbar_init(&o_empty[i], 256);
        }
        tc_mbar_init(p_full, 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(tmem_slot)), "n"(FA_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tm_s = tmem_base, tm_o = tmem_base + 128;           /* + b * 64 */

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
            const uint64_t dq = tc_smem_desc(s32(sm + FA_OFF_Q)), dp = tc_smem_desc(s32(sm + FA_OFF_P));
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
                    tc_mbar_wait(p_full, j & 1);
                    tc_mbar_wait(&o_empty[s], (u & 1) ^ 1);
                    tc_fence_after();
                    if (tc_elect_one()) {
                        const uint64_t dv = dv0 + (uint64_t)(s * ((3 * FA_VTILE) >> 4));
#pragma unroll
                        for (int pr = 0; pr < 6; pr++) {
#pragma unroll
                            for (int k = 0; k < FA_BK / 16; k++)
                                tc_umma_bf16(tm_o + s * 64, dp + (uint64_t)(FA_PA(pr) * (FA_QPLANE >> 4) + 2 * k),
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
        /* softmax warps 2..9: a warp may only touch TMEM lanes [32*(warp%4), +32); warps w and w+4 share the rows of a lane
         * quarter and split the 64 columns (keys of S, head dims of O) in halves.  The only thing the halves exchange per key
         * block is their partial row maximum (shared memory + one named barrier); their partial row sums meet at the end. */
        const int qd = warp & 3, hf = (warp - 2) >> 2;
        const int r = qd * 32 + lane;                                   /* row of the query tile */
        const uint32_t lane_off = ((uint32_t)(qd * 32) << 16) + (uint32_t)(hf * 32);
        const int g = q_offset + q0 + r;                                /* index of this query in the key buffer */
        int lo = 0;
        if (window > 0 && g - window + 1 > 0) lo = g - window + 1;
        const int hi = min(g, seq_k - 1);
        /* key blocks every real row of the tile sees completely need no mask */
        int lo_max = 0;
        if (window > 0 && g_last - window + 1 > 0) lo_max = g_last - window + 1;
        const int hi_min = min(g_first, seq_k - 1);
        float o[32];
#pragma unroll
        for (int d = 0; d < 32; d++) o[d] = 0.f;
        float m = -1e30f, l = 0.f, alpha_prev = 1.f;
        const uint32_t prow = s32(sm + FA_OFF_P) + r * 128;
        const int sw = r & 7;
        float *xch = reinterpret_cast<float *>(sm + FA_OFF_XCH);        /* [2 parities][2 halves][128 rows] */

        for (int j = 0; j < nb; j++) {
            const int b = j & 1;
            const int k0 = k_lo + j * FA_BK;
            uint32_t sv[32];
            tc_mbar_wait(&s_full[b], (j >> 1) & 1);
            tc_fence_after();
            tc_tmem_ld32(tm_s + b * 64 + lane_off, sv);
            tc_fence_before();
            tc_mbar_arrive(&s_empty[b]);

            float mx = -1e30f;
            if (k0 >= lo_max && k0 + FA_BK - 1 <= hi_min) {             /* interior block (uniform over the CTA) */
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const float a = __uint_as_float(sv[c]) * scale;
                    sv[c] = __float_as_uint(a);
                    mx = fmaxf(mx, a);
                }
            } else {
                const int c_lo = lo - k0 - hf * 32, c_hi = hi - k0 - hf * 32;
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    float a = __uint_as_float(sv[c]) * scale;
                    a = (c >= c_lo && c <= c_hi) ? a : -1e30f;
                    sv[c] = __float_as_uint(a);
                    mx = fmaxf(mx, a);
                }
            }
            xch[(b * 2 + hf) * FA_BQ + r] = mx;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mx = fmaxf(mx, xch[(b * 2 + (hf ^ 1)) * FA_BQ + r]);
            const float mn = fmaxf(m, mx);
            const float alpha = fa_exp(m - mn);
            float rs = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const float a = __uint_as_float(sv[c]);
                const float p = a > -1e29f ? fa_exp(a - mn) : 0.f;
                sv[c] = __float_as_uint(p);
                rs += p;
            }
            l = l * alpha + rs;
            m = mn;

            if (j >= 1) tc_mbar_wait(&o_full[b ^ 1], ((j - 1) >> 1) & 1);     /* (P V)(j-1) retired: P may be overwritten */
            /* P(j) -> three bf16 planes, rows of 128 B, 16-byte chunk c of row r at chunk (c ^ (r & 7)) (SWIZZLE_128B) */
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) {
                uint32_t w0[4], w1[4], w2[4];
#pragma unroll
                for (int e2 = 0; e2 < 4; e2++) {
                    float x = __uint_as_float(sv[c4 * 8 + e2 * 2]), y = __uint_as_float(sv[c4 * 8 + e2 * 2 + 1]);
                    w0[e2] = tc_pack_bf16x2(x, y);
                    x -= vb_bf16_lo(w0[e2]); y -= vb_bf16_hi(w0[e2]);   /* exact */
                    w1[e2] = tc_pack_bf16x2(x, y);
                    x -= vb_bf16_lo(w1[e2]); y -= vb_bf16_hi(w1[e2]);
                    w2[e2] = tc_pack_bf16x2(x, y);
                }
                const uint32_t dst = prow + (((hf * 4 + c4) ^ sw) << 4);
                fa_sts128(dst, w0);
                fa_sts128(dst + FA_QPLANE, w1);
                fa_sts128(dst + 2 * FA_QPLANE, w2);
            }
            tc_fence_proxy_async();
            tc_mbar_arrive(p_full);

            if (j >= 1) {                                               /* O = O * alpha(j-1) + (P V)(j-1) */
                tc_fence_after();
                tc_tmem_ld32(tm_o + (b ^ 1) * 64 + lane_off, sv);
                tc_fence_before();
                tc_mbar_arrive(&o_empty[b ^ 1]);
#pragma unroll
                for (int d = 0; d < 32; d++) o[d] = fmaf(o[d], alpha_prev, __uint_as_float(sv[d]));
            }
            alpha_prev = alpha;
        }
        if (nb > 0) {
            const int b = (nb - 1) & 1;
            uint32_t sv[32];
            tc_mbar_wait(&o_full[b], ((nb - 1) >> 1) & 1);
            tc_fence_after();
            tc_tmem_ld32(tm_o + b * 64 + lane_off, sv);
#pragma unroll
            for (int d = 0; d < 32; d++) o[d] = fmaf(o[d], alpha_prev, __uint_as_float(sv[d]));
        }
        /* row sum = the two halves' partial sums (same running maximum in both) */
        float *lx = reinterpret_cast<float *>(sm + FA_OFF_XCH) + 4 * FA_BQ;
        lx[hf * FA_BQ + r] = l;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        l += lx[(hf ^ 1) * FA_BQ + r];
        if (q0 + r < seq_q) {
            const float inv = l > 0.f ? 1.0f / l : 0.f;
            if (oplanes) {
                uint16_t *dst = oplanes + (size_t)(q0 + r) * cols + hoff + hf * 32;
                const size_t plane = (size_t)seq_q * cols;
#pragma unroll
                for (int d = 0; d < 32; d += 8) {
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
                float *dst = out + (size_t)(q0 + r) * ldo + hoff + hf * 32;
#pragma unroll
                for (int d = 0; d < 32; d += 4)
                    *reinterpret_cast<float4 *>(dst + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(FA_TMEM_COLS));
    }
}

/* The same kernel with P in tensor memory: the softmax warps store P(j) as packed bf16 pairs with tcgen05.st and P V reads its A
 * operand from TMEM (tcgen05.mma [d], [a], b-desc).  No shared-memory P buffer, no STS / fence.proxy.async, and P is double-buffered
 * (2 x 96 columns), so writing P(j+1) does not wait for (P V)(j): the serial chain P write -> P V -> P write of k_attn_tc is gone.
 * Selected by VOX_CUDA_ATTN_P=tmem. */
__global__ void __launch_bounds__(FA_THREADS, 1)
k_attn_tc_t(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
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
            tc_mbar_init(&s_full[i], 1); tc_mbar_init(&s_empty[i], 256);
            tc_mbar_init(&o_full[i], 1); tc_mbar_init(&o_empty[i], 256);
        }
        tc_mbar_init(&p_full[0], 256); tc_mbar_init(&p_full[1], 256);
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
        /* softmax warps 2..9: a warp may only touch TMEM lanes [32*(warp%4), +32); warps w and w+4 share the rows of a lane
         * quarter and split the 64 columns (keys of S, head dims of O) in halves.  The only thing the halves exchange per key
         * block is their partial row maximum (shared memory + one named barrier); their partial row sums meet at the end. */
        const int qd = warp & 3, hf = (warp - 2) >> 2;
        const int r = qd * 32 + lane;                                   /* row of the query tile */
        const uint32_t lane_off = ((uint32_t)(qd * 32) << 16) + (uint32_t)(hf * 32);
        const int g = q_offset + q0 + r;                                /* index of this query in the key buffer */
        int lo = 0;
        if (window > 0 && g - window + 1 > 0) lo = g - window + 1;
        const int hi = min(g, seq_k - 1);
        /* key blocks every real row of the tile sees completely need no mask */
        int lo_max = 0;
        if (window > 0 && g_last - window + 1 > 0) lo_max = g_last - window + 1;
        const int hi_min = min(g_first, seq_k - 1);
        float o[32];
#pragma unroll
        for (int d = 0; d < 32; d++) o[d] = 0.f;
        float m = -1e30f, l = 0.f, alpha_prev = 1.f;
        float *xch = reinterpret_cast<float *>(sm + FA_OFF_XCH);        /* [2 parities][2 halves][128 rows] */

        for (int j = 0; j < nb; j++) {
            const int b = j & 1;
            const int k0 = k_lo + j * FA_BK;
            uint32_t sv[32];
            tc_mbar_wait(&s_full[b], (j >> 1) & 1);
            tc_fence_after();
            tc_tmem_ld32(tm_s + b * 64 + lane_off, sv);
            tc_fence_before();
            tc_mbar_arrive(&s_empty[b]);

            float mx = -1e30f;
            if (k0 >= lo_max && k0 + FA_BK - 1 <= hi_min) {             /* interior block (uniform over the CTA) */
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const float a = __uint_as_float(sv[c]) * scale;
                    sv[c] = __float_as_uint(a);
                    mx = fmaxf(mx, a);
                }
            } else {
                const int c_lo = lo - k0 - hf * 32, c_hi = hi - k0 - hf * 32;
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    float a = __uint_as_float(sv[c]) * scale;
                    a = (c >= c_lo && c <= c_hi) ? a : -1e30f;
                    sv[c] = __float_as_uint(a);
                    mx = fmaxf(mx, a);
                }
            }
            xch[(b * 2 + hf) * FA_BQ + r] = mx;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mx = fmaxf(mx, xch[(b * 2 + (hf ^ 1)) * FA_BQ + r]);
            const float mn = fmaxf(m, mx);
            const float alpha = fa_exp(m - mn);
            float rs = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const float a = __uint_as_float(sv[c]);
                const float p = a > -1e29f ? fa_exp(a - mn) : 0.f;
                sv[c] = __float_as_uint(p);
                rs += p;
            }
            l = l * alpha + rs;
            m = mn;

            /* P(j) -> three bf16 planes in TMEM (the A operand of P V): this thread's 32 keys are 16 packed words per plane.
             * Buffer b was last read by (P V)(j-2), whose completion this thread observed in iteration j-1 (o_full). */
            {
                uint32_t w0[16], w1[16], w2[16];
#pragma unroll
                for (int e2 = 0; e2 < 16; e2++) {
                    float x = __uint_as_float(sv[2 * e2]), y = __uint_as_float(sv[2 * e2 + 1]);
                    w0[e2] = tc_pack_bf16x2(x, y);
                    x -= vb_bf16_lo(w0[e2]); y -= vb_bf16_hi(w0[e2]);   /* exact */
                    w1[e2] = tc_pack_bf16x2(x, y);
                    x -= vb_bf16_lo(w1[e2]); y -= vb_bf16_hi(w1[e2]);
                    w2[e2] = tc_pack_bf16x2(x, y);
                }
                const uint32_t pa = tm_p + b * 96 + ((uint32_t)(qd * 32) << 16) + (uint32_t)(hf * 16);
                tc_tmem_st16(pa, w0);
                tc_tmem_st16(pa + 32, w1);
                tc_tmem_st16(pa + 64, w2);
                tc_tmem_wait_st();
            }
            tc_fence_before();
            tc_mbar_arrive(&p_full[b]);

            if (j >= 1) tc_mbar_wait(&o_full[b ^ 1], ((j - 1) >> 1) & 1);
            if (j >= 1) {                                               /* O = O * alpha(j-1) + (P V)(j-1) */
                tc_fence_after();
                tc_tmem_ld32(tm_o + (b ^ 1) * 64 + lane_off, sv);
                tc_fence_before();
                tc_mbar_arrive(&o_empty[b ^ 1]);
#pragma unroll
                for (int d = 0; d < 32; d++) o[d] = fmaf(o[d], alpha_prev, __uint_as_float(sv[d]));
            }
            alpha_prev = alpha;
        }
        if (nb > 0) {
            const int b = (nb - 1) & 1;
            uint32_t sv[32];
            tc_mbar_wait(&o_full[b], ((nb - 1) >> 1) & 1);
            tc_fence_after();
            tc_tmem_ld32(tm_o + b * 64 + lane_off, sv);
#pragma unroll
            for (int d = 0; d < 32; d++) o[d] = fmaf(o[d], alpha_prev, __uint_as_float(sv[d]));
        }
        /* row sum = the two halves' partial sums (same running maximum in both) */
        float *lx = reinterpret_cast<float *>(sm + FA_OFF_XCH) + 4 * FA_BQ;
        lx[hf * FA_BQ + r] = l;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        l += lx[(hf ^ 1) * FA_BQ + r];
        if (q0 + r < seq_q) {
            const float inv = l > 0.f ? 1.0f / l : 0.f;
            if (oplanes) {
                uint16_t *dst = oplanes + (size_t)(q0 + r) * cols + hoff + hf * 32;
                const size_t plane = (size_t)seq_q * cols;
#pragma unroll
                for (int d = 0; d < 32; d += 8) {
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
                float *dst = out + (size_t)(q0 + r) * ldo + hoff + hf * 32;
#pragma unroll
                for (int d = 0; d < 32; d += 4)
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

/* 0 = use the CUDA-core kernels (VOX_CUDA_ATTN=simt), 1 = tensor cores */
int vb_attn_tc_enabled(void) {
    static int v = -1;
    if (v < 0) { const char *s = getenv("VOX_CUDA_ATTN"); v = (s && (s[0] == 's' || s[0] == '0')) ? 0 : 1; }
    return v;
}

int vb_attn_tc_usable(int seq_q, int seq_k, int n_heads, int n_kv_heads, int head_dim, int ldq, int ldkv, int ldo) {
    return head_dim == FA_HD && n_heads == n_kv_heads && seq_q >= 32 && seq_k >= 1 && (ldq % 4) == 0 && (ldkv % 4) == 0 && (ldo % 4) == 0 &&
           ((n_heads * head_dim) % 64) == 0;
}

/* where P(j) lives between the softmax warps and P V: shared memory (k_attn_tc, the default) or tensor memory (k_attn_tc_t,
 * VOX_CUDA_ATTN_P=tmem).  Read per call so that a test can run both in one process.  Measured (profiles/r02_encoder.md): both
 * give bit-identical results; the TMEM form is 2.3 % faster on the 60 s encoder pass (33.03 vs 33.81 ms) -- the serial
 * P write -> P V -> P write chain it removes was not the bound -- and has not been through the whole GPU suite, so it stays opt-in. */
#define FA_P_DEFAULT_TMEM 0
static int attn_p_in_tmem(void) {
    const char *s = getenv("VOX_CUDA_ATTN_P");
    if (!s) return FA_P_DEFAULT_TMEM;
    return s[0] == 't';
}

/* V^T planes, tensor maps, launch.  qp: [3][seq_q][cols], kp: [3][seq_k][cols] complete. */
static void attn_tc_launch(VbEngine *e, float *out, int ldo, const uint16_t *qp, const uint16_t *kp, const float *V, int ldkv,
                           int seq_q, int seq_k, int n_heads, float scale, int window, int q_offset, uint16_t *oplanes) {
    static unsigned int attr_done = 0;                                  /* one bit per device */
    const unsigned int dev_bit = 1u << (e->device & 31);
    const int cols = n_heads * FA_HD;
    const int nk_pad = (seq_k + 63) & ~63;
    uint16_t *vt = (uint16_t *)vb_ws(e, VB_WS_ATT_VT, (size_t)3 * cols * nk_pad * 2 + 256);
    dim3 tg(nk_pad / 64, cols / 32);
    k_vt_planes<<<tg, 256, 0, e->stream>>>(V, ldkv, seq_k, cols, nk_pad, vt);
    VB_CUDA_OK(cudaGetLastError());
    CUtensorMap tmQ, tmK, tmV;
    vb_tc_make_map(&tmQ, qp, (uint64_t)cols, (uint64_t)3 * seq_q, (uint64_t)cols * 2, FA_BQ);
    vb_tc_make_map(&tmK, kp, (uint64_t)cols, (uint64_t)3 * seq_k, (uint64_t)cols * 2, FA_BK);
    vb_tc_make_map(&tmV, vt, (uint64_t)nk_pad, (uint64_t)3 * cols, (uint64_t)nk_pad * 2, FA_HD);
    if (!(attr_done & dev_bit)) {
        VB_CUDA_OK(cudaFuncSetAttribute(k_attn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_attn_tc_t, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES));
        attr_done |= dev_bit;
    }
    dim3 grid((seq_q + FA_BQ - 1) / FA_BQ, n_heads);
    if (attn_p_in_tmem())
        k_attn_tc_t<<<grid, FA_THREADS, FA_SMEM_BYTES, e->stream>>>(tmQ, tmK, tmV, out, ldo, seq_q, seq_k, cols, scale, window, q_offset, oplanes);
    else
        k_attn_tc<<<grid, FA_THREADS, FA_SMEM_BYTES, e->stream>>>(tmQ, tmK, tmV, out, ldo, seq_q, seq_k, cols, scale, window, q_offset, oplanes);
    VB_CUDA_OK(cudaGetLastError());
    vb_launch_count(e, 2);
}

/* the plane buffers of a call: the same sizes wherever they are requested, so a producer's rows survive until the consumer */
uint16_t *vb_attn_tc_qplanes(VbEngine *e, int seq_q, int n_heads) {
    return (uint16_t *)vb_ws(e, VB_WS_ATT_QP, (size_t)3 * seq_q * n_heads * FA_HD * 2 + 256);
}
uint16_t *vb_attn_tc_kplanes(VbEngine *e, int seq_k, int n_heads) {
    return (uint16_t *)vb_ws(e, VB_WS_ATT_KP, (size_t)3 * seq_k * n_heads * FA_HD * 2 + 256);
}

void vb_attention_tc(VbEngine *e, float *out, int ldo, const float *Q, int ldq, const float *K, const float *V, int ldkv,
                     int seq_q, int seq_k, int n_heads, float scale, int window, int q_offset, uint16_t *oplanes) {
    const int cols = n_heads * FA_HD;
    uint16_t *qp = vb_attn_tc_qplanes(e, seq_q, n_heads);
    uint16_t *kp = vb_attn_tc_kplanes(e, seq_k, n_heads);
    vb_tc_split_planes(e, Q, ldq, seq_q, cols, 3, qp);
    vb_tc_split_planes(e, K, ldkv, seq_k, cols, 3, kp);
    vb_launch_count(e, 2);
    attn_tc_launch(e, out, ldo, qp, kp, V, ldkv, seq_q, seq_k, n_heads, scale, window, q_offset, oplanes);
}

/* Q planes and the K planes of the call's own rows [q_offset, seq_k) were written by the wq|wk|wv epilogue (vb_gemm_tc_qkv_rope);
 * the rows before them (cache tail / halo of a sharded run, f32 in K) are split here. */
void vb_attention_tc_pre(VbEngine *e, float *out, int ldo, const float *K, const float *V, int ldkv,
                         int seq_q, int seq_k, int n_heads, float scale, int window, int q_offset, uint16_t *oplanes) {
    const int cols = n_heads * FA_HD;
    uint16_t *qp = vb_attn_tc_qplanes(e, seq_q, n_heads);
    uint16_t *kp = vb_attn_tc_kplanes(e, seq_k, n_heads);
    if (q_offset > 0) {
        vb_tc_split_planes_strided(e, K, ldkv, q_offset, cols, 3, kp, (size_t)seq_k * cols);
        vb_launch_count(e, 1);
    }
    attn_tc_launch(e, out, ldo, qp, kp, V, ldkv, seq_q, seq_k, n_heads, scale, window, q_offset, oplanes);
}
