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
 * registers (expf), so the result agrees with the CUDA-core kernels (k_attn_tile64 / k_attn_warp) to f32 rounding.
 *
 * One CTA = 128 queries of one head; keys in blocks of 64.  192 threads:
 *   warp 0     TMA: Q planes once (3 x [128 x 64] bf16), then per key block K planes (3 x [64 keys x 64 d]) into a 3-slot ring
 *              and V^T planes (3 x [64 d x 64 keys]) into a 2-slot ring, all SWIZZLE_128B boxes
 *   warp 1     one thread issues the MMAs: S(j+1) = Q K(j+1)^T is issued BEFORE P(j) V(j), so the tensor pipe works on the next
 *              scores while the softmax warps turn S(j) into P(j); 24 + 24 tcgen05.mma (M128 N64 K16) per key block
 *   warps 2-5  thread = query row: tcgen05.ld S(j) (64 columns), mask, running max / sum, P(j) -> three bf16 planes written
 *              to shared memory in the SWIZZLE_128B K-major layout the MMA reads, then the deferred O = O*alpha + (P V)(j-1)
 *              from TMEM into 64 registers
 * TMEM: S double-buffered (2 x 64 columns), P V block result double-buffered (2 x 64 columns).
 *
 * V^T: the B operand of P V has to be K-major, i.e. [d][key] with keys contiguous; k_vt_planes transposes V while splitting
 * it.  Q and K planes come from the GEMM's splitter (vb_tc_split_planes).
 */
#include "vb_tc.cuh"
#include <stdlib.h>

#define FA_BQ 128
#define FA_BK 64
#define FA_HD 64
#define FA_THREADS 192
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
#define FA_SMEM_BYTES (FA_OFF_BAR + 256 + 1024 /*align*/)
#define FA_TMEM_COLS 256

/* the six plane pairs (a plane, b plane), largest first: (0,0) (0,1) (1,0) (1,1) (0,2) (2,0) */
#define FA_PA(pr) ((0x201100u >> (4 * (pr))) & 0xFu)
#define FA_PB(pr) ((0x021010u >> (4 * (pr))) & 0xFu)

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
          float *__restrict__ out, int ldo, int seq_q, int seq_k, int cols, float scale, int window, int q_offset) {
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
            tc_mbar_init(&s_full[i], 1); tc_mbar_init(&s_empty[i], 128);
            tc_mbar_init(&o_full[i], 1); tc_mbar_init(&o_empty[i], 128);
        }
        tc_mbar_init(p_full, 128);
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
        if (lane == 0 && nb > 0) {
            const uint32_t idesc = tc_idesc(FA_BQ, FA_BK);              /* M128 N64 for both products */
            const uint32_t q_addr = s32(sm + FA_OFF_Q), p_addr = s32(sm + FA_OFF_P);
            tc_mbar_wait(q_full, 0);
            for (int j = -1; j < nb; j++) {
                if (j + 1 < nb) {                                       /* S(j+1) = Q K(j+1)^T */
                    const int jj = j + 1, s = jj % FA_K_SLOTS, b = jj & 1;
                    tc_mbar_wait(&k_full[s], (jj / FA_K_SLOTS) & 1);
                    tc_mbar_wait(&s_empty[b], ((jj >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t k_addr = s32(sm + FA_OFF_K + s * 3 * FA_KTILE);
#pragma unroll
                    for (int pr = 0; pr < 6; pr++) {
                        const uint32_t a = q_addr + FA_PA(pr) * FA_QPLANE, bb = k_addr + FA_PB(pr) * FA_KTILE;
#pragma unroll
                        for (int k = 0; k < FA_HD / 16; k++)
                            tc_umma_bf16(tm_s + b * 64, tc_smem_desc(a + k * 32), tc_smem_desc(bb + k * 32), idesc, (pr | k) ? 1u : 0u);
                    }
                    tc_umma_commit(&k_empty[s]);
                    tc_umma_commit(&s_full[b]);
                }
                if (j >= 0) {                                           /* (P V)(j) */
                    const int s = j & 1, u = j >> 1;
                    tc_mbar_wait(&v_full[s], u & 1);
                    tc_mbar_wait(p_full, j & 1);
                    tc_mbar_wait(&o_empty[s], (u & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t v_addr = s32(sm + FA_OFF_V + s * 3 * FA_VTILE);
#pragma unroll
                    for (int pr = 0; pr < 6; pr++) {
                        const uint32_t a = p_addr + FA_PA(pr) * FA_QPLANE, bb = v_addr + FA_PB(pr) * FA_VTILE;
#pragma unroll
                        for (int k = 0; k < FA_BK / 16; k++)
                            tc_umma_bf16(tm_o + s * 64, tc_smem_desc(a + k * 32), tc_smem_desc(bb + k * 32), idesc, (pr | k) ? 1u : 0u);
                    }
                    tc_umma_commit(&v_empty[s]);
                    tc_umma_commit(&o_full[s]);
                }
            }
        }
    } else {
        /* softmax warps 2..5: a warp may only touch TMEM lanes [32*(warp%4), +32) */
        const int qd = warp & 3;
        const int r = qd * 32 + lane;                                   /* row of the query tile */
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const int g = q_offset + q0 + r;                                /* index of this query in the key buffer */
        int lo = 0;
        if (window > 0 && g - window + 1 > 0) lo = g - window + 1;
        const int hi = min(g, seq_k - 1);
        float o[FA_HD];
#pragma unroll
        for (int d = 0; d < FA_HD; d++) o[d] = 0.f;
        float m = -1e30f, l = 0.f, alpha_prev = 1.f;
        const uint32_t prow = s32(sm + FA_OFF_P) + r * 128;
        const int sw = r & 7;

        for (int j = 0; j < nb; j++) {
            const int b = j & 1;
            const int c_lo = lo - (k_lo + j * FA_BK), c_hi = hi - (k_lo + j * FA_BK);
            uint32_t s0[32], s1[32];
            tc_mbar_wait(&s_full[b], (j >> 1) & 1);
            tc_fence_after();
            tc_tmem_ld32_nowait(tm_s + b * 64 + lane_off, s0);
            tc_tmem_ld32_nowait(tm_s + b * 64 + 32 + lane_off, s1);
            tc_tmem_wait_ld();
            tc_fence_before();
            tc_mbar_arrive(&s_empty[b]);

            float mx = -1e30f;
#pragma unroll
            for (int c = 0; c < 32; c++) {
                float a = __uint_as_float(s0[c]) * scale, bq = __uint_as_float(s1[c]) * scale;
                a = (c >= c_lo && c <= c_hi) ? a : -1e30f;
                bq = (c + 32 >= c_lo && c + 32 <= c_hi) ? bq : -1e30f;
                s0[c] = __float_as_uint(a); s1[c] = __float_as_uint(bq);
                mx = fmaxf(mx, fmaxf(a, bq));
            }
            const float mn = fmaxf(m, mx);
            const float alpha = expf(m - mn);
            float rs = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const float a = __uint_as_float(s0[c]), bq = __uint_as_float(s1[c]);
                const float pa = a > -1e29f ? expf(a - mn) : 0.f;
                const float pb = bq > -1e29f ? expf(bq - mn) : 0.f;
                s0[c] = __float_as_uint(pa); s1[c] = __float_as_uint(pb);
                rs += pa + pb;
            }
            l = l * alpha + rs;
            m = mn;

            if (j >= 1) tc_mbar_wait(&o_full[b ^ 1], ((j - 1) >> 1) & 1);     /* (P V)(j-1) retired: P may be overwritten */
            /* P(j) -> three bf16 planes, rows of 128 B, 16-byte chunk c of row r at chunk (c ^ (r & 7)) (SWIZZLE_128B) */
#pragma unroll
            for (int cc = 0; cc < 8; cc++) {
                uint32_t w0[4], w1[4], w2[4];
#pragma unroll
                for (int e2 = 0; e2 < 4; e2++) {
                    const int c = cc * 8 + e2 * 2;
                    const float x = __uint_as_float(c < 32 ? s0[c & 31] : s1[c & 31]);
                    const float y = __uint_as_float(c + 1 < 32 ? s0[(c + 1) & 31] : s1[(c + 1) & 31]);
                    uint32_t x0, x1, x2, y0, y1, y2;
                    tc_split3(x, x0, x1, x2);
                    tc_split3(y, y0, y1, y2);
                    w0[e2] = x0 | (y0 << 16); w1[e2] = x1 | (y1 << 16); w2[e2] = x2 | (y2 << 16);
                }
                const uint32_t dst = prow + ((cc ^ sw) << 4);
                fa_sts128(dst, w0);
                fa_sts128(dst + FA_QPLANE, w1);
                fa_sts128(dst + 2 * FA_QPLANE, w2);
            }
            tc_fence_proxy_async();
            tc_mbar_arrive(p_full);

            if (j >= 1) {                                               /* O = O * alpha(j-1) + (P V)(j-1) */
                tc_fence_after();
                tc_tmem_ld32_nowait(tm_o + (b ^ 1) * 64 + lane_off, s0);
                tc_tmem_ld32_nowait(tm_o + (b ^ 1) * 64 + 32 + lane_off, s1);
                tc_tmem_wait_ld();
                tc_fence_before();
                tc_mbar_arrive(&o_empty[b ^ 1]);
#pragma unroll
                for (int d = 0; d < 32; d++) {
                    o[d] = fmaf(o[d], alpha_prev, __uint_as_float(s0[d]));
                    o[d + 32] = fmaf(o[d + 32], alpha_prev, __uint_as_float(s1[d]));
                }
            }
            alpha_prev = alpha;
        }
        if (nb > 0) {
            const int b = (nb - 1) & 1;
            uint32_t s0[32], s1[32];
            tc_mbar_wait(&o_full[b], ((nb - 1) >> 1) & 1);
            tc_fence_after();
            tc_tmem_ld32_nowait(tm_o + b * 64 + lane_off, s0);
            tc_tmem_ld32_nowait(tm_o + b * 64 + 32 + lane_off, s1);
            tc_tmem_wait_ld();
#pragma unroll
            for (int d = 0; d < 32; d++) {
                o[d] = fmaf(o[d], alpha_prev, __uint_as_float(s0[d]));
                o[d + 32] = fmaf(o[d + 32], alpha_prev, __uint_as_float(s1[d]));
            }
        }
        if (q0 + r < seq_q) {
            const float inv = l > 0.f ? 1.0f / l : 0.f;
            float *dst = out + (size_t)(q0 + r) * ldo + hoff;
#pragma unroll
            for (int d = 0; d < FA_HD; d += 4)
                *reinterpret_cast<float4 *>(dst + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(FA_TMEM_COLS));
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

void vb_attention_tc(VbEngine *e, float *out, int ldo, const float *Q, int ldq, const float *K, const float *V, int ldkv,
                     int seq_q, int seq_k, int n_heads, float scale, int window, int q_offset) {
    static unsigned int attr_done = 0;                                  /* one bit per device */
    const unsigned int dev_bit = 1u << (e->device & 31);
    const int cols = n_heads * FA_HD;
    const int nk_pad = (seq_k + 63) & ~63;
    uint16_t *qp = (uint16_t *)vb_ws(e, VB_WS_ATT_QP, (size_t)3 * seq_q * cols * 2 + 256);
    uint16_t *kp = (uint16_t *)vb_ws(e, VB_WS_ATT_KP, (size_t)3 * seq_k * cols * 2 + 256);
    uint16_t *vt = (uint16_t *)vb_ws(e, VB_WS_ATT_VT, (size_t)3 * cols * nk_pad * 2 + 256);
    vb_tc_split_planes(e, Q, ldq, seq_q, cols, 3, qp);
    vb_tc_split_planes(e, K, ldkv, seq_k, cols, 3, kp);
    dim3 tg(nk_pad / 64, cols / 32);
    k_vt_planes<<<tg, 256, 0, e->stream>>>(V, ldkv, seq_k, cols, nk_pad, vt);
    VB_CUDA_OK(cudaGetLastError());
    CUtensorMap tmQ, tmK, tmV;
    vb_tc_make_map(&tmQ, qp, (uint64_t)cols, (uint64_t)3 * seq_q, (uint64_t)cols * 2, FA_BQ);
    vb_tc_make_map(&tmK, kp, (uint64_t)cols, (uint64_t)3 * seq_k, (uint64_t)cols * 2, FA_BK);
    vb_tc_make_map(&tmV, vt, (uint64_t)nk_pad, (uint64_t)3 * cols, (uint64_t)nk_pad * 2, FA_HD);
    if (!(attr_done & dev_bit)) {
        VB_CUDA_OK(cudaFuncSetAttribute(k_attn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES));
        attr_done |= dev_bit;
    }
    dim3 grid((seq_q + FA_BQ - 1) / FA_BQ, n_heads);
    k_attn_tc<<<grid, FA_THREADS, FA_SMEM_BYTES, e->stream>>>(tmQ, tmK, tmV, out, ldo, seq_q, seq_k, cols, scale, window, q_offset);
    VB_CUDA_OK(cudaGetLastError());
    vb_launch_count(e, 4);
}
