/*
 * vb_gemm_tc.cu -- M>1 linears on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
 *
 * Computes C[M,N] = A[M,K] * W[N,K]^T (+bias, epilogue) for the encoder, adapter, conv stem and decoder
 * prefill, i.e. reference vox_linear*_bf16 with seq_len > 1 (voxtral_kernels.c:197-264), where the
 * reference widens the bf16 weights to f32 and calls cblas_sgemm on f32 activations.
 *
 * Numerics: the weights are exact bf16.  The f32 activations are split into bf16 planes
 * x = hi + lo (+ lo2), hi = bf16(x), lo = bf16(x - hi), ...; each plane x bf16 weight product is exact in the
 * f32 accumulator, so two planes carry 16 mantissa bits (relative 2^-17 per element) and three planes carry all
 * 24 -- the default, since the third plane costs ~4% of the encoder time and makes the products f32-exact.  All planes accumulate into the SAME TMEM tile, so
 * a split GEMM is simply a K loop that is `nsplit` times longer.
 *
 * Two kernels, same warp roles (192 threads: warp 0 TMA producer, warp 1 TMEM allocation + MMA issue, warps 2-5 epilogue:
 * tcgen05.ld 32 lanes x 32 columns -> registers -> epilogue -> global):
 *   k_gemm_tc   one 128 x 128 output tile per CTA, 4-stage ring of (A plane tile, W tile), planes one after the other
 *               (round 1; calls with fewer than 512 rows)
 *   k_gemm_tc2  persistent, 128 x 256 tiles, a stage = W tile + the A tiles of all planes, accumulator double-buffered in
 *               TMEM, epilogues that write the next kernel's operands (see the comment above it; profiles/r02_encoder.md)
 * Entry points: vb_gemm_tc (f32 activations: splits, then) -> vb_gemm_tc_planes (pre-split activations); vb_gemm_tc_qkv_rope
 * (the encoder's wq|wk|wv with bias + RoPE + K/V append + Q/K planes as its epilogue).
 */
#include "vb_tc.cuh"
#include <string.h>

#define TC_BM 128
#define TC_BN 128
#define TC_BK 64
#define TC_STAGES 4
#define TC_THREADS 192
#define TC_A_BYTES (TC_BM * TC_BK * 2)
#define TC_B_BYTES (TC_BN * TC_BK * 2)
#define TC_STAGE_BYTES (TC_A_BYTES + TC_B_BYTES)
#define TC_SMEM_BYTES (TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/)

/* ------------------------------------------------------------------ activation split: f32 [M,lda] -> bf16 planes [nsplit][M][K] */
__global__ void k_split_planes(const float *__restrict__ A, int lda, int M, int K, int nsplit, uint16_t *__restrict__ planes, size_t plane) {
    long long idx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (idx >= (long long)M * K) return;
    int m = (int)(idx / K), k = (int)(idx % K);                       /* K % 4 == 0 */
    const float4 v = *reinterpret_cast<const float4 *>(A + (size_t)m * lda + k);
    float x[4] = { v.x, v.y, v.z, v.w };
    uint16_t out[3][4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float r = x[j];
#pragma unroll
        for (int p = 0; p < 3; p++) {
            uint32_t h = f2bf_rne(r);
            out[p][j] = (uint16_t)h;
            r -= __uint_as_float(h << 16);                            /* exact: r and hi share the leading bits */
        }
    }
    for (int p = 0; p < nsplit; p++) {
        uint2 w;
        w.x = (uint32_t)out[p][0] | ((uint32_t)out[p][1] << 16);
        w.y = (uint32_t)out[p][2] | ((uint32_t)out[p][3] << 16);
        *reinterpret_cast<uint2 *>(planes + p * plane + (size_t)m * K + k) = w;
    }
}

void vb_tc_split_planes_strided(VbEngine *e, const float *A, int lda, int M, int K, int nsplit, uint16_t *planes, size_t plane_elems) {
    if (M <= 0) return;
    long long quads = ((long long)M * K + 3) / 4;
    k_split_planes<<<(int)((quads + 255) / 256), 256, 0, e->stream>>>(A, lda, M, K, nsplit, planes, plane_elems);
    VB_CUDA_OK(cudaGetLastError());
}
void vb_tc_split_planes(VbEngine *e, const float *A, int lda, int M, int K, int nsplit, uint16_t *planes) {
    vb_tc_split_planes_strided(e, A, lda, M, K, nsplit, planes, (size_t)M * K);
}

/* ------------------------------------------------------------------ the GEMM */
template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
          const float *__restrict__ bias, float *__restrict__ C, int ldc, int M, int N, int K, int nsplit) {
    extern __shared__ uint8_t tc_smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(tiles + TC_STAGES * TC_STAGE_BYTES);
    uint64_t *empty = full + TC_STAGES;
    uint64_t *tmem_full = empty + TC_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;
    const int kblocks = K / TC_BK, iters = kblocks * nsplit;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; i++) { tc_mbar_init(&full[i], 1); tc_mbar_init(&empty[i], 1); }
        tc_mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                                   /* TMEM: 128 columns of f32 accumulators */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(tmem_slot)), "n"(TC_BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < iters; it++) {
                const int s = it % TC_STAGES, p = it / kblocks, kb = it % kblocks;
                tc_mbar_wait(&empty[s], ((it / TC_STAGES) & 1) ^ 1);
                tc_mbar_expect(&full[s], TC_STAGE_BYTES);
                uint8_t *a = tiles + s * TC_STAGE_BYTES, *b = a + TC_A_BYTES;
                tc_tma_load_2d(a, &tmA, kb * TC_BK, p * M + m0, &full[s]);   /* planes are stacked along the row axis */
                tc_tma_load_2d(b, &tmW, kb * TC_BK, n0, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc_idesc(TC_BM, TC_BN);
            for (int it = 0; it < iters; it++) {
                const int s = it % TC_STAGES;
                tc_mbar_wait(&full[s], (it / TC_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = s32(tiles + s * TC_STAGE_BYTES), b_addr = a_addr + TC_A_BYTES;
#pragma unroll
                for (int k = 0; k < TC_BK / 16; k++) {
                    uint64_t ad = tc_smem_desc(a_addr + k * 32), bd = tc_smem_desc(b_addr + k * 32);
                    tc_umma_bf16(tmem_base, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
                tc_umma_commit(&empty[s]);                              /* frees the stage when these MMAs retire */
            }
            tc_umma_commit(tmem_full);                                  /* accumulator complete */
        }
    } else {
        /* epilogue warps 2..5: a warp may only touch TMEM lanes [32*(warp%4), +32) */
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        tc_mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < TC_BN; c0 += 32) {
            uint32_t r[32];
            tc_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            if (row < M) {
                if (EPI == VB_EPI_SWIGLU) {
                    float *dst = C + (size_t)row * ldc + ((n0 + c0) >> 1);
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float4 o;
                        o.x = vb_silu(__uint_as_float(r[j + 0])) * __uint_as_float(r[j + 1]);
                        o.y = vb_silu(__uint_as_float(r[j + 2])) * __uint_as_float(r[j + 3]);
                        o.z = vb_silu(__uint_as_float(r[j + 4])) * __uint_as_float(r[j + 5]);
                        o.w = vb_silu(__uint_as_float(r[j + 6])) * __uint_as_float(r[j + 7]);
                        *reinterpret_cast<float4 *>(dst + (j >> 1)) = o;
                    }
                } else {
                    float *dst = C + (size_t)row * ldc + n0 + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                        if (bias) {
                            const float4 bv = *reinterpret_cast<const float4 *>(bias + n0 + c0 + j);
                            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                        }
                        if (EPI == VB_EPI_GELU) { o.x = vb_gelu_tanh(o.x); o.y = vb_gelu_tanh(o.y); o.z = vb_gelu_tanh(o.z); o.w = vb_gelu_tanh(o.w); }
                        if (EPI == VB_EPI_RESIDUAL) {
                            const float4 cv = *reinterpret_cast<const float4 *>(dst + j);
                            o.x += cv.x; o.y += cv.y; o.z += cv.z; o.w += cv.w;
                        }
                        *reinterpret_cast<float4 *>(dst + j) = o;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(TC_BN));
    }
}

/* ------------------------------------------------------------------ the persistent GEMM (the default when N % 256 == 0)
 *
 * k_gemm_tc above is bound by the L2 -> shared-memory feed, not by the tensor pipe: a 128 x 128 tile re-reads the W tile once
 * per activation plane and needs 32 KB per 256 MMA cycles = 128 B/cycle/SM, three times what the L2 delivers to 148 SMs
 * (~42 B/cycle/SM, B300_MICROARCH.md "LTS throughput cap"); it measured 36-44 % tensor-pipe activity (profiles/r01_encoder.md).
 * Here:
 *   - tile 128 x 256, and one pipeline stage = the W tile (256 rows x 64 k, 32 KB) + the A tiles of ALL planes (3 x 16 KB):
 *     the W tile is read once for the three planes: 80 KB per 1536 MMA cycles = 53 B/cycle/SM;
 *   - persistent CTAs (one per SM) walk the tiles n-fastest, so the TMA of the next tile's first stages overlaps the tail
 *     of the current one, and barrier init / TMEM allocation are paid once;
 *   - the accumulator is double-buffered in TMEM (2 x 256 columns): the epilogue warps drain tile i while the MMA thread
 *     is already accumulating tile i+1.
 */
#define T2_BN 256
/* Calls with fewer rows stay on k_gemm_tc: with <= 4 row tiles there is nothing to pipeline across.  The two kernels add the
 * plane products in a different order (here k-block-major, there plane-major), so their results differ at the level of the
 * tensor core's f32 accumulation (~1e-5 of the row scale) -- both inside the stated tolerance, but not bit-identical.  Measured
 * consequence (gpurun_out/r02c_stream_*.log): with the persistent kernel on the 1-s feeds of the 172-s continuous fixture, the
 * near-tie at step 1864 (reference top-2 margin 2.7e-5) decodes to the other id; with the plane-major order all 2154 ids match.
 * The persistent kernel therefore takes the long one-shot / sharded calls, where its throughput matters. */
#define T2_MIN_M 512
#define T2_STAGES 2
#define T2_W_BYTES (T2_BN * TC_BK * 2)                      /* 32 KB */
#define T2_STAGE_BYTES (T2_W_BYTES + 3 * TC_A_BYTES)        /* 80 KB */
#define T2_SMEM_BYTES (T2_STAGES * T2_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/)

/* What an epilogue of the persistent kernel writes besides / instead of C (fused producers of the long encoder calls). */
struct VbTcSink {
    uint16_t *oplanes;              /* SWIGLU: SiLU(g)*u as [3][M][N/2] bf16 planes (the A operand of w2) instead of C */
    /* QKV_ROPE (encoder wq|wk|wv, head_dim 64, 2048 columns per block): bias, RoPE, then
     *   q -> qplanes [3][M][2048];  k -> kdst f32 rows (dst_row0 + m) AND kplanes [3][seq_k][2048] rows (dst_row0 + m);  v -> vdst f32 */
    const float2 *rope;             /* [M][32] (cos, sin) of angle (float)(pos0 + m) * inv_freq[d] */
    uint16_t *qplanes, *kplanes;
    float *kdst, *vdst;
    int dst_row0, seq_k;
};
#define VB_EPI_QKV_ROPE 4

template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tc2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
           const float *__restrict__ bias, float *__restrict__ C, int ldc, int M, int N, int K, int nsplit,
           const VbTcSink sink) {
    uint16_t *__restrict__ oplanes = sink.oplanes;
    extern __shared__ uint8_t tc_smem_raw[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(tiles + T2_STAGES * T2_STAGE_BYTES);
    uint64_t *empty = full + T2_STAGES;
    uint64_t *acc_full = empty + T2_STAGES;                  /* [2] accumulator complete (tcgen05.commit) */
    uint64_t *acc_empty = acc_full + 2;                      /* [2] accumulator drained (128 epilogue threads) */
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = K / TC_BK;
    const int tiles_n = N / T2_BN, tiles_m = (M + TC_BM - 1) / TC_BM;
    const int total = tiles_n * tiles_m;

    if (threadIdx.x == 0) {
        for (int i = 0; i < T2_STAGES; i++) { tc_mbar_init(&full[i], 1); tc_mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; i++) { tc_mbar_init(&acc_full[i], 1); tc_mbar_init(&acc_empty[i], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                         /* the whole TMEM: 2 accumulators of 256 f32 columns */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t stage_tx = T2_W_BYTES + (uint32_t)nsplit * TC_A_BYTES;
            int it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int m0 = (t / tiles_n) * TC_BM, n0 = (t % tiles_n) * T2_BN;
                for (int kb = 0; kb < kblocks; kb++, it++) {
                    const int s = it % T2_STAGES;
                    tc_mbar_wait(&empty[s], ((it / T2_STAGES) & 1) ^ 1);
                    tc_mbar_expect(&full[s], stage_tx);
                    uint8_t *w = tiles + s * T2_STAGE_BYTES, *a = w + T2_W_BYTES;
                    tc_tma_load_2d(w, &tmW, kb * TC_BK, n0, &full[s]);
                    for (int p = 0; p < nsplit; p++)                      /* planes are stacked along the row axis */
                        tc_tma_load_2d(a + p * TC_A_BYTES, &tmA, kb * TC_BK, p * M + m0, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        /* the whole warp walks the schedule converged; one elected lane issues (vb_tc.cuh: tc_elect_one) */
        const uint32_t idesc = tc_idesc(TC_BM, T2_BN);
        const uint64_t d0 = tc_smem_desc(s32(tiles));                   /* + (byte offset >> 4) in the start-address field */
        int it = 0, tl = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, tl++) {
            const int ab = tl & 1;
            tc_mbar_wait(&acc_empty[ab], ((tl >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)ab * T2_BN;
            for (int kb = 0; kb < kblocks; kb++, it++) {
                const int s = it % T2_STAGES;
                tc_mbar_wait(&full[s], (it / T2_STAGES) & 1);
                tc_fence_after();
                if (tc_elect_one()) {
                    const uint64_t dw = d0 + (uint64_t)(s * (T2_STAGE_BYTES >> 4)), da = dw + (uint64_t)(T2_W_BYTES >> 4);
                    for (int p = 0; p < nsplit; p++) {
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; k++)
                            tc_umma_bf16(d, da + (uint64_t)(p * (TC_A_BYTES >> 4) + 2 * k), dw + (uint64_t)(2 * k), idesc,
                                         (kb > 0 || p > 0 || k > 0) ? 1u : 0u);
                    }
                    tc_umma_commit(&empty[s]);                          /* frees the stage when these MMAs retire */
                    if (kb == kblocks - 1) tc_umma_commit(&acc_full[ab]);   /* accumulator complete */
                }
                __syncwarp();
            }
        }
    } else {
        /* epilogue warps 2..5: a warp may only touch TMEM lanes [32*(warp%4), +32) */
        const int q = warp & 3;
        int tl = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, tl++) {
            const int ab = tl & 1;
            const int m0 = (t / tiles_n) * TC_BM, n0 = (t % tiles_n) * T2_BN;
            const int row = m0 + q * 32 + lane;
            tc_mbar_wait(&acc_full[ab], (tl >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < T2_BN; c0 += 32) {
                uint32_t r[32];
                tc_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * T2_BN + c0), r);
                if (row < M) {
                    if (EPI == VB_EPI_QKV_ROPE) {
                        /* voxtral_encoder.c:533-553: q/k/v = x W^T + b, RoPE on q and k, k and v appended to the cache.  32 columns =
                         * half a head = 16 (even, odd) pairs; arithmetic as k_rope_split: y0 = fma(x0, c, -(x1 s)), y1 = fma(x0, s, x1 c) */
                        const int n = n0 + c0, region = n >> 11, c = n & 2047;
                        float x[32];
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bv = *reinterpret_cast<const float4 *>(bias + n + j);
                            x[j] = __uint_as_float(r[j]) + bv.x; x[j + 1] = __uint_as_float(r[j + 1]) + bv.y;
                            x[j + 2] = __uint_as_float(r[j + 2]) + bv.z; x[j + 3] = __uint_as_float(r[j + 3]) + bv.w;
                        }
                        if (region == 2) {
                            float *dst = sink.vdst + (size_t)(sink.dst_row0 + row) * 2048 + c;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
                        } else {
                            const float2 *t = sink.rope + (size_t)row * 32 + ((c & 63) >> 1);
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                const float2 cs = t[j];
                                const float x0 = x[2 * j], x1 = x[2 * j + 1];
                                x[2 * j]     = __fmaf_rn(x0, cs.x, -__fmul_rn(x1, cs.y));
                                x[2 * j + 1] = __fmaf_rn(x0, cs.y, __fmul_rn(x1, cs.x));
                            }
                            uint16_t *pd; size_t plane;
                            if (region == 1) {
                                float *dst = sink.kdst + (size_t)(sink.dst_row0 + row) * 2048 + c;
#pragma unroll
                                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
                                pd = sink.kplanes + (size_t)(sink.dst_row0 + row) * 2048 + c; plane = (size_t)sink.seq_k * 2048;
                            } else {
                                pd = sink.qplanes + (size_t)row * 2048 + c; plane = (size_t)M * 2048;
                            }
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint32_t w0[4], w1[4], w2[4];
#pragma unroll
                                for (int e2 = 0; e2 < 4; e2++) {
                                    float a = x[j + 2 * e2], b2 = x[j + 2 * e2 + 1];
                                    w0[e2] = tc_pack_bf16x2(a, b2);
                                    a -= vb_bf16_lo(w0[e2]); b2 -= vb_bf16_hi(w0[e2]);
                                    w1[e2] = tc_pack_bf16x2(a, b2);
                                    a -= vb_bf16_lo(w1[e2]); b2 -= vb_bf16_hi(w1[e2]);
                                    w2[e2] = tc_pack_bf16x2(a, b2);
                                }
                                *reinterpret_cast<uint4 *>(pd + j) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
                                *reinterpret_cast<uint4 *>(pd + j + plane) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
                                *reinterpret_cast<uint4 *>(pd + j + 2 * plane) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
                            }
                        }
                    } else if (EPI == VB_EPI_SWIGLU && oplanes != nullptr) {
                        /* 16 outputs of this row -> three bf16 planes, 32 B each */
                        const size_t nh = (size_t)(N >> 1);
                        uint16_t *dst = oplanes + (size_t)row * nh + ((n0 + c0) >> 1);
                        uint32_t w0[8], w1[8], w2[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float x = vb_silu(__uint_as_float(r[4 * j + 0])) * __uint_as_float(r[4 * j + 1]);
                            float y = vb_silu(__uint_as_float(r[4 * j + 2])) * __uint_as_float(r[4 * j + 3]);
                            w0[j] = tc_pack_bf16x2(x, y);
                            x -= vb_bf16_lo(w0[j]); y -= vb_bf16_hi(w0[j]);
                            w1[j] = tc_pack_bf16x2(x, y);
                            x -= vb_bf16_lo(w1[j]); y -= vb_bf16_hi(w1[j]);
                            w2[j] = tc_pack_bf16x2(x, y);
                        }
                        const size_t plane = (size_t)M * nh;
                        *reinterpret_cast<uint4 *>(dst) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
                        *reinterpret_cast<uint4 *>(dst + 8) = make_uint4(w0[4], w0[5], w0[6], w0[7]);
                        *reinterpret_cast<uint4 *>(dst + plane) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
                        *reinterpret_cast<uint4 *>(dst + plane + 8) = make_uint4(w1[4], w1[5], w1[6], w1[7]);
                        *reinterpret_cast<uint4 *>(dst + 2 * plane) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
                        *reinterpret_cast<uint4 *>(dst + 2 * plane + 8) = make_uint4(w2[4], w2[5], w2[6], w2[7]);
                    } else if (EPI == VB_EPI_SWIGLU) {
                        float *dst = C + (size_t)row * ldc + ((n0 + c0) >> 1);
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            float4 o;
                            o.x = vb_silu(__uint_as_float(r[j + 0])) * __uint_as_float(r[j + 1]);
                            o.y = vb_silu(__uint_as_float(r[j + 2])) * __uint_as_float(r[j + 3]);
                            o.z = vb_silu(__uint_as_float(r[j + 4])) * __uint_as_float(r[j + 5]);
                            o.w = vb_silu(__uint_as_float(r[j + 6])) * __uint_as_float(r[j + 7]);
                            *reinterpret_cast<float4 *>(dst + (j >> 1)) = o;
                        }
                    } else {
                        float *dst = C + (size_t)row * ldc + n0 + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 o = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                            if (bias) {
                                const float4 bv = *reinterpret_cast<const float4 *>(bias + n0 + c0 + j);
                                o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                            }
                            if (EPI == VB_EPI_GELU) { o.x = vb_gelu_tanh(o.x); o.y = vb_gelu_tanh(o.y); o.z = vb_gelu_tanh(o.z); o.w = vb_gelu_tanh(o.w); }
                            if (EPI == VB_EPI_RESIDUAL) {
                                const float4 cv = *reinterpret_cast<const float4 *>(dst + j);
                                o.x += cv.x; o.y += cv.y; o.z += cv.z; o.w += cv.w;
                            }
                            *reinterpret_cast<float4 *>(dst + j) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            tc_mbar_arrive(&acc_empty[ab]);                              /* 128 arrivals: this accumulator may be overwritten */
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(512));
    }
}

/* ------------------------------------------------------------------ host */
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = NULL;

void vb_tc_make_map(CUtensorMap *map, const void *base, uint64_t inner_elems, uint64_t rows, uint64_t row_pitch_bytes, uint32_t box_rows) {
    if (!g_encode) {
        cudaDriverEntryPointQueryResult q;
        void *fn = NULL;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
            VB_FAIL("cuTensorMapEncodeTiled unavailable (driver too old for TMA)");
        }
        g_encode = (PFN_encodeTiled)fn;
    }
    cuuint64_t dims[2] = { inner_elems, rows }, strides[1] = { row_pitch_bytes };
    cuuint32_t box[2] = { 64, box_rows }, estr[2] = { 1, 1 };
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "voxtral_b200: cuTensorMapEncodeTiled failed (%d)\n", (int)r); VB_FAIL("cuTensorMapEncodeTiled failed"); }
}

int vb_gemm_tc_usable(int M, int N, int K) {
    return M >= 1 && (N % TC_BN) == 0 && (K % TC_BK) == 0;
}

int vb_gemm_nsplit(void) {
    static int n = 0;
    if (!n) { const char *s = getenv("VOX_CUDA_GEMM_SPLIT"); n = s ? atoi(s) : 3; if (n < 1) n = 1; if (n > 3) n = 3; }
    return n;
}

/* 2 = persistent 128 x 256 kernel (default), 1 = one 128 x 128 tile per CTA (VOX_CUDA_GEMM=v1) */
static int gemm_variant(void) {
    static int v = 0;
    if (!v) { const char *s = getenv("VOX_CUDA_GEMM"); v = (s && (s[0] == '1' || (s[0] == 'v' && s[1] == '1'))) ? 1 : 2; }
    return v;
}

/* Fused producers (RMSNorm, attention and SwiGLU epilogues writing bf16 planes directly) are used for the calls that take
 * the persistent kernel; VOX_CUDA_FUSE=0 keeps the separate k_split_planes passes (A/B, validation). */
int vb_gemm_tc_fused_ok(int M) {
    const char *s = getenv("VOX_CUDA_FUSE");                          /* read per call: a test flips it between two passes */
    const char *g = getenv("VOX_CUDA_GEMM");
    return !(s && s[0] == '0') && gemm_variant() == 2 && !(g && !strcmp(g, "simt")) && M >= T2_MIN_M;
}

/* The GEMM on activations that are already split: planes [3][M][K] bf16 (plane p, row m at (p*M + m)*K).
 * oplanes (SWIGLU, persistent kernel only): the result is written as [3][M][N/2] bf16 planes instead of C. */
void vb_gemm_tc_planes(VbEngine *e, const uint16_t *planes, const uint16_t *W, const float *bias, float *C, int ldc,
                       int M, int N, int K, int epi, uint16_t *oplanes) {
    static unsigned int attr_done = 0;                              /* one bit per device: function attributes are per device */
    const unsigned int dev_bit = 1u << (e->device & 31);
    const int nsplit = vb_gemm_nsplit();
    if (!(attr_done & dev_bit)) {
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<VB_EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<VB_EPI_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<VB_EPI_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<VB_EPI_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc2<VB_EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc2<VB_EPI_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc2<VB_EPI_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc2<VB_EPI_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc2<VB_EPI_QKV_ROPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES));
        attr_done |= dev_bit;
    }
    CUtensorMap tmA, tmW;
    vb_tc_make_map(&tmA, planes, (uint64_t)K, (uint64_t)nsplit * M, (uint64_t)K * 2, TC_BM);
    dim3 block(TC_THREADS);
    VbTcSink sink;
    memset(&sink, 0, sizeof sink);
    sink.oplanes = oplanes;
    if (gemm_variant() == 2 && (N % T2_BN) == 0 && M >= T2_MIN_M) {
        vb_tc_make_map(&tmW, W, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, T2_BN);
        const int total = (N / T2_BN) * ((M + TC_BM - 1) / TC_BM);
        dim3 grid(total < e->sm_count ? total : e->sm_count);
        switch (epi) {
        case VB_EPI_STORE:    k_gemm_tc2<VB_EPI_STORE><<<grid, block, T2_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit, sink); break;
        case VB_EPI_GELU:     k_gemm_tc2<VB_EPI_GELU><<<grid, block, T2_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit, sink); break;
        case VB_EPI_RESIDUAL: k_gemm_tc2<VB_EPI_RESIDUAL><<<grid, block, T2_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit, sink); break;
        case VB_EPI_SWIGLU:   k_gemm_tc2<VB_EPI_SWIGLU><<<grid, block, T2_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit, sink); break;
        }
    } else {
        if (oplanes) VB_FAIL("vb_gemm_tc_planes: plane output needs the persistent kernel");
        vb_tc_make_map(&tmW, W, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, TC_BN);
        dim3 grid(N / TC_BN, (M + TC_BM - 1) / TC_BM);
        switch (epi) {
        case VB_EPI_STORE:    k_gemm_tc<VB_EPI_STORE><<<grid, block, TC_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit); break;
        case VB_EPI_GELU:     k_gemm_tc<VB_EPI_GELU><<<grid, block, TC_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit); break;
        case VB_EPI_RESIDUAL: k_gemm_tc<VB_EPI_RESIDUAL><<<grid, block, TC_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit); break;
        case VB_EPI_SWIGLU:   k_gemm_tc<VB_EPI_SWIGLU><<<grid, block, TC_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit); break;
        }
    }
    VB_CUDA_OK(cudaGetLastError());
    vb_launch_count(e, 1);
}

void vb_gemm_tc(VbEngine *e, const float *A, int lda, const uint16_t *W, const float *bias, float *C, int ldc,
                int M, int N, int K, int epi) {
    const int nsplit = vb_gemm_nsplit();
    uint16_t *planes = (uint16_t *)vb_ws(e, VB_WS_GEMM_PLANES, (size_t)nsplit * M * K * 2 + 256);
    vb_tc_split_planes(e, A, lda, M, K, nsplit, planes);
    vb_launch_count(e, 1);
    vb_gemm_tc_planes(e, planes, W, bias, C, ldc, M, N, K, epi, nullptr);
}

/* ------------------------------------------------------------------ encoder wq|wk|wv with RoPE + K/V append + Q/K planes as the epilogue */
__global__ void k_rope_table(float2 *__restrict__ t, const float *__restrict__ inv_freq, int M, int half, int pos0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * half) return;
    const int m = i / half, d = i % half;
    float sn, cs;
    sincosf((float)(pos0 + m) * inv_freq[d], &sn, &cs);                /* the angle and the call of k_rope_split (vb_ops.cu) */
    t[i] = make_float2(cs, sn);
}

int vb_gemm_tc_qkv_ok(int M) {
    const char *s = getenv("VOX_CUDA_FUSE_QKV");
    return vb_gemm_tc_fused_ok(M) && !(s && s[0] == '0');
}

/* planes: RMSNorm output as [3][M][1280] planes.  q planes -> qplanes [3][M][2048]; rotated k -> kdst rows dst_row0.. (f32) and
 * kplanes [3][seq_k][2048] rows dst_row0..; v -> vdst rows dst_row0.. (f32).  Encoder geometry only (32 heads x 64). */
void vb_gemm_tc_qkv_rope(VbEngine *e, const uint16_t *planes, const uint16_t *W, const float *bias, int M, int K,
                         const float *inv_freq, int pos0, uint16_t *qplanes, uint16_t *kplanes, float *kdst, float *vdst,
                         int dst_row0, int seq_k) {
    const int N = 3 * 2048, nsplit = vb_gemm_nsplit();
    if (!bias || M < T2_MIN_M) VB_FAIL("vb_gemm_tc_qkv_rope: needs a bias and a long call");
    float2 *table = (float2 *)vb_ws(e, VB_WS_ENC_ROPE, (size_t)M * 32 * sizeof(float2) + 256);
    k_rope_table<<<(M * 32 + 255) / 256, 256, 0, e->stream>>>(table, inv_freq, M, 32, pos0);
    VB_CUDA_OK(cudaGetLastError());
    static unsigned int attr_done = 0;
    const unsigned int dev_bit = 1u << (e->device & 31);
    if (!(attr_done & dev_bit)) {
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc2<VB_EPI_QKV_ROPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES));
        attr_done |= dev_bit;
    }
    CUtensorMap tmA, tmW;
    vb_tc_make_map(&tmA, planes, (uint64_t)K, (uint64_t)nsplit * M, (uint64_t)K * 2, TC_BM);
    vb_tc_make_map(&tmW, W, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, T2_BN);
    VbTcSink sink;
    memset(&sink, 0, sizeof sink);
    sink.rope = table; sink.qplanes = qplanes; sink.kplanes = kplanes; sink.kdst = kdst; sink.vdst = vdst;
    sink.dst_row0 = dst_row0; sink.seq_k = seq_k;
    const int total = (N / T2_BN) * ((M + TC_BM - 1) / TC_BM);
    dim3 grid(total < e->sm_count ? total : e->sm_count), block(TC_THREADS);
    k_gemm_tc2<VB_EPI_QKV_ROPE><<<grid, block, T2_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, nullptr, 0, M, N, K, nsplit, sink);
    VB_CUDA_OK(cudaGetLastError());
    vb_launch_count(e, 2);
}
