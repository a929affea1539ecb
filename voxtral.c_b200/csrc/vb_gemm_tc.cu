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
 * Kernel anatomy (one 128 x 128 output tile per CTA, 192 threads):
 *   warp 0    TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B boxes of 128 rows x 64 bf16) for the A plane
 *             tile and the W tile into a 4-stage shared-memory ring, completion on mbarriers (expect_tx)
 *   warp 1    TMEM allocation + single-thread tcgen05.mma.cta_group::1.kind::f16 issue (M=128, N=128, K=16 x4
 *             per stage), tcgen05.commit releases the stage / publishes the accumulator
 *   warps 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> bias / GELU / residual / SiLU(g)*u
 *             -> global f32
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
__global__ void k_split_planes(const float *__restrict__ A, int lda, int M, int K, int nsplit, uint16_t *__restrict__ planes) {
    long long idx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (idx >= (long long)M * K) return;
    int m = (int)(idx / K), k = (int)(idx % K);                       /* K % 4 == 0 */
    const float4 v = *reinterpret_cast<const float4 *>(A + (size_t)m * lda + k);
    float x[4] = { v.x, v.y, v.z, v.w };
    const size_t plane = (size_t)M * K;
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

void vb_tc_split_planes(VbEngine *e, const float *A, int lda, int M, int K, int nsplit, uint16_t *planes) {
    long long quads = ((long long)M * K + 3) / 4;
    k_split_planes<<<(int)((quads + 255) / 256), 256, 0, e->stream>>>(A, lda, M, K, nsplit, planes);
    VB_CUDA_OK(cudaGetLastError());
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
#define T2_STAGES 2
#define T2_W_BYTES (T2_BN * TC_BK * 2)                      /* 32 KB */
#define T2_STAGE_BYTES (T2_W_BYTES + 3 * TC_A_BYTES)        /* 80 KB */
#define T2_SMEM_BYTES (T2_STAGES * T2_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/)

template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tc2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
           const float *__restrict__ bias, float *__restrict__ C, int ldc, int M, int N, int K, int nsplit) {
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
        if (lane == 0) {
            const uint32_t idesc = tc_idesc(TC_BM, T2_BN);
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
                    const uint32_t w_addr = s32(tiles + s * T2_STAGE_BYTES), a_addr = w_addr + T2_W_BYTES;
                    for (int p = 0; p < nsplit; p++) {
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; k++)
                            tc_umma_bf16(d, tc_smem_desc(a_addr + p * TC_A_BYTES + k * 32), tc_smem_desc(w_addr + k * 32), idesc,
                                         (kb > 0 || p > 0 || k > 0) ? 1u : 0u);
                    }
                    tc_umma_commit(&empty[s]);                          /* frees the stage when these MMAs retire */
                }
                tc_umma_commit(&acc_full[ab]);                          /* accumulator complete */
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

void vb_gemm_tc(VbEngine *e, const float *A, int lda, const uint16_t *W, const float *bias, float *C, int ldc,
                int M, int N, int K, int epi) {
    static unsigned int attr_done = 0;                              /* one bit per device: function attributes are per device */
    const unsigned int dev_bit = 1u << (e->device & 31);
    const int nsplit = vb_gemm_nsplit();
    uint16_t *planes = (uint16_t *)vb_ws(e, VB_WS_GEMM_PLANES, (size_t)nsplit * M * K * 2 + 256);
    vb_tc_split_planes(e, A, lda, M, K, nsplit, planes);
    if (!(attr_done & dev_bit)) {
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<VB_EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<VB_EPI_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<VB_EPI_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<VB_EPI_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc2<VB_EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc2<VB_EPI_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc2<VB_EPI_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc2<VB_EPI_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES));
        attr_done |= dev_bit;
    }
    CUtensorMap tmA, tmW;
    vb_tc_make_map(&tmA, planes, (uint64_t)K, (uint64_t)nsplit * M, (uint64_t)K * 2, TC_BM);
    dim3 block(TC_THREADS);
    if (gemm_variant() == 2 && (N % T2_BN) == 0) {
        vb_tc_make_map(&tmW, W, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, T2_BN);
        const int total = (N / T2_BN) * ((M + TC_BM - 1) / TC_BM);
        dim3 grid(total < e->sm_count ? total : e->sm_count);
        switch (epi) {
        case VB_EPI_STORE:    k_gemm_tc2<VB_EPI_STORE><<<grid, block, T2_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit); break;
        case VB_EPI_GELU:     k_gemm_tc2<VB_EPI_GELU><<<grid, block, T2_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit); break;
        case VB_EPI_RESIDUAL: k_gemm_tc2<VB_EPI_RESIDUAL><<<grid, block, T2_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit); break;
        case VB_EPI_SWIGLU:   k_gemm_tc2<VB_EPI_SWIGLU><<<grid, block, T2_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit); break;
        }
    } else {
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
    vb_launch_count(e, 2);
}
