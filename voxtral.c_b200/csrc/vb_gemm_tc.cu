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
#include "vb_ops.cuh"
#include <cuda.h>
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

/* ------------------------------------------------------------------ PTX */
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void tc_mbar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 4000000000ll) __trap();            /* never hang the GPU on a pipeline bug */
    }
}
__device__ __forceinline__ void tc_tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(s32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tc_umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tc_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

/* K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
 * start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64).
 * Rows are 128 B (64 bf16); 8-row swizzle atoms are 1024 B apart (SBO); LBO is unused for this layout (1). */
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)1u << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1u << 46;
    d |= (uint64_t)2u << 61;
    return d;
}
/* Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1,
 * A,B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29). */
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

/* ------------------------------------------------------------------ activation split: f32 [M,lda] -> bf16 planes [nsplit][M][K] */
__device__ __forceinline__ uint32_t f2bf_rne(float f) {
    uint32_t u = __float_as_uint(f);
    return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
}
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

/* ------------------------------------------------------------------ host */
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = NULL;

static void make_map(CUtensorMap *map, const void *base, uint64_t inner_elems, uint64_t rows, uint64_t row_pitch_bytes) {
    if (!g_encode) {
        cudaDriverEntryPointQueryResult q;
        void *fn = NULL;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
            VB_FAIL("cuTensorMapEncodeTiled unavailable (driver too old for TMA)");
        }
        g_encode = (PFN_encodeTiled)fn;
    }
    cuuint64_t dims[2] = { inner_elems, rows }, strides[1] = { row_pitch_bytes };
    cuuint32_t box[2] = { TC_BK, TC_BM }, estr[2] = { 1, 1 };
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

void vb_gemm_tc(VbEngine *e, const float *A, int lda, const uint16_t *W, const float *bias, float *C, int ldc,
                int M, int N, int K, int epi) {
    static unsigned int attr_done = 0;                              /* one bit per device: function attributes are per device */
    const unsigned int dev_bit = 1u << (e->device & 31);
    const int nsplit = vb_gemm_nsplit();
    uint16_t *planes = (uint16_t *)vb_ws(e, VB_WS_GEMM_PLANES, (size_t)nsplit * M * K * 2 + 256);
    long long quads = ((long long)M * K + 3) / 4;
    k_split_planes<<<(int)((quads + 255) / 256), 256, 0, e->stream>>>(A, lda, M, K, nsplit, planes);
    CUtensorMap tmA, tmW;
    make_map(&tmA, planes, (uint64_t)K, (uint64_t)nsplit * M, (uint64_t)K * 2);
    make_map(&tmW, W, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2);
    if (!(attr_done & dev_bit)) {
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<VB_EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<VB_EPI_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<VB_EPI_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        VB_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<VB_EPI_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        attr_done |= dev_bit;
    }
    dim3 grid(N / TC_BN, (M + TC_BM - 1) / TC_BM), block(TC_THREADS);
    switch (epi) {
    case VB_EPI_STORE:    k_gemm_tc<VB_EPI_STORE><<<grid, block, TC_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit); break;
    case VB_EPI_GELU:     k_gemm_tc<VB_EPI_GELU><<<grid, block, TC_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit); break;
    case VB_EPI_RESIDUAL: k_gemm_tc<VB_EPI_RESIDUAL><<<grid, block, TC_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit); break;
    case VB_EPI_SWIGLU:   k_gemm_tc<VB_EPI_SWIGLU><<<grid, block, TC_SMEM_BYTES, e->stream>>>(tmA, tmW, bias, C, ldc, M, N, K, nsplit); break;
    }
    VB_CUDA_OK(cudaGetLastError());
    vb_launch_count(e, 2);
}
