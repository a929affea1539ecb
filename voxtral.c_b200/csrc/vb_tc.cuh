/*
 * vb_tc.cuh -- the sm_100a tensor-core plumbing shared by the tcgen05 kernels (vb_gemm_tc.cu, vb_attn_tc.cu):
 * mbarrier / TMA / tcgen05 PTX wrappers, shared-memory and instruction descriptors, the f32 -> bf16-plane split
 * and the host-side tensor-map encoder.
 */
#ifndef VB_TC_CUH
#define VB_TC_CUH

#include "vb_ops.cuh"
#include <cuda.h>

/* ------------------------------------------------------------------ PTX */
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void tc_mbar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s32(b)) : "memory");
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
/* D[tmem] (+)= A[tmem] * B[smem]: A as packed 16-bit pairs, row m in lane m, k elements 2 per column (cute SM100_MMA_F16BF16_TS) */
__device__ __forceinline__ void tc_umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z), "r"(z), "r"(z), "r"(z) : "memory");
}
__device__ __forceinline__ void tc_umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(s32(bar)) : "memory");
}
/* One lane of a converged warp (cute::elect_one_sync).  The MMA / TMA issue loops run with the whole warp converged and only the
 * issuing instructions under this predicate: descriptors and loop state then live in uniform registers.  With the loop inside an
 * `if (lane == 0)` the compiler has to move every descriptor into uniform registers through an elect/broadcast sequence
 * (9-14 instructions per tcgen05.mma): harmless at 12 MMAs per 1536 cycles (GEMM), the bottleneck at 48 short MMAs per key block
 * (attention, profiles/r02_encoder.md). */
__device__ __forceinline__ uint32_t tc_elect_one() {
    uint32_t pred = 0, laneid = 0;
    asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}"
                 : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFFu));
    return pred;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
/* generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads, TMA) */
__device__ __forceinline__ void tc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/* 32 lanes x 32 consecutive columns of TMEM -> 32 registers per thread (thread i of the warp = lane base + i) */
__device__ __forceinline__ void tc_tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
/* 16 registers per thread -> 32 lanes x 16 consecutive columns of TMEM */
__device__ __forceinline__ void tc_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tc_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
/* 32 lanes x 16 consecutive columns -> 16 registers per thread (waits for the load) */
__device__ __forceinline__ void tc_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    tc_tmem_ld32_nowait(taddr, r);
    tc_tmem_wait_ld();
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

/* ------------------------------------------------------------------ f32 -> bf16 planes: x = p0 + p1 + p2 (each step exact) */
__device__ __forceinline__ uint32_t f2bf_rne(float f) {
    uint32_t u = __float_as_uint(f);
    return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
}
/* two f32 -> one word of two bf16 (RNE, one F2FP): x in the low half, y in the high half */
__device__ __forceinline__ uint32_t tc_pack_bf16x2(float x, float y) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(y), "f"(x));
    return d;
}
__device__ __forceinline__ void tc_split3(float x, uint32_t &p0, uint32_t &p1, uint32_t &p2) {
    p0 = f2bf_rne(x);
    float r = x - __uint_as_float(p0 << 16);                          /* exact: r and p0 share the leading bits */
    p1 = f2bf_rne(r);
    r -= __uint_as_float(p1 << 16);
    p2 = f2bf_rne(r);
}

/* f32 [M, lda] (K columns used, K % 4 == 0) -> bf16 planes [nsplit][M][K] on the engine's stream (vb_gemm_tc.cu) */
void vb_tc_split_planes(VbEngine *e, const float *A, int lda, int M, int K, int nsplit, uint16_t *planes);
/* the same with the planes plane_elems apart (rows of a larger [nsplit][rows][K] buffer) */
void vb_tc_split_planes_strided(VbEngine *e, const float *A, int lda, int M, int K, int nsplit, uint16_t *planes, size_t plane_elems);

/* host: 2-D bf16 tensor map with SWIZZLE_128B boxes of 64 columns x box_rows rows (vb_gemm_tc.cu) */
void vb_tc_make_map(CUtensorMap *map, const void *base, uint64_t inner_elems, uint64_t rows, uint64_t row_pitch_bytes, uint32_t box_rows);

#endif
