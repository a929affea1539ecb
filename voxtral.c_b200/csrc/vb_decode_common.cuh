/*
 * vb_decode_common.cuh -- pieces shared by the decode drivers:
 *   vb_decode.cu          one kernel per phase, replayed as a CUDA graph (validation / fallback path)
 *   vb_decode_persist.cu  round 1's persistent cooperative kernel (direct streaming loads)
 *   vb_decode_v2.cu       the persistent kernel with a decoupled TMA weight stream and up to 8 activation columns (default)
 */
#ifndef VB_DECODE_COMMON_CUH
#define VB_DECODE_COMMON_CUH
#include "vb_ops.cuh"

#define DT 512                       /* threads per decode CTA */
#define DW (DT / 32)
#define DEC_DIM   VOX_DEC_DIM
#define DEC_HID   VOX_DEC_HIDDEN
#define HD        VOX_DEC_HEAD_DIM
#define NS_PER_CTA 2                 /* split-S groups per CTA in decode attention */

struct DecParams {
    const uint16_t *tok_emb;
    const uint16_t *wqkv[VOX_DEC_LAYERS], *wo[VOX_DEC_LAYERS], *w13[VOX_DEC_LAYERS], *w2[VOX_DEC_LAYERS];
    const float *attn_norm[VOX_DEC_LAYERS], *ffn_norm[VOX_DEC_LAYERS];
    const float *ada;               /* [26][3072] */
    const float *final_norm;
    const float *inv_freq;          /* [64] */
    float *kv_k, *kv_v;             /* [26][8192][1024] */
    float *x, *q, *attn_out, *gate, *logits;
    float *part_m, *part_l, *part_o;
    unsigned long long *argmax;
    VbDecState *st;
    const float *const *adapter_pp; /* device slot holding the adapter base pointer */
    int *tokens;
    int use_embed_kernel;
};

/* ---------------------------------------------------------------- helpers */
__device__ __forceinline__ uint4 ldg_stream16(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float dot8(const uint4 w, const float *x, float acc) {
    acc = fmaf(vb_bf16_lo(w.x), x[0], acc); acc = fmaf(vb_bf16_hi(w.x), x[1], acc);
    acc = fmaf(vb_bf16_lo(w.y), x[2], acc); acc = fmaf(vb_bf16_hi(w.y), x[3], acc);
    acc = fmaf(vb_bf16_lo(w.z), x[4], acc); acc = fmaf(vb_bf16_hi(w.z), x[5], acc);
    acc = fmaf(vb_bf16_lo(w.w), x[6], acc); acc = fmaf(vb_bf16_hi(w.w), x[7], acc);
    return acc;
}

template <int R> struct Log2;
template <> struct Log2<4>  { static const int v = 2; };
template <> struct Log2<8>  { static const int v = 3; };
template <> struct Log2<16> { static const int v = 4; };
template <> struct Log2<32> { static const int v = 5; };

/* Sum v[i] over the 32 lanes for all i<R with ~R shuffles.  Returns the total of
 * index (lane >> (5-log2 R)); lanes sharing that index hold the same value. */
template <int R>
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[R], int lane) {
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
    for (int o = (16 >> Log2<R>::v); o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;
}

/* Thread t (< NT) loads its CPT*8 activation values: chunk c covers k = (c*NT+t)*8 .. +7 */
template <int CPT>
__device__ __forceinline__ void load_x_cols(float (&xr)[CPT * 8], const float *__restrict__ x, int NT) {
    const int t = threadIdx.x;
#pragma unroll
    for (int c = 0; c < CPT; c++) {
        if (t < NT) {
            const float4 *p = reinterpret_cast<const float4 *>(x + (size_t)(c * NT + t) * 8);
            float4 a = p[0], b = p[1];
            xr[c * 8 + 0] = a.x; xr[c * 8 + 1] = a.y; xr[c * 8 + 2] = a.z; xr[c * 8 + 3] = a.w;
            xr[c * 8 + 4] = b.x; xr[c * 8 + 5] = b.y; xr[c * 8 + 6] = b.z; xr[c * 8 + 7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) xr[c * 8 + j] = 0.f;
        }
    }
}

/* Block-wide sum (all DT threads call). */
__device__ __forceinline__ float block_sum(float v, float *red /* [DW] */) {
    v = vb_warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < DW; i++) t += red[i];
    return t;
}

/* RMSNorm of the register-resident vector (voxtral_kernels.c:346-363), optional (1+ada). */
template <int CPT>
__device__ __forceinline__ void rmsnorm_cols(float (&xr)[CPT * 8], const float *__restrict__ w,
                                             const float *__restrict__ ada, int NT, int hidden, float *red) {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < CPT * 8; j++) ss = fmaf(xr[j], xr[j], ss);
    float tot = block_sum(ss, red);
    float rinv = 1.0f / sqrtf(tot / (float)hidden + VOX_DEC_NORM_EPS);
    const int t = threadIdx.x;
    if (t < NT) {
#pragma unroll
        for (int c = 0; c < CPT; c++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                int k = (c * NT + t) * 8 + j;
                float v = xr[c * 8 + j] * rinv * w[k];
                if (ada) v *= (1.0f + ada[k]);
                xr[c * 8 + j] = v;
            }
    }
}

/* y[row] = W[row,:] . x for rows [row0, row0+nrows); epi(row, value, lane, valid) runs in warp 0. */
template <int CPT, int R, typename Epi>
__device__ __forceinline__ void gemv_rows(const uint16_t *__restrict__ W, int K, int NT, int row0, int nrows,
                                          const float (&xr)[CPT * 8], float (*red)[R], Epi epi) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool active = t < NT;
    const uint16_t *wt = W + (size_t)t * 8;
    for (int rb = 0; rb < nrows; rb += R) {
        const int nr = min(R, nrows - rb);
        uint4 w[R][CPT];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int c = 0; c < CPT; c++) {
                if (active && r < nr) w[r][c] = ldg_stream16(wt + (size_t)(row0 + rb + r) * K + (size_t)c * NT * 8);
                else w[r][c] = make_uint4(0u, 0u, 0u, 0u);
            }
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            float a = 0.f;
#pragma unroll
            for (int c = 0; c < CPT; c++) a = dot8(w[r][c], &xr[c * 8], a);
            acc[r] = a;
        }
        float tot = warp_transpose_reduce<R>(acc, lane);
        if ((lane & ((32 >> Log2<R>::v) - 1)) == 0) red[warp][lane >> (5 - Log2<R>::v)] = tot;
        __syncthreads();
        if (warp == 0) {
            float s = 0.f;
            if (lane < R) {
#pragma unroll
                for (int wv = 0; wv < DW; wv++) s += red[wv][lane];
            }
            epi(row0 + rb + lane, s, lane, lane < nr);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void cta_rows(int total_units, int &u0, int &n) {
    /* contiguous, balanced partition of `total_units` over the grid */
    long long a = (long long)total_units * blockIdx.x / gridDim.x;
    long long b = (long long)total_units * (blockIdx.x + 1) / gridDim.x;
    u0 = (int)a; n = (int)(b - a);
}


__device__ __forceinline__ unsigned long long pack_cand(float v, int idx) {
    /* order-preserving float key in the high word, inverted index in the low word:
     * max() over packed values = largest value, ties -> smallest index (voxtral_decoder.c:697-704) */
    unsigned int u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)idx);
}
__device__ __forceinline__ int cand_index(unsigned long long c) { return (int)(0xFFFFFFFFu - (unsigned int)(c & 0xFFFFFFFFull)); }

DecParams vb_make_dec_params(VbEngine *e, int use_embed_kernel);
#endif
