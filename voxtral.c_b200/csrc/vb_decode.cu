/*
 * vb_decode.cu -- the autoregressive decoder step (HOT LOOP C of SURVEY.md section 3.1)
 * as CUDA kernels: reference voxtral_decoder.c:586-706 (one token through 26 layers,
 * final norm, tied-embedding logits, argmax) plus the embedding add of
 * voxtral.c:1057-1061, all driven by device-resident state so that consecutive
 * steps need no host round trip.
 *
 * Data flow per step (everything f32 except the bf16 weights, like the reference):
 *   x = adapter[row] + tok_emb[prev]                                   k_dec_embed
 *   26 x { RMSNorm -> [wq|wk|wv] GEMV -> RoPE -> KV ring write         k_dec_qkv
 *          split-S GQA attention over the ring + combine               k_dec_attn_partial/_combine
 *          wo GEMV + residual                                          k_dec_wo
 *          RMSNorm*(1+ada) -> [w1|w3] GEMV -> SiLU(g)*u                k_dec_w13
 *          w2 GEMV + residual }                                        k_dec_w2
 *   RMSNorm -> 131072x3072 GEMV -> per-CTA argmax                      k_dec_logits
 *   global argmax, state advance, token append                         k_dec_finish
 *
 * GEMV design (memory-bound: 6.86 GB of weights per step): every CTA owns a
 * contiguous slab of output rows; inside the CTA each thread owns a fixed set of
 * k-columns (8 consecutive k per 16-byte load) for ALL rows of the slab, so the
 * activation vector lives in registers, every row is read with perfectly
 * coalesced 128-bit streaming loads (ld.global.nc.L1::no_allocate), and R rows
 * are in flight per thread.  Partial sums are combined with a warp transpose-
 * reduce (R values in ~R shuffles) and one shared-memory pass.
 */
#include "vb_decode_common.cuh"
#include <string.h>

/* ---------------------------------------------------------------- kernels */
__global__ void __launch_bounds__(256) k_dec_embed(DecParams p) {
    const VbDecState st = *p.st;
    if (st.eos) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= DEC_DIM) return;
    const float *adapter = *p.adapter_pp;
    float a = adapter[(size_t)st.adapter_row * DEC_DIM + i];
    float t = __uint_as_float((uint32_t)p.tok_emb[(size_t)st.token * DEC_DIM + i] << 16);
    p.x[i] = a + t;                                      /* voxtral.c:1057-1061 */
}

__global__ void __launch_bounds__(DT, 1) k_dec_qkv(DecParams p, int layer) {
    __shared__ float red[DW][16];
    __shared__ float sred[DW];
    const VbDecState st = *p.st;
    if (st.eos) return;
    const int NT = DEC_DIM / 8;                          /* 384 */
    float xr[8];
    load_x_cols<1>(xr, p.x, NT);
    rmsnorm_cols<1>(xr, p.attn_norm[layer], nullptr, NT, DEC_DIM, sred);
    int u0, n; cta_rows(VB_DEC_QKV / 2, u0, n);
    const int pos = st.pos, slot = st.pos & (VB_KV_SLOTS - 1);
    float *kdst = p.kv_k + ((size_t)layer * VB_KV_SLOTS + slot) * VB_DEC_KV;
    float *vdst = p.kv_v + ((size_t)layer * VB_KV_SLOTS + slot) * VB_DEC_KV;
    const float *inv_freq = p.inv_freq;
    float *q = p.q;
    gemv_rows<1, 16>(p.wqkv[layer], DEC_DIM, NT, u0 * 2, n * 2, xr, red,
        [&](int row, float v, int lane, bool valid) {
            float other = __shfl_xor_sync(0xffffffffu, v, 1);
            if (!valid) return;
            if (row < VB_DEC_Q + VB_DEC_KV) {            /* RoPE on q and k, voxtral_kernels.c:503-526 */
                int d = (row & (HD - 1)) >> 1;
                float sn, cs;
                sincosf((float)pos * inv_freq[d], &sn, &cs);
                float y = (row & 1) ? (other * sn + v * cs) : (v * cs - other * sn);
                if (row < VB_DEC_Q) q[row] = y; else kdst[row - VB_DEC_Q] = y;
            } else {
                vdst[row - VB_DEC_Q - VB_DEC_KV] = v;
            }
        });
}

/* Split-S GQA attention over the KV ring: warp = (S-group, kv head); the 4 query heads
 * of a kv head share every K/V row read.  Order-independent online softmax, so the ring
 * needs no unrolling: valid slots are [0, min(pos+1, 8192)).  (voxtral_kernels.c:412-482) */
__global__ void __launch_bounds__(DT, 1) k_dec_attn_partial(DecParams p, int layer) {
    const VbDecState st = *p.st;
    if (st.eos) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kvh = warp & 7, sub = warp >> 3;
    const int NS = gridDim.x * NS_PER_CTA, sg = blockIdx.x * NS_PER_CTA + sub;
    const int n_valid = min(st.pos + 1, VB_KV_SLOTS);
    const int s0 = (int)((long long)n_valid * sg / NS), s1 = (int)((long long)n_valid * (sg + 1) / NS);
    const float scale = 1.0f / sqrtf((float)HD);

    float4 qv[4];
#pragma unroll
    for (int hq = 0; hq < 4; hq++)
        qv[hq] = *reinterpret_cast<const float4 *>(p.q + (kvh * 4 + hq) * HD + lane * 4);
    float m[4], l[4]; float4 o[4];
#pragma unroll
    for (int hq = 0; hq < 4; hq++) { m[hq] = -1e30f; l[hq] = 0.f; o[hq] = make_float4(0.f, 0.f, 0.f, 0.f); }

    const float *kb = p.kv_k + (size_t)layer * VB_KV_SLOTS * VB_DEC_KV + kvh * HD + lane * 4;
    const float *vb = p.kv_v + (size_t)layer * VB_KV_SLOTS * VB_DEC_KV + kvh * HD + lane * 4;
    for (int s = s0; s < s1; s++) {
        float4 k4 = *reinterpret_cast<const float4 *>(kb + (size_t)s * VB_DEC_KV);
        float4 v4 = *reinterpret_cast<const float4 *>(vb + (size_t)s * VB_DEC_KV);
        float sc[4];
#pragma unroll
        for (int hq = 0; hq < 4; hq++)
            sc[hq] = qv[hq].x * k4.x + qv[hq].y * k4.y + qv[hq].z * k4.z + qv[hq].w * k4.w;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
#pragma unroll
            for (int hq = 0; hq < 4; hq++) sc[hq] += __shfl_xor_sync(0xffffffffu, sc[hq], off);
#pragma unroll
        for (int hq = 0; hq < 4; hq++) {
            float sv = sc[hq] * scale;
            float mn = fmaxf(m[hq], sv);
            float c = expf(m[hq] - mn), pw = expf(sv - mn);
            l[hq] = l[hq] * c + pw;
            o[hq].x = o[hq].x * c + pw * v4.x; o[hq].y = o[hq].y * c + pw * v4.y;
            o[hq].z = o[hq].z * c + pw * v4.z; o[hq].w = o[hq].w * c + pw * v4.w;
            m[hq] = mn;
        }
    }
#pragma unroll
    for (int hq = 0; hq < 4; hq++) {
        int h = kvh * 4 + hq;
        size_t pi = (size_t)sg * VOX_DEC_HEADS + h;
        if (lane == 0) { p.part_m[pi] = m[hq]; p.part_l[pi] = l[hq]; }
        *reinterpret_cast<float4 *>(p.part_o + pi * HD + lane * 4) = o[hq];
    }
}

/* Combine the split-S partials: grid = (32 heads, 4 dim-quarters), 512 threads = 32 dims x 16 partial lanes. */
__global__ void __launch_bounds__(DT) k_dec_attn_combine(DecParams p, int NS) {
    __shared__ float sm_m[16], sm_num[16][33], sm_den[16][33];
    const VbDecState st = *p.st;
    if (st.eos) return;
    const int h = blockIdx.x, dq = blockIdx.y;
    const int dl = threadIdx.x & 31, pl = threadIdx.x >> 5;          /* dim within quarter, partial lane (= warp) */
    const int d = dq * 32 + dl;
    float M = -1e30f;
    for (int i = threadIdx.x; i < NS; i += DT) M = fmaxf(M, p.part_m[(size_t)i * VOX_DEC_HEADS + h]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
    if (dl == 0) sm_m[pl] = M;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; i++) M = fmaxf(M, sm_m[i]);
    float num = 0.f, den = 0.f;
    for (int i = pl; i < NS; i += 16) {
        size_t pi = (size_t)i * VOX_DEC_HEADS + h;
        float li = p.part_l[pi];
        if (li > 0.f) {
            float w = expf(p.part_m[pi] - M);
            den = fmaf(w, li, den);
            num = fmaf(w, p.part_o[pi * HD + d], num);
        }
    }
    sm_num[pl][dl] = num; sm_den[pl][dl] = den;
    __syncthreads();
    if (pl == 0) {
        float n2 = 0.f, d2 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i++) { n2 += sm_num[i][dl]; d2 += sm_den[i][dl]; }
        p.attn_out[h * HD + d] = d2 > 0.f ? n2 / d2 : 0.f;
    }
}

__global__ void __launch_bounds__(DT, 1) k_dec_wo(DecParams p, int layer) {
    __shared__ float red[DW][16];
    const VbDecState st = *p.st;
    if (st.eos) return;
    const int NT = VB_DEC_Q / 8;                         /* 512 */
    float xr[8];
    load_x_cols<1>(xr, p.attn_out, NT);
    int r0, n; cta_rows(DEC_DIM, r0, n);
    float *x = p.x;
    gemv_rows<1, 16>(p.wo[layer], VB_DEC_Q, NT, r0, n, xr, red,
        [&](int row, float v, int, bool valid) { if (valid) x[row] += v; });
}

__global__ void __launch_bounds__(DT, 1) k_dec_w13(DecParams p, int layer) {
    __shared__ float red[DW][16];
    __shared__ float sred[DW];
    const VbDecState st = *p.st;
    if (st.eos) return;
    const int NT = DEC_DIM / 8;
    float xr[8];
    load_x_cols<1>(xr, p.x, NT);
    rmsnorm_cols<1>(xr, p.ffn_norm[layer], p.ada + (size_t)layer * DEC_DIM, NT, DEC_DIM, sred);
    int u0, n; cta_rows(DEC_HID, u0, n);
    float *gate = p.gate;
    gemv_rows<1, 16>(p.w13[layer], DEC_DIM, NT, u0 * 2, n * 2, xr, red,
        [&](int row, float v, int, bool valid) {
            float other = __shfl_xor_sync(0xffffffffu, v, 1);
            if (valid && !(row & 1)) gate[row >> 1] = vb_silu(v) * other;   /* voxtral_decoder.c:682-686 */
        });
}

__global__ void __launch_bounds__(DT, 1) k_dec_w2(DecParams p, int layer) {
    __shared__ float red[DW][4];
    const VbDecState st = *p.st;
    if (st.eos) return;
    const int NT = DEC_HID / 8 / 3;                      /* 384, 3 chunks per thread */
    float xr[24];
    load_x_cols<3>(xr, p.gate, NT);
    int r0, n; cta_rows(DEC_DIM, r0, n);
    float *x = p.x;
    gemv_rows<3, 4>(p.w2[layer], DEC_HID, NT, r0, n, xr, red,
        [&](int row, float v, int, bool valid) { if (valid) x[row] += v; });
}

__global__ void __launch_bounds__(DT, 1) k_dec_logits(DecParams p) {
    __shared__ float red[DW][16];
    __shared__ float sred[DW];
    const VbDecState st = *p.st;
    if (st.eos) return;
    const int NT = DEC_DIM / 8;
    float xr[8];
    load_x_cols<1>(xr, p.x, NT);
    rmsnorm_cols<1>(xr, p.final_norm, nullptr, NT, DEC_DIM, sred);
    int r0, n; cta_rows(VOX_VOCAB_SIZE, r0, n);
    float *logits = p.logits;
    unsigned long long best = 0ull;
    gemv_rows<1, 16>(p.tok_emb, DEC_DIM, NT, r0, n, xr, red,
        [&](int row, float v, int, bool valid) {
            if (!valid) return;
            logits[row] = v;
            unsigned long long c = pack_cand(v, row);
            if (c > best) best = c;
        });
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            if (other > best) best = other;
        }
        if (threadIdx.x == 0) p.argmax[blockIdx.x] = best;
    }
}

__global__ void __launch_bounds__(256) k_dec_finish(DecParams p, int n_cands) {
    __shared__ unsigned long long sh[8];
    VbDecState st = *p.st;
    if (st.eos) return;
    unsigned long long best = 0ull;
    for (int i = threadIdx.x; i < n_cands; i += 256) { unsigned long long c = p.argmax[i]; if (c > best) best = c; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        if (other > best) best = other;
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) if (sh[i] > best) best = sh[i];
        int tok = cand_index(best);
        p.tokens[st.n_out] = tok;
        st.n_out += 1; st.token = tok; st.pos += 1; st.adapter_row += 1;
        if (tok == VB_TOKEN_EOS) st.eos = 1;
        *p.st = st;
    }
}

/* Standalone GEMV used by the host-pointer parity seam (vox_linear*_bf16 with seq_len == 1):
 * exercises the very same gemv_rows core as the decode step. */
template <int CPT>
__global__ void __launch_bounds__(DT, 1)
k_gemv_generic(float *__restrict__ y, const float *__restrict__ x, const uint16_t *__restrict__ W,
               const float *__restrict__ bias, int K, int N, int NT) {
    __shared__ float red[DW][8];
    float xr[CPT * 8];
    load_x_cols<CPT>(xr, x, NT);
    int r0, n; cta_rows(N, r0, n);
    gemv_rows<CPT, 8>(W, K, NT, r0, n, xr, red,
        [&](int row, float v, int, bool valid) { if (valid) y[row] = bias ? v + bias[row] : v; });
}

/* Small-M linears (2 <= M < 8: the live-stream encoder calls of main.c's -I 0.1 mode, where a call has ~5 positions):
 * the same streaming GEMV core with MB activation rows in registers, so the weights are read ONCE, coalesced, by all SMs,
 * instead of going through a 64x64-tile GEMM whose grid is a fraction of the machine (measured: 33 ms per 5-position encoder
 * call before, see profiles/r02_live.md).  y[m][row] = W[row,:] . x[m] (+bias) with the GEMM epilogues of vb_ops.cuh;
 * 16 / MB rows per batch so that a batch always reduces 16 (row, m) values.  C and A are row-major with pitches ldc / lda. */
template <int CPT, int MB, int EPI>
__global__ void __launch_bounds__(DT, 1)
k_gemv_cols(float *__restrict__ C, int ldc, const float *__restrict__ A, int lda, const uint16_t *__restrict__ W,
            const float *__restrict__ bias, int K, int N, int NT, int M, int WG, int NG) {
    constexpr int R = 16 / MB;                                    /* rows per batch */
    __shared__ float red[2][DW][16];
    /* A row of K <= 2048 occupies only NT = K/8 threads; the CTA then runs NG groups of WG warps, each group on its own
     * batches of rows (K = 1280: 3 groups of 5 warps instead of 5 busy warps out of 16).  All groups step together, so one
     * __syncthreads per batch serves them all. */
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int g = warp / WG, tl = t - g * WG * 32;
    const bool active = g < NG && tl < NT;
    float xr[MB][CPT * 8];
#pragma unroll
    for (int m = 0; m < MB; m++)
#pragma unroll
        for (int c = 0; c < CPT; c++) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (active && m < M) {
                const float4 *p = reinterpret_cast<const float4 *>(A + (size_t)m * lda + (size_t)(c * NT + tl) * 8);
                a = p[0]; b = p[1];
            }
            xr[m][c * 8 + 0] = a.x; xr[m][c * 8 + 1] = a.y; xr[m][c * 8 + 2] = a.z; xr[m][c * 8 + 3] = a.w;
            xr[m][c * 8 + 4] = b.x; xr[m][c * 8 + 5] = b.y; xr[m][c * 8 + 6] = b.z; xr[m][c * 8 + 7] = b.w;
        }
    int u0, nu; cta_rows(N / 2, u0, nu);                          /* row pairs: SwiGLU keeps (gate, up) in one CTA */
    const int row0 = u0 * 2, nrows = nu * 2;
    const uint16_t *wt = W + (size_t)(active ? tl : 0) * 8;
    /* U batches of R rows are loaded up front (8 rows = 128 bytes per thread in flight for CPT = 1: with MB = 8 a single batch would
     * be 2 rows and the kernel latency-bound), then reduced one after the other */
    constexpr int U = CPT == 1 ? (8 / R > 0 ? 8 / R : 1) : 1;
    int buf = 0;
    for (int base = 0; base < nrows; base += NG * U * R) {
        const int rb = base + g * U * R;                          /* this group's rows (may lie past the end: masked) */
        uint4 w[U * R][CPT];
#pragma unroll
        for (int r = 0; r < U * R; r++)
#pragma unroll
            for (int c = 0; c < CPT; c++)
                w[r][c] = (active && rb + r < nrows) ? ldg_stream16(wt + (size_t)(row0 + rb + r) * K + (size_t)c * NT * 8) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (base + u * R >= nrows) break;                     /* uniform over the CTA: group 0 has the lowest rows */
            float acc[16];
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int m = 0; m < MB; m++) {
                    float a = 0.f;
#pragma unroll
                    for (int c = 0; c < CPT; c++) a = dot8(w[u * R + r][c], &xr[m][c * 8], a);
                    acc[r * MB + m] = a;
                }
            const float tot = warp_transpose_reduce<16>(acc, lane);
            if (!(lane & 1)) red[buf][warp][lane >> 1] = tot;
            __syncthreads();
            if (g < NG && warp == g * WG) {
                float v = 0.f;
                if (lane < 16)
                    for (int wv = 0; wv < WG; wv++) v += red[buf][g * WG + wv][lane];
                const int r = lane / MB, m = lane % MB, row = row0 + rb + u * R + r;
                const float other = __shfl_xor_sync(0xffffffffu, v, MB);       /* row ^ 1, same m: the (gate, up) partner */
                if (lane < 16 && rb + u * R + r < nrows && m < M) {
                    if (EPI == VB_EPI_SWIGLU) {
                        if (!(row & 1)) C[(size_t)m * ldc + (row >> 1)] = vb_silu(v) * other;
                    } else {
                        if (bias) v += bias[row];
                        if (EPI == VB_EPI_GELU) v = vb_gelu_tanh(v);
                        if (EPI == VB_EPI_RESIDUAL) v += C[(size_t)m * ldc + row];
                        C[(size_t)m * ldc + row] = v;
                    }
                }
            }
            buf ^= 1;
        }
    }
}

template <int CPT, int MB>
static void gemv_cols_launch(VbEngine *e, float *C, int ldc, const float *A, int lda, const uint16_t *W, const float *bias,
                             int K, int N, int NT, int M, int epi) {
    const int G = e->sm_count;
    const int WG = (NT + 31) / 32, NG = DW / WG > 0 ? DW / WG : 1;
    switch (epi) {
    case VB_EPI_STORE:    k_gemv_cols<CPT, MB, VB_EPI_STORE><<<G, DT, 0, e->stream>>>(C, ldc, A, lda, W, bias, K, N, NT, M, WG, NG); break;
    case VB_EPI_GELU:     k_gemv_cols<CPT, MB, VB_EPI_GELU><<<G, DT, 0, e->stream>>>(C, ldc, A, lda, W, bias, K, N, NT, M, WG, NG); break;
    case VB_EPI_RESIDUAL: k_gemv_cols<CPT, MB, VB_EPI_RESIDUAL><<<G, DT, 0, e->stream>>>(C, ldc, A, lda, W, bias, K, N, NT, M, WG, NG); break;
    case VB_EPI_SWIGLU:   k_gemv_cols<CPT, MB, VB_EPI_SWIGLU><<<G, DT, 0, e->stream>>>(C, ldc, A, lda, W, bias, K, N, NT, M, WG, NG); break;
    }
}

/* returns 0 if the shape is not covered (the caller falls back to the tiled GEMM) */
extern "C" int vb_gemv_cols_dev(VbEngine *e, const float *A, int lda, const uint16_t *W, const float *bias, float *C, int ldc,
                                int M, int N, int K, int epi) {
    if (M < 1 || M > 8 || (K % 8) || (N % 2) || (lda % 4)) return 0;
    const int chunks = K / 8;
    int cpt = 1;
    while (cpt <= 2 && (chunks % cpt || chunks / cpt > DT)) cpt++;
    if (cpt > 2) return 0;
    const int NT = chunks / cpt;
    const int per = cpt == 1 ? 8 : 4;                             /* activation rows per launch (register budget) */
    for (int m0 = 0; m0 < M; m0 += per) {
        const int mb = M - m0 < per ? M - m0 : per;
        const float *a = A + (size_t)m0 * lda; float *c = C + (size_t)m0 * ldc;
        if (cpt == 1) {
            if (mb <= 1) gemv_cols_launch<1, 1>(e, c, ldc, a, lda, W, bias, K, N, NT, mb, epi);
            else if (mb <= 2) gemv_cols_launch<1, 2>(e, c, ldc, a, lda, W, bias, K, N, NT, mb, epi);
            else if (mb <= 4) gemv_cols_launch<1, 4>(e, c, ldc, a, lda, W, bias, K, N, NT, mb, epi);
            else gemv_cols_launch<1, 8>(e, c, ldc, a, lda, W, bias, K, N, NT, mb, epi);
        } else {
            if (mb <= 1) gemv_cols_launch<2, 1>(e, c, ldc, a, lda, W, bias, K, N, NT, mb, epi);
            else if (mb <= 2) gemv_cols_launch<2, 2>(e, c, ldc, a, lda, W, bias, K, N, NT, mb, epi);
            else gemv_cols_launch<2, 4>(e, c, ldc, a, lda, W, bias, K, N, NT, mb, epi);
        }
        VB_CUDA_OK(cudaGetLastError());
        vb_launch_count(e, 1);
    }
    return 1;
}

/* ---------------------------------------------------------------- host side */
DecParams vb_make_dec_params(VbEngine *e, int use_embed_kernel) {
    DecParams p;
    memset(&p, 0, sizeof p);
    p.tok_emb = e->d_tok_emb;
    for (int l = 0; l < VOX_DEC_LAYERS; l++) {
        p.wqkv[l] = e->dec[l].wqkv; p.wo[l] = e->dec[l].wo; p.w13[l] = e->dec[l].w13; p.w2[l] = e->dec[l].w2;
        p.attn_norm[l] = e->dec[l].attn_norm; p.ffn_norm[l] = e->dec[l].ffn_norm;
    }
    p.ada = e->d_ada_scale; p.final_norm = e->d_dec_norm; p.inv_freq = e->d_dec_inv_freq;
    p.kv_k = e->d_kv_k; p.kv_v = e->d_kv_v;
    p.x = e->d_x; p.q = e->d_q; p.attn_out = e->d_attn_out; p.gate = e->d_gate; p.logits = e->d_logits;
    p.part_m = e->d_part_m; p.part_l = e->d_part_l; p.part_o = e->d_part_o;
    p.argmax = e->d_argmax; p.st = e->d_state;
    p.adapter_pp = (const float *const *)(e->d_state + 1);   /* pointer slot lives right after the state */
    p.tokens = e->d_tokens;
    p.use_embed_kernel = use_embed_kernel;
    return p;
}

extern "C" void vb_decoder_free(VbEngine *e) {
    if (e->h_tokens_pinned) { cudaFreeHost(e->h_tokens_pinned); e->h_tokens_pinned = NULL; }
}

extern "C" int vb_decoder_alloc(VbEngine *e) {
    if (e->d_kv_k) return 0;
    const size_t weight_bytes_before = e->weight_bytes;     /* activations/KV are not "weights" */
    size_t kv = (size_t)VOX_DEC_LAYERS * VB_KV_SLOTS * VB_DEC_KV * sizeof(float);
    e->d_kv_k = (float *)vb_dev_alloc_owned(e, kv);
    e->d_kv_v = (float *)vb_dev_alloc_owned(e, kv);
    e->kv_bytes = 2 * kv;
    VB_CUDA_OK(cudaMemsetAsync(e->d_kv_k, 0, kv, e->stream));
    VB_CUDA_OK(cudaMemsetAsync(e->d_kv_v, 0, kv, e->stream));
    e->d_state = (VbDecState *)vb_dev_alloc_owned(e, sizeof(VbDecState) + 64);
    VB_CUDA_OK(cudaMemsetAsync(e->d_state, 0, sizeof(VbDecState) + 64, e->stream));
    e->d_x = (float *)vb_dev_alloc_owned(e, DEC_DIM * 4);
    e->d_q = (float *)vb_dev_alloc_owned(e, VB_DEC_Q * 4);
    e->d_attn_out = (float *)vb_dev_alloc_owned(e, VB_DEC_Q * 4);
    e->d_gate = (float *)vb_dev_alloc_owned(e, DEC_HID * 4);
    e->d_logits = (float *)vb_dev_alloc_owned(e, (size_t)VOX_VOCAB_SIZE * 4);
    int NS = e->sm_count * NS_PER_CTA;
    e->d_part_m = (float *)vb_dev_alloc_owned(e, (size_t)NS * VOX_DEC_HEADS * 4);
    e->d_part_l = (float *)vb_dev_alloc_owned(e, (size_t)NS * VOX_DEC_HEADS * 4);
    e->d_part_o = (float *)vb_dev_alloc_owned(e, (size_t)NS * VOX_DEC_HEADS * HD * 4);
    e->d_argmax = (unsigned long long *)vb_dev_alloc_owned(e, (size_t)e->sm_count * 8);
    e->tokens_cap = 65536;
    e->d_tokens = (int *)vb_dev_alloc_owned(e, (size_t)e->tokens_cap * 4);
    VB_CUDA_OK(cudaMallocHost((void **)&e->h_tokens_pinned, (size_t)e->tokens_cap * 4));
    e->d_embed_in = (float *)vb_dev_alloc_owned(e, DEC_DIM * 4);
    e->weight_bytes = weight_bytes_before;
    return 0;
}

/* Enqueue the kernels of one decode step on e->stream. */
static void enqueue_step(VbEngine *e, const DecParams &p) {
    const int G = e->sm_count;
    cudaStream_t s = e->stream;
    if (p.use_embed_kernel) k_dec_embed<<<(DEC_DIM + 255) / 256, 256, 0, s>>>(p);
    for (int l = 0; l < VOX_DEC_LAYERS; l++) {
        k_dec_qkv<<<G, DT, 0, s>>>(p, l);
        k_dec_attn_partial<<<G, DT, 0, s>>>(p, l);
        k_dec_attn_combine<<<dim3(VOX_DEC_HEADS, 4), DT, 0, s>>>(p, G * NS_PER_CTA);
        k_dec_wo<<<G, DT, 0, s>>>(p, l);
        k_dec_w13<<<G, DT, 0, s>>>(p, l);
        k_dec_w2<<<G, DT, 0, s>>>(p, l);
    }
    k_dec_logits<<<G, DT, 0, s>>>(p);
    k_dec_finish<<<1, 256, 0, s>>>(p, G);
}
#define STEP_KERNELS (VOX_DEC_LAYERS * 6 + 2)

static void set_state(VbEngine *e, int pos, int token, int adapter_row, const float *d_adapter) {
    struct { VbDecState st; const float *adapter; } h;
    memset(&h, 0, sizeof h);
    h.st.pos = pos; h.st.token = token; h.st.adapter_row = adapter_row;
    h.adapter = d_adapter;
    VB_CUDA_OK(cudaMemcpyAsync(e->d_state, &h, sizeof h, cudaMemcpyHostToDevice, e->stream));
    VB_CUDA_OK(cudaStreamSynchronize(e->stream));   /* h is on the stack */
}

/* returns the decode driver to use: 1 CUDA graph of per-phase kernels, 3 direct-load persistent kernel, 5 v2 persistent kernel */
static int decode_driver(VbEngine *e) {
    if (e->decode_mode == 0) {
        const char *vd = getenv("VOX_CUDA_VERIFY");
        if (vd && e->verify_depth == 0) e->verify_depth = atoi(vd);
        const char *m = getenv("VOX_CUDA_DECODE");
        if (m && !strcmp(m, "graph")) e->decode_mode = 1;
        else if (m && !strcmp(m, "persist")) e->decode_mode = 3;
        else if (m && !strcmp(m, "v2")) e->decode_mode = 5;
        else e->decode_mode = vb_decoder_v2_supported(e) ? 5 : vb_decoder_persist_supported(e) ? 3 : 1;
    }
    if (e->decode_mode != 1 && e->decode_mode != 3 && e->decode_mode != 5) e->decode_mode = vb_decoder_v2_supported(e) ? 5 : vb_decoder_persist_supported(e) ? 3 : 1;
    if (e->decode_mode == 3 && !vb_decoder_persist_supported(e)) {
        VB_FAIL("persistent decode kernel requested but cooperative launch is unavailable");
    }
    if (e->decode_mode == 5 && !vb_decoder_v2_supported(e)) {
        VB_FAIL("v2 decode kernel requested but unavailable on this device");
    }
    return e->decode_mode;
}

extern "C" int vb_decoder_run_steps(VbEngine *e, const float *d_adapter, int adapter_row, int n_steps,
                                    int prev_token, int pos, int *out_tokens_host) {
    if (n_steps <= 0) return 0;
    vb_decoder_alloc(e);
    const int driver = decode_driver(e);
    const int mega = driver >= 2;
    const int max_chunk = mega ? 2048 : e->tokens_cap;          /* bound the lifetime of one persistent launch */
    int done = 0;
    double total_ms = 0;
    while (done < n_steps) {
        int chunk = n_steps - done;
        if (chunk > max_chunk) chunk = max_chunk;
        int verify_launch = 0;
        if (mega) {
            VB_CUDA_OK(cudaEventRecord(e->ev0, e->stream));
            if (driver == 5) {
                VbV2Col col = { e, d_adapter, adapter_row + done, chunk, prev_token, pos + done };
                /* exact multi-token decoding while the drafted positions cannot wrap the KV ring (vb_decode_v2.cu, verify mode) */
                const int depth = e->verify_depth > 1 && pos + done + chunk + e->verify_depth <= VB_KV_SLOTS ? e->verify_depth : 1;
                verify_launch = depth > 1;
                if (vb_decoder_v2_launch(e, &col, depth, chunk, depth > 1, NULL) != 0) VB_FAIL("v2 decode launch failed");
            } else vb_decoder_persist_launch(e, d_adapter, adapter_row + done, chunk, prev_token, pos + done);
            VB_CUDA_OK(cudaEventRecord(e->ev1, e->stream));
        } else {
            set_state(e, pos + done, prev_token, adapter_row + done, d_adapter);
            if (!e->step_graph_ready) {
                DecParams p = vb_make_dec_params(e, 1);
                cudaGraph_t g;
                VB_CUDA_OK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
                enqueue_step(e, p);
                VB_CUDA_OK(cudaStreamEndCapture(e->stream, &g));
                VB_CUDA_OK(cudaGraphInstantiate(&e->step_graph, g, 0));
                VB_CUDA_OK(cudaGraphDestroy(g));
                e->step_graph_ready = 1;
            }
            VB_CUDA_OK(cudaEventRecord(e->ev0, e->stream));
            for (int i = 0; i < chunk; i++) VB_CUDA_OK(cudaGraphLaunch(e->step_graph, e->stream));
            VB_CUDA_OK(cudaEventRecord(e->ev1, e->stream));
        }
        VbDecState st;
        VB_CUDA_OK(cudaMemcpyAsync(&st, driver == 5 ? e->v2.st : e->d_state, sizeof st, cudaMemcpyDeviceToHost, e->stream));
        cudaError_t serr = cudaStreamSynchronize(e->stream);
        if (serr != cudaSuccess) {
            fprintf(stderr, "voxtral_b200: decode kernel failed: %s (mode %s, wait-guard code %d)\n", cudaGetErrorString(serr),
                    driver == 1 ? "graph" : driver == 5 ? "v2" : "persist", driver == 5 && e->v2.err_host ? *e->v2.err_host : -1);
            vb_cuda_fail(serr, __FILE__, __LINE__);
        }
        VB_CUDA_OK(cudaMemcpy(e->h_tokens_pinned, e->d_tokens, (size_t)st.n_out * 4, cudaMemcpyDeviceToHost));
        float ms = 0; cudaEventElapsedTime(&ms, e->ev0, e->ev1); total_ms += ms;
        if (verify_launch) { e->verify_passes += st.pad[0]; e->verify_tokens += st.n_out; }
        if (!mega) e->launches += (unsigned long long)st.n_out * STEP_KERNELS;
        memcpy(out_tokens_host + done, e->h_tokens_pinned, (size_t)st.n_out * 4);
        done += st.n_out;
        if (st.n_out > 0) prev_token = e->h_tokens_pinned[st.n_out - 1];
        if (st.eos || st.n_out < chunk) break;
    }
    e->last_decode_ms = total_ms; e->last_decode_steps = done;
    e->total_decode_ms += total_ms; e->total_decode_steps += done;
    return done;
}

/* One step from an explicit input embedding (vox_decoder_forward semantics). */
extern "C" int vb_decoder_step_from_embed(VbEngine *e, const float *d_embed, int pos, float *logits_host) {
    vb_decoder_alloc(e);
    set_state(e, pos, 0, 0, nullptr);
    VB_CUDA_OK(cudaMemcpyAsync(e->d_x, d_embed, DEC_DIM * 4, cudaMemcpyDeviceToDevice, e->stream));
    DecParams p = vb_make_dec_params(e, 0);
    enqueue_step(e, p);
    VB_CUDA_OK(cudaGetLastError());
    e->launches += STEP_KERNELS - 1;
    int tok = 0;
    VB_CUDA_OK(cudaMemcpyAsync(&tok, e->d_tokens, 4, cudaMemcpyDeviceToHost, e->stream));
    if (logits_host)
        VB_CUDA_OK(cudaMemcpyAsync(logits_host, e->d_logits, (size_t)VOX_VOCAB_SIZE * 4, cudaMemcpyDeviceToHost, e->stream));
    VB_CUDA_OK(cudaStreamSynchronize(e->stream));
    return tok;
}

extern "C" void vb_gemv_bf16_dev(VbEngine *e, float *y, const float *x, const uint16_t *W, const float *bias, int K, int N) {
    int chunks = K / 8, cpt = 1;
    while (cpt <= 4 && (chunks % cpt || chunks / cpt > DT)) cpt++;
    if (K % 8 || cpt > 4) { fprintf(stderr, "voxtral_b200: GEMV K=%d unsupported\n", K); VB_FAIL("GEMV shape unsupported"); }
    int NT = chunks / cpt, G = e->sm_count;
    switch (cpt) {
    case 1: k_gemv_generic<1><<<G, DT, 0, e->stream>>>(y, x, W, bias, K, N, NT); break;
    case 2: k_gemv_generic<2><<<G, DT, 0, e->stream>>>(y, x, W, bias, K, N, NT); break;
    case 3: k_gemv_generic<3><<<G, DT, 0, e->stream>>>(y, x, W, bias, K, N, NT); break;
    case 4: k_gemv_generic<4><<<G, DT, 0, e->stream>>>(y, x, W, bias, K, N, NT); break;
    }
    VB_CUDA_OK(cudaGetLastError());
    vb_launch_count(e, 1);
}

/* Prefill: seq_len prompt embeddings through the 26 layers, filling the KV ring; no logits
 * (voxtral_decoder.c:410-558). */
extern "C" void vb_decoder_prefill_dev(VbEngine *e, const float *d_embeds, int n, int start_pos) {
    if (n <= 0) return;
    vb_decoder_alloc(e);
    if (start_pos + n > VB_KV_SLOTS) {
        /* the M > 1 attention addresses keys by physical row: a prompt may not cross the end of the ring (the stream API
         * prefills at position 0 only; 8192 slots).  An error for the caller's guard, not a process abort. */
        fprintf(stderr, "voxtral_b200: prefill of %d tokens at position %d would wrap the KV ring\n", n, start_pos);
        VB_FAIL("prefill would wrap the KV ring");
    }
    float *x = vb_ws(e, 0, (size_t)n * DEC_DIM * 4);
    float *xn = vb_ws(e, 1, (size_t)n * DEC_DIM * 4);
    float *qkv = vb_ws(e, 2, (size_t)n * VB_DEC_QKV * 4);
    float *att = vb_ws(e, 3, (size_t)n * VB_DEC_Q * 4);
    float *g = vb_ws(e, 4, (size_t)n * DEC_HID * 4);
    VB_CUDA_OK(cudaMemcpyAsync(x, d_embeds, (size_t)n * DEC_DIM * 4, cudaMemcpyDeviceToDevice, e->stream));
    const float scale = 1.0f / sqrtf((float)HD);
    for (int l = 0; l < VOX_DEC_LAYERS; l++) {
        float *kl = e->d_kv_k + (size_t)l * VB_KV_SLOTS * VB_DEC_KV;
        float *vl = e->d_kv_v + (size_t)l * VB_KV_SLOTS * VB_DEC_KV;
        vb_rmsnorm_rows(e, xn, x, e->dec[l].attn_norm, nullptr, n, DEC_DIM, VOX_DEC_NORM_EPS);
        vb_gemm_bf16w(e, xn, DEC_DIM, e->dec[l].wqkv, nullptr, qkv, VB_DEC_QKV, n, VB_DEC_QKV, DEC_DIM, VB_EPI_STORE);
        vb_rope_split(e, qkv, VB_DEC_QKV, n, VOX_DEC_HEADS, VOX_DEC_KV_HEADS, HD, e->d_dec_inv_freq, start_pos,
                      kl, vl, start_pos, VB_KV_SLOTS - 1);
        vb_attention_rows(e, att, VB_DEC_Q, qkv, VB_DEC_QKV, kl, vl, VB_DEC_KV, n, start_pos + n,
                          VOX_DEC_HEADS, VOX_DEC_KV_HEADS, HD, scale, VOX_DEC_WINDOW, start_pos);
        vb_gemm_bf16w(e, att, VB_DEC_Q, e->dec[l].wo, nullptr, x, DEC_DIM, n, DEC_DIM, VB_DEC_Q, VB_EPI_RESIDUAL);
        vb_rmsnorm_rows(e, xn, x, e->dec[l].ffn_norm, e->d_ada_scale + (size_t)l * DEC_DIM, n, DEC_DIM, VOX_DEC_NORM_EPS);
        vb_gemm_bf16w(e, xn, DEC_DIM, e->dec[l].w13, nullptr, g, DEC_HID, n, 2 * DEC_HID, DEC_DIM, VB_EPI_SWIGLU);
        vb_gemm_bf16w(e, g, DEC_HID, e->dec[l].w2, nullptr, x, DEC_DIM, n, DEC_DIM, DEC_HID, VB_EPI_RESIDUAL);
    }
}


/* ---------------------------------------------------------------- alternatives (vox_stream_set_alt) on the device */
/* What stream_fill_alts() needs from a step's 131072 logits (voxtral.c:911-966): Z = sum_i exp(l_i - l_best) and the
 * (up to) three largest text-range logits other than the best token, first index winning ties.  One CTA re-reads the
 * logits from L2 (512 KB) instead of shipping them to the host for a 131072-wide expf loop. */
#define ALT_THREADS 1024
__global__ void __launch_bounds__(ALT_THREADS) k_alt_candidates(const float *__restrict__ logits, int best, int text_min,
                                                                float *__restrict__ out /* [0]=Z, [1..3]=exp(l-l_best), [4..6]=index as float bits */) {
    __shared__ float s_sum[ALT_THREADS / 32];
    __shared__ unsigned long long s_best[ALT_THREADS / 32];
    __shared__ int s_used[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float lbest = logits[best];
    float z = 0.f;
    for (int i = tid; i < VOX_VOCAB_SIZE; i += ALT_THREADS) z += expf(logits[i] - lbest);
    z = vb_warp_sum(z);
    if (lane == 0) s_sum[warp] = z;
    if (tid < 3) s_used[tid] = -1;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int w = 0; w < ALT_THREADS / 32; w++) t += s_sum[w]; out[0] = t; }
    for (int r = 0; r < 3; r++) {
        unsigned long long c = 0ull;
        const int u0 = s_used[0], u1 = s_used[1];
        for (int i = text_min + tid; i < VOX_VOCAB_SIZE; i += ALT_THREADS) {
            if (i == best || i == u0 || i == u1) continue;
            unsigned long long k = pack_cand(logits[i], i);
            if (k > c) c = k;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { unsigned long long other = __shfl_xor_sync(0xffffffffu, c, o); if (other > c) c = other; }
        if (lane == 0) s_best[warp] = c;
        __syncthreads();
        if (tid == 0) {
            unsigned long long m = 0ull;
            for (int w = 0; w < ALT_THREADS / 32; w++) if (s_best[w] > m) m = s_best[w];
            const int idx = m ? cand_index(m) : -1;
            s_used[r] = idx;
            out[1 + r] = idx >= 0 ? expf(logits[idx] - lbest) : 0.f;
            out[4 + r] = __int_as_float(idx);
        }
        __syncthreads();
    }
}

/* host view: z, e[3] (exp(l_i - l_best), descending), idx[3] (-1 = none) for the logits of the last decode step */
extern "C" void vb_alt_candidates(VbEngine *e, int best, int text_min, float *z, float ev[3], int idx[3]) {
    float *d_out = vb_ws(e, VB_WS_ALT, 8 * sizeof(float));
    k_alt_candidates<<<1, ALT_THREADS, 0, e->stream>>>(e->d_logits, best, text_min, d_out);
    e->launches += 1;
    float h[8];
    vb_d2h_sync(e, h, d_out, sizeof h);
    *z = h[0];
    for (int r = 0; r < 3; r++) { ev[r] = h[1 + r]; memcpy(&idx[r], &h[4 + r], 4); }
}
