/*
 * vb_ops.cu -- CUDA kernels for the M>1 building blocks and the host-pointer
 * "kernel dispatch surface" (every function of /root/reference voxtral_kernels.h:18-163).
 *
 * The wrappers at the bottom keep the reference's host-pointer semantics
 * (H2D -> kernel -> D2H); they are the per-op parity seam used by tests/, not
 * the fast path.  The fast path calls the vb_* launchers with device pointers.
 */
#include "vb_ops.cuh"
#include <math.h>
#include <string.h>

void vb_launch_count(VbEngine *e, int n) { e->launches += (unsigned long long)n; }

/* ======================================================================
 * SIMT GEMM  C = A * W^T   (f32 activations, bf16 or f32 weights, f32 accumulate)
 * Reference semantics: vox_linear*_bf16 / vox_matmul_t_bf16, voxtral_kernels.c:197-264
 * (bf16 -> f32 by <<16, products and sums in f32).
 * ==================================================================== */
#define GB_M 64
#define GB_N 64
#define GB_K 16

template <typename WT> __device__ __forceinline__ float vb_w2f(WT w);
template <> __device__ __forceinline__ float vb_w2f<float>(float w) { return w; }
template <> __device__ __forceinline__ float vb_w2f<uint16_t>(uint16_t w) { return __uint_as_float((uint32_t)w << 16); }

template <typename WT, int EPI>
__global__ void __launch_bounds__(256)
k_gemm_simt(const float *__restrict__ A, int lda, const WT *__restrict__ W, const float *__restrict__ bias,
            float *__restrict__ C, int ldc, int M, int N, int K) {
    __shared__ float As[GB_K][GB_M + 4];
    __shared__ float Ws[GB_K][GB_N + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
    const int ty = tid / 16, tx = tid % 16;
    const int lrow = tid / 4, lk = (tid % 4) * 4;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += GB_K) {
        {   /* A tile */
            int m = m0 + lrow;
            float v[4] = { 0.f, 0.f, 0.f, 0.f };
            if (m < M) {
                const float *src = A + (size_t)m * lda + k0 + lk;
#pragma unroll
                for (int j = 0; j < 4; j++) if (k0 + lk + j < K) v[j] = src[j];
            }
#pragma unroll
            for (int j = 0; j < 4; j++) As[lk + j][lrow] = v[j];
        }
        {   /* W tile */
            int n = n0 + lrow;
            float v[4] = { 0.f, 0.f, 0.f, 0.f };
            if (n < N) {
                const WT *src = W + (size_t)n * K + k0 + lk;
#pragma unroll
                for (int j = 0; j < 4; j++) if (k0 + lk + j < K) v[j] = vb_w2f<WT>(src[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) Ws[lk + j][lrow] = v[j];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GB_K; k++) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) w[j] = Ws[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; i++) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        if (EPI == VB_EPI_SWIGLU) {
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
                int n = n0 + tx * 4 + j;
                if (n + 1 < N) {
                    float g = acc[i][j], u = acc[i][j + 1];
                    C[(size_t)m * ldc + (n >> 1)] = vb_silu(g) * u;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int n = n0 + tx * 4 + j;
                if (n >= N) continue;
                float v = acc[i][j];
                if (bias) v += bias[n];
                if (EPI == VB_EPI_GELU) v = vb_gelu_tanh(v);
                if (EPI == VB_EPI_RESIDUAL) v += C[(size_t)m * ldc + n];
                C[(size_t)m * ldc + n] = v;
            }
        }
    }
}

template <typename WT>
static void gemm_dispatch(VbEngine *e, const float *A, int lda, const WT *W, const float *bias,
                          float *C, int ldc, int M, int N, int K, int epi) {
    if (M <= 0 || N <= 0) return;
    dim3 grid((N + GB_N - 1) / GB_N, (M + GB_M - 1) / GB_M), block(256);
    switch (epi) {
    case VB_EPI_STORE:    k_gemm_simt<WT, VB_EPI_STORE><<<grid, block, 0, e->stream>>>(A, lda, W, bias, C, ldc, M, N, K); break;
    case VB_EPI_GELU:     k_gemm_simt<WT, VB_EPI_GELU><<<grid, block, 0, e->stream>>>(A, lda, W, bias, C, ldc, M, N, K); break;
    case VB_EPI_RESIDUAL: k_gemm_simt<WT, VB_EPI_RESIDUAL><<<grid, block, 0, e->stream>>>(A, lda, W, bias, C, ldc, M, N, K); break;
    case VB_EPI_SWIGLU:   k_gemm_simt<WT, VB_EPI_SWIGLU><<<grid, block, 0, e->stream>>>(A, lda, W, bias, C, ldc, M, N, K); break;
    }
    VB_CUDA_OK(cudaGetLastError());
    vb_launch_count(e, 1);
}

/* M>1 bf16-weight GEMMs go to the tensor cores (vb_gemm_tc.cu) whenever the shape tiles (N%128, K%64 -- every
 * linear of the model does); VOX_CUDA_GEMM=simt forces the f32 CUDA-core kernel (validation). */
void vb_gemm_tc(VbEngine *e, const float *A, int lda, const uint16_t *W, const float *bias, float *C, int ldc,
                int M, int N, int K, int epi);
int vb_gemm_tc_usable(int M, int N, int K);
static int gemm_use_tc(void) {
    static int mode = -1;
    if (mode < 0) { const char *s = getenv("VOX_CUDA_GEMM"); mode = (s && !strcmp(s, "simt")) ? 0 : 1; }
    return mode;
}
void vb_gemm_bf16w(VbEngine *e, const float *A, int lda, const uint16_t *W, const float *bias,
                   float *C, int ldc, int M, int N, int K, int epi) {
    if (M <= 0 || N <= 0) return;
    if (gemm_use_tc() && M >= 8 && vb_gemm_tc_usable(M, N, K) && (lda % 4) == 0 && (ldc % 4) == 0)
        vb_gemm_tc(e, A, lda, W, bias, C, ldc, M, N, K, epi);
    else if (gemm_use_tc() && M < 8 && vb_gemv_cols_dev(e, A, lda, W, bias, C, ldc, M, N, K, epi))
        return;                                              /* a handful of rows: stream the weights once, GEMV style */
    else
        gemm_dispatch<uint16_t>(e, A, lda, W, bias, C, ldc, M, N, K, epi);
}
void vb_gemm_f32w(VbEngine *e, const float *A, int lda, const float *W, const float *bias,
                  float *C, int ldc, int M, int N, int K, int epi) {
    gemm_dispatch<float>(e, A, lda, W, bias, C, ldc, M, N, K, epi);
}

/* C[M,N] = A[M,K] * B[K,N]  (vox_matmul, voxtral_kernels.c:54-69; no pipeline caller) */
__global__ void k_matmul_nn(const float *A, const float *B, float *C, int M, int K, int N) {
    int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
    if (n >= N || m >= M) return;
    float s = 0.f;
    for (int k = 0; k < K; k++) s = fmaf(A[(size_t)m * K + k], B[(size_t)k * N + n], s);
    C[(size_t)m * N + n] = s;
}

/* ======================================================================
 * RMSNorm rows   (voxtral_kernels.c:346-363; ada multiply voxtral_decoder.c:508-515)
 * ==================================================================== */
__global__ void __launch_bounds__(256)
k_rmsnorm_rows(float *__restrict__ out, const float *__restrict__ x, const float *__restrict__ w,
               const float *__restrict__ ada, int hidden, float eps) {
    __shared__ float red[8];
    const float *xr = x + (size_t)blockIdx.x * hidden;
    float *orow = out + (size_t)blockIdx.x * hidden;
    float ss = 0.f;
    for (int i = threadIdx.x; i < hidden; i += 256) { float v = xr[i]; ss = fmaf(v, v, ss); }
    ss = vb_warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) tot += red[i];
    float rinv = 1.0f / sqrtf(tot / (float)hidden + eps);
    for (int i = threadIdx.x; i < hidden; i += 256) {
        float v = xr[i] * rinv * w[i];
        if (ada) v *= (1.0f + ada[i]);
        orow[i] = v;
    }
}

void vb_rmsnorm_rows(VbEngine *e, float *out, const float *x, const float *w, const float *ada,
                     int rows, int hidden, float eps) {
    if (rows <= 0) return;
    k_rmsnorm_rows<<<rows, 256, 0, e->stream>>>(out, x, w, ada, hidden, eps);
    VB_CUDA_OK(cudaGetLastError());
    vb_launch_count(e, 1);
}

/* The same normalisation written as three bf16 planes [3][rows][hidden] (x = p0 + p1 + p2, vb_tc.cuh): the A operand of the
 * tcgen05 GEMM that follows, without the f32 round trip through HBM.  First loop identical to k_rmsnorm_rows (same partial
 * sums, same rinv), so plane p0+p1+p2 of an element is exactly the f32 value k_rmsnorm_rows would have stored. */
__global__ void __launch_bounds__(256)
k_rmsnorm_rows_planes(uint16_t *__restrict__ planes, const float *__restrict__ x, const float *__restrict__ w, int rows, int hidden, float eps) {
    __shared__ float red[8];
    const float *xr = x + (size_t)blockIdx.x * hidden;
    float ss = 0.f;
    for (int i = threadIdx.x; i < hidden; i += 256) { float v = xr[i]; ss = fmaf(v, v, ss); }
    ss = vb_warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) tot += red[i];
    float rinv = 1.0f / sqrtf(tot / (float)hidden + eps);
    const size_t plane = (size_t)rows * hidden;
    uint16_t *orow = planes + (size_t)blockIdx.x * hidden;
    for (int i = 2 * threadIdx.x; i < hidden; i += 512) {               /* hidden is even */
        float a = xr[i] * rinv * w[i], b = xr[i + 1] * rinv * w[i + 1];
        uint32_t p0 = vb_pack_bf16x2(a, b);
        a -= vb_bf16_lo(p0); b -= vb_bf16_hi(p0);
        uint32_t p1 = vb_pack_bf16x2(a, b);
        a -= vb_bf16_lo(p1); b -= vb_bf16_hi(p1);
        uint32_t p2 = vb_pack_bf16x2(a, b);
        *reinterpret_cast<uint32_t *>(orow + i) = p0;
        *reinterpret_cast<uint32_t *>(orow + plane + i) = p1;
        *reinterpret_cast<uint32_t *>(orow + 2 * plane + i) = p2;
    }
}

void vb_rmsnorm_rows_planes(VbEngine *e, uint16_t *planes, const float *x, const float *w, int rows, int hidden, float eps) {
    if (rows <= 0) return;
    if (hidden & 1) VB_FAIL("vb_rmsnorm_rows_planes: odd hidden size");
    k_rmsnorm_rows_planes<<<rows, 256, 0, e->stream>>>(planes, x, w, rows, hidden, eps);
    VB_CUDA_OK(cudaGetLastError());
    vb_launch_count(e, 1);
}

/* ======================================================================
 * RoPE + K/V scatter   (voxtral_kernels.c:488-526; cache writes voxtral_decoder.c:482-488,
 * voxtral_encoder.c:547-553).  angle = (float)pos * inv_freq[d] in f32, as :494-497.
 * ==================================================================== */
__global__ void k_rope_split(float *__restrict__ qkv, int ldq, int M, int n_q, int n_kv, int hd,
                             const float *__restrict__ inv_freq, int pos0,
                             float *__restrict__ kdst, float *__restrict__ vdst, int dst_row0, int slot_mask) {
    const int half = hd / 2;
    const int pairs_per_row = (n_q + n_kv) * half;          /* q pairs then k pairs */
    const int v_elems = n_kv * hd;
    const int per_row = pairs_per_row + v_elems;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)M * per_row) return;
    int m = (int)(idx / per_row), r = (int)(idx % per_row);
    float *row = qkv + (size_t)m * ldq;
    int drow = dst_row0 + m;
    if (slot_mask >= 0) drow &= slot_mask;
    if (r < pairs_per_row) {
        int head = r / half, d = r % half;
        float ang = (float)(pos0 + m) * inv_freq[d];
        float sn, cs;
        sincosf(ang, &sn, &cs);
        float *p = row + head * hd + 2 * d;
        float x0 = p[0], x1 = p[1];
        float y0 = x0 * cs - x1 * sn, y1 = x0 * sn + x1 * cs;
        if (head < n_q) { p[0] = y0; p[1] = y1; }
        else {
            float *kd = kdst + (size_t)drow * (n_kv * hd) + (head - n_q) * hd + 2 * d;
            kd[0] = y0; kd[1] = y1;
        }
    } else {
        int c = r - pairs_per_row;
        vdst[(size_t)drow * v_elems + c] = row[(n_q + n_kv) * hd + c];
    }
}

void vb_rope_split(VbEngine *e, float *qkv, int ldq, int M, int n_q, int n_kv, int hd, const float *inv_freq,
                   int pos0, float *kdst, float *vdst, int dst_row0, int slot_mask) {
    if (M <= 0) return;
    long long total = (long long)M * ((n_q + n_kv) * (hd / 2) + n_kv * hd);
    int blocks = (int)((total + 255) / 256);
    k_rope_split<<<blocks, 256, 0, e->stream>>>(qkv, ldq, M, n_q, n_kv, hd, inv_freq, pos0, kdst, vdst, dst_row0, slot_mask);
    VB_CUDA_OK(cudaGetLastError());
    vb_launch_count(e, 1);
}

/* ======================================================================
 * Causal sliding-window attention, one warp per (query, head).
 * Semantics: voxtral_kernels.c:412-482 -- keys [max(0,g-W+1), min(g,seq_k-1)], g = q_offset+i
 * (physical index), online softmax, GQA kv_h = h / (H/Hkv).
 * ==================================================================== */
template <int EPL>   /* elements per lane = head_dim / 32 */
__global__ void __launch_bounds__(256)
k_attn_warp(float *__restrict__ out, int ldo, const float *__restrict__ Q, int ldq,
            const float *__restrict__ K, const float *__restrict__ V, int ldkv,
            int seq_q, int seq_k, int n_heads, int n_kv_heads, float scale, int window, int q_offset) {
    const int hd = EPL * 32;
    long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (long long)seq_q * n_heads) return;
    const int i = (int)(w / n_heads), h = (int)(w % n_heads);
    const int kvh = h / (n_heads / n_kv_heads);
    const int g = q_offset + i;
    int k_start = 0;
    if (window > 0 && g - window + 1 > 0) k_start = g - window + 1;
    int k_end = g + 1;
    if (k_end > seq_k) k_end = seq_k;

    float q[EPL], o[EPL];
#pragma unroll
    for (int t = 0; t < EPL; t++) { q[t] = Q[(size_t)i * ldq + h * hd + lane * EPL + t]; o[t] = 0.f; }
    float mx = -1e30f, sum = 0.f;
    for (int j = k_start; j < k_end; j++) {
        const float *kr = K + (size_t)j * ldkv + kvh * hd + lane * EPL;
        const float *vr = V + (size_t)j * ldkv + kvh * hd + lane * EPL;
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < EPL; t++) s = fmaf(q[t], kr[t], s);
        s = vb_warp_sum(s) * scale;
        if (s > mx) {
            float c = expf(mx - s);
            sum = sum * c + 1.0f;
#pragma unroll
            for (int t = 0; t < EPL; t++) o[t] = o[t] * c + vr[t];
            mx = s;
        } else {
            float p = expf(s - mx);
            sum += p;
#pragma unroll
            for (int t = 0; t < EPL; t++) o[t] = fmaf(p, vr[t], o[t]);
        }
    }
    float inv = sum > 0.f ? 1.0f / sum : 0.f;
#pragma unroll
    for (int t = 0; t < EPL; t++) out[(size_t)i * ldo + h * hd + lane * EPL + t] = o[t] * inv;
}

/* Few queries against a long key range (the live stream's encoder calls: ~5 new positions x a 750-key window, where one warp
 * per (query, head) is 160 warps walking 750 keys one after the other: 157 us per layer, profiles/r02_live.md).  One CTA per
 * (query, head); its 8 warps take interleaved blocks of 4 keys (4 K/V rows in flight per warp, one softmax rescale per block),
 * then the 8 partial (max, sum, o[]) states are merged.  Same masking and f32 arithmetic as k_attn_warp; the summation
 * order differs (per-warp partial sums), i.e. results agree to f32 rounding, like the tiled kernel's. */
template <int EPL>
__global__ void __launch_bounds__(256)
k_attn_split(float *__restrict__ out, int ldo, const float *__restrict__ Q, int ldq,
             const float *__restrict__ K, const float *__restrict__ V, int ldkv,
             int seq_q, int seq_k, int n_heads, int n_kv_heads, float scale, int window, int q_offset) {
    constexpr int hd = EPL * 32;
    __shared__ float s_m[8], s_l[8], s_o[8][hd];
    const int i = blockIdx.x / n_heads, h = blockIdx.x % n_heads;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kvh = h / (n_heads / n_kv_heads);
    const int g = q_offset + i;
    int k_start = 0;
    if (window > 0 && g - window + 1 > 0) k_start = g - window + 1;
    int k_end = g + 1;
    if (k_end > seq_k) k_end = seq_k;

    float q[EPL], o[EPL];
#pragma unroll
    for (int t = 0; t < EPL; t++) { q[t] = Q[(size_t)i * ldq + h * hd + lane * EPL + t]; o[t] = 0.f; }
    float mx = -1e30f, sum = 0.f;
    const size_t col = (size_t)kvh * hd + lane * EPL;
    for (int j0 = k_start + warp * 4; j0 < k_end; j0 += 32) {
        float kk[4][EPL], vv[4][EPL], s[4];
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const bool ok = j0 + jj < k_end;
            const size_t off = (size_t)(ok ? j0 + jj : j0) * ldkv + col;
#pragma unroll
            for (int t = 0; t < EPL; t++) { kk[jj][t] = K[off + t]; vv[jj][t] = V[off + t]; }
        }
        float mb = -1e30f;
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            float d = 0.f;
#pragma unroll
            for (int t = 0; t < EPL; t++) d = fmaf(q[t], kk[jj][t], d);
            d = vb_warp_sum(d) * scale;
            s[jj] = j0 + jj < k_end ? d : -1e30f;
            mb = fmaxf(mb, s[jj]);
        }
        if (mb > mx) {
            const float c = expf(mx - mb);
            sum *= c;
#pragma unroll
            for (int t = 0; t < EPL; t++) o[t] *= c;
            mx = mb;
        }
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const float p = s[jj] > -1e29f ? expf(s[jj] - mx) : 0.f;
            sum += p;
#pragma unroll
            for (int t = 0; t < EPL; t++) o[t] = fmaf(p, vv[jj][t], o[t]);
        }
    }
    if (lane == 0) { s_m[warp] = mx; s_l[warp] = sum; }
#pragma unroll
    for (int t = 0; t < EPL; t++) s_o[warp][lane * EPL + t] = o[t];
    __syncthreads();
    for (int d = threadIdx.x; d < hd; d += 256) {
        float M = -1e30f;
#pragma unroll
        for (int w = 0; w < 8; w++) M = fmaxf(M, s_m[w]);
        float L = 0.f, acc = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const float c = s_l[w] > 0.f ? expf(s_m[w] - M) : 0.f;
            L = fmaf(s_l[w], c, L);
            acc = fmaf(s_o[w][d], c, acc);
        }
        out[(size_t)i * ldo + h * hd + d] = L > 0.f ? acc / L : 0.f;
    }
}

/* Tiled flash attention for the encoder shape (head_dim 64, MHA): one CTA = 64 queries of one head, K/V streamed in
 * 64-key tiles through shared memory, S = QK^T and O += PV as 4x4 register blocks, online softmax per row.
 * Same masking as k_attn_warp (keys [max(0,g-W+1), min(g,seq_k-1)], g = q_offset+i); exact f32 arithmetic. */
#define AT_BQ 64
#define AT_BK 64
#define AT_HD 64
#define AT_LD 68                      /* padded leading dimension (floats) of the transposed tiles */
__global__ void __launch_bounds__(256, 2)
k_attn_tile64(float *__restrict__ out, int ldo, const float *__restrict__ Q, int ldq,
              const float *__restrict__ K, const float *__restrict__ V, int ldkv,
              int seq_q, int seq_k, float scale, int window, int q_offset) {
    extern __shared__ float at_smem[];
    float *Qt = at_smem;                         /* [d][q]  */
    float *Kt = Qt + AT_HD * AT_LD;              /* [d][k]  */
    float *Vs = Kt + AT_HD * AT_LD;              /* [k][d]  */
    float *Pt = Vs + AT_BK * AT_LD;              /* [k][q]  */
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int h = blockIdx.y, q0 = blockIdx.x * AT_BQ;
    const int hoff = h * AT_HD;

    for (int i = tid; i < AT_BQ * (AT_HD / 4); i += 256) {          /* Q tile, transposed, pre-scaled */
        int r = i % AT_BQ, c4 = (i / AT_BQ) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < seq_q) v = *reinterpret_cast<const float4 *>(Q + (size_t)(q0 + r) * ldq + hoff + c4);
        Qt[(c4 + 0) * AT_LD + r] = v.x * scale; Qt[(c4 + 1) * AT_LD + r] = v.y * scale;
        Qt[(c4 + 2) * AT_LD + r] = v.z * scale; Qt[(c4 + 3) * AT_LD + r] = v.w * scale;
    }
    float o[4][4], m[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { m[i] = -1e30f; l[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) o[i][j] = 0.f; }

    const int g_first = q_offset + q0, g_last = q_offset + min(q0 + AT_BQ, seq_q) - 1;
    int k_lo = 0;
    if (window > 0 && g_first - window + 1 > 0) k_lo = g_first - window + 1;
    int k_hi = min(g_last + 1, seq_k);                              /* exclusive */
    k_lo = (k_lo / AT_BK) * AT_BK;

    for (int kt = k_lo; kt < k_hi; kt += AT_BK) {
        __syncthreads();                                            /* previous tile fully consumed (also covers Qt) */
        for (int i = tid; i < AT_BK * (AT_HD / 4); i += 256) {
            {   /* K: lanes walk the key index so the transposed shared-memory stores are conflict free
                 * (the strided global reads are L2 hits: every K row is used by ~13 query tiles) */
                int r = i % AT_BK, c4 = (i / AT_BK) * 4;
                float4 kv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kt + r < seq_k) kv = *reinterpret_cast<const float4 *>(K + (size_t)(kt + r) * ldkv + hoff + c4);
                Kt[(c4 + 0) * AT_LD + r] = kv.x; Kt[(c4 + 1) * AT_LD + r] = kv.y;
                Kt[(c4 + 2) * AT_LD + r] = kv.z; Kt[(c4 + 3) * AT_LD + r] = kv.w;
            }
            {   /* V: row-major, coalesced */
                int r = i / (AT_HD / 4), c4 = (i % (AT_HD / 4)) * 4;
                float4 vv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kt + r < seq_k) vv = *reinterpret_cast<const float4 *>(V + (size_t)(kt + r) * ldkv + hoff + c4);
                *reinterpret_cast<float4 *>(Vs + r * AT_LD + c4) = vv;
            }
        }
        __syncthreads();
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) s[i][j] = 0.f;
#pragma unroll 8
        for (int d = 0; d < AT_HD; d++) {
            const float4 a = *reinterpret_cast<const float4 *>(Qt + d * AT_LD + ty * 4);
            const float4 b = *reinterpret_cast<const float4 *>(Kt + d * AT_LD + tx * 4);
            const float av[4] = { a.x, a.y, a.z, a.w }, bv[4] = { b.x, b.y, b.z, b.w };
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
        }
        /* mask + online softmax (rows are shared by the 16 threads with the same ty = one half-warp) */
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int g = q_offset + q0 + ty * 4 + i;
            int lo = 0;
            if (window > 0 && g - window + 1 > 0) lo = g - window + 1;
            float mx = -1e30f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int kj = kt + tx * 4 + j;
                const bool ok = kj >= lo && kj <= g && kj < seq_k;
                s[i][j] = ok ? s[i][j] : -1e30f;
                mx = fmaxf(mx, s[i][j]);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float mn = fmaxf(m[i], mx);
            const float corr = expf(m[i] - mn);
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float p = s[i][j] > -1e29f ? expf(s[i][j] - mn) : 0.f;
                s[i][j] = p; rs += p;
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
            l[i] = l[i] * corr + rs;
            m[i] = mn;
#pragma unroll
            for (int j = 0; j < 4; j++) o[i][j] *= corr;
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            *reinterpret_cast<float4 *>(Pt + (tx * 4 + j) * AT_LD + ty * 4) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < AT_BK; k++) {
            const float4 a = *reinterpret_cast<const float4 *>(Pt + k * AT_LD + ty * 4);
            const float4 b = *reinterpret_cast<const float4 *>(Vs + k * AT_LD + tx * 4);
            const float av[4] = { a.x, a.y, a.z, a.w }, bv[4] = { b.x, b.y, b.z, b.w };
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) o[i][j] = fmaf(av[i], bv[j], o[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = q0 + ty * 4 + i;
        if (r < seq_q) {
            const float inv = l[i] > 0.f ? 1.0f / l[i] : 0.f;
            *reinterpret_cast<float4 *>(out + (size_t)r * ldo + hoff + tx * 4) =
                make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
        }
    }
}

void vb_attention_rows(VbEngine *e, float *out, int ldo, const float *Q, int ldq, const float *K,
                       const float *V, int ldkv, int seq_q, int seq_k, int n_heads, int n_kv_heads,
                       int head_dim, float scale, int window, int q_offset) {
    if (seq_q <= 0) return;
    if (vb_attn_tc_enabled() && vb_attn_tc_usable(seq_q, seq_k, n_heads, n_kv_heads, head_dim, ldq, ldkv, ldo)) {
        vb_attention_tc(e, out, ldo, Q, ldq, K, V, ldkv, seq_q, seq_k, n_heads, scale, window, q_offset, nullptr);
        return;
    }
    if (head_dim == AT_HD && n_heads == n_kv_heads && seq_q >= 16 && (ldq % 4) == 0 && (ldkv % 4) == 0 && (ldo % 4) == 0) {
        static unsigned int attr_done = 0;                          /* one bit per device */
        const unsigned int dev_bit = 1u << (e->device & 31);
        const int smem = 4 * AT_HD * AT_LD * (int)sizeof(float);
        if (!(attr_done & dev_bit)) { VB_CUDA_OK(cudaFuncSetAttribute(k_attn_tile64, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_done |= dev_bit; }
        dim3 grid((seq_q + AT_BQ - 1) / AT_BQ, n_heads);
        k_attn_tile64<<<grid, 256, smem, e->stream>>>(out, ldo, Q, ldq, K, V, ldkv, seq_q, seq_k, scale, window, q_offset);
        VB_CUDA_OK(cudaGetLastError());
        vb_launch_count(e, 1);
        return;
    }
    if (seq_q < 16 && seq_k >= 64 && (long long)seq_q * n_heads <= 65535) {
        const int blocks = seq_q * n_heads;
#define ATT_SPLIT(E) case E: k_attn_split<E><<<blocks, 256, 0, e->stream>>>(out, ldo, Q, ldq, K, V, ldkv, seq_q, seq_k, \
                         n_heads, n_kv_heads, scale, window, q_offset); break;
        switch (head_dim % 32 ? 0 : head_dim / 32) {
            ATT_SPLIT(1) ATT_SPLIT(2) ATT_SPLIT(3) ATT_SPLIT(4) ATT_SPLIT(8)
        default: VB_FAIL("attention head_dim unsupported (need 32,64,96,128,256)");
        }
#undef ATT_SPLIT
        VB_CUDA_OK(cudaGetLastError());
        vb_launch_count(e, 1);
        return;
    }
    long long warps = (long long)seq_q * n_heads;
    int blocks = (int)((warps * 32 + 255) / 256);
#define ATT_CASE(E) case E: k_attn_warp<E><<<blocks, 256, 0, e->stream>>>(out, ldo, Q, ldq, K, V, ldkv, seq_q, seq_k, \
                         n_heads, n_kv_heads, scale, window, q_offset); break;
    switch (head_dim % 32 ? 0 : head_dim / 32) {
        ATT_CASE(1) ATT_CASE(2) ATT_CASE(3) ATT_CASE(4) ATT_CASE(8)
    default: VB_FAIL("attention head_dim unsupported (need 32,64,96,128,256)");
    }
#undef ATT_CASE
    VB_CUDA_OK(cudaGetLastError());
    vb_launch_count(e, 1);
}

/* ======================================================================
 * Elementwise / small kernels (voxtral_kernels.c:30-48, 369-406)
 * ==================================================================== */
enum { EW_ADD, EW_MUL, EW_AXPY, EW_SCALE, EW_SILU, EW_GELU };
template <int OP>
__global__ void k_elementwise(float *a, const float *b, float s, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = a[i];
    if (OP == EW_ADD) v += b[i];
    if (OP == EW_MUL) v *= b[i];
    if (OP == EW_AXPY) v += s * b[i];
    if (OP == EW_SCALE) v *= s;
    if (OP == EW_SILU) v = vb_silu(v);
    if (OP == EW_GELU) v = vb_gelu_tanh(v);
    a[i] = v;
}

__global__ void __launch_bounds__(256) k_softmax_rows(float *x, int cols) {
    __shared__ float red[8];
    __shared__ float bcast;
    float *row = x + (size_t)blockIdx.x * cols;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < cols; i += 256) m = fmaxf(m, row[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) { float t = red[0]; for (int i = 1; i < 8; i++) t = fmaxf(t, red[i]); bcast = t; }
    __syncthreads();
    m = bcast;
    float s = 0.f;
    for (int i = threadIdx.x; i < cols; i += 256) { float v = expf(row[i] - m); row[i] = v; s += v; }
    s = vb_warp_sum(s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; i++) t += red[i]; bcast = 1.0f / t; }
    __syncthreads();
    float inv = bcast;
    for (int i = threadIdx.x; i < cols; i += 256) row[i] *= inv;
}

/* Naive conv1d with symmetric padding (vox_conv1d, voxtral_kernels.c:270-291; no pipeline caller). */
__global__ void k_conv1d_naive(float *out, const float *in, const float *w, const float *bias,
                               int cin, int cout, int length, int ks, int stride, int padding, int out_len) {
    int ol = blockIdx.x * blockDim.x + threadIdx.x, oc = blockIdx.y;
    if (ol >= out_len || oc >= cout) return;
    float s = bias ? bias[oc] : 0.f;
    for (int ic = 0; ic < cin; ic++)
        for (int k = 0; k < ks; k++) {
            int il = ol * stride - padding + k;
            if (il >= 0 && il < length) s = fmaf(in[(size_t)ic * length + il], w[((size_t)oc * cin + ic) * ks + k], s);
        }
    out[(size_t)oc * out_len + ol] = s;
}

/* Causal conv on channel-major host tensors (vox_causal_conv1d, voxtral_kernels.c:293-340):
 * left pad = ks - stride, out_len = ceil((L - ks + pad)/stride + 1), OOB taps read 0. */
__global__ void k_causal_conv1d_cm(float *out, const float *in, const float *w, const float *bias,
                                   int cin, int cout, int length, int ks, int stride, int out_len) {
    int ol = blockIdx.x * blockDim.x + threadIdx.x, oc = blockIdx.y;
    if (ol >= out_len || oc >= cout) return;
    int left = ks - stride;
    float s = 0.f;
    for (int ic = 0; ic < cin; ic++)
        for (int k = 0; k < ks; k++) {
            int il = ol * stride - left + k;
            if (il >= 0 && il < length) s = fmaf(in[(size_t)ic * length + il], w[((size_t)oc * cin + ic) * ks + k], s);
        }
    if (bias) s += bias[oc];
    out[(size_t)oc * out_len + ol] = s;
}

/* ======================================================================
 * Host-pointer wrappers: the reference's kernel dispatch surface
 * ==================================================================== */
namespace {
struct Staged {
    VbEngine *e;
    void *ptrs[8]; int n;
    explicit Staged(const char *what) : n(0) { vb_require_gpu(what); e = vb_default_engine(); VB_CUDA_OK(cudaSetDevice(e->device)); }
    float *in(const void *h, size_t bytes) {
        void *d = vb_dev_alloc(bytes);
        if (h) VB_CUDA_OK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, e->stream));
        ptrs[n++] = d;
        return (float *)d;
    }
    void out(void *h, const void *d, size_t bytes) {
        VB_CUDA_OK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, e->stream));
    }
    ~Staged() { cudaStreamSynchronize(e->stream); for (int i = 0; i < n; i++) cudaFree(ptrs[i]); }
};
template <int OP> void ew(const char *name, float *a, const float *b, float s, int n) {
    if (n <= 0) return;
    Staged st(name);
    float *da = st.in(a, (size_t)n * 4), *db = b ? st.in(b, (size_t)n * 4) : nullptr;
    k_elementwise<OP><<<(n + 255) / 256, 256, 0, st.e->stream>>>(da, db, s, n);
    vb_launch_count(st.e, 1);
    st.out(a, da, (size_t)n * 4);
}
}  // namespace

extern "C" {

int vox_verbose = 0;
int vox_monitor = 0;

void vox_add_inplace(float *a, const float *b, int n) { ew<EW_ADD>("vox_add_inplace", a, b, 0.f, n); }
void vox_mul_inplace(float *a, const float *b, int n) { ew<EW_MUL>("vox_mul_inplace", a, b, 0.f, n); }
void vox_axpy(float *a, float s, const float *b, int n) { ew<EW_AXPY>("vox_axpy", a, b, s, n); }
void vox_scale(float *x, float s, int n) { ew<EW_SCALE>("vox_scale", x, nullptr, s, n); }
void vox_silu(float *x, int n) { ew<EW_SILU>("vox_silu", x, nullptr, 0.f, n); }
void vox_gelu(float *x, int n) { ew<EW_GELU>("vox_gelu", x, nullptr, 0.f, n); }
void vox_copy(float *dst, const float *src, int n) { if (n > 0) memcpy(dst, src, (size_t)n * sizeof(float)); }

void vox_matmul(float *C, const float *A, const float *B, int M, int K, int N) {
    Staged st("vox_matmul");
    float *dA = st.in(A, (size_t)M * K * 4), *dB = st.in(B, (size_t)K * N * 4), *dC = st.in(nullptr, (size_t)M * N * 4);
    dim3 grid((N + 127) / 128, M);
    k_matmul_nn<<<grid, 128, 0, st.e->stream>>>(dA, dB, dC, M, K, N);
    vb_launch_count(st.e, 1);
    st.out(C, dC, (size_t)M * N * 4);
}

void vox_linear(float *y, const float *x, const float *W, const float *b, int seq_len, int in_dim, int out_dim) {
    Staged st("vox_linear");
    float *dx = st.in(x, (size_t)seq_len * in_dim * 4), *dW = st.in(W, (size_t)out_dim * in_dim * 4);
    float *db = b ? st.in(b, (size_t)out_dim * 4) : nullptr, *dy = st.in(nullptr, (size_t)seq_len * out_dim * 4);
    vb_gemm_f32w(st.e, dx, in_dim, dW, db, dy, out_dim, seq_len, out_dim, in_dim, VB_EPI_STORE);
    st.out(y, dy, (size_t)seq_len * out_dim * 4);
}
void vox_linear_nobias(float *y, const float *x, const float *W, int seq_len, int in_dim, int out_dim) {
    vox_linear(y, x, W, NULL, seq_len, in_dim, out_dim);
}
void vox_matmul_t(float *C, const float *A, const float *B, int M, int K, int N) {
    vox_linear(C, A, B, NULL, M, K, N);
}

void vb_gemv_bf16_dev(VbEngine *e, float *y, const float *x, const uint16_t *W, const float *bias, int K, int N);

void vox_linear_bf16(float *y, const float *x, const uint16_t *W_bf16, const float *b,
                     int seq_len, int in_dim, int out_dim) {
    Staged st("vox_linear_bf16");
    float *dx = st.in(x, (size_t)seq_len * in_dim * 4);
    /* weights that belong to a loaded model are already resident in HBM */
    const uint16_t *dW = (const uint16_t *)vb_find_mirror(st.e, W_bf16);
    if (!dW) dW = (const uint16_t *)st.in(W_bf16, (size_t)out_dim * in_dim * 2);
    float *db = b ? st.in(b, (size_t)out_dim * 4) : nullptr, *dy = st.in(nullptr, (size_t)seq_len * out_dim * 4);
    if (seq_len == 1 && in_dim % 256 == 0)
        vb_gemv_bf16_dev(st.e, dy, dx, dW, db, in_dim, out_dim);      /* the decode-step GEMV path */
    else
        vb_gemm_bf16w(st.e, dx, in_dim, dW, db, dy, out_dim, seq_len, out_dim, in_dim, VB_EPI_STORE);
    st.out(y, dy, (size_t)seq_len * out_dim * 4);
}
void vox_linear_nobias_bf16(float *y, const float *x, const uint16_t *W_bf16, int seq_len, int in_dim, int out_dim) {
    vox_linear_bf16(y, x, W_bf16, NULL, seq_len, in_dim, out_dim);
}
void vox_matmul_t_bf16(float *C, const float *A, const uint16_t *B_bf16, int M, int K, int N) {
    vox_linear_bf16(C, A, B_bf16, NULL, M, K, N);
}

void vox_conv1d(float *out, const float *in, const float *weight, const float *bias,
                int cin, int cout, int length, int ks, int stride, int padding) {
    int out_len = (length + 2 * padding - ks) / stride + 1;
    if (out_len <= 0) return;
    Staged st("vox_conv1d");
    float *din = st.in(in, (size_t)cin * length * 4), *dw = st.in(weight, (size_t)cout * cin * ks * 4);
    float *db = bias ? st.in(bias, (size_t)cout * 4) : nullptr, *dout = st.in(nullptr, (size_t)cout * out_len * 4);
    dim3 grid((out_len + 127) / 128, cout);
    k_conv1d_naive<<<grid, 128, 0, st.e->stream>>>(dout, din, dw, db, cin, cout, length, ks, stride, padding, out_len);
    vb_launch_count(st.e, 1);
    st.out(out, dout, (size_t)cout * out_len * 4);
}

void vox_causal_conv1d(float *out, const float *in, const float *weight, const float *bias,
                       int cin, int cout, int length, int ks, int stride) {
    int pad_total = ks - stride;
    float nf = ((float)length - ks + pad_total) / (float)stride + 1.0f;
    int out_len = (int)ceilf(nf);
    if (out_len <= 0) return;
    Staged st("vox_causal_conv1d");
    float *din = st.in(in, (size_t)cin * length * 4), *dw = st.in(weight, (size_t)cout * cin * ks * 4);
    float *db = bias ? st.in(bias, (size_t)cout * 4) : nullptr, *dout = st.in(nullptr, (size_t)cout * out_len * 4);
    dim3 grid((out_len + 127) / 128, cout);
    k_causal_conv1d_cm<<<grid, 128, 0, st.e->stream>>>(dout, din, dw, db, cin, cout, length, ks, stride, out_len);
    vb_launch_count(st.e, 1);
    st.out(out, dout, (size_t)cout * out_len * 4);
}

void vox_rms_norm(float *out, const float *x, const float *weight, int seq_len, int hidden, float eps) {
    if (seq_len <= 0) return;
    Staged st("vox_rms_norm");
    float *dx = st.in(x, (size_t)seq_len * hidden * 4), *dw = st.in(weight, (size_t)hidden * 4);
    float *dout = st.in(nullptr, (size_t)seq_len * hidden * 4);
    vb_rmsnorm_rows(st.e, dout, dx, dw, nullptr, seq_len, hidden, eps);
    st.out(out, dout, (size_t)seq_len * hidden * 4);
}

void vox_softmax(float *x, int rows, int cols) {
    if (rows <= 0 || cols <= 0) return;
    Staged st("vox_softmax");
    float *dx = st.in(x, (size_t)rows * cols * 4);
    k_softmax_rows<<<rows, 256, 0, st.e->stream>>>(dx, cols);
    vb_launch_count(st.e, 1);
    st.out(x, dx, (size_t)rows * cols * 4);
}

void vox_causal_attention(float *out, const float *Q, const float *K, const float *V,
                          int seq_q, int seq_k, int n_heads, int n_kv_heads,
                          int head_dim, float scale, int window_size, int q_offset) {
    if (seq_q <= 0) return;
    Staged st("vox_causal_attention");
    size_t qb = (size_t)seq_q * n_heads * head_dim * 4, kb = (size_t)seq_k * n_kv_heads * head_dim * 4;
    float *dQ = st.in(Q, qb), *dK = st.in(K, kb), *dV = st.in(V, kb), *dO = st.in(nullptr, qb);
    vb_attention_rows(st.e, dO, n_heads * head_dim, dQ, n_heads * head_dim, dK, dV, n_kv_heads * head_dim,
                      seq_q, seq_k, n_heads, n_kv_heads, head_dim, scale, window_size, q_offset);
    st.out(out, dO, qb);
}

/* RoPE tables are host f32 in the reference API (voxtral_kernels.c:488-501): computed here with the
 * same expressions; the device kernels use the same inv_freq values and angle arithmetic. */
void vox_compute_rope_freqs(float *freqs, const int *pos, int seq, int dim, float theta) {
    int half = dim / 2;
    for (int s = 0; s < seq; s++) {
        float p = (float)pos[s];
        for (int d = 0; d < half; d++) {
            float freq = 1.0f / powf(theta, (float)(2 * d) / (float)dim);
            float ang = p * freq;
            freqs[((size_t)s * half + d) * 2] = cosf(ang);
            freqs[((size_t)s * half + d) * 2 + 1] = sinf(ang);
        }
    }
}

__global__ void k_apply_rope_table(float *x, const float *freqs, int seq, int heads, int hd) {
    int half = hd / 2;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)seq * heads * half) return;
    int d = (int)(idx % half), h = (int)((idx / half) % heads), s = (int)(idx / ((long long)half * heads));
    float cs = freqs[((size_t)s * half + d) * 2], sn = freqs[((size_t)s * half + d) * 2 + 1];
    float *p = x + ((size_t)s * heads + h) * hd + 2 * d;
    float x0 = p[0], x1 = p[1];
    p[0] = x0 * cs - x1 * sn;
    p[1] = x0 * sn + x1 * cs;
}

void vox_apply_rope(float *x, const float *freqs, int seq, int heads, int head_dim) {
    if (seq <= 0) return;
    Staged st("vox_apply_rope");
    size_t xb = (size_t)seq * heads * head_dim * 4, fb = (size_t)seq * head_dim * 4;
    float *dx = st.in(x, xb), *df = st.in(freqs, fb);
    long long total = (long long)seq * heads * (head_dim / 2);
    k_apply_rope_table<<<(int)((total + 255) / 256), 256, 0, st.e->stream>>>(dx, df, seq, heads, head_dim);
    vb_launch_count(st.e, 1);
    st.out(x, dx, xb);
}

}  /* extern "C" */
