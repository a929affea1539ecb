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

void vb_gemm_bf16w(VbEngine *e, const float *A, int lda, const uint16_t *W, const float *bias,
                   float *C, int ldc, int M, int N, int K, int epi) {
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

void vb_attention_rows(VbEngine *e, float *out, int ldo, const float *Q, int ldq, const float *K,
                       const float *V, int ldkv, int seq_q, int seq_k, int n_heads, int n_kv_heads,
                       int head_dim, float scale, int window, int q_offset) {
    if (seq_q <= 0) return;
    long long warps = (long long)seq_q * n_heads;
    int blocks = (int)((warps * 32 + 255) / 256);
#define ATT_CASE(E) case E: k_attn_warp<E><<<blocks, 256, 0, e->stream>>>(out, ldo, Q, ldq, K, V, ldkv, seq_q, seq_k, \
                         n_heads, n_kv_heads, scale, window, q_offset); break;
    switch (head_dim / 32) {
        ATT_CASE(1) ATT_CASE(2) ATT_CASE(3) ATT_CASE(4) ATT_CASE(8)
    default:
        fprintf(stderr, "voxtral_b200: attention head_dim %d unsupported (need 32,64,96,128,256)\n", head_dim);
        abort();
    }
#undef ATT_CASE
    if (head_dim % 32) { fprintf(stderr, "voxtral_b200: attention head_dim %d not a multiple of 32\n", head_dim); abort(); }
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
