/*
 * vb_stream_dev.cu -- small device helpers used by the host-side stream state machine
 * (vb_stream.c): async copies on the engine stream, the prompt-embedding kernel and the
 * one-shot conv stem of vox_encoder_forward.
 */
#include "vb_ops.cuh"

extern "C" void vb_d2d(VbEngine *e, void *dst, const void *src, size_t bytes) {
    if (bytes) VB_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, e->stream));
}
extern "C" void vb_dzero(VbEngine *e, void *dst, size_t bytes) {
    if (bytes) VB_CUDA_OK(cudaMemsetAsync(dst, 0, bytes, e->stream));
}
extern "C" void vb_h2d(VbEngine *e, void *dst, const void *src, size_t bytes) {
    if (bytes) VB_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, e->stream));
}
extern "C" void vb_d2h_sync(VbEngine *e, void *dst, const void *src, size_t bytes) {
    if (bytes) VB_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream));
    VB_CUDA_OK(cudaStreamSynchronize(e->stream));
}
extern "C" void vb_sync(VbEngine *e) { VB_CUDA_OK(cudaStreamSynchronize(e->stream)); }

/* prompt_embeds[i] = adapter[i] + tok_embed(i == 0 ? BOS : STREAMING_PAD)   (voxtral.c:990-999) */
__global__ void k_build_prompt(float *out, const float *adapter, const uint16_t *tok_emb, int n, int bos, int pad) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * VOX_DEC_DIM) return;
    int i = (int)(idx / VOX_DEC_DIM), j = (int)(idx % VOX_DEC_DIM);
    int tok = i == 0 ? bos : pad;
    out[idx] = adapter[idx] + __uint_as_float((uint32_t)tok_emb[(size_t)tok * VOX_DEC_DIM + j] << 16);
}

extern "C" void vb_build_prompt_dev(VbEngine *e, float *d_out, const float *d_adapter, int n, int bos, int pad) {
    long long total = (long long)n * VOX_DEC_DIM;
    k_build_prompt<<<(int)((total + 255) / 256), 256, 0, e->stream>>>(d_out, d_adapter, e->d_tok_emb, n, bos, pad);
    VB_CUDA_OK(cudaGetLastError());
    vb_launch_count(e, 1);
}

/* Whole-sequence conv stem (voxtral_encoder.c:150-181): conv0 k3 s1 (left pad 2) + GELU,
 * conv1 k3 s2 (left pad 1, right tap of an odd-length input reads 0) + GELU.
 * d_mel: [F,128] position-major; d_out: [ceil(F/2),1280]. */
extern "C" void vb_conv_stem_full_dev(VbEngine *e, const float *d_mel, int F, float *d_out, int *out_len) {
    int P = (F + 1) / 2;
    *out_len = P;
    if (F <= 0) return;
    float *in0 = vb_ws(e, 10, (size_t)(F + 2) * VOX_MEL_BINS * 4);
    vb_dzero(e, in0, (size_t)2 * VOX_MEL_BINS * 4);
    vb_d2d(e, in0 + 2 * VOX_MEL_BINS, d_mel, (size_t)F * VOX_MEL_BINS * 4);
    float *c0 = vb_ws(e, 11, (size_t)(F + 3) * VOX_ENC_DIM * 4);
    vb_dzero(e, c0, (size_t)VOX_ENC_DIM * 4);                                 /* left pad row */
    vb_conv_view_dev(e, in0, VOX_MEL_BINS, 1, F, e->d_conv0_wk, e->d_conv0_b, c0 + VOX_ENC_DIM, VOX_ENC_DIM);
    vb_dzero(e, c0 + (size_t)(F + 1) * VOX_ENC_DIM, (size_t)2 * VOX_ENC_DIM * 4);   /* right zero taps */
    vb_conv_view_dev(e, c0, VOX_ENC_DIM, 2, P, e->d_conv1_wk, e->d_conv1_b, d_out, VOX_ENC_DIM);
}

/* Conv stem of encoder positions [p0, p1) only (stream semantics: P = F/2 positions, left zero padding at the start of the
 * recording): conv1 output j reads conv0 rows 2j-1..2j+1, conv0 row i reads mel frames i-2..i, so the slice needs mel frames
 * 2*p0-3 .. 2*p1-1 -- a 3-frame halo recomputed locally by a sequence-sharded run (vb_dist.c).  d_mel holds frames
 * [mel_first, ...) of the recording (mel_first = max(0, 2*p0-3) is enough).  Row for row the same GEMM arithmetic as the
 * whole-sequence stem.  d_out: [p1-p0, 1280]. */
extern "C" void vb_conv_stem_range_dev(VbEngine *e, const float *d_mel, int mel_first, int F, int p0, int p1, float *d_out) {
    const int M = p1 - p0;
    if (M <= 0 || 2 * p1 - 1 > F - 1) return;
    const int c_start = 2 * p0 - 1, n0 = 2 * M + 1;                 /* conv0 rows [c_start, c_start + n0) */
    const int m_start = c_start - 2, nm = n0 + 2;                    /* mel frames [m_start, m_start + nm) */
    const int neg = m_start < 0 ? -m_start : 0;
    float *in0 = vb_ws(e, 10, (size_t)nm * VOX_MEL_BINS * 4);
    if (neg) vb_dzero(e, in0, (size_t)neg * VOX_MEL_BINS * 4);
    vb_d2d(e, in0 + (size_t)neg * VOX_MEL_BINS, d_mel + (size_t)(m_start + neg - mel_first) * VOX_MEL_BINS, (size_t)(nm - neg) * VOX_MEL_BINS * 4);
    float *c0 = vb_ws(e, 11, (size_t)n0 * VOX_ENC_DIM * 4);
    vb_conv_view_dev(e, in0, VOX_MEL_BINS, 1, n0, e->d_conv0_wk, e->d_conv0_b, c0, VOX_ENC_DIM);
    if (c_start < 0) vb_dzero(e, c0, (size_t)VOX_ENC_DIM * 4);      /* conv1's left zero pad row, not conv0 of zeros */
    vb_conv_view_dev(e, c0, VOX_ENC_DIM, 2, M, e->d_conv1_wk, e->d_conv1_b, d_out, VOX_ENC_DIM);
}
