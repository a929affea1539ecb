/*
 * vb_dist.c -- ONE long recording over the GPUs of a node: sequence-sharded encoder (BASELINE.json configs[4],
 * SURVEY.md section 8e).  Host C; one process per GPU; the only collective on the data path is NCCL over NVLink,
 * issued on the engine's own stream -- there is no host synchronisation inside the layer loop.
 *
 * The encoder is exact under sharding: rank r owns a contiguous, 4-aligned range [p0, p1) of encoder positions.
 *   mel, conv     only the rank's own rows: conv1 output j reads conv0 rows 2j-1..2j+1, conv0 row i reads mel frames i-2..i,
 *                 frame t reads 400 samples at 160 t of the padded signal -- a 3-frame halo that the rank recomputes locally,
 *                 so only its slice of the PCM crosses PCIe (voxtral_audio.c:454-513, voxtral.c:537-715 on a slice)
 *   layer l       [RMSNorm -> wq|wk|wv -> RoPE at GLOBAL positions] for the own rows, then rank r sends its LAST 750
 *                 K and V rows to rank r+1 (ncclSend/ncclRecv in one group): that is all the window-750 attention of
 *                 the next rank can see across the boundary (voxtral_encoder.c:388-406), and layer-l K/V of a position
 *                 depend only on that position's layer-(l-1) state, so the ranks run in lock step with no serial chain
 *   adapter       own rows; then ONE ncclAllGather of [T_max,3072] and a compaction to position order on every rank
 * The reference has no multi-device code at all ("Parallelism strategies: none"); its single-device
 * semantics are voxtral_encoder.c:452-636, which vb_encoder.cu implements and this file only slices.
 *
 * NCCL is resolved with dlopen at vox_cuda_dist_init (the unchanged reference CLI links this library without NCCL).
 * The 128-byte unique id comes from vox_cuda_dist_unique_id() on one rank and reaches the others by whatever channel the
 * launcher has (bench.py: a torch.distributed broadcast; a C launcher: a file or a socket).
 */
#define _GNU_SOURCE
#include "vb_engine.h"

#include <dlfcn.h>
#include <string.h>

typedef struct { char internal[128]; } vb_nccl_id;
typedef void *vb_nccl_comm;
#define VB_NCCL_FLOAT 7

typedef struct VbDist {
    void *lib;
    vb_nccl_comm comm;
    int rank, world;
    int (*GetUniqueId)(vb_nccl_id *);
    int (*CommInitRank)(vb_nccl_comm *, int, vb_nccl_id, int);
    int (*CommDestroy)(vb_nccl_comm);
    int (*Send)(const void *, size_t, int, int, vb_nccl_comm, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, vb_nccl_comm, cudaStream_t);
    int (*AllGather)(const void *, void *, size_t, int, vb_nccl_comm, cudaStream_t);
    int (*GroupStart)(void);
    int (*GroupEnd)(void);
    const char *(*GetErrorString)(int);
} VbDist;

static VbDist g_nccl;                       /* function table, shared by all contexts of the process */

static int nccl_load(void) {
    if (g_nccl.lib) return 0;
    const char *names[] = { getenv("VOX_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    void *h = NULL;
    for (int i = 0; i < 3 && !h; i++) if (names[i]) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "voxtral_b200: cannot load NCCL (%s)\n", dlerror()); return -1; }
#define SYM(field, name) do { *(void **)&g_nccl.field = dlsym(h, name); if (!g_nccl.field) { fprintf(stderr, "voxtral_b200: NCCL symbol %s missing\n", name); dlclose(h); return -1; } } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(AllGather, "ncclAllGather");
    SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = h;
    return 0;
}

#define NCCL_OK(call) do { int r__ = (call); if (r__ != 0) { fprintf(stderr, "voxtral_b200: NCCL error at %s:%d: %s\n", __FILE__, __LINE__, \
    g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?"); vb_cuda_fail(cudaErrorUnknown, __FILE__, __LINE__); } } while (0)

/* ---------------------------------------------------------------- the shard plan (pure host arithmetic) */
/* [p0, p1) of `rank`: contiguous, complete over the 4-aligned prefix of n_positions, boundaries on multiples of 4 so that
 * no adapter group straddles two ranks; sizes differ by at most 4.  halo = K/V rows needed from the left neighbour. */
int vox_cuda_shard_plan(int n_positions, int world, int rank, int *p0, int *p1, int *halo) {
    if (world < 1 || rank < 0 || rank >= world || n_positions < 0) return -1;
    const long long tokens = n_positions / VOX_DOWNSAMPLE;
    const int a = (int)(tokens * rank / world) * VOX_DOWNSAMPLE, b = (int)(tokens * (rank + 1) / world) * VOX_DOWNSAMPLE;
    if (p0) *p0 = a;
    if (p1) *p1 = b;
    if (halo) *halo = a < VOX_ENC_WINDOW ? a : VOX_ENC_WINDOW;
    return 0;
}

int vox_cuda_dist_unique_id(void *out128) {
    if (!out128 || nccl_load() != 0) return -1;
    vb_nccl_id id;
    if (g_nccl.GetUniqueId(&id) != 0) return -1;
    memcpy(out128, &id, sizeof id);
    return 0;
}

int vox_cuda_dist_init(vox_ctx_t *ctx, int rank, int world, const void *id128) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world || nccl_load() != 0) return -1;
    VbEngine *e = vb_engine(ctx);
    if (e->dist) return 0;
    VbDist *d = calloc(1, sizeof *d);
    if (!d) return -1;
    *d = g_nccl;
    d->rank = rank; d->world = world;
    vb_nccl_id id;
    memcpy(&id, id128, sizeof id);
    if (cudaSetDevice(e->device) != cudaSuccess || d->CommInitRank(&d->comm, world, id, rank) != 0) {
        fprintf(stderr, "voxtral_b200: ncclCommInitRank failed (rank %d of %d)\n", rank, world);
        free(d);
        return -1;
    }
    e->dist = d;
    return 0;
}

void vox_cuda_dist_shutdown(vox_ctx_t *ctx) {
    if (!ctx) return;
    VbEngine *e = vb_engine(ctx);
    VbDist *d = (VbDist *)e->dist;
    if (!d) return;
    cudaStreamSynchronize(e->stream);
    d->CommDestroy(d->comm);
    free(d);
    e->dist = NULL;
}

/* ---------------------------------------------------------------- sharded encode */
enum { WSD_X = VB_WS_DIST_X, WSD_KB = VB_WS_DIST_X + 1, WSD_VB = VB_WS_DIST_X + 2 };   /* workspace slots owned by this file (vb_engine.h) */

/* Complete recording -> this rank's adapter rows are computed, all ranks' rows are gathered: *d_adapter_out ([T,3072] f32,
 * device memory owned by the ctx, valid until the next call) holds all T adapter rows in position order on EVERY rank.
 * pcm: host mono 16 kHz, the whole recording (each rank reads only the slice its own frames need).
 * encode_ms: device time of the call on this rank (CUDA events on the engine's stream), max it over ranks yourself.
 * Works with world == 1 (no NCCL needed): the same code path unsharded -- the comparison baseline. */
int vox_cuda_encode_sharded(vox_ctx_t *ctx, const float *pcm, int n_samples, float **d_adapter_out, int *n_tokens,
                            int *n_positions, double *encode_ms) {
    if (!ctx || !pcm || n_samples <= 0 || !d_adapter_out || !n_tokens) return -1;
    VbEngine *e = vb_engine(ctx);
    VbDist *d = (VbDist *)e->dist;
    const int world = d ? d->world : 1, rank = d ? d->rank : 0;
    VB_API_GUARD({ return -1; });
    VB_CUDA_OK(cudaSetDevice(e->device));
    VB_CUDA_OK(cudaEventRecord(e->ev0, e->stream));

    /* the recording's frame count as the stream path would produce it (left pad, flush padding, finish: voxtral.c:1203,1593-1606) */
    const int F = vb_mel_recording_frames(n_samples, ctx->delay_tokens);
    const int P = F / 2;                                   /* stream path: an odd last frame never gets a partner */
    int p0, p1, h;
    vox_cuda_shard_plan(P, world, rank, &p0, &p1, &h);
    const int M = p1 - p0;
    if (world > 1) {
        int q0, q1;
        for (int r = 0; r < world; r++) {
            vox_cuda_shard_plan(P, world, r, &q0, &q1, NULL);
            if (q1 - q0 < VOX_ENC_WINDOW) { fprintf(stderr, "vox_cuda_encode_sharded: recording too short for %d ranks (every shard must hold one attention window)\n", world); VB_API_END; return -1; }
        }
    }

    /* mel frames and conv stem of the rank's own positions only: conv1 output j needs mel frames 2j-3..2j+1, so the rank reads
     * just its slice of the PCM (plus a 3-frame halo it recomputes) */
    float *x = vb_ws(e, WSD_X, (size_t)(M > 0 ? M : 1) * VOX_ENC_DIM * 4);
    float *kb = vb_ws(e, WSD_KB, (size_t)(h + M) * VB_ENC_ATT * 4);
    float *vv = vb_ws(e, WSD_VB, (size_t)(h + M) * VB_ENC_ATT * 4);
    {
        const int f0 = 2 * p0 - 3 > 0 ? 2 * p0 - 3 : 0, f1 = 2 * p1;
        float *d_mel = kb;                                 /* (2M+3) x 128 floats fit in the K scratch, which layer 0 fills later */
        vb_mel_recording_range(e, pcm, n_samples, f0, f1, d_mel);
        vb_conv_stem_range_dev(e, d_mel, f0, F, p0, p1, x);
    }

    /* 32 layers with a K/V halo exchange between the two halves of each */
    int nxt = 0;                                          /* rows the right neighbour needs from this rank */
    if (rank + 1 < world) vox_cuda_shard_plan(P, world, rank + 1, NULL, NULL, &nxt);
    for (int l = 0; l < VOX_ENC_LAYERS; l++) {
        vb_enc_layer_qkv_dev(e, l, x, M, p0, kb, vv, h);
        if (world > 1) {
            NCCL_OK(d->GroupStart());
            if (nxt > 0) {
                NCCL_OK(d->Send(kb + (size_t)(h + M - nxt) * VB_ENC_ATT, (size_t)nxt * VB_ENC_ATT, VB_NCCL_FLOAT, rank + 1, d->comm, e->stream));
                NCCL_OK(d->Send(vv + (size_t)(h + M - nxt) * VB_ENC_ATT, (size_t)nxt * VB_ENC_ATT, VB_NCCL_FLOAT, rank + 1, d->comm, e->stream));
            }
            if (h > 0) {
                NCCL_OK(d->Recv(kb, (size_t)h * VB_ENC_ATT, VB_NCCL_FLOAT, rank - 1, d->comm, e->stream));
                NCCL_OK(d->Recv(vv, (size_t)h * VB_ENC_ATT, VB_NCCL_FLOAT, rank - 1, d->comm, e->stream));
            }
            NCCL_OK(d->GroupEnd());
        }
        vb_enc_layer_rest_dev(e, l, x, M, kb, vv, h);
    }
    vox_cuda_encoder_final_norm(ctx, x, M);

    /* adapter rows of this rank, then all ranks' rows in position order */
    const int T = P / VOX_DOWNSAMPLE;
    int t_max = 0;
    for (int r = 0; r < world; r++) { int q0, q1; vox_cuda_shard_plan(P, world, r, &q0, &q1, NULL); if ((q1 - q0) / VOX_DOWNSAMPLE > t_max) t_max = (q1 - q0) / VOX_DOWNSAMPLE; }
    const size_t row = (size_t)VOX_DEC_DIM * 4;
    if (e->dist_adapter_cap < T) {
        VB_CUDA_OK(cudaStreamSynchronize(e->stream));
        cudaFree(e->d_dist_adapter); e->d_dist_adapter = NULL; e->dist_adapter_cap = 0;
        e->d_dist_adapter = vb_dev_alloc((size_t)(T + 8) * row);
        e->dist_adapter_cap = T + 8;
    }
    if (world == 1) {
        vb_adapter_dev(e, x, M, e->d_dist_adapter);
    } else {
        float *mine = vb_ws(e, WSD_KB, (size_t)t_max * row);                       /* K/V scratch is free again */
        float *all = vb_ws(e, WSD_VB, (size_t)t_max * world * row);
        vb_dzero(e, mine, (size_t)t_max * row);
        vb_adapter_dev(e, x, M, mine);
        NCCL_OK(d->AllGather(mine, all, (size_t)t_max * VOX_DEC_DIM, VB_NCCL_FLOAT, d->comm, e->stream));
        for (int r = 0; r < world; r++) {
            int q0, q1;
            vox_cuda_shard_plan(P, world, r, &q0, &q1, NULL);
            vb_d2d(e, e->d_dist_adapter + (size_t)(q0 / VOX_DOWNSAMPLE) * VOX_DEC_DIM, all + (size_t)r * t_max * VOX_DEC_DIM,
                   (size_t)((q1 - q0) / VOX_DOWNSAMPLE) * row);
        }
    }
    VB_CUDA_OK(cudaEventRecord(e->ev1, e->stream));
    VB_CUDA_OK(cudaEventSynchronize(e->ev1));
    float ms = 0.f;
    VB_CUDA_OK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    if (encode_ms) *encode_ms = ms;
    if (n_positions) *n_positions = P;
    *d_adapter_out = e->d_dist_adapter;
    *n_tokens = T;
    VB_API_END;
    return 0;
}

/* Greedy decode of a complete adapter sequence on this rank (the inherently sequential part: replicas only):
 * prompt prefill (voxtral.c:990-1012) then the device-side loop.  Returns the number of ids written (<= max_ids). */
int vox_cuda_decode_adapter(vox_ctx_t *ctx, const float *d_adapter, int n_tokens, int *out_ids, int max_ids) {
    if (!ctx || !d_adapter || !out_ids) return -1;
    VbEngine *e = vb_engine(ctx);
    const int prompt_len = 1 + 32 + ctx->delay_tokens, pre = prompt_len - 1;
    if (n_tokens < prompt_len) return 0;
    VB_API_GUARD({ return -1; });
    VB_CUDA_OK(cudaSetDevice(e->device));
    vox_cuda_reset_caches(ctx);
    float *prompt = vb_ws(e, 17, (size_t)pre * VOX_DEC_DIM * 4);
    vb_build_prompt_dev(e, prompt, d_adapter, pre, 1, 32);
    vox_cuda_decoder_prefill(ctx, prompt, pre);
    int want = n_tokens - pre;
    if (want > max_ids) want = max_ids;
    int got = vox_cuda_decoder_steps(ctx, d_adapter, pre, want, 32, out_ids);
    VB_API_END;
    return got;
}
