"""Sequence-sharded encoding of ONE long recording over the GPUs of a node (BASELINE.json configs[4]) -- the PROTOTYPE and
the CPU-testable mirror of the host logic.  The product path is host C: csrc/vb_dist.c (vox_cuda_encode_sharded, NCCL on the
engine's stream, no host sync in the layer loop); tests/test_cpu_dist.py checks that its shard plan equals plan_shards() here
and runs this module's exchange schedule over gloo.

Exact, not approximate: rank r owns a contiguous range of encoder positions (aligned to the 4x adapter groups).  In every
layer all ranks first compute q/k/v (+RoPE at GLOBAL positions) for their own rows, then rank r sends its last 750 K/V
rows to rank r+1 (NCCL send/recv over NVLink) -- the only thing the sliding-window attention of the next rank's first 750
queries can see across the boundary -- then each rank finishes the layer.  K/V of a position depend only on that position's
previous-layer state, so the ranks run in lock-step with no serial dependency chain.  The adapter rows are all-gathered and
rank 0 runs the (inherently sequential) decoder.  mel + conv stem are recomputed on every rank (0.01 GFLOP/position).

Host-side plan (pure python, unit-tested on CPU): `plan_shards`.  Device work: C ABI calls of include/voxtral_b200.h
(vox_cuda_encoder_layer_qkv / _rest, vox_cuda_adapter, ...) on torch-owned device buffers.
"""
import ctypes as C
from typing import List, Tuple

ENC_WINDOW = 750


def plan_shards(n_positions: int, world: int) -> List[Tuple[int, int]]:
    """[start, end) encoder-position range per rank: contiguous, complete over the 4-aligned prefix, boundaries on
    multiples of 4 (adapter groups never straddle ranks), sizes differing by at most 4."""
    tokens = n_positions // 4
    out = []
    for r in range(world):
        a = tokens * r // world
        b = tokens * (r + 1) // world
        out.append((4 * a, 4 * b))
    return out


def halo_rows(start: int) -> int:
    """K/V rows a rank needs from its left neighbour(s)."""
    return min(ENC_WINDOW, start)


def exchange_halo(dist, rank, world, shards, kb, vv, h, M):
    """One layer's exchange: rank r's last halo_rows(start of r+1) K/V rows -> rows [0, h) of rank r+1's buffers.
    kb / vv: [h + M, 2048] torch tensors (any device the process group supports); rows [h, h+M) are this rank's own."""
    if world <= 1:
        return
    ops = []
    if rank + 1 < world:
        nxt = halo_rows(shards[rank + 1][0])
        ops += [dist.P2POp(dist.isend, kb[h + M - nxt:h + M], rank + 1), dist.P2POp(dist.isend, vv[h + M - nxt:h + M], rank + 1)]
    if rank > 0:
        ops += [dist.P2POp(dist.irecv, kb[0:h], rank - 1), dist.P2POp(dist.irecv, vv[0:h], rank - 1)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def gather_adapter(dist, world, shards, a_r):
    """All ranks' adapter rows, in position order.  a_r: [T_max, 3072], rows beyond this rank's own count are padding."""
    import torch
    if world <= 1:
        return a_r[:(shards[0][1] - shards[0][0]) // 4].contiguous()
    parts = [torch.empty_like(a_r) for _ in range(world)]
    dist.all_gather(parts, a_r)
    return torch.cat([parts[r][:(shards[r][1] - shards[r][0]) // 4] for r in range(world)], dim=0).contiguous()


def stream_mel_device(vb, eng, pcm, delay_tokens=6):
    """mel frames exactly as the stream path sees a complete recording (left pad, flush padding, finish)."""
    L = vb.lib()
    mel = L.vox_mel_ctx_init(32 * 1280)
    L.vox_mel_feed(mel, pcm.ctypes.data_as(C.POINTER(C.c_float)), pcm.size)
    align = (1280 - pcm.size % 1280) % 1280
    L.vox_cuda_mel_feed_zeros(mel, align + (delay_tokens + 1 + 10) * 1280)
    L.vox_mel_finish(mel, 0)
    n = C.c_int()
    d = L.vox_cuda_mel_device_frames(mel, C.byref(n))
    return mel, d, n.value


def sharded_encode(vb, eng, pcm, dist, rank, world):
    """Returns (adapter rows as a torch CUDA tensor [T,3072] on every rank, timing dict)."""
    import torch
    L = vb.lib()
    ctx = eng.ctx
    dev = torch.device("cuda", torch.cuda.current_device())
    t_all = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    e0, e1 = ev(), ev()
    e0.record()
    mel, d_mel, F = stream_mel_device(vb, eng, pcm)
    P = F // 2
    x_all = torch.empty(((F + 1) // 2, 1280), dtype=torch.float32, device=dev)
    L.vox_cuda_mel_conv_stem(ctx, d_mel, F, x_all.data_ptr())
    L.vox_cuda_sync(ctx)
    L.vox_mel_free(mel)
    shards = plan_shards(P, world)
    p0, p1 = shards[rank]
    M = p1 - p0
    h = halo_rows(p0)
    if world > 1:
        assert all(b - a >= ENC_WINDOW for a, b in shards), "every shard must hold at least one attention window"
    x = x_all[p0:p1].contiguous()
    del x_all
    kb = torch.zeros((h + M, 2048), dtype=torch.float32, device=dev)
    vv = torch.zeros((h + M, 2048), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()        # the library writes these buffers on ITS stream: the memsets on torch's stream must have landed
    for layer in range(32):
        L.vox_cuda_encoder_layer_qkv(ctx, layer, x.data_ptr(), M, p0, kb.data_ptr(), vv.data_ptr(), h)
        L.vox_cuda_sync(ctx)
        if world > 1:
            exchange_halo(dist, rank, world, shards, kb, vv, h, M)
            torch.cuda.synchronize()
        L.vox_cuda_encoder_layer_rest(ctx, layer, x.data_ptr(), M, kb.data_ptr(), vv.data_ptr(), h)
        L.vox_cuda_sync(ctx)
    L.vox_cuda_encoder_final_norm(ctx, x.data_ptr(), M)
    T_max = max((b - a) // 4 for a, b in shards)
    a_r = torch.zeros((T_max, 3072), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    L.vox_cuda_adapter(ctx, x.data_ptr(), M, a_r.data_ptr())
    L.vox_cuda_sync(ctx)
    adapter = gather_adapter(dist, world, shards, a_r)
    e1.record(); torch.cuda.synchronize()
    t_all["encode_ms"] = e0.elapsed_time(e1)
    t_all["positions"] = P
    t_all["shard"] = (p0, p1)
    return adapter, t_all


def decode_from_adapter(vb, eng, adapter, delay_tokens=6):
    """Greedy decode of a complete adapter sequence on this rank (prefill 38 prompt rows, then device-side loop)."""
    import numpy as np
    import torch
    L = vb.lib()
    ctx = eng.ctx
    T = adapter.shape[0]
    prompt_len = 1 + 32 + delay_tokens
    if T < prompt_len:
        return np.zeros(0, np.int32)
    eng.reset_caches()
    pre = prompt_len - 1
    prompt = torch.empty((pre, 3072), dtype=torch.float32, device=adapter.device)
    torch.cuda.synchronize()
    L.vox_cuda_build_prompt(ctx, prompt.data_ptr(), adapter.data_ptr(), pre)
    L.vox_cuda_decoder_prefill(ctx, prompt.data_ptr(), pre)
    out = np.zeros(T - pre, np.int32)
    n = L.vox_cuda_decoder_steps(ctx, adapter.data_ptr(), pre, T - pre, 32, out.ctypes.data_as(C.POINTER(C.c_int)))
    return out[:n]
