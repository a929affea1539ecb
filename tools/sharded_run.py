#!/usr/bin/env python3
"""Sequence-sharded encoder check, launched with torchrun (one rank per GPU; works with one rank too):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
      tools/sharded_run.py [seconds]

Every rank calls vox_cuda_encode_sharded (host C, vb_dist.c: per-layer K/V halo over ncclSend/ncclRecv on the engine's
stream, adapter ncclAllGather); rank 0 then decodes (vox_cuda_decode_adapter) and checks the result against (a) the same
call unsharded on a forked context and (b) the ordinary streaming API -- adapter rows to 1e-4 of their scale and identical
greedy token ids.  Prints one JSON line with the encoder-phase time (max over ranks)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import ensure_synth_model, read_wav_f32, synth_wav  # noqa: E402
import vbload  # noqa: E402


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["VOX_CUDA_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ensure_synth_model(); synth_wav(seconds)
    if world > 1:
        dist.barrier()
    vb = vbload.load()
    multi = vbload.load_submodule("multi")
    L = vb.lib()
    eng = vb.Engine(ensure_synth_model())
    pcm = read_wav_f32(synth_wav(seconds))
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            assert L.vox_cuda_dist_unique_id(buf) == 0
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda(local)
        dist.broadcast(uid, 0)
        assert L.vox_cuda_dist_init(eng.ctx, rank, world, bytes(uid.cpu().tolist())) == 0

    def encode(ctx):
        d_ad = C.c_void_p(); T = C.c_int(); P = C.c_int(); ms = C.c_double()
        assert L.vox_cuda_encode_sharded(ctx, pcm.ctypes.data_as(fp), pcm.size, C.byref(d_ad), C.byref(T), C.byref(P), C.byref(ms)) == 0
        return d_ad, T.value, P.value, ms.value

    encode(eng.ctx)                                                     # warm-up
    if world > 1:
        dist.barrier()
    d_ad, T, P, ms = encode(eng.ctx)
    enc_ms = multi.reduce_max([ms], dist if world > 1 else None, torch.device("cuda", local))[0]
    out = {"seconds": seconds, "world": world, "positions": P, "adapter_tokens": T,
           "encode_ms_max_over_ranks": enc_ms, "positions_per_s": P / (enc_ms / 1e3)}
    if rank == 0:
        ids = np.zeros(T, np.int32)
        n = L.vox_cuda_decode_adapter(eng.ctx, d_ad, T, ids.ctypes.data_as(ip), T)
        ids = ids[:n]
        a = np.empty((T, 3072), np.float32)
        L.vox_cuda_memcpy_d2h(eng.ctx, a.ctypes.data_as(C.c_void_p), d_ad, a.nbytes)
        f = eng.fork()                                                  # no communicator: the same call, unsharded
        encode(f.ctx)
        d1, T1, P1, ms1 = encode(f.ctx)
        b = np.empty((T1, 3072), np.float32)
        L.vox_cuda_memcpy_d2h(f.ctx, b.ctypes.data_as(C.c_void_p), d1, b.nbytes)
        s = f.stream(); s.feed(pcm); s.finish(); ids_stream = s.token_ids().copy(); counts = s.counts(); s.close()
        f.close()
        diff = float(np.abs(a - b).max()) if T == T1 else float("inf"); scale = float(np.abs(b).max())
        out.update({"adapter_max_abs_diff_vs_unsharded": diff, "adapter_scale": scale,
                    "tokens": int(len(ids)), "tokens_equal_stream_api": bool(np.array_equal(ids, ids_stream)),
                    "stream_adapter_tokens": counts["adapter_tokens"], "encode_ms_single_gpu": ms1,
                    "encoder_speedup": ms1 / enc_ms})
        print(json.dumps(out), flush=True)
        assert diff <= 1e-4 * max(scale, 1.0), (diff, scale)
        assert out["tokens_equal_stream_api"] and counts["adapter_tokens"] == T
    if world > 1:
        dist.barrier()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
