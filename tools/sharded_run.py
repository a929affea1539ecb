#!/usr/bin/env python3
"""Sequence-sharded encoder demo/check, launched with torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
      tools/sharded_run.py [seconds]

Every rank encodes its slice (K/V halo exchange per layer over NCCL, adapter all-gather); rank 0 then decodes and checks the
result against (a) the same layer API run unsharded on one GPU and (b) the ordinary streaming API -- adapter rows to 1e-4 of
their scale and identical greedy token ids.  Prints one JSON line with the encoder-phase time (max over ranks)."""
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
    sh = vbload.load_submodule("sharded")
    multi = vbload.load_submodule("multi")
    eng = vb.Engine(ensure_synth_model())
    pcm = read_wav_f32(synth_wav(seconds))
    for _ in range(2):                                                  # warm-up + timed
        adapter, t = sh.sharded_encode(vb, eng, pcm, dist if world > 1 else None, rank, world)
    enc_ms = multi.reduce_max([t["encode_ms"]], dist if world > 1 else None, torch.device("cuda", local))[0]
    out = {"seconds": seconds, "world": world, "positions": t["positions"], "adapter_tokens": int(adapter.shape[0]),
           "encode_ms_max_over_ranks": enc_ms, "positions_per_s": t["positions"] / (enc_ms / 1e3)}
    if rank == 0:
        ids = sh.decode_from_adapter(vb, eng, adapter)
        # (a) unsharded run of the same layer API, (b) the streaming API
        single, t1 = sh.sharded_encode(vb, eng, pcm, None, 0, 1)
        single, t1 = sh.sharded_encode(vb, eng, pcm, None, 0, 1)
        diff = float((adapter - single).abs().max()); scale = float(single.abs().max())
        s = eng.stream(); s.feed(pcm); s.finish(); ids_stream = s.token_ids().copy(); counts = s.counts(); s.close()
        out.update({"adapter_max_abs_diff_vs_unsharded": diff, "adapter_scale": scale,
                    "tokens": int(len(ids)), "tokens_equal_stream_api": bool(np.array_equal(ids, ids_stream)),
                    "stream_adapter_tokens": counts["adapter_tokens"], "encode_ms_single_gpu": t1["encode_ms"],
                    "encoder_speedup": t1["encode_ms"] / enc_ms})
        print(json.dumps(out), flush=True)
        assert diff <= 1e-4 * max(scale, 1.0), (diff, scale)
        assert out["tokens_equal_stream_api"] and counts["adapter_tokens"] == adapter.shape[0]
    eng.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
