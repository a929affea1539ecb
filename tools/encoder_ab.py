#!/usr/bin/env python3
"""Encoder A/B on the one-shot shape: the same clip through the tcgen05 attention + persistent GEMM (default) and through the
round-1 kernels (VOX_CUDA_ATTN=simt VOX_CUDA_GEMM=v1), one process per variant; prints encoder ms / positions and the md5 of
the greedy ids (which must agree).   python tools/encoder_ab.py [seconds] [passes]"""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import ensure_synth_model, read_wav_f32, synth_wav
    import vbload
    seconds, passes = float(sys.argv[2]), int(sys.argv[3])
    vb = vbload.load()
    eng = vb.Engine(ensure_synth_model())
    pcm = read_wav_f32(synth_wav(seconds))
    for i in range(passes):
        i0 = eng.info()
        s = eng.stream()
        s.feed(pcm)
        s.finish()
        ids = s.token_ids().copy()
        s.close()
        i1 = eng.info()
        ems = i1["total_encoder_ms"] - i0["total_encoder_ms"]
        pos = i1["total_encoder_positions"] - i0["total_encoder_positions"]
        print(f"  pass {i}: encoder {ems:.2f} ms for {pos} positions = {pos / ems:.1f} k positions/s; {len(ids)} ids md5 "
              f"{hashlib.md5(ids.tobytes()).hexdigest()[:12]}", flush=True)
    eng.close()
    sys.exit(0)

seconds = sys.argv[1] if len(sys.argv) > 1 else "60"
passes = sys.argv[2] if len(sys.argv) > 2 else "3"
only = set(sys.argv[3].split(",")) if len(sys.argv) > 3 else None      # e.g. 0,4: variant indices
for idx, (name, env) in enumerate((("default: tcgen05 attention + persistent GEMM + fused producers (RoPE/planes as epilogues)", {}),
                  ("... without the wq|wk|wv epilogue fusion", {"VOX_CUDA_FUSE_QKV": "0"}),
                  ("... without any fused producer (separate k_split_planes passes)", {"VOX_CUDA_FUSE": "0"}),
                  ("tcgen05 attention + 128x128 GEMM (round 1)", {"VOX_CUDA_GEMM": "v1"}),
                  ("round-1 kernels (CUDA-core attention, 128x128 GEMM)", {"VOX_CUDA_ATTN": "simt", "VOX_CUDA_GEMM": "v1"}))):
    if only is not None and str(idx) not in only:
        continue
    print(f"{name} ({seconds} s clip):", flush=True)
    subprocess.call([sys.executable, os.path.abspath(__file__), "--child", seconds, passes], env={**os.environ, **env})
