#!/usr/bin/env python3
"""Decode-kernel sweep on one loaded engine: python tools/v2_sweep.py [seconds] [config ...]
A config is mode[:streams][:ENV=VAL,...], e.g. persist  v2  v2:1:VOX_CUDA_V2_INFLIGHT=2  v2:8  v2:1:VOX_CUDA_V2_PROF=100"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import ensure_synth_model, read_wav_f32, synth_wav  # noqa: E402
import vbload  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
configs = sys.argv[2:] or ["persist", "v2"]
vb = vbload.load()
pcm = read_wav_f32(synth_wav(seconds))
eng = vb.Engine(ensure_synth_model())
forks = []
h = lambda ids: hashlib.md5(ids.tobytes()).hexdigest()[:10]
KEYS = ("VOX_CUDA_V2_INFLIGHT", "VOX_CUDA_V2_DYNAMIC", "VOX_CUDA_V2_PROF", "VOX_CUDA_V2_DBG", "VOX_CUDA_V2_LL")
for cfg in configs:
    parts = cfg.split(":")
    mode, n = parts[0], int(parts[1]) if len(parts) > 1 and parts[1] else 1
    for k in KEYS:
        os.environ.pop(k, None)
    verify = 1
    if len(parts) > 2:
        for kv in parts[2].split(","):
            k, v = kv.split("=")
            if k == "VERIFY":
                verify = int(v)
            else:
                os.environ[k] = v
    eng.set_verify_depth(verify)
    while len(forks) < n - 1:
        forks.append(eng.fork())
    engines = [eng] + forks[:n - 1]
    eng.set_decode_mode(mode)
    best = None
    for rep in range(2):
        i0 = eng.info()
        streams = [e.stream() for e in engines]
        eng.timer_start()
        if n == 1:
            streams[0].feed(pcm); streams[0].finish()
        else:
            for s in streams:
                s.set_deferred(1); s.feed(pcm)
            vb.streams_decode(streams)
            for s in streams:
                s.finish()
            vb.streams_decode(streams)
        ms = eng.timer_stop_ms()
        i1 = eng.info()
        ids = [s.token_ids() for s in streams]
        for s in streams:
            s.close()
        dsteps = i1["total_decode_steps"] - i0["total_decode_steps"]
        dms = i1["total_decode_kernel_ms"] - i0["total_decode_kernel_ms"]
        os.environ.pop("VOX_CUDA_V2_PROF", None)          # profile the first repetition only
        if best is None or dms < best[0]:
            best = (dms, dsteps, ms)
    dms, dsteps, ms = best
    if verify > 1:
        print(f"[sweep] verify depth {verify}: {i1['verify_tokens'] - i0['verify_tokens']} tokens in {i1['verify_passes'] - i0['verify_passes']} weight passes")
    print(f"[sweep] {cfg:40s} {seconds:g}s x {n}: decode {dms / max(dsteps, 1):.4f} ms/step ({dsteps} steps), pass {ms:.1f} ms, "
          f"aggregate RTF {n * seconds / (ms / 1e3):.1f}, ids {sorted(set(h(x) for x in ids))}", flush=True)
for f in forks:
    f.close()
eng.close()
