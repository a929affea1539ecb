#!/usr/bin/env python3
"""Stress the column modes of the v2 decode kernel: many batched / verify runs with varying column counts and clip lengths on one
loaded engine, every result compared with the single-stream ids.  python tools/stress_columns.py [iterations] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import ensure_synth_model, read_wav_f32, synth_wav  # noqa: E402
import vbload  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
vb = vbload.load()
eng = vb.Engine(ensure_synth_model())
eng.set_decode_mode("v2")
clips = {sec: read_wav_f32(synth_wav(sec)) for sec in (2, 2.03, 4)}


def solo(pcm, depth=1):
    eng.set_verify_depth(depth)
    s = eng.stream(); s.feed(pcm); s.finish(); ids = s.token_ids().copy(); s.close()
    eng.set_verify_depth(1)
    return ids


want = {sec: solo(p) for sec, p in clips.items()}
forks = [eng.fork() for _ in range(7)]
bad = 0
for it in range(iters):
    n = int(rng.choice([2, 3, 4, 5, 8]))
    secs = [float(rng.choice(list(clips))) for _ in range(n)]
    engines = [eng] + forks[:n - 1]
    streams = [e.stream() for e in engines]
    for s, sec in zip(streams, secs):
        s.set_deferred(1); s.feed(clips[sec])
    r1 = vb.streams_decode(streams)
    for s in streams:
        s.finish()
    r2 = vb.streams_decode(streams)
    ok = r1 >= 0 and r2 >= 0 and all(np.array_equal(s.token_ids(), want[sec]) for s, sec in zip(streams, secs))
    for s in streams:
        s.close()
    depth = int(rng.choice([2, 3, 4, 8]))
    sec = float(rng.choice(list(clips)))
    okv = np.array_equal(solo(clips[sec], depth), want[sec])
    print(f"iter {it}: {n} streams {secs} -> {'ok' if ok else 'MISMATCH/FAIL ' + str((r1, r2))}; verify depth {depth} on {sec}s -> {'ok' if okv else 'MISMATCH'}", flush=True)
    bad += (not ok) + (not okv)
for f in forks:
    f.close()
eng.close()
print("stress:", "all ok" if bad == 0 else f"{bad} failures")
sys.exit(1 if bad else 0)
