#!/usr/bin/env python3
"""One pass of the pipeline on N seconds of synthetic audio -- the command profiled under ncu.
   python tools/profile_run.py [seconds] [passes] [graph|mega]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import ensure_synth_model, read_wav_f32, synth_wav  # noqa: E402
import vbload  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
vb = vbload.load()
model = ensure_synth_model()
pcm = read_wav_f32(synth_wav(seconds))
eng = vb.Engine(model)
if len(sys.argv) > 3:
    eng.set_decode_mode(sys.argv[3])
for _ in range(passes):
    s = eng.stream()
    eng.timer_start()
    s.feed(pcm)
    s.finish()
    ms = eng.timer_stop_ms()
    ids = s.token_ids()
    print(f"{seconds:g}s audio: {len(ids)} decoder steps, {ms:.1f} ms device time, RTF {seconds / (ms / 1e3):.1f}, info {eng.info()}")
    s.close()
eng.close()
