#!/usr/bin/env python3
"""Same recording through several decode drivers: token ids must agree; prints device ms per decode step for each.
   python tools/compare_drivers.py [seconds] [mode ...]      (modes: graph persist v2)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import ensure_synth_model, read_wav_f32, synth_wav  # noqa: E402
import vbload  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
modes = sys.argv[2:] or ["persist", "v2"]
vb = vbload.load()
eng = vb.Engine(ensure_synth_model())
pcm = read_wav_f32(synth_wav(seconds))
ref = None
for mode in modes:
    eng.set_decode_mode(mode)
    for rep in range(2):
        i0 = eng.info()
        s = eng.stream(); eng.timer_start(); s.feed(pcm); s.finish(); ms = eng.timer_stop_ms()
        ids = s.token_ids().copy(); s.close()
        i1 = eng.info()
    dms = i1["total_decode_kernel_ms"] - i0["total_decode_kernel_ms"]; dst = i1["total_decode_steps"] - i0["total_decode_steps"]
    print(f"{mode:8s} {len(ids)} tokens, pass {ms:.1f} ms, decode {dms / max(dst, 1):.4f} ms/step over {dst} steps", flush=True)
    if ref is None:
        ref = ids
    else:
        same = np.array_equal(ref, ids)
        first = -1 if same else int(np.argmax(ref[:min(len(ref), len(ids))] != ids[:min(len(ref), len(ids))]))
        print(f"         tokens equal to {modes[0]}: {same}" + ("" if same else f" (first difference at {first})"), flush=True)
eng.close()
