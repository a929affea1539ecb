#!/usr/bin/env python3
"""Dump the greedy token ids of a synthetic recording (for draft-acceptance statistics).  python tools/dump_tokens.py [seconds]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import ensure_synth_model, read_wav_f32, synth_wav
import vbload
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
vb = vbload.load()
eng = vb.Engine(ensure_synth_model())
s = eng.stream(); s.feed(read_wav_f32(synth_wav(seconds))); s.finish()
ids = s.token_ids().copy(); s.close(); eng.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"tokens_{seconds:g}s.npy"), ids)
rep = float((ids[1:] == ids[:-1]).mean())
print(f"{len(ids)} tokens, P(next == current) = {rep:.3f}, distinct = {len(set(ids.tolist()))}")
