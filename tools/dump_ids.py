#!/usr/bin/env python3
"""One one-shot pass of N seconds of the synthetic recording; saves the greedy ids to gpurun_out/ids_<tag>_<N>s.npy
(for an offline comparison with a reference trace).   [env switches] python tools/dump_ids.py <seconds> <tag>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import ensure_synth_model, read_wav_f32, synth_wav  # noqa: E402
import vbload  # noqa: E402

seconds, tag = float(sys.argv[1]), sys.argv[2]
vb = vbload.load()
eng = vb.Engine(ensure_synth_model())
s = eng.stream()
s.feed(read_wav_f32(synth_wav(seconds)))
s.finish()
ids = s.token_ids().copy()
s.close()
eng.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = os.path.join(ROOT, "gpurun_out", f"ids_{tag}_{seconds:g}s.npy")
np.save(out, ids.astype(np.int32))
print(out, len(ids))
