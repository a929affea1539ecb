#!/usr/bin/env python3
"""Live-stream shape (main.c --stdin / --from-mic with -I 0.1): 0.1 s feeds with a 0.1 s processing interval.  After the first
chunk every vox_stream_feed() runs one tiny encoder call (~10 mel frames -> 5 positions) and 1-2 decode steps.
   python tools/live_run.py [seconds]
Prints the host wall time per feed call (median / p90 / max), the kernel launches per call and the real-time margin."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import ensure_synth_model, read_wav_f32, synth_wav  # noqa: E402
import vbload  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
vb = vbload.load()
pcm = read_wav_f32(synth_wav(seconds))
eng = vb.Engine(ensure_synth_model())
for rep in range(2):                                   # first repetition warms workspaces and kernels
    i0 = eng.info()
    s = eng.stream()
    s.set_interval(0.1)
    s.set_continuous(1)
    times, launches = [], []
    for off in range(0, pcm.size, 1600):
        l0 = eng.info()["kernel_launches"]
        t0 = time.perf_counter()
        s.feed(pcm[off:off + 1600])
        s.get()
        times.append((time.perf_counter() - t0) * 1e3)
        launches.append(eng.info()["kernel_launches"] - l0)
    s.finish()
    ids = s.token_ids()
    s.close()
    i1 = eng.info()
t = np.array(times[20:]); la = np.array(launches[20:])   # steady state: after the prompt delay
print(f"live feeding, 0.1 s chunks over {seconds:g} s: {len(times)} feeds, {len(ids)} decoder steps; per feed (steady state): "
      f"median {np.median(t):.2f} ms, p90 {np.percentile(t, 90):.2f} ms, max {t.max():.2f} ms of a 100 ms budget; "
      f"kernel launches per feed: median {int(np.median(la))}, max {int(la.max())}")
enc_ms = i1["total_encoder_ms"] - i0["total_encoder_ms"]; enc_pos = i1["total_encoder_positions"] - i0["total_encoder_positions"]
dec_ms = i1["total_decode_kernel_ms"] - i0["total_decode_kernel_ms"]; dec_steps = i1["total_decode_steps"] - i0["total_decode_steps"]
print(f"  of which (host-observed) encoder+adapter calls {enc_ms:.0f} ms total = {enc_ms / max(len(times), 1):.2f} ms per feed ({enc_pos} positions), "
      f"decode kernels {dec_ms:.0f} ms = {dec_ms / max(dec_steps, 1):.2f} ms per step over {dec_steps} steps; sum of feed times {sum(times):.0f} ms")
eng.close()
