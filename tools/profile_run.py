#!/usr/bin/env python3
"""One pass of the pipeline on N seconds of synthetic audio -- the command profiled under ncu.
   python tools/profile_run.py [seconds] [passes] [auto|graph|persist|v2] [n_streams]
With n_streams > 1 the streams run on forked contexts in deferred mode and vox_cuda_streams_decode() advances all decoders
in one persistent kernel (one weight pass for all of them)."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import ensure_synth_model, read_wav_f32, synth_wav  # noqa: E402
import vbload  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mode = sys.argv[3] if len(sys.argv) > 3 else "auto"
n_streams = int(sys.argv[4]) if len(sys.argv) > 4 else 1
vb = vbload.load()
model = ensure_synth_model()
pcm = read_wav_f32(synth_wav(seconds))
eng = vb.Engine(model)
eng.set_decode_mode(mode)
engines = [eng] + [eng.fork() for _ in range(n_streams - 1)]
h = lambda ids: hashlib.md5(ids.tobytes()).hexdigest()[:12]
for _ in range(passes):
    i0 = eng.info()
    streams = [e.stream() for e in engines]
    eng.timer_start()
    if n_streams == 1:
        streams[0].feed(pcm)
        streams[0].finish()
    else:
        for s in streams:
            s.set_deferred(1)
            s.feed(pcm)
        vb.streams_decode(streams)
        for s in streams:
            s.finish()
        vb.streams_decode(streams)
    ms = eng.timer_stop_ms()
    i1 = eng.info()
    ids = [s.token_ids() for s in streams]
    dsteps = i1["total_decode_steps"] - i0["total_decode_steps"]
    dms = i1["total_decode_kernel_ms"] - i0["total_decode_kernel_ms"]
    print(f"{seconds:g}s audio x {n_streams} stream(s), mode {mode}: {len(ids[0])} decoder steps/stream, {ms:.1f} ms device time, "
          f"aggregate RTF {n_streams * seconds / (ms / 1e3):.1f}, decode {dms / max(dsteps, 1):.3f} ms/step over {dsteps} steps, "
          f"ids md5 {[h(x) for x in ids]}", flush=True)
    for s in streams:
        s.close()
for e in engines[1:]:
    e.close()
eng.close()
