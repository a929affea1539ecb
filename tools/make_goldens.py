#!/usr/bin/env python3
"""Turn a oracle/_ref/ref_trace output directory into a compact fixture under tests/golden/.

  python tools/make_goldens.py <trace_dir> <name> [slim]     ->  tests/golden/<name>.npz

The trace comes from the UNMODIFIED reference (oracle/ref_trace.c interposes its model-block entry
points), run in the build container on the seeded synthetic checkpoint + PCM:
  tools/_build/make_synth_model /dev/shm/voxsynth ; python tools/make_synth_tekken.py /dev/shm/voxsynth
  tools/_build/make_synth_wav /dev/shm/voxsynth/s2.wav 2
  oracle/_ref/ref_trace /dev/shm/voxsynth /dev/shm/voxsynth/s2.wav /tmp/trace_s2 [feed_chunk]
Full tensors are megabytes, so the fixture keeps: all token ids, per-step top-8 logits and 64 probe
logits, and for every block boundary per-row sums plus a few complete rows.
"""
import json
import os
import sys

import numpy as np


def rows_digest(a, keep=(0, 1, -2, -1)):
    idx = sorted(set(i % len(a) for i in keep))
    return {"shape": np.array(a.shape), "row_sum": a.sum(axis=1, dtype=np.float64).astype(np.float32),
            "row_abs": np.abs(a).sum(axis=1, dtype=np.float64).astype(np.float32),
            "rows_idx": np.array(idx), "rows": a[idx]}


def main(trace, name, slim=False):
    """slim: long traces keep ids, top-8 logits, text and drain counts only (no block digests, no probe logits)."""
    man = json.load(open(os.path.join(trace, "trace.json")))
    out = {"samples": np.array(man["samples"]), "feed_chunk": np.array(man["feed_chunk"]),
           "probe_ids": np.array(man["probe_ids"], dtype=np.int32)}
    f32 = lambda fn, w: np.fromfile(os.path.join(trace, fn), dtype=np.float32).reshape(-1, w)
    out["tokens"] = np.fromfile(os.path.join(trace, "tokens.i32"), dtype=np.int32)
    out["top_val"] = f32("logits_top.f32", 8)
    out["top_idx"] = np.fromfile(os.path.join(trace, "logits_top.i32"), dtype=np.int32).reshape(-1, 8)
    if not slim:
        out["probe_val"] = f32("logits_probe.f32", 64)
    out["text"] = np.frombuffer(open(os.path.join(trace, "text.txt"), "rb").read(), dtype=np.uint8)
    for k in range(0 if slim else man["encoder_calls"]):
        for kind, w in (("enc_in", 1280), ("enc_out", 1280)):
            for kk, v in rows_digest(f32(f"{kind}_{k}.f32", w)).items():
                out[f"{kind}_{k}_{kk}"] = v
    if not slim:
        for k in range(man["adapter_calls"]):
            for kk, v in rows_digest(f32(f"adapter_{k}.f32", 3072)).items():
                out[f"adapter_{k}_{kk}"] = v
        for kk, v in rows_digest(f32("prefill_embed.f32", 3072)).items():
            out[f"prefill_embed_{kk}"] = v
        se = f32("step_embed.f32", 3072)
        for kk, v in rows_digest(se, keep=(0, 1, 2, -1)).items():
            out[f"step_embed_{kk}"] = v
    out["n_prefills"] = np.array(man.get("prefills", 0))
    # scenario traces (TRACE_* knobs of ref_trace): how many positions each drain returned, and the alternatives table
    dp = os.path.join(trace, "drain.txt")
    if os.path.exists(dp):
        rows = [l.split() for l in open(dp).read().splitlines() if l.strip()]
        out["drain_tag"] = np.array([r[0] for r in rows]); out["drain_n"] = np.array([int(r[1]) for r in rows], dtype=np.int32)
    ap = os.path.join(trace, "alt.txt")
    if os.path.exists(ap) and os.path.getsize(ap):
        out["alt"] = np.frombuffer(open(ap, "rb").read(), dtype=np.uint8)
    out["n_encoder_calls"] = np.array(man["encoder_calls"])
    out["n_adapter_calls"] = np.array(man["adapter_calls"])
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", name + ".npz")
    np.savez_compressed(dst, **out)
    print(dst, os.path.getsize(dst), "bytes;", len(out["tokens"]), "steps")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], slim=len(sys.argv) > 3 and sys.argv[3] == "slim")
