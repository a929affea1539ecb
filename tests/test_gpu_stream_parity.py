"""Whole-pipeline parity through the streaming C API: PCM in -> greedy token ids out, against golden
traces of the UNMODIFIED reference (tests/golden/*.npz, produced by oracle/ref_trace + tools/make_goldens.py
on the same seeded checkpoint and PCM), plus chunking-invariance properties that need no oracle.
"""
import numpy as np
import pytest

from conftest import golden, read_wav_f32, synth_wav

pytestmark = pytest.mark.gpu


def run_stream(engine, pcm, chunk=None, interval=None):
    s = engine.stream()
    if interval is not None:
        s.set_interval(interval)
    if chunk is None:
        s.feed(pcm)
    else:
        for off in range(0, pcm.size, chunk):
            s.feed(pcm[off:off + chunk])
    s.finish()
    text = b"".join(s.get())
    ids = s.token_ids().copy()
    counts = s.counts()
    s.close()
    return ids, text, counts


def margin_ok_positions(g, thr=2e-3):
    return (g["top_val"][:, 0] - g["top_val"][:, 1]) > thr


def check_against(g, ids, text):
    ref_ids = g["tokens"]
    assert len(ids) == len(ref_ids), (len(ids), len(ref_ids))
    # greedy equality; a mismatch is only tolerated at a step where the reference's own top-1/top-2
    # margin is below 2e-3 (and then nothing after it is comparable)
    ok = margin_ok_positions(g)
    for i, (a, b) in enumerate(zip(ids, ref_ids)):
        if a != b:
            assert not ok[i], f"token mismatch at step {i}: {a} vs {b} (reference margin {g['top_val'][i,0]-g['top_val'][i,1]:.3e})"
            # a flip where the reference's own top-2 margin is below 2e-3 is legitimate (the reference flips between its own
            # builds there, runtest.sh:24-26), but it makes every later step incomparable: report it as an expected failure
            # so it stays visible instead of hiding the rest of the run behind a skip
            pytest.xfail(f"near-tie flip at step {i} of {len(ref_ids)} (reference margin {g['top_val'][i,0]-g['top_val'][i,1]:.3e}); "
                         f"{i} steps matched, later steps are not comparable")
    assert text == g["text"].tobytes()


def test_oneshot_tokens_match_reference(engine):
    g = golden("synth_s2_oneshot")
    pcm = read_wav_f32(synth_wav(2))
    assert pcm.size == int(g["samples"])
    ids, text, counts = run_stream(engine, pcm)
    print("ids", ids.tolist())
    assert counts["adapter_tokens"] == 74 and counts["mel_frames"] == 592
    check_against(g, ids, text)


def test_decode_drivers_agree(engine):
    """The per-phase CUDA-graph driver and the persistent kernels are schedules of the same math."""
    pcm = read_wav_f32(synth_wav(2))
    g = golden("synth_s2_oneshot")
    out = {}
    for mode in ("graph", "persist", "v2"):
        engine.set_decode_mode(mode)
        out[mode], text, _ = run_stream(engine, pcm)
        check_against(g, out[mode], text)
    engine.set_decode_mode("auto")
    assert out["graph"].tolist() == out["persist"].tolist() == out["v2"].tolist()


def test_chunked_1s_tokens_match_reference(engine):
    g = golden("synth_s2_chunk1s")
    pcm = read_wav_f32(synth_wav(2))
    ids, text, _ = run_stream(engine, pcm, chunk=16000)
    check_against(g, ids, text)


def test_two_long_feeds_match_oneshot_reference(engine):
    """60 s fed as two 30-s halves: both encoder calls are long (persistent GEMM, tcgen05 attention, fused epilogues) and the
    second one attends over the 750-row K/V tail of the first -- the cached rows go through the tail split of the K planes
    (vb_attention_tc_pre), the path a sharded run's halo rows take.  Ids against the reference's one-shot trace of the clip."""
    g = golden("synth_s60_oneshot")
    pcm = read_wav_f32(synth_wav(60))
    ids, text, _ = run_stream(engine, pcm, chunk=pcm.size // 2)
    check_against(g, ids, text)


def test_10min_oneshot_matches_reference(engine):
    """The benchmark's own workload (BASELINE configs[2]: 10 minutes, one feed -- one encoder call over 30196 positions, 7511
    decoder steps up to KV length 7549) against the unmodified reference's trace of the same recording (~3 h of host time to make,
    tools/make_goldens.py ... slim).  bench.py compares the same ids in every run (`parity_prefix`)."""
    g = golden("synth_s600_oneshot")
    pcm = read_wav_f32(synth_wav(600))
    assert pcm.size == int(g["samples"])
    ids, text, counts = run_stream(engine, pcm)
    assert counts["mel_frames"] == 60392 and counts["adapter_tokens"] == 7549
    ref_ids, margin, runner_up = g["tokens"], g["top_val"][:, 0] - g["top_val"][:, 1], g["top_idx"][:, 1]
    assert len(ids) == len(ref_ids) == 7511
    bad = np.nonzero(ids != ref_ids)[0]
    # Unlike check_against this does not stop at the first near-tie: on this checkpoint the decoder re-converges after a flipped id,
    # so every later step stays comparable.  A differing id must be the reference's runner-up at a step where the reference's own
    # top-2 margin is below 2e-3 (the near-tie rule of this file), and there may be only a handful of them.
    for i in bad:
        print(f"near-tie at step {i}: reference {ref_ids[i]} (margin {margin[i]:.3e}), engine {ids[i]}")
        assert margin[i] < 2e-3 and ids[i] == runner_up[i], f"token mismatch at step {i}: {ids[i]} vs {ref_ids[i]} (reference margin {margin[i]:.3e})"
    assert bad.size <= 4, f"{bad.size} near-tie flips in 7511 steps"
    if bad.size == 0:
        assert text == g["text"].tobytes()


def test_chunking_invariance(engine):
    """The incremental path must not depend on how the caller slices the audio (same mel frames, same
    conv/encoder rows up to f32 reordering) -- the tiny-interval run exercises the conv tails, the odd
    conv0 residual, the 4x adapter residual and the encoder KV tail on every call."""
    pcm = read_wav_f32(synth_wav(2))
    a, ta, ca = run_stream(engine, pcm)
    b, tb, cb = run_stream(engine, pcm, chunk=1600, interval=0.1)
    assert ca["adapter_tokens"] == cb["adapter_tokens"] and ca["mel_frames"] == cb["mel_frames"]
    assert a.tolist() == b.tolist()
    assert ta == tb


def test_feed_after_finish_and_empty(engine):
    s = engine.stream()
    assert s.feed(np.zeros(0, np.float32)) == -1          # n <= 0 -> -1 (voxtral.c:1237)
    s.feed(np.zeros(1600, np.float32))
    assert s.finish() == 0
    assert s.finish() == -1 and s.feed(np.zeros(10, np.float32)) == -1
    s.close()
