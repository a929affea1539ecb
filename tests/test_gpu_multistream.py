"""Several streams sharing one weight pass (vox_cuda_ctx_fork + deferred streams + vox_cuda_streams_decode, vb_decode_v2.cu):
every stream must decode to exactly the ids it produces alone, and to the reference's trace of that clip."""
import numpy as np
import pytest

from conftest import golden, read_wav_f32, synth_wav
from test_gpu_stream_parity import check_against, run_stream

pytestmark = pytest.mark.gpu


def run_batched(vb, engine, pcms, chunk=None):
    engines = [engine] + [engine.fork() for _ in pcms[1:]]
    streams = [e.stream() for e in engines]
    try:
        for s in streams:
            s.set_deferred(1)
        if chunk is None:
            for s, pcm in zip(streams, pcms):
                s.feed(pcm)
            assert vb.streams_decode(streams) >= 0
        else:                                            # live-like: every stream is fed a slice, then one batched decode
            n = max(p.size for p in pcms)
            for off in range(0, n, chunk):
                for s, pcm in zip(streams, pcms):
                    if off < pcm.size:
                        s.feed(pcm[off:off + chunk])
                assert vb.streams_decode(streams) >= 0
        for s in streams:
            s.finish()
        assert vb.streams_decode(streams) >= 0
        out = [(s.token_ids().copy(), b"".join(s.get()), s.counts()) for s in streams]
    finally:
        for s in streams:
            s.close()
        for e in engines[1:]:
            e.close()
    return out


@pytest.mark.parametrize("n", [2, 3, 8])
def test_batched_streams_equal_single_stream(vb, engine, n):
    """n copies of two different clips (32000 and 32480 samples): ids of every column == the single-stream run == the reference."""
    engine.set_decode_mode("v2")
    a, b = read_wav_f32(synth_wav(2)), read_wav_f32(synth_wav(2.03))
    ga, gb = golden("synth_s2_oneshot"), golden("synth_s2p03_oneshot")
    solo_a, text_a, _ = run_stream(engine, a)
    solo_b, text_b, _ = run_stream(engine, b)
    check_against(ga, solo_a, text_a)
    check_against(gb, solo_b, text_b)
    pcms = [a if i % 2 == 0 else b for i in range(n)]
    out = run_batched(vb, engine, pcms)
    for i, (ids, text, counts) in enumerate(out):
        want, wtext = (solo_a, text_a) if i % 2 == 0 else (solo_b, text_b)
        assert ids.tolist() == want.tolist(), f"column {i} of {n} differs from the single-stream run"
        assert text == wtext
    engine.set_decode_mode("auto")


def test_batched_streams_fed_in_slices(vb, engine):
    """0.5-s slices to 4 streams with a batched decode after every round of feeds: same ids as one-shot feeding."""
    engine.set_decode_mode("v2")
    a, b = read_wav_f32(synth_wav(2)), read_wav_f32(synth_wav(2.03))
    solo_a, _, _ = run_stream(engine, a)
    solo_b, _, _ = run_stream(engine, b)
    out = run_batched(vb, engine, [a, b, b, a], chunk=8000)
    for (ids, _, _), want in zip(out, [solo_a, solo_b, solo_b, solo_a]):
        assert ids.tolist() == want.tolist()
    engine.set_decode_mode("auto")


def test_fork_is_independent_of_parent(vb, engine):
    """A fork decodes alone (non-deferred) to the same ids as the parent, and leaves the parent's caches untouched."""
    a = read_wav_f32(synth_wav(2))
    want, _, _ = run_stream(engine, a)
    f = engine.fork()
    try:
        got, _, _ = run_stream(f, a)
    finally:
        f.close()
    assert got.tolist() == want.tolist()
    again, _, _ = run_stream(engine, a)
    assert again.tolist() == want.tolist()
