"""Exact multi-token decoding (vox_cuda_set_verify_depth, vb_decode_v2.cu verify mode): several consecutive positions of one
stream per weight pass, longest prefix of right drafts accepted.  Whatever the drafts are, the ids must be those of plain greedy
decoding, i.e. the reference's trace."""
import pytest

from conftest import golden, read_wav_f32, synth_wav
from test_gpu_stream_parity import check_against, run_stream

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("depth", [2, 3, 4, 8])
def test_verify_mode_ids_equal_plain_greedy(engine, depth):
    engine.set_decode_mode("v2")
    pcm = read_wav_f32(synth_wav(2))
    g = golden("synth_s2_oneshot")
    plain, text, _ = run_stream(engine, pcm)
    check_against(g, plain, text)
    i0 = engine.info()
    engine.set_verify_depth(depth)
    try:
        ids, text2, _ = run_stream(engine, pcm)
        ids_chunked, _, _ = run_stream(engine, pcm, chunk=4000)
    finally:
        engine.set_verify_depth(1)
        engine.set_decode_mode("auto")
    i1 = engine.info()
    assert ids.tolist() == plain.tolist() and text2 == text
    assert ids_chunked.tolist() == plain.tolist()
    passes, toks = i1["verify_passes"] - i0["verify_passes"], i1["verify_tokens"] - i0["verify_tokens"]
    assert toks >= len(plain) and 0 < passes <= toks          # every pass yields at least one token
    print(f"verify depth {depth}: {toks} tokens in {passes} weight passes ({toks / passes:.3f} per pass)")


def test_verify_mode_30s(engine):
    """30 s clip (386 steps, reference trace): accepted prefixes never change an id, also across launches and repeated tokens."""
    engine.set_decode_mode("v2")
    g = golden("synth_s30_oneshot")
    pcm = read_wav_f32(synth_wav(30))
    engine.set_verify_depth(4)
    try:
        ids, text, _ = run_stream(engine, pcm)
    finally:
        engine.set_verify_depth(1)
        engine.set_decode_mode("auto")
    check_against(g, ids, text)
