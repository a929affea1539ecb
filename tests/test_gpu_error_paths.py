"""Reference error semantics at the C boundary (voxtral.c:132-158, 1199-1200, 1237; voxtral_decoder.c:621): an out-of-memory
condition inside the library becomes the entry point's error return, not a process abort.  Allocation failures are injected
with vox_cuda_debug_fail_alloc_after (the n-th device allocation from now on fails)."""
import numpy as np
import pytest

from conftest import read_wav_f32, synth_wav
from test_gpu_stream_parity import run_stream

pytestmark = pytest.mark.gpu


def test_vox_load_returns_null_on_device_oom(vb, model_dir):
    L = vb.lib()
    for n in (0, 5, 200):                     # first allocation, inside the encoder layers, inside the decoder layers
        L.vox_cuda_debug_fail_alloc_after(n)
        try:
            ctx = L.vox_load(model_dir.encode())
        finally:
            L.vox_cuda_debug_fail_alloc_after(-1)
        assert not ctx, f"vox_load must return NULL when allocation {n} fails"


def test_stream_calls_return_minus_one_on_device_oom(vb, engine):
    L = vb.lib()
    pcm = read_wav_f32(synth_wav(2))
    want, _, _ = run_stream(engine, pcm)
    # 1. vox_stream_init: its first device allocation fails -> NULL
    L.vox_cuda_debug_fail_alloc_after(0)
    try:
        s = L.vox_stream_init(engine.ctx)
    finally:
        L.vox_cuda_debug_fail_alloc_after(-1)
    assert not s
    # 2. vox_stream_feed: the adapter-row buffer cannot be allocated -> -1, and the stream stays failed
    s = engine.stream()
    L.vox_cuda_debug_fail_alloc_after(0)
    try:
        rc = s.feed(pcm)
    finally:
        L.vox_cuda_debug_fail_alloc_after(-1)
    assert rc == -1
    assert s.feed(pcm[:1600]) == -1 and s.flush() == -1 and s.finish() == -1
    s.close()
    # 3. the context is still usable for a new stream, with the same ids as before
    again, _, _ = run_stream(engine, pcm)
    assert again.tolist() == want.tolist()
