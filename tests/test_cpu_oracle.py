"""Pins the CPU restatement (oracle/vox_oracle.c) against the UNMODIFIED reference compiled into
oracle/_ref/libvoxref.so, function by function, on seeded inputs.  CPU only.

The reference build is -O3 -ffast-math (+ OpenBLAS for M>1), so "same arithmetic" means: bit-exact for the
purely elementwise f32 ops, ~1e-6 relative where the reference's summation order is not canonical.
"""
import ctypes as C
import os
import subprocess
import wave

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)
u16p = C.POINTER(C.c_uint16)


def P(a):
    return a.ctypes.data_as(fp)


@pytest.fixture(scope="module")
def orc():
    so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    src = os.path.join(ROOT, "oracle", "vox_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    L = C.CDLL(so)
    L.orc_stream_mel.restype = C.c_int
    L.orc_causal_conv1d_out_len.restype = C.c_int
    L.orc_argmax.restype = C.c_int
    return L


def bf16(rng, shape, scale):
    w = (rng.uniform(-1, 1, size=shape) * scale).astype(np.float32)
    u = w.view(np.uint32)
    return np.ascontiguousarray(((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16))


def rel(a, b):
    return float(np.abs(a - b).max() / max(float(np.abs(b).max()), 1e-9))


def test_elementwise_and_activations(orc, ref):
    rng = np.random.default_rng(0)
    x = (rng.normal(size=4000) * 3).astype(np.float32)
    y = rng.normal(size=4000).astype(np.float32)
    for mine, theirs in (("orc_add", "vox_add_inplace"), ("orc_mul", "vox_mul_inplace")):
        a, b = x.copy(), x.copy()
        getattr(orc, mine)(P(a), P(y), C.c_int(a.size)); getattr(ref.L, theirs)(P(b), P(y), a.size)
        assert np.array_equal(a, b)
    for mine, theirs in (("orc_silu", "vox_silu"), ("orc_gelu", "vox_gelu")):
        a, b = x.copy(), x.copy()
        getattr(orc, mine)(P(a), C.c_int(a.size)); getattr(ref.L, theirs)(P(b), b.size)
        assert rel(a, b) < 2e-6, mine
    a, b = x.reshape(4, 1000).copy(), x.reshape(4, 1000).copy()
    orc.orc_softmax(P(a), 4, 1000); ref.L.vox_softmax(P(b), 4, 1000)
    assert rel(a, b) < 2e-6


@pytest.mark.parametrize("M,K,N,bias", [(1, 3072, 64, False), (1, 1280, 48, True), (7, 512, 40, True)])
def test_linear_bf16(orc, ref, M, K, N, bias):
    rng = np.random.default_rng(M * 100 + N)
    x = rng.normal(size=(M, K)).astype(np.float32)
    W = bf16(rng, (N, K), np.sqrt(3.0 / K))
    b = (rng.normal(size=N) * 0.1).astype(np.float32) if bias else None
    ya, yb = np.empty((M, N), np.float32), np.empty((M, N), np.float32)
    orc.orc_linear_bf16(P(ya), P(x), W.ctypes.data_as(u16p), P(b) if bias else None, M, K, N)
    ref.L.vox_linear_bf16(P(yb), P(x), W.ctypes.data_as(u16p), P(b) if bias else None, M, K, N)
    assert rel(ya, yb) < 3e-6


def test_rms_norm_rope_attention(orc, ref):
    rng = np.random.default_rng(5)
    x = rng.normal(size=(5, 3072)).astype(np.float32)
    w = rng.uniform(0.9, 1.1, size=3072).astype(np.float32)
    a, b = np.empty_like(x), np.empty_like(x)
    orc.orc_rms_norm(P(a), P(x), P(w), 5, 3072, C.c_float(1e-5)); ref.L.vox_rms_norm(P(b), P(x), P(w), 5, 3072, 1e-5)
    assert rel(a, b) < 2e-6
    pos = np.array([0, 1, 37, 900, 8191], dtype=np.int32)
    fa, fb = np.empty((5, 128), np.float32), np.empty((5, 128), np.float32)
    orc.orc_rope_freqs(P(fa), pos.ctypes.data_as(ip), 5, 128, C.c_float(1e6))
    ref.L.vox_compute_rope_freqs(P(fb), pos.ctypes.data_as(ip), 5, 128, 1e6)
    assert np.abs(fa - fb).max() < 4 * 8192 * 2.0 ** -23 + 2e-6          # a few ulps of the f32 ANGLE (fast-math libmvec in the ref)
    q = rng.normal(size=(5, 8 * 128)).astype(np.float32)
    a, b = q.copy(), q.copy()
    orc.orc_apply_rope(P(a), P(fb), 5, 8, 128); ref.L.vox_apply_rope(P(b), P(fb), 5, 8, 128)
    assert rel(a, b) < 1e-6
    for (sq, sk, H, Hkv, hd, win, off) in [(3, 40, 8, 2, 32, 16, 37), (6, 6, 4, 4, 64, 750, 0), (1, 30, 32, 8, 128, 8192, 29)]:
        Q = rng.normal(size=(sq, H * hd)).astype(np.float32)
        K = rng.normal(size=(sk, Hkv * hd)).astype(np.float32)
        V = rng.normal(size=(sk, Hkv * hd)).astype(np.float32)
        a, b = np.empty_like(Q), np.empty_like(Q)
        sc = 1.0 / np.sqrt(hd)
        orc.orc_causal_attention(P(a), P(Q), P(K), P(V), sq, sk, H, Hkv, hd, C.c_float(sc), win, off)
        ref.L.vox_causal_attention(P(b), P(Q), P(K), P(V), sq, sk, H, Hkv, hd, sc, win, off)
        assert rel(a, b) < 3e-6


@pytest.mark.parametrize("cin,cout,L,stride", [(16, 8, 21, 1), (16, 8, 21, 2), (16, 8, 22, 2)])
def test_causal_conv1d(orc, ref, cin, cout, L, stride):
    rng = np.random.default_rng(7)
    x = rng.normal(size=(cin, L)).astype(np.float32)
    w = (rng.normal(size=(cout, cin, 3)) * 0.1).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    n = orc.orc_causal_conv1d_out_len(L, 3, stride)
    assert n == (L if stride == 1 else (L + 1) // 2)
    ya, yb = np.zeros((cout, n), np.float32), np.zeros((cout, n), np.float32)
    orc.orc_causal_conv1d(P(ya), P(x), P(w), P(b), cin, cout, L, 3, stride)
    ref.L.vox_causal_conv1d(P(yb), P(x), P(w), P(b), cin, cout, L, 3, stride)
    assert rel(ya, yb) < 3e-6


def _ref_stream_mel(ref, pcm, delay_tokens=6):
    ctx = ref.L.vox_mel_ctx_init(32 * 1280)
    ref.L.vox_mel_feed(ctx, P(pcm), pcm.size)
    align = (1280 - pcm.size % 1280) % 1280
    z = np.zeros(align + (delay_tokens + 1 + 10) * 1280, np.float32)
    ref.L.vox_mel_feed(ctx, P(z), z.size)
    ref.L.vox_mel_finish(ctx, 0)
    n = C.c_int()
    p = ref.L.vox_mel_data(ctx, C.byref(n))
    out = np.ctypeslib.as_array(p, shape=(n.value, 128)).copy()
    ref.L.vox_mel_free(ctx)
    return out


def test_stream_mel_matches_reference(orc, ref):
    rng = np.random.default_rng(9)
    for n in (1000, 16000, 20001):
        pcm = (rng.normal(size=n) * 0.1).astype(np.float32)
        want = _ref_stream_mel(ref, pcm)
        frames = orc.orc_stream_mel(None, None, n, 6)
        assert frames == want.shape[0]
        got = np.empty((frames, 128), np.float32)
        orc.orc_stream_mel(P(got), P(pcm), n, 6)
        assert np.abs(got - want).max() < 5e-4


def test_stream_mel_jfk_checksum(orc, ref):
    """the reference's streaming mel of samples/jfk.wav is 1496 frames, sum -18876.35."""
    path = "/root/reference/samples/jfk.wav"
    if not os.path.exists(path):
        pytest.skip("reference samples not present on this machine")
    with wave.open(path) as w:
        pcm = (np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).astype(np.float32) / np.float32(32768.0))
    assert pcm.size == 176000
    frames = orc.orc_stream_mel(None, None, pcm.size, 6)
    assert frames == 1496
    got = np.empty((frames, 128), np.float32)
    orc.orc_stream_mel(P(got), P(np.ascontiguousarray(pcm)), pcm.size, 6)
    assert abs(float(got.sum(dtype=np.float64)) - (-18876.350776)) < 0.5
    assert abs(float(got.min()) - (-0.625)) < 1e-6
    want = _ref_stream_mel(ref, np.ascontiguousarray(pcm))
    # real speech has bins whose power is pure rounding residue (1e-9 of the frame energy); their log moves by a few
    # percent with the summation order (-ffast-math reference), so: tight on average, loose on the worst bin
    d = np.abs(got - want)
    assert d.mean() < 2e-5 and d.max() < 1e-2


def test_counts_and_time_conditioning(orc, ref):
    for n, want in ((176000, (1496, 748, 187, 149)), (480000, (3392, 1696, 424, 386)), (9600000, (60392, 30196, 7549, 7511)),
                    (57600000, (360392, 180196, 45049, 45011)), (32000, (592, 296, 74, 36))):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        orc.orc_stream_counts(n, 6, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        assert (a.value, b.value, c.value, d.value) == want       # SURVEY.md section 8 table + golden synth_s2
    x = np.array([0.5, 3.0, 3.0, -1.0], np.float32)
    assert orc.orc_argmax(P(x), 4) == 1                            # first maximum wins


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference headers (build container only)")
def test_decoder_step_restatement_matches_reference_forward(orc, ref, tmp_path):
    """orc_decoder_layer_step x 26 + final norm + tied-embedding logits + argmax vs the reference's own vox_decoder_forward at
    the real dimensions, three consecutive positions (tests/c/pin_decoder_step.c fills a public vox_ctx_t by hand).  This is
    the restatement tests/test_gpu_kv_ring_wrap.py uses as its checker."""
    exe = str(tmp_path / "pin_dec")
    odir, rdir = os.path.join(ROOT, "oracle", "_build"), os.path.join(ROOT, "oracle", "_ref")
    subprocess.check_call(["gcc", "-O2", "-I/root/reference", "-I" + os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "c", "pin_decoder_step.c"),
                           "-o", exe, "-L" + rdir, "-lvoxref", "-L" + odir, "-loracle", "-Wl,-rpath," + rdir, "-Wl,-rpath," + odir, "-lm"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr[-500:]
    lines = [l.split() for l in r.stdout.splitlines() if l.startswith("step")]
    assert len(lines) == 3
    for l in lines:
        diff, scale, same = float(l[3]), float(l[5]), int(l[7])
        assert same == 1 and diff < 5e-6 * scale * 2          # measured 2.3e-5 on logits of scale 8: -ffast-math summation order


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference headers (build container only)")
def test_encoder_layer_and_adapter_restatements_match_reference(orc, ref, tmp_path):
    """orc_encoder_layer x 32 + final norm vs vox_encoder_forward_incremental over two calls (7 then 5 rows: K/V cache carry,
    logical RoPE positions), and orc_adapter vs vox_adapter_forward, at the real dimensions (tests/c/pin_encoder_adapter.c)."""
    exe = str(tmp_path / "pin_enc")
    odir, rdir = os.path.join(ROOT, "oracle", "_build"), os.path.join(ROOT, "oracle", "_ref")
    subprocess.check_call(["gcc", "-O2", "-I/root/reference", "-I" + os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "c", "pin_encoder_adapter.c"),
                           "-o", exe, "-L" + rdir, "-lvoxref", "-L" + odir, "-loracle", "-Wl,-rpath," + rdir, "-Wl,-rpath," + odir, "-lm"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr[-500:]
    checks = {l.split()[0]: (float(l.split()[2]), float(l.split()[4])) for l in r.stdout.splitlines() if "max_abs_diff" in l}
    assert set(checks) == {"encoder_full", "encoder_call_1", "encoder_call_2", "adapter"}
    for name, (diff, scale) in checks.items():
        assert diff < 2e-5 * max(scale, 1.0), name             # measured 5e-6..6e-6: OpenBLAS / -ffast-math summation order
    assert "enc_cache_len 12" in r.stdout
