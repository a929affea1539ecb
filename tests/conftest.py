import ctypes as C
import os
import subprocess
import sys
import wave

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbload  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SYNTH_DIR = os.environ.get("VOX_SYNTH_DIR", "/dev/shm/voxsynth_b200")
MODEL_MD5_SEED = "b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _tool(name, src, extra=()):
    exe = os.path.join(ROOT, "tools", "_build", name)
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(os.path.join(ROOT, "tools", src)):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", *extra, "-o", exe, os.path.join(ROOT, "tools", src), "-lm"])
    return exe


def ensure_synth_model(path=SYNTH_DIR):
    """Seeded synthetic checkpoint + tokenizer (bit-identical wherever it is generated)."""
    st = os.path.join(path, "consolidated.safetensors")
    if not (os.path.exists(st) and os.path.getsize(st) > 8_000_000_000):
        os.makedirs(path, exist_ok=True)
        subprocess.check_call([_tool("make_synth_model", "make_synth_model.c", ["-fopenmp"]), path])
    if not os.path.exists(os.path.join(path, "tekken.json")):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_synth_tekken.py"), path])
    return path


def synth_wav(seconds, path=SYNTH_DIR):
    os.makedirs(path, exist_ok=True)
    fn = os.path.join(path, f"synth_{seconds:g}s.wav")
    if not os.path.exists(fn):
        subprocess.check_call([_tool("make_synth_wav", "make_synth_wav.c"), fn, f"{seconds:g}"])
    return fn


def read_wav_f32(fn):
    with wave.open(fn) as w:
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
    return (pcm.astype(np.float32) / np.float32(32768.0)).astype(np.float32)


@pytest.fixture(scope="session")
def vb():
    m = vbload.load()
    if not os.path.exists(m.LIB_PATH):
        m.build()
    return m


class RefLib:
    """The UNMODIFIED reference, compiled by oracle/Makefile into oracle/_ref/libvoxref.so."""

    def __init__(self, path):
        self.L = C.CDLL(path)          # RTLD_LOCAL: must not interpose the engine's identical symbol names
        fp, i, f, vp = C.POINTER(C.c_float), C.c_int, C.c_float, C.c_void_p
        u16p = C.POINTER(C.c_uint16)
        ip = C.POINTER(C.c_int)
        sig = {
            "vox_load": (vp, [C.c_char_p]), "vox_free": (None, [vp]),
            "vox_rms_norm": (None, [fp, fp, fp, i, i, f]),
            "vox_linear_bf16": (None, [fp, fp, u16p, fp, i, i, i]),
            "vox_linear_nobias_bf16": (None, [fp, fp, u16p, i, i, i]),
            "vox_linear": (None, [fp, fp, fp, fp, i, i, i]),
            "vox_matmul": (None, [fp, fp, fp, i, i, i]),
            "vox_silu": (None, [fp, i]), "vox_gelu": (None, [fp, i]), "vox_softmax": (None, [fp, i, i]),
            "vox_add_inplace": (None, [fp, fp, i]), "vox_mul_inplace": (None, [fp, fp, i]),
            "vox_axpy": (None, [fp, f, fp, i]), "vox_scale": (None, [fp, f, i]),
            "vox_causal_attention": (None, [fp, fp, fp, fp, i, i, i, i, i, f, i, i]),
            "vox_compute_rope_freqs": (None, [fp, ip, i, i, f]),
            "vox_apply_rope": (None, [fp, fp, i, i, i]),
            "vox_conv1d": (None, [fp, fp, fp, fp, i, i, i, i, i, i]),
            "vox_causal_conv1d": (None, [fp, fp, fp, fp, i, i, i, i, i]),
            "vox_mel_spectrogram": (fp, [fp, i, ip]),
            "vox_mel_ctx_init": (vp, [i]), "vox_mel_feed": (i, [vp, fp, i]), "vox_mel_finish": (i, [vp, i]),
            "vox_mel_data": (fp, [vp, ip]), "vox_mel_free": (None, [vp]),
            "vox_mel_frame_offset": (i, [vp]), "vox_mel_discard_before": (None, [vp, i]),
            "vox_load_wav": (fp, [C.c_char_p, ip]),
            "vox_tokenizer_load": (vp, [C.c_char_p]), "vox_tokenizer_decode": (C.c_char_p, [vp, i]),
            "vox_tokenizer_free": (None, [vp]),
            "vox_decoder_forward": (i, [vp, fp, fp]), "vox_decoder_prefill": (None, [vp, fp, i]),
            "vox_encoder_forward_incremental": (fp, [vp, fp, i, ip]),
            "vox_encoder_forward": (fp, [vp, fp, i, ip]),
            "vox_copy": (None, [fp, fp, i]), "vox_matmul_t": (None, [fp, fp, fp, i, i, i]),
            "vox_linear_nobias": (None, [fp, fp, fp, i, i, i]), "vox_matmul_t_bf16": (None, [fp, fp, u16p, i, i, i]),
            "vox_adapter_forward": (fp, [vp, fp, i, ip]),
        }
        for n, (r, a) in sig.items():
            fn = getattr(self.L, n)
            fn.restype, fn.argtypes = r, a
        self.free = C.CDLL(None).free
        self.free.argtypes = [C.c_void_p]


@pytest.fixture(scope="session")
def ref():
    p = os.path.join(ROOT, "oracle", "_ref", "libvoxref.so")
    if not os.path.exists(p):
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
        else:
            pytest.skip("oracle/_ref/libvoxref.so not built and /root/reference absent")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", str(os.cpu_count() or 8))
    return RefLib(p)


@pytest.fixture(scope="session")
def model_dir():
    return ensure_synth_model()


@pytest.fixture(scope="session")
def engine(vb, model_dir):
    e = vb.Engine(model_dir)
    yield e
    e.close()


def golden(name):
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"fixture {name}.npz not generated (tools/make_goldens.py)")
    return np.load(path)
