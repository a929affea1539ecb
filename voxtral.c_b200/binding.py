"""ctypes binding of libvoxtral_b200.so (C ABI: include/voxtral_b200.h).

Mirrors the reference's own API names (vox_load, vox_stream_*, vox_linear_bf16, ...), so the
parity tests read like calls into /root/reference.  Nothing here computes: every function is a
thin call into the shared library, and the library itself refuses to run without a CUDA device.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libvoxtral_b200.so")
HEADER = os.path.join(REPO_ROOT, "include", "voxtral_b200.h")

_lib = None
c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


def build(verbose=False):
    """Compile the CUDA/C sources for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", PKG_DIR, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libvoxtral_b200.so failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout[-2000:])
    return LIB_PATH


class CudaInfo(C.Structure):
    _fields_ = [("device", C.c_int), ("sm_count", C.c_int), ("cc_major", C.c_int), ("cc_minor", C.c_int),
                ("weight_bytes_hbm", C.c_size_t), ("kv_bytes_hbm", C.c_size_t),
                ("kernel_launches", C.c_ulonglong), ("last_decode_kernel_ms", C.c_double),
                ("last_decode_steps", C.c_int), ("last_encoder_kernel_ms", C.c_double),
                ("last_encoder_positions", C.c_int), ("last_mel_kernel_ms", C.c_double),
                ("total_decode_kernel_ms", C.c_double), ("total_decode_steps", C.c_longlong),
                ("total_encoder_ms", C.c_double), ("total_encoder_positions", C.c_longlong), ("load_ms", C.c_double),
                ("verify_passes", C.c_longlong), ("verify_tokens", C.c_longlong)]


def lib():
    """Load the shared library (fails loudly if it was not built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no fallback implementation)")
    L = C.CDLL(LIB_PATH)   # RTLD_LOCAL: the reference oracle exports the same names
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    u16p = C.POINTER(C.c_uint16)
    sig = {
        "vox_load": (vp, [C.c_char_p]), "vox_free": (None, [vp]), "vox_set_delay": (None, [vp, i]),
        "vox_stream_init": (vp, [vp]), "vox_stream_feed": (i, [vp, c_float_p, i]),
        "vox_stream_finish": (i, [vp]), "vox_stream_flush": (i, [vp]),
        "vox_stream_get": (i, [vp, C.POINTER(C.c_char_p), i]),
        "vox_stream_get_alt": (i, [vp, C.POINTER(C.c_char_p), i, i]),
        "vox_transcribe": (vp, [vp, C.c_char_p]), "vox_transcribe_stdin": (vp, [vp]),
        "vox_decoder_kv_cache_preallocate": (i, [vp, i]), "vox_encoder_kv_cache_preallocate": (i, [vp, i]),
        "vox_stream_set_alt": (None, [vp, i, f]), "vox_set_processing_interval": (None, [vp, f]),
        "vox_stream_set_continuous": (None, [vp, i]), "vox_stream_free": (None, [vp]),
        "vox_transcribe_audio": (vp, [vp, c_float_p, i]),
        "vox_encoder_forward": (c_float_p, [vp, c_float_p, i, c_int_p]),
        "vox_encoder_forward_incremental": (c_float_p, [vp, c_float_p, i, c_int_p]),
        "vox_adapter_forward": (c_float_p, [vp, c_float_p, i, c_int_p]),
        "vox_decoder_forward": (i, [vp, c_float_p, c_float_p]),
        "vox_decoder_prefill": (None, [vp, c_float_p, i]),
        "vox_add_inplace": (None, [c_float_p, c_float_p, i]), "vox_mul_inplace": (None, [c_float_p, c_float_p, i]),
        "vox_axpy": (None, [c_float_p, f, c_float_p, i]), "vox_scale": (None, [c_float_p, f, i]),
        "vox_copy": (None, [c_float_p, c_float_p, i]),
        "vox_matmul": (None, [c_float_p, c_float_p, c_float_p, i, i, i]),
        "vox_matmul_t": (None, [c_float_p, c_float_p, c_float_p, i, i, i]),
        "vox_linear": (None, [c_float_p, c_float_p, c_float_p, c_float_p, i, i, i]),
        "vox_linear_nobias": (None, [c_float_p, c_float_p, c_float_p, i, i, i]),
        "vox_linear_nobias_bf16": (None, [c_float_p, c_float_p, u16p, i, i, i]),
        "vox_linear_bf16": (None, [c_float_p, c_float_p, u16p, c_float_p, i, i, i]),
        "vox_matmul_t_bf16": (None, [c_float_p, c_float_p, u16p, i, i, i]),
        "vox_conv1d": (None, [c_float_p, c_float_p, c_float_p, c_float_p, i, i, i, i, i, i]),
        "vox_causal_conv1d": (None, [c_float_p, c_float_p, c_float_p, c_float_p, i, i, i, i, i]),
        "vox_rms_norm": (None, [c_float_p, c_float_p, c_float_p, i, i, f]),
        "vox_silu": (None, [c_float_p, i]), "vox_gelu": (None, [c_float_p, i]),
        "vox_softmax": (None, [c_float_p, i, i]),
        "vox_causal_attention": (None, [c_float_p, c_float_p, c_float_p, c_float_p, i, i, i, i, i, f, i, i]),
        "vox_compute_rope_freqs": (None, [c_float_p, c_int_p, i, i, f]),
        "vox_apply_rope": (None, [c_float_p, c_float_p, i, i, i]),
        "vox_load_wav": (c_float_p, [C.c_char_p, c_int_p]),
        "vox_mel_spectrogram": (c_float_p, [c_float_p, i, c_int_p]),
        "vox_mel_ctx_init": (vp, [i]), "vox_mel_feed": (i, [vp, c_float_p, i]), "vox_mel_finish": (i, [vp, i]),
        "vox_mel_data": (c_float_p, [vp, c_int_p]), "vox_mel_frame_offset": (i, [vp]),
        "vox_mel_discard_before": (None, [vp, i]), "vox_mel_free": (None, [vp]),
        "vox_tokenizer_load": (vp, [C.c_char_p]), "vox_tokenizer_free": (None, [vp]),
        "vox_tokenizer_decode": (C.c_char_p, [vp, i]),
        "vox_cuda_get_info": (i, [vp, C.POINTER(CudaInfo)]), "vox_cuda_version": (C.c_char_p, []),
        "vox_cuda_reset_caches": (None, [vp]), "vox_cuda_set_decode_mode": (None, [vp, i]),
        "vox_cuda_mel_conv_stem": (i, [vp, vp, i, vp]),
        "vox_cuda_encoder_layer_qkv": (i, [vp, i, vp, i, i, vp, vp, i]),
        "vox_cuda_encoder_layer_rest": (i, [vp, i, vp, i, vp, vp, i]),
        "vox_cuda_encoder_final_norm": (i, [vp, vp, i]), "vox_cuda_adapter": (i, [vp, vp, i, vp]),
        "vox_cuda_sync": (None, [vp]), "vox_cuda_debug_copy_kv": (i, [vp, i, c_float_p, c_float_p]),
        "vox_cuda_debug_copy_logits": (i, [vp, c_float_p]), "vox_cuda_mel_device_frames": (vp, [vp, c_int_p]),
        "vox_cuda_mel_feed_zeros": (i, [vp, i]), "vox_cuda_build_prompt": (i, [vp, vp, vp, i]),
        "vox_cuda_decoder_prefill": (i, [vp, vp, i]), "vox_cuda_decoder_steps": (i, [vp, vp, i, i, i, c_int_p]),
        "vox_cuda_encoder_step": (i, [vp, vp, i]),
        "vox_cuda_stream_feed_device": (i, [vp, vp, i]),
        "vox_cuda_malloc": (vp, [vp, C.c_size_t]), "vox_cuda_free": (None, [vp, vp]),
        "vox_cuda_memcpy_h2d": (i, [vp, vp, vp, C.c_size_t]), "vox_cuda_memcpy_d2h": (i, [vp, vp, vp, C.c_size_t]),
        "vox_cuda_timer_start": (None, [vp]), "vox_cuda_timer_stop_ms": (C.c_double, [vp]),
        "vox_cuda_stream_token_ids": (i, [vp, c_int_p, i]),
        "vox_cuda_stream_counts": (i, [vp, c_int_p, c_int_p, c_int_p]),
        "vox_cuda_debug_fail_alloc_after": (None, [C.c_longlong]), "vox_cuda_set_verify_depth": (None, [vp, i]),
        "vox_cuda_ctx_fork": (vp, [vp]), "vox_cuda_stream_set_deferred": (None, [vp, i]),
        "vox_cuda_streams_decode": (i, [C.POINTER(vp), i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    L.free_ = C.CDLL(None).free
    L.free_.argtypes = [C.c_void_p]
    _lib = L
    return L


def declared_symbols():
    """Every function/variable name include/voxtral_b200.h declares (for the export test)."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b((?:vox|safetensors?)_[a-z0-9_]+)\s*\(", src))
    names |= set(re.findall(r"extern\s+int\s+(vox_[a-z_]+)\s*;", src))
    return sorted(n for n in names if not n.endswith("_t"))


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def fptr(a):
    return a.ctypes.data_as(c_float_p)


def take(ptr, shape):
    """Copy a malloc'd float buffer returned by the C API into numpy and free it."""
    n = int(np.prod(shape))
    out = np.ctypeslib.as_array(ptr, shape=(n,)).copy().reshape(shape)
    lib().free_(C.cast(ptr, C.c_void_p))
    return out


class Engine:
    """vox_ctx_t wrapper."""

    def __init__(self, model_dir, verbose=0, _ctx=None):
        L = lib()
        C.c_int.in_dll(L, "vox_verbose").value = verbose
        self.ctx = _ctx if _ctx is not None else L.vox_load(model_dir.encode())
        if not self.ctx:
            raise RuntimeError(f"vox_load({model_dir}) failed (no GPU, or bad checkpoint)")
        self.model_dir = model_dir

    def fork(self):
        """A second context on the same weights (vox_cuda_ctx_fork); close it before the parent."""
        c = lib().vox_cuda_ctx_fork(self.ctx)
        if not c:
            raise RuntimeError("vox_cuda_ctx_fork failed")
        return Engine(self.model_dir, _ctx=c)

    def close(self):
        if self.ctx:
            lib().vox_free(self.ctx)
            self.ctx = None

    def info(self):
        ci = CudaInfo()
        lib().vox_cuda_get_info(self.ctx, C.byref(ci))
        return {k: getattr(ci, k) for k, _ in CudaInfo._fields_}

    def stream(self):
        return Stream(self)

    def to_device(self, arr):
        a = np.ascontiguousarray(arr)
        d = lib().vox_cuda_malloc(self.ctx, a.nbytes)
        lib().vox_cuda_memcpy_h2d(self.ctx, d, a.ctypes.data_as(C.c_void_p), a.nbytes)
        return d

    def dev_free(self, d):
        lib().vox_cuda_free(self.ctx, d)

    def timer_start(self):
        lib().vox_cuda_timer_start(self.ctx)

    def timer_stop_ms(self):
        return lib().vox_cuda_timer_stop_ms(self.ctx)

    def set_delay(self, delay_ms):
        lib().vox_set_delay(self.ctx, int(delay_ms))

    def set_decode_mode(self, mode):
        lib().vox_cuda_set_decode_mode(self.ctx, {"auto": 0, "graph": 1, "persist": 3, "v2": 5}[mode])

    def set_verify_depth(self, depth):
        lib().vox_cuda_set_verify_depth(self.ctx, int(depth))

    def reset_caches(self):
        lib().vox_cuda_reset_caches(self.ctx)

    def decoder_prefill(self, embeds):
        e = np.ascontiguousarray(embeds, dtype=np.float32)
        lib().vox_decoder_prefill(self.ctx, fptr(e), e.shape[0])

    def decoder_forward(self, embed):
        e = np.ascontiguousarray(embed, dtype=np.float32)
        logits = np.empty(131072, dtype=np.float32)
        tok = lib().vox_decoder_forward(self.ctx, fptr(e), fptr(logits))
        return tok, logits

    def encoder_forward_incremental(self, x_new):
        x = np.ascontiguousarray(x_new, dtype=np.float32)
        n = C.c_int()
        p = lib().vox_encoder_forward_incremental(self.ctx, fptr(x), x.shape[0], C.byref(n))
        return take(p, (n.value, 1280))

    def adapter_forward(self, enc_out):
        x = np.ascontiguousarray(enc_out, dtype=np.float32)
        n = C.c_int()
        p = lib().vox_adapter_forward(self.ctx, fptr(x), x.shape[0], C.byref(n))
        return take(p, (n.value, 3072))


def streams_decode(streams):
    """vox_cuda_streams_decode over Stream objects put in deferred mode: one weight pass for all of them."""
    arr = (C.c_void_p * len(streams))(*[s.s for s in streams])
    return lib().vox_cuda_streams_decode(arr, len(streams))


class Stream:
    """vox_stream_t wrapper."""

    def __init__(self, eng):
        self.eng = eng
        self.s = lib().vox_stream_init(eng.ctx)
        if not self.s:
            raise RuntimeError("vox_stream_init failed (tekken.json missing?)")

    def feed(self, pcm):
        a = np.ascontiguousarray(pcm, dtype=np.float32)
        return lib().vox_stream_feed(self.s, fptr(a), a.size)

    def feed_device(self, d_ptr, n):
        return lib().vox_cuda_stream_feed_device(self.s, d_ptr, n)

    def finish(self):
        return lib().vox_stream_finish(self.s)

    def flush(self):
        return lib().vox_stream_flush(self.s)

    def set_interval(self, seconds):
        lib().vox_set_processing_interval(self.s, seconds)

    def set_deferred(self, on):
        lib().vox_cuda_stream_set_deferred(self.s, int(on))

    def set_continuous(self, on):
        lib().vox_stream_set_continuous(self.s, int(on))

    def set_alt(self, n_alt, cutoff):
        lib().vox_stream_set_alt(self.s, n_alt, cutoff)

    def get_alt(self, n_alt):
        """[[best, alt1, ...], ...] per token position; missing alternatives are None (voxtral.h:264-269)."""
        out = []
        buf = (C.c_char_p * (64 * n_alt))()
        while True:
            n = lib().vox_stream_get_alt(self.s, buf, 64, n_alt)
            if n <= 0:
                break
            out.extend([buf[i * n_alt + k] for k in range(n_alt)] for i in range(n))
        return out

    def get(self):
        out = []
        buf = (C.c_char_p * 64)()
        while True:
            n = lib().vox_stream_get(self.s, buf, 64)
            if n <= 0:
                break
            out.extend(buf[i] for i in range(n))
        return out

    def token_ids(self):
        n = lib().vox_cuda_stream_token_ids(self.s, None, 0)
        ids = np.zeros(max(n, 1), dtype=np.int32)
        lib().vox_cuda_stream_token_ids(self.s, ids.ctypes.data_as(c_int_p), n)
        return ids[:n]

    def counts(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib().vox_cuda_stream_counts(self.s, C.byref(a), C.byref(b), C.byref(c))
        return {"mel_frames": a.value, "adapter_tokens": b.value, "decoder_steps": c.value}

    def close(self):
        if self.s:
            lib().vox_stream_free(self.s)
            self.s = None
