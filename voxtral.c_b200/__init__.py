"""voxtral.c_b200 -- B200 (sm_100a) engine behind the antirez/voxtral.c C API.

The product is ``libvoxtral_b200.so`` (hand-written CUDA + host C, see ``csrc/``); this
package is only the ctypes view of its C ABI (``include/voxtral_b200.h``) that tests,
``bench.py`` and ``__graft_entry__.py`` use.  The directory name contains a dot, so load it
with ``vbload.load()`` (repo root) rather than a plain ``import``.
"""
from .binding import (  # noqa: F401
    LIB_PATH, PKG_DIR, REPO_ROOT, Engine, Stream, build, lib, declared_symbols, have_gpu, streams_decode,
)
