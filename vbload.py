"""Import helper: the package directory is literally ``voxtral.c_b200`` (dot in the name)."""
import importlib.util
import os
import sys

_NAME = "voxtral_c_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    root = os.path.dirname(os.path.abspath(__file__))
    pkg = os.path.join(root, "voxtral.c_b200")
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(pkg, "__init__.py"), submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


def load_submodule(name):
    """Import voxtral.c_b200/<name>.py without touching the shared library (pure-python helpers)."""
    full = _NAME + "_" + name
    if full in sys.modules:
        return sys.modules[full]
    root = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location(full, os.path.join(root, "voxtral.c_b200", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod
