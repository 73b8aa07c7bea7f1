"""The C-ABI library loads and exports every symbol include/saena_b200.h declares (no compute)."""
import ctypes
import os
import re

from saena_b200 import build, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "saena_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(saena_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build_library()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/saena_b200.h but not exported"
    assert sorted(native.EXPORTED_SYMBOLS) == names


def test_sass_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", build.build_library()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_init_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    try:
        native.Context()
    except native.NativeError as e:
        assert "CUDA" in str(e) or "cuda" in str(e)
    else:
        raise AssertionError("Context() must fail without a CUDA device: there is no CPU path")
