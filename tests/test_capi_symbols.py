"""The C-ABI library loads without a GPU (no link-time dependency on the driver), exports every entry point that
include/imagestitch.h declares, the ctypes table binds all of them, and creating a context without a device fails
loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "imagestitch.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(is_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from imagestitch_b200 import build as B, capi
    B.build()
    lib = capi.load()
    names = _declared()
    assert len(names) > 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in imagestitch.h but not exported: {missing}"
    unbound = [n for n in names if n not in capi.SYMBOLS]
    assert not unbound, f"declared in imagestitch.h but not bound in capi.SYMBOLS: {unbound}"
    extra = [n for n in capi.SYMBOLS if n not in names]
    assert not extra, f"bound in capi.SYMBOLS but not declared in imagestitch.h: {extra}"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from imagestitch_b200 import capi
    lib = capi.load()
    h = C.c_void_p()
    assert lib.is_ctx_create(0, C.byref(h)) < 0 and not h.value
