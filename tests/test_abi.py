"""The C-ABI library must load (no GPU needed) and export exactly the symbols include/fullbatch_b200.h declares; the
ctypes binding must cover all of them.  No compute call is made here."""
import ctypes
import os
import re

import pytest

from fullbatchtraining_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "fullbatch_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(fb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = header_symbols()
    assert len(names) >= 20
    handle = lib.load()
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    assert sorted(lib.EXPORTS) == names, "ctypes binding and header disagree"
    assert handle.fb_version() >= 100


def test_struct_sizes_match_the_header_layout():
    assert ctypes.sizeof(lib.Tap) == 8
    assert ctypes.sizeof(lib.WgradTap) == 4
    assert ctypes.sizeof(lib.WprepEntry) == 80
    assert ctypes.sizeof(lib.ConvGemmArgs) % 8 == 0


def test_errors_are_reported_not_thrown():
    handle = lib.load()
    rc = handle.fb_flat_scale(None, 0, 1.0, None)  # argument validation happens before any CUDA call
    assert rc == 1001
    assert "fb_flat_scale" in lib.last_error()
    with pytest.raises(RuntimeError):
        lib.check(rc, "fb_flat_scale")
