"""CPU tests: the C-ABI library loads and exports every symbol include/jetb200.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

from jet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "jetb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_list_agree():
    assert _declared() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail("jet_b200/lib/libjetb200.so missing: run __graft_entry__.build()")
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), name


def test_version_string_and_error_channel():
    L = _lib.lib()
    assert b"sm_100a" in L.jb_version()
    assert isinstance(L.jb_last_error(), bytes)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libjetb200.so")
    with pytest.raises(_lib.JetB200Error, match="no CPU fallback"):
        _lib.lib()
