"""The C-ABI library loads and exports every symbol include/vqb200.h declares (no compute
calls: there is no GPU in the authoring container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vqb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vqb_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = _declared_symbols()
    for required in ("vqb_forward", "vqb_backward", "vqb_gather", "vqb_one_hot", "vqb_encode_host",
                     "vqb_workspace_bytes", "vqb_query", "vqb_error_string"):
        assert required in names


def test_library_exports_every_declared_symbol():
    import vqb200
    lib = vqb200._lib.load()
    raw = ctypes.CDLL(vqb200._lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(raw, name), f"{name} declared in include/vqb200.h but not exported"
        assert name in vqb200._lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.vqb_version() == 100


def test_pure_host_entry_points():
    import vqb200
    lib = vqb200._lib.load()
    assert lib.vqb_workspace_bytes(256, 32) > 256 * 4
    assert lib.vqb_workspace_bytes(256, 32) % 256 == 0
    assert lib.vqb_workspace_bytes(0, 32) == 0
    assert "invalid argument" in vqb200._lib.error_string(-1)
    assert "workspace" in vqb200._lib.error_string(-2)
    # argument validation happens before any CUDA call
    assert lib.vqb_forward(0, None, 4, 1, 0, 0, 0, 1, None, 4, 0.25, None, None, None, None, None, None,
                           None, 0, 0, None) == -1
    assert lib.vqb_gather(0, None, -1, None, 4, 4, None, None, None) == -1


def test_missing_library_fails_loudly(monkeypatch):
    import vqb200
    monkeypatch.setattr(vqb200._lib, "_lib", None)
    monkeypatch.setattr(vqb200._lib, "LIB_PATH", "/nonexistent/libvqb200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vqb200._lib.load()
