"""CPU: the C-ABI library loads and exports every symbol include/radnet_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "radnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(radnet_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from rock_art_radnet_b200 import _lib
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"


def test_version_and_sizes_without_gpu():
    from rock_art_radnet_b200 import _lib
    lib = _lib.load()
    assert lib.radnet_version() == _lib.ABI_VERSION
    assert lib.radnet_det_record_bytes(300) == 16 + 300 * 24
    assert lib.radnet_det_record_bytes(7) % 16 == 0
    assert lib.radnet_error_name(0) == b"RADNET_OK" and lib.radnet_error_name(-2) == b"RADNET_E_CUDA"
    ws = lib.radnet_sort_nms_i32_workspace_bytes(64, 12996, 38, 38, 300)
    assert ws >= 64 * 12996 * 16 and ws % 256 == 0
    assert lib.radnet_nms_f64_workspace_bytes(1000, 300) > 1000 * 8 * 3
    assert lib.radnet_rpn_targets_workspace_bytes(64, 20, 38, 38, 9) >= 16
    assert lib.radnet_rpn_subsample_workspace_bytes(64, 38, 38, 9) >= 64 * 12996 * 37


def test_argument_errors_are_reported_not_thrown():
    from rock_art_radnet_b200 import _lib
    lib = _lib.load()
    rc = lib.radnet_roi_pool(None, 1, 38, 38, 1024, None, 0, None, None, 1, 14, None, None)
    assert rc == -1 and b"null" in lib.radnet_last_error_string()
    try:
        _lib.call("radnet_decode_clip_i32", None, None, 1, 1, 1, 1, None, ctypes.c_float(4.0), 1, None, None, None, None)
    except _lib.RadnetError as e:
        assert e.code == -1
    else:
        raise AssertionError("expected RadnetError")


def test_missing_library_fails_loudly(monkeypatch):
    from rock_art_radnet_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libradnet_b200.so")
    try:
        _lib.load()
    except ImportError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("expected ImportError")
