"""CPU suite: the C-ABI library builds, loads, and exports every symbol include/csb200.h declares (no compute)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "csb200.h")).read()
    return sorted(set(re.findall(r"CSB_API\s+[\w\s\*]+?\b(csb_\w+)\s*\(", txt)))


def test_header_declares_entry_points():
    names = _declared()
    assert "csb_pointcloud_render" in names and "csb_disocclusion_fill" in names and len(names) >= 15


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for n in _declared():
        assert hasattr(lib, n), f"{n} declared in include/csb200.h but not exported"
    assert lib.csb_version() >= 100


def test_invalid_arguments_return_status_not_crash(built_lib):
    lib = ctypes.CDLL(built_lib)
    lib.csb_last_error.restype = ctypes.c_char_p
    st = lib.csb_pointcloud_render(None, None, 1, 10, 4, 8, 8, ctypes.c_double(512.0), ctypes.c_double(40.0), None, None, None, None, None, None, None)
    assert st == 1 and b"null" in lib.csb_last_error()
    assert lib.csb_render_acc_channels(4) == 8 and lib.csb_render_acc_channels(3) == 4 and lib.csb_render_acc_channels(68) == 72


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "cartoonsegmentation_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b|importlib.*oracle|libkb_oracle|oracle/", src, re.M), f"{f} uses the oracle"


def test_ops_fail_loudly_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cartoonsegmentation_b200._lib import CsbError
    from cartoonsegmentation_b200.anime_3dkenburns.common import fill_disocclusion
    with pytest.raises(CsbError):
        fill_disocclusion(torch.zeros(1, 4, 8, 8), torch.zeros(1, 1, 8, 8))
