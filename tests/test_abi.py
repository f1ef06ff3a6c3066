"""The C-ABI library loads on a GPU-less host and exports exactly what include/lssvc_b200.h declares."""
import ctypes
import os
import re

from lssvc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lssvc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lssvc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_binding_covers_header():
    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_library_reports_version_and_errors_without_gpu():
    lib = _lib.load()
    assert lib.lssvc_abi_version() == 3
    # argument validation happens before any CUDA call
    rc = lib.lssvc_conv_tc(None, None)
    assert rc == -1 and b"null descriptor" in lib.lssvc_last_error()
    assert lib.lssvc_launch_count() >= 0
