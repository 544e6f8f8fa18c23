"""The C-ABI library loads and exports every symbol include/fdm_b200.h declares.
No compute calls (there is no GPU in the CPU tier)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fdm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fdmb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_header():
    import __graft_entry__ as ge
    path = os.path.join(ROOT, "fdm_b200", "libfdm_b200.so")
    if not os.path.exists(path):
        ge.build()
    L = C.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_no_gpu_fails_loudly():
    import fdm_b200
    L = fdm_b200.lib()
    if L.fdmb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fdm_b200.FdmB200Error):
        fdm_b200.LaplCube(0.1, 0.1, 0.1, 1.6, 1.6, 1.6, 15, 15, 15)


def test_invalid_size_message():
    # the reference aborts with verify((1<<n) == N) (src/fft.cpp:67); we return FDMB_ERR_INVALID
    import fdm_b200
    with pytest.raises(fdm_b200.FdmB200Error) as e:
        fdm_b200.LaplCube(0.1, 0.1, 0.1, 3.3, 3.3, 3.3, 32, 32, 32)
    assert "powers of two" in str(e.value) or "CUDA" in str(e.value) or "cuda" in str(e.value)


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "fdm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("test oracle", ""), os.path.join(dirpath, f)
