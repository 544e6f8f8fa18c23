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


def test_argument_validation_precedes_cuda():
    """Geometry errors are reported as such (FDMB_ERR_INVALID with the reference's rule in the message), before any
    CUDA call -- so they read the same on a box without a device."""
    import fdm_b200
    cases = [
        (lambda: fdm_b200.VelocityPlotter(0.1, 0.1, 0.1, 31, 32, 31, 0, 3.1, 0, 3.2, 0, 3.1, yperiodic=True), "periodic y needs periodic z"),
        (lambda: fdm_b200.VelocityPlotter(0.1, 0.1, 0.1, 1, 31, 31, 0, 3.1, 0, 3.1, 0, 3.1), "must be >= 2"),
        (lambda: fdm_b200.VelocityPlotter(0.1, 0.1, 0.1, 31, 30, 31, 0, 3.1, 0, 3.0, 0, 3.1), "powers of two"),
        (lambda: fdm_b200.NBodyPM(n=2), "n >= 4"),
        (lambda: fdm_b200.LaplRect(0.1, 0.1, 1.0, 1.0, 16, 16), "powers of two"),
        (lambda: fdm_b200.LaplRect(0.01, 0.1, 40.0, 0.8, 4000, 7), "shared-memory tile"),
    ]
    for make, needle in cases:
        with pytest.raises(fdm_b200.FdmB200Error) as e:
            make()
        assert "code -1" in str(e.value) and needle in str(e.value), str(e.value)
