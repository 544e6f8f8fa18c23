"""the stand-in tensor of fdm_b200/cxx/fdm_compat_tensor.h (used by the drop-in headers and example
drivers outside the reference tree) behaves like the reference's src/tensor.h on the constructions of
ut/ut_tensor.cpp:46-118 -- offsets, range-intersection assignment, periodic wrap -- plus use(), index(), maxabs(),
norm2(): the same program is built against both headers and the outputs compared."""
import os
import signal
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cxx", "tensor_check.cpp")


def build_ours(out, extra=()):
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "fdm_b200", "cxx"), *extra,
                    SRC, "-o", out], check=True, capture_output=True, text=True)
    return out


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_compat_tensor_matches_the_reference(tmp_path):
    ref_exe = str(tmp_path / "t_ref")
    subprocess.run(["/usr/bin/g++", "-std=c++20", "-O1", "-DUSE_REFERENCE_TENSOR", "-I/root/reference/src", SRC, "-o",
                    ref_exe], check=True, capture_output=True, text=True)
    a = subprocess.run([ref_exe], capture_output=True, text=True, check=True).stdout
    b = subprocess.run([build_ours(str(tmp_path / "t_ours"))], capture_output=True, text=True, check=True).stdout
    assert a == b and "p=x size 125" in a


def test_compat_tensor_bounds_check_aborts(tmp_path):
    # check = true: an out-of-range index is the reference's verify() -> abort (src/tensor.h:124-126, src/verify.h:10-18)
    src = tmp_path / "oob.cpp"
    src.write_text('#include "fdm_compat_tensor.h"\nint main() { fdm::tensor<double, 2, true> t({0, 1, 0, 1}); return (int)t[2][0]; }\n')
    exe = str(tmp_path / "oob")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "fdm_b200", "cxx"), str(src), "-o", exe],
                   check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == -signal.SIGABRT and "verify(" in r.stderr
