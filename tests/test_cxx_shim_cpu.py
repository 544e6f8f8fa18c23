"""The drop-in C++ class headers compile (against our compat headers and, when the reference
tree is present, against the reference's own tensor.h/config.h) and fail like the reference
(verify-style abort) when there is no device."""
import os
import signal
import subprocess

import numpy as np
import pytest

from tests import cxx_build


def test_shim_builds_with_compat_headers(tmp_path):
    exe = cxx_build.build(str(tmp_path / "shim_check"))
    assert os.path.exists(exe)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_shim_compiles_against_reference_headers(tmp_path):
    cxx_build.build(str(tmp_path / "shim_check.o"), with_reference_headers=True)


def test_shim_aborts_without_device(tmp_path):
    import fdm_b200
    if fdm_b200.lib().fdmb_device_count() > 0:
        pytest.skip("a GPU is present")
    exe = cxx_build.build(str(tmp_path / "shim_check"))
    rhs = tmp_path / "rhs.bin"
    np.zeros(15 ** 3).tofile(rhs)
    r = subprocess.run([exe, "cube", "15", str(rhs), str(tmp_path / "ans.bin")], capture_output=True, text=True)
    assert r.returncode == -signal.SIGABRT
    assert "verify(" in r.stderr
