"""GPU: the drop-in C++ classes (fdm_b200/cxx) called like the reference's callers, checked
against the oracle.  Bar 1e-12 rel-L2 (fp64); 1e-5 for the float instantiation."""
import math
import subprocess

import numpy as np
import pytest

from oracle import fdm_oracle as O
from tests import cxx_build

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    return cxx_build.build(str(tmp_path_factory.mktemp("cxx") / "shim_check"))


def run_cube(exe, tmp_path, mode, n, rhs):
    rhs.tofile(tmp_path / "rhs.bin")
    subprocess.run([exe, mode, str(n), str(tmp_path / "rhs.bin"), str(tmp_path / "ans.bin")], check=True)
    return np.fromfile(tmp_path / "ans.bin").reshape(n, n, n)


def test_cxx_lapl_cube_dirichlet(exe, tmp_path):
    n = 31; dx = 1.0 / n; l = 1 + dx
    rhs = O.synthetic_rhs((n, n, n), seed=5)
    a = run_cube(exe, tmp_path, "cube", n, rhs)
    assert O.rel_l2(a, O.LaplCube(dx, dx, dx, l, l, l, n, n, n).solve(rhs)) < 1e-12


def test_cxx_lapl_cube_periodic(exe, tmp_path):
    n = 32; dx = 2 * math.pi / n; l = 2 * math.pi
    rhs = O.synthetic_rhs((n, n, n), seed=6); rhs -= rhs.mean()
    a = run_cube(exe, tmp_path, "cubep", n, rhs)
    assert O.rel_l2(a, O.LaplCube(dx, dx, dx, l, l, l, n, n, n, True).solve(rhs)) < 1e-12


def test_cxx_lapl_cube_float(exe, tmp_path):
    n = 15; dx = 1.0 / n; l = 1 + dx
    rhs = O.synthetic_rhs((n, n, n), seed=7).astype(np.float32).astype(np.float64)
    a = run_cube(exe, tmp_path, "cubef", n, rhs)
    assert O.rel_l2(a, O.LaplCube(dx, dx, dx, l, l, l, n, n, n).solve(rhs)) < 1e-5


def test_cxx_ns_cube(exe, tmp_path):
    n, steps = 15, 8
    r = subprocess.run([exe, "ns", str(n), str(steps), str(tmp_path / "ns"), "--ns:Re=250", "--ns:dt=0.01"],
                       check=True, capture_output=True, text=True)
    assert f"time_index {steps}" in r.stdout
    po = O.NSCube(nx=n, nz=n, Re=250.0, dt=0.01)
    for _ in range(steps):
        po.step()
    got = np.concatenate([np.fromfile(tmp_path / f"ns_{f}.bin") for f in "uvwp"])
    want = np.concatenate([po.fields()[f].ravel() for f in "uvwp"])
    assert O.rel_l2(got, want) < 1e-12
