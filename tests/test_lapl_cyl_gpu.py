"""GPU parity tests for LaplCyl3FFT2 through the C ABI (bar: fp64 rel-L2 <= 1e-12)."""
import math

import numpy as np
import pytest

from oracle import fdm_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def fb():
    import fdm_b200
    assert fdm_b200.lib().fdmb_device_count() > 0, "GPU tests need a CUDA device"
    return fdm_b200


def geom(nr, nz, nphi, zp):
    # the NSCyl construction of the solver (ns_cyl.h:95-97): staggered in r, h in [0,10], r in [pi/2, pi]
    R0, R1 = math.pi / 2, math.pi
    dr = (R1 - R0) / nr; dz = 10.0 / nz
    return (dr, dz, R0 - dr / 2, R1 - R0 + dr, 10.0 if zp else 10.0 + dz, nr, nz, nphi, zp)


def test_cyl_golden(fb, golden):
    for zp in (False, True):
        args = geom(16, 16 if zp else 15, 16, zp)
        tag = "p" if zp else "d"
        a = fb.LaplCyl3FFT2(*args).solve(golden[f"cyl_{tag}_rhs"])
        assert O.rel_l2(a, golden[f"cyl_{tag}_ans"]) < TOL


@pytest.mark.parametrize("nr,nz,nphi", [(32, 31, 32), (128, 127, 128), (40, 31, 32), (17, 7, 64), (64, 255, 16)])
def test_cyl_dirichlet_vs_oracle(fb, nr, nz, nphi):
    args = geom(nr, nz, nphi, False)
    rhs = O.synthetic_rhs((nphi, nz, nr), seed=nr + nz)
    a = fb.LaplCyl3FFT2(*args).solve(rhs)
    assert O.rel_l2(a, O.LaplCyl3FFT2(*args).solve(rhs)) < TOL


@pytest.mark.parametrize("nr,nz,nphi", [(32, 32, 32), (24, 64, 16)])
def test_cyl_zperiodic_vs_oracle(fb, nr, nz, nphi):
    args = geom(nr, nz, nphi, True)
    rhs = O.synthetic_rhs((nphi, nz, nr), seed=nr + nz + 1)
    a = fb.LaplCyl3FFT2(*args).solve(rhs)
    assert O.rel_l2(a, O.LaplCyl3FFT2(*args).solve(rhs)) < TOL


def test_cyl_vs_compiled_reference(fb, ref):
    # BASELINE configs[3] solver size: nr = 128, nz = 127, nphi = 128
    args = geom(128, 127, 128, False)
    rhs = O.synthetic_rhs((128, 127, 128), seed=99)
    a = fb.LaplCyl3FFT2(*args).solve(rhs)
    assert O.rel_l2(a, ref.LaplCyl3FFT2(*args).solve(rhs)) < TOL


def test_cyl_residual_property(fb):
    """Apply the discrete cylindrical operator (lapl_cyl.cpp:151-159 + periodic phi, Dirichlet z)
    to the answer: must reproduce the rhs (size-independent check, no oracle involved)."""
    nr, nz, nphi = 128, 127, 128
    dr, dz, r0, lr, lz, *_ = geom(nr, nz, nphi, False)
    rhs = O.synthetic_rhs((nphi, nz, nr), seed=5)
    u = fb.LaplCyl3FFT2(dr, dz, r0, lr, lz, nr, nz, nphi).solve(rhs)
    dphi = 2 * math.pi / nphi
    r = r0 + np.arange(1, nr + 1) * dr
    up = np.zeros((nphi, nz + 2, nr + 2)); up[:, 1:-1, 1:-1] = u
    c = up[:, 1:-1, 1:-1]
    lap = ((r - 0.5 * dr) / dr**2 / r) * up[:, 1:-1, :-2] + ((r + 0.5 * dr) / dr**2 / r) * up[:, 1:-1, 2:] - 2 / dr**2 * c
    lap += (up[:, 2:, 1:-1] - 2 * c + up[:, :-2, 1:-1]) / dz**2
    lap += (np.roll(c, -1, 0) - 2 * c + np.roll(c, 1, 0)) / dphi**2 / r**2
    assert O.rel_l2(lap, rhs) < 1e-10


@pytest.mark.parametrize("nr,nz,nphi", [(431, 7, 8), (512, 7, 16), (1000, 3, 8)])
def test_cyl_long_radial_systems(fb, nr, nz, nphi):
    """nr beyond the full shared-memory tile of the r sweep: lower tiles, same arithmetic."""
    args = geom(nr, nz, nphi, False)
    rhs = O.synthetic_rhs((nphi, nz, nr), seed=nr)
    assert O.rel_l2(fb.LaplCyl3FFT2(*args).solve(rhs), O.LaplCyl3FFT2(*args).solve(rhs)) < TOL


def test_cyl_errors(fb):
    with pytest.raises(fb.FdmB200Error):
        fb.LaplCyl3FFT2(0.1, 0.1, 1.0, 3.3, 3.3, 32, 32, 32)         # Dirichlet z needs nz = 2^k - 1
    with pytest.raises(fb.FdmB200Error):
        fb.LaplCyl3FFT2(0.1, 0.1, 1.0, 3.3, 3.1, 32, 31, 30)         # nphi must be 2^k
