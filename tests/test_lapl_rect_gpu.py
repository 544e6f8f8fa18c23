"""GPU parity tests for LaplRect / LaplRectFFT2 through the C ABI (bar: fp64 rel-L2 <= 1e-12)."""
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


def geom(nx, ny, yp=False, xp=False, dx=0.1, dy=0.05):
    return (dx, dy, dx * (nx if xp else nx + 1), dy * (ny if yp else ny + 1), nx, ny)


def test_rect_golden(fb, golden):
    # tests/golden/make_golden.py: 31 x 15 Dirichlet, outputs of the compiled reference
    g = geom(31, 15)
    assert O.rel_l2(fb.LaplRect(*g).solve(golden["rect_rhs"]), golden["rect_ans"]) < TOL
    assert O.rel_l2(fb.LaplRectFFT2(*g).solve(golden["rect_rhs"]), golden["rectfft2_ans"]) < TOL


@pytest.mark.parametrize("nx,ny,yp", [(31, 15, False), (40, 31, False), (255, 127, False), (17, 64, True),
                                      (400, 255, False), (2, 3, False)])
def test_rect_vs_oracle(fb, nx, ny, yp):
    g = geom(nx, ny, yp)
    rhs = O.synthetic_rhs((ny, nx), seed=nx + ny)
    a = fb.LaplRect(*g, yperiodic=yp).solve(rhs)
    assert O.rel_l2(a, O.LaplRect(*g, yperiodic=yp).solve(rhs)) < TOL


@pytest.mark.parametrize("nx,ny,yp", [(511, 31, False), (1023, 16, True), (860, 7, False), (2047, 7, False)])
def test_rect_long_tridiagonals(fb, nx, ny, yp):
    """Systems too long for the full 32-row shared-memory tile run with a lower tile (16, 8, 4 rows): the
    reference's own 511 x 511 case (ut/ut_lapl_rect.cpp:384-455) and the plotter's LaplRect at 511^3 / 1023^3."""
    g = geom(nx, ny, yp, dx=0.01)
    rhs = O.synthetic_rhs((ny, nx), seed=nx)
    assert O.rel_l2(fb.LaplRect(*g, yperiodic=yp).solve(rhs), O.LaplRect(*g, yperiodic=yp).solve(rhs)) < TOL


def test_rect_tridiagonal_length_limit(fb):
    with pytest.raises(fb.FdmB200Error):
        fb.LaplRect(*geom(4000, 7))          # not even four systems fit in shared memory


@pytest.mark.parametrize("nx,ny,yp,xp", [(31, 15, False, False), (127, 255, False, False), (63, 64, True, False),
                                         (64, 32, True, True), (1023, 7, False, False), (3, 3, False, False)])
def test_rectfft2_vs_oracle(fb, nx, ny, yp, xp):
    g = geom(nx, ny, yp, xp)
    rhs = O.synthetic_rhs((ny, nx), seed=nx + 3 * ny)
    a = fb.LaplRectFFT2(*g, yperiodic=yp, xperiodic=xp).solve(rhs)
    assert O.rel_l2(a, O.LaplRectFFT2(*g, yperiodic=yp, xperiodic=xp).solve(rhs)) < TOL


def test_rect_aliased_eigenvalues(fb):
    # nx == ny with dx != dy: LaplRectFFT2 aliases lm_x to lm_y (lapl_rect.cpp:36-40) -- reproduce, do not fix
    g = geom(31, 31, dx=0.1, dy=0.03)
    rhs = O.synthetic_rhs((31, 31), seed=5)
    a = fb.LaplRectFFT2(*g).solve(rhs)
    assert O.rel_l2(a, O.LaplRectFFT2(*g).solve(rhs)) < TOL


def cyl_scales(nx, x1, dx):
    # src/velocity_plot.h:113-127
    j = np.arange(nx + 1, dtype=np.float64)
    r = x1 + j * dx - dx / 2
    r[0] = 1.0
    lm = 1.0 / r / r; U = (r + dx / 2) / r; L = (r - dx / 2) / r
    lm[0] = U[0] = L[0] = 1.0
    return lm, L, U


@pytest.mark.parametrize("kind", ["rect", "fft2"])
def test_rect_cyl_scales_vs_compiled_reference(fb, ref, kind):
    nx, ny = 63, 31
    dx = (math.pi / 2) / nx; dy = 10.0 / ny
    g = (dx, dy, math.pi / 2 + dx, 10.0 + dy, nx, ny)
    lm, L, U = cyl_scales(nx, math.pi / 2, dx)
    rhs = O.synthetic_rhs((ny, nx), seed=77)
    R = ref.LaplRect(kind, *g, 0)
    R.set_scales(lm, L, U)
    S = (fb.LaplRect if kind == "rect" else fb.LaplRectFFT2)(*g)
    S.set_scales(lm, L, U)
    assert O.rel_l2(S.solve(rhs), R.solve(rhs)) < TOL


def test_rect_equals_rectfft2(fb):
    # ut/ut_lapl_rect.cpp:384-455: the gtsv path and the two-transform path agree (unit scales)
    g = geom(63, 31)
    rhs = O.synthetic_rhs((31, 63), seed=9)
    assert O.rel_l2(fb.LaplRect(*g).solve(rhs), fb.LaplRectFFT2(*g).solve(rhs)) < 1e-11


def test_rect_invalid_size(fb):
    with pytest.raises(fb.FdmB200Error):
        fb.LaplRect(0.1, 0.1, 1.0, 1.0, 16, 16)     # ny + 1 = 17 is not a power of two
