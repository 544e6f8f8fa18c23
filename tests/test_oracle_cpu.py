"""The oracle restatement (oracle/fdm_oracle.py) against the reference's golden vectors,
the live compiled reference (oracle/_ref, when present) and the reference's own
known-answer constructions.  CPU only."""
import math

import numpy as np
import pytest

from oracle import fdm_oracle as O

TOL = 1e-13


def test_fft_golden(golden):
    for N in (32, 128):
        s = golden[f"sFFT_{N}_in"]
        assert O.rel_l2(O.sFFT(s[1:N], 0.37), golden[f"sFFT_{N}_out"][1:N]) < TOL
        s = golden[f"cFFT_{N}_in"]
        assert O.rel_l2(O.cFFT(s, 0.37), golden[f"cFFT_{N}_out"]) < TOL
        s = golden[f"pFFT_1_{N}_in"]
        assert O.rel_l2(O.pFFT_1(s[:N], 0.37), golden[f"pFFT_1_{N}_out"][:N]) < TOL
        assert O.rel_l2(O.pFFT(s[:N], 0.37), golden[f"pFFT_{N}_out"][:N]) < TOL


def test_fft_known_answers():
    # ut/ut_fft.cpp:191-228: sFFT o sFFT * (2/N) = id, and vs the O(N^2) definition (asp_fft.cpp:308-319)
    N = 64
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, N - 1)
    assert O.rel_l2(O.sFFT(O.sFFT(x, 1.0), 2.0 / N), x) < 1e-14
    j = np.arange(1, N)
    direct = np.array([np.sum(x * np.sin(math.pi * k * j / N)) for k in range(1, N)])
    assert O.rel_l2(O.sFFT(x, 1.0), direct) < 1e-13
    # ut/ut_fft.cpp:53-78: pFFT o pFFT_1 * (2/N) = id
    x = rng.uniform(-1, 1, N)
    assert O.rel_l2(O.pFFT(O.pFFT_1(x, 2.0 / N), 1.0), x) < 1e-14
    # delta input: sFFT(delta_{j=1}) = sin(pi k / N)
    d = np.zeros(N - 1); d[0] = 1.0
    assert np.allclose(O.sFFT(d, 1.0), np.sin(math.pi * np.arange(1, N) / N), atol=1e-15)


def test_lapl_cube_golden(golden):
    n = 15; dx = 1.0 / n; l = 1 + dx
    a = O.LaplCube(dx, dx, dx, l, l, l, n, n, n).solve(golden["cube_d15_rhs"])
    assert O.rel_l2(a, golden["cube_d15_ans"]) < TOL
    a = O.LaplCube(0.1, 0.2, 0.3, 1.6, 3.2, 4.8, n, n, n).solve(golden["cube_d15_rhs"])
    assert O.rel_l2(a, golden["cube_aniso15_ans"]) < TOL      # eigenvalue aliasing quirk
    a = O.LaplCube(0.1, 0.2, 0.3, 3.2, 3.2, 2.4, 31, 15, 7).solve(golden["cube_ragged_rhs"])
    assert O.rel_l2(a, golden["cube_ragged_ans"]) < TOL
    n = 16; dx = 2 * math.pi / n; l = 2 * math.pi
    a = O.LaplCube(dx, dx, dx, l, l, l, n, n, n, True).solve(golden["cube_p16_rhs"])
    assert O.rel_l2(a, golden["cube_p16_ans"]) < TOL


def test_lapl_cube_analytic():
    # ut/ut_lapl_cube.cpp:39-125: u = sin^2 x + cos^2 y + sin^2 z on [0,1]^3, ghosts folded into RHS, < 1e-4
    n = 31; d = 1.0 / (n + 1)
    c = np.arange(0, n + 2) * d
    Z, Y, X = np.meshgrid(c, c, c, indexing="ij")
    u = np.sin(X) ** 2 + np.cos(Y) ** 2 + np.sin(Z) ** 2
    f = 2 * np.cos(2 * X) - 2 * np.cos(2 * Y) + 2 * np.cos(2 * Z)
    rhs = f[1:-1, 1:-1, 1:-1].copy()
    d2 = d * d
    rhs[0] -= u[0, 1:-1, 1:-1] / d2; rhs[-1] -= u[-1, 1:-1, 1:-1] / d2
    rhs[:, 0] -= u[1:-1, 0, 1:-1] / d2; rhs[:, -1] -= u[1:-1, -1, 1:-1] / d2
    rhs[:, :, 0] -= u[1:-1, 1:-1, 0] / d2; rhs[:, :, -1] -= u[1:-1, 1:-1, -1] / d2
    a = O.LaplCube(d, d, d, 1.0, 1.0, 1.0, n, n, n).solve(rhs)
    err = np.max(np.abs(a - u[1:-1, 1:-1, 1:-1])) / np.max(np.abs(u))
    assert err < 1e-4


def test_ns_cube_golden(golden):
    ns = O.NSCube(nx=15, nz=15, Re=100.0, dt=0.01)
    done = 0
    for steps in (1, 2, 10):
        for _ in range(steps - done):
            ns.step()
        done = steps
        cat_o = np.concatenate([ns.fields()[f].ravel() for f in "uvwp"])
        cat_g = np.concatenate([golden[f"nscube15_s{steps}_{f}"] for f in "uvwp"])
        assert O.rel_l2(cat_o, cat_g) < 1e-12
        for f in "uvwp":
            assert O.rel_l2(ns.fields()[f], golden[f"nscube15_s{steps}_{f}"]) < 1e-11, (steps, f)


def test_cyl_rect_golden(golden):
    for zp in (False, True):
        nr, nz, nphi = 16, (16 if zp else 15), 16
        R0, R1 = math.pi / 2, math.pi
        dr = (R1 - R0) / nr; dz = 10.0 / nz
        tag = "p" if zp else "d"
        a = O.LaplCyl3FFT2(dr, dz, R0 - dr / 2, R1 - R0 + dr, 10.0 if zp else 10.0 + dz, nr, nz, nphi, zp).solve(
            golden[f"cyl_{tag}_rhs"])
        assert O.rel_l2(a, golden[f"cyl_{tag}_ans"]) < TOL
    nx, ny = 31, 15; dx, dy = 0.1, 0.05
    a = O.LaplRect(dx, dy, dx * (nx + 1), dy * (ny + 1), nx, ny).solve(golden["rect_rhs"])
    assert O.rel_l2(a, golden["rect_ans"]) < 1e-12
    a = O.LaplRectFFT2(dx, dy, dx * (nx + 1), dy * (ny + 1), nx, ny).solve(golden["rect_rhs"])
    assert O.rel_l2(a, golden["rectfft2_ans"]) < TOL
    # ut/ut_lapl_rect.cpp:384-455: LaplRect (gtsv) == LaplRectFFT2
    assert O.rel_l2(golden["rect_ans"], golden["rectfft2_ans"]) < 1e-12


# ---- live comparison with the compiled reference (skipped when oracle/_ref is absent) ----

def test_live_reference_cube(ref):
    n = 63; dx = 1.0 / n; l = 1 + dx
    rhs = O.synthetic_rhs((n, n, n), seed=7)
    assert O.rel_l2(O.LaplCube(dx, dx, dx, l, l, l, n, n, n).solve(rhs),
                    ref.LaplCube(dx, dx, dx, l, l, l, n, n, n).solve(rhs)) < TOL


def test_live_reference_ns_cube(ref):
    R = ref.NSCube(nx=31, nz=31, Re=250.0, dt=0.01)
    P = O.NSCube(nx=31, nz=31, Re=250.0, dt=0.01)
    for _ in range(5):
        R.step(); P.step()
    cat_o = np.concatenate([P.fields()[f].ravel() for f in "uvwp"])
    cat_r = np.concatenate([R.field(f) for f in "uvwp"])
    assert O.rel_l2(cat_o, cat_r) < 1e-13


def test_lapack_restatement_vs_scipy(ref):
    """oracle/lapack_gt.c (linked into oracle/_ref) against SciPy's LAPACK, via LaplRect (gtsv)
    and LaplCyl3FFT2 (gttrf/gttrs) with non-trivial scale hooks."""
    import scipy.linalg.lapack as la
    rng = np.random.default_rng(3)
    nx, ny = 31, 15; dx, dy = 0.1, 0.05
    sc = [rng.uniform(0.5, 1.5, nx + 1) for _ in range(3)]
    R = ref.LaplRect("rect", dx, dy, dx * (nx + 1), dy * (ny + 1), nx, ny, 0)
    R.set_scales(*sc)
    P = O.LaplRect(dx, dy, dx * (nx + 1), dy * (ny + 1), nx, ny)
    P.lm_y_scale, P.L_scale, P.U_scale = sc
    rhs = rng.uniform(-1, 1, (ny, nx))
    assert O.rel_l2(P.solve(rhs), R.solve(rhs)) < 1e-12
    # direct check of the Thomas restatement against dgtsv
    n = 40
    dl = rng.uniform(0.1, 1, n - 1); du = rng.uniform(0.1, 1, n - 1); d = -(3 + rng.uniform(0, 1, n))
    b = rng.uniform(-1, 1, n)
    x_la = la.dgtsv(dl.copy(), d.copy(), du.copy(), b.copy())[3]
    Lr = np.concatenate([[0.0], dl]); Ur = np.concatenate([du, [0.0]])
    assert O.rel_l2(O.tridiag_solve(Lr, d, Ur, b), x_la) < 1e-14
