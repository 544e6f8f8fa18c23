"""Particle-mesh N-body step: the numpy restatement against the compiled, unmodified
test/nbody.cpp and the committed golden vectors; the device arithmetic (pm_math.h) emulated on the host.  No GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import fdm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BOX = dict(x0=-10.0, y0=-10.0, z0=-10.0, l=20.0)
GOLDEN = os.path.join(ROOT, "tests", "golden", "golden_nbody_v1.npz")


def restatement(n, x, v, mass, dt=0.001, G=1.0, deposit_all=False):
    P = O.NBodyPM(BOX["x0"], BOX["y0"], BOX["z0"], BOX["l"], n, dt, G, deposit_all=deposit_all)
    P.set_bodies(x, v, mass)
    return P


@pytest.mark.parametrize("n,N", [(16, 500), (32, 2000)])
def test_restatement_vs_compiled_reference(ref, n, N, capfd):
    R = ref.NBody(n=n, N=N, **BOX)
    P = restatement(n, R.bodies("x"), R.bodies("v"), R.bodies("mass"))
    assert P.mass == R.total_mass()
    R.step(1)
    P.step(1)
    for g in ("f", "rhs", "psi", "E"):
        assert O.rel_l2(getattr(P, g), R.grid(g)) < 1e-12, g
    # the reference deposits only the bodies in all-even / all-odd cells: the deposited mass is well below the total
    assert (R.grid("f") - R.grid("f").min()).sum() < 0.5 * R.total_mass()
    R.step(9)
    P.step(9)
    for b in ("x", "v", "a", "aprev"):
        assert O.rel_l2(getattr(P, b), R.bodies(b)) < 1e-12, b
    capfd.readouterr()       # the reference prints one line per step (:505-507)


def test_restatement_vs_golden():
    g = np.load(GOLDEN)
    P = restatement(int(g["n"]), g["x0"], g["v0"], g["mass"])
    P.step(1)
    assert O.rel_l2(P.psi, g["psi1"]) < 1e-12 and O.rel_l2(P.E, g["E1"]) < 1e-12
    P.step(4)
    for b in ("x", "v", "a"):
        assert O.rel_l2(getattr(P, b), g[b + "5"]) < 1e-12, b


@pytest.fixture(scope="module")
def emul():
    here = os.path.join(ROOT, "tests", "host_emul")
    src, lib = os.path.join(here, "pm_host.cpp"), os.path.join(here, "libpm_host.so")
    hdrs = [os.path.join(ROOT, "fdm_b200", "csrc", h) for h in ("pm_math.h", "vplot_math.h")]
    if (not os.path.exists(lib)) or os.path.getmtime(lib) < max(os.path.getmtime(p) for p in [src] + hdrs):
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", src, "-o", lib], check=True)
    L = C.CDLL(lib)
    dp = C.POINTER(C.c_double)
    L.emul_pm_deposit.argtypes = [C.c_int, C.c_longlong] + [C.c_double] * 6 + [C.c_int] + [dp] * 4
    L.emul_pm_gather_move.argtypes = [C.c_int, C.c_longlong] + [C.c_double] * 5 + [dp] * 6 + [C.c_int]
    return L


@pytest.mark.parametrize("deposit_all", [False, True])
def test_device_arithmetic_emulated_on_the_host(emul, deposit_all):
    """fdm_b200/csrc/pm_math.h (what the CUDA kernels call) against the restatement, three steps, with the
    restatement's periodic LaplCube between the two halves.  Includes bodies in the last cell of every axis."""
    n, N = 16, 700
    rng = np.random.default_rng(5)
    x = rng.uniform(-10, 10, (N, 3))
    x[:8] = np.array([[9.99, -10.0, 9.5], [-10.0, 9.999, -10.0], [9.9, 9.9, 9.9], [0.0, 0.0, 0.0], [-10, -10, -10],
                      [8.75, 8.75, 8.75], [-8.75, 9.99999, 1.25], [1.25, 2.5, 9.9999999]])
    v = rng.uniform(-300, 300, (N, 3))           # fast enough to cross the periodic faces within a step
    mass = rng.uniform(0.2, 1.7, N)
    P = restatement(n, x, v, mass, dt=0.002, deposit_all=deposit_all)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))     # noqa: E731
    xs, vs = np.ascontiguousarray(x.T), np.ascontiguousarray(v.T)         # [3][N] like the device arrays
    a_s, ap = np.zeros((3, N)), np.zeros((3, N))
    f, rhs, E = np.empty((n, n, n)), np.empty((n, n, n)), np.empty((n, n, n, 3))
    for step in range(3):
        emul.emul_pm_deposit(n, N, 20.0, -10.0, -10.0, -10.0, 1.0, P.mass, int(deposit_all), p(xs), p(mass), p(f), p(rhs))
        psi = np.ascontiguousarray(P.solver.solve(rhs)).reshape(n, n, n)
        emul.emul_pm_gather_move(n, N, 20.0, -10.0, -10.0, -10.0, 0.002, p(psi), p(E), p(xs), p(vs), p(a_s), p(ap), 1)
        P.step(1)
        assert O.rel_l2(f, P.f) < 1e-13 and O.rel_l2(rhs, P.rhs) < 1e-13
        assert O.rel_l2(E, P.E) < 1e-12
        assert O.rel_l2(a_s.T, P.a) < 1e-12 and O.rel_l2(xs.T, P.x) < 1e-13 and O.rel_l2(vs.T, P.v) < 1e-13
        assert xs.min() >= -10.0 and xs.max() < 10.0
    assert np.abs(xs.T - x).max() > 1.0          # bodies really moved and wrapped
