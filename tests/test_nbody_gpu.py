"""GPU parity tests for the particle-mesh N-body step through the C ABI.
Bar: fp64 rel-L2 <= 1e-12 against the compiled, unmodified test/nbody.cpp on the bodies it seeds."""
import os

import numpy as np
import pytest

from oracle import fdm_oracle as O
from tests.test_nbody_cpu import BOX, GOLDEN, restatement

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def fb():
    import fdm_b200
    assert fdm_b200.lib().fdmb_device_count() > 0, "GPU tests need a CUDA device"
    return fdm_b200


@pytest.mark.parametrize("n,N", [(16, 500), (32, 2000), (64, 20000)])
def test_pm_vs_compiled_reference(fb, ref, n, N, capfd):
    R = ref.NBody(n=n, N=N, **BOX)
    P = fb.NBodyPM(n=n, **BOX)
    P.set_bodies(R.bodies("x"), R.bodies("v"), R.bodies("mass"))
    assert P.N == N
    R.step(1)
    P.step(1)
    for g in ("f", "rhs", "psi", "E"):
        assert O.rel_l2(P.grid(g), R.grid(g)) < TOL, g
    R.step(9)
    P.step(9)
    for b in ("x", "v", "a", "aprev"):
        assert O.rel_l2(P.bodies(b), R.bodies(b)) < TOL, b
    capfd.readouterr()


def test_pm_golden(fb):
    g = np.load(GOLDEN)
    P = fb.NBodyPM(n=int(g["n"]), **BOX)
    P.set_bodies(g["x0"], g["v0"], g["mass"])
    P.step(1)
    assert O.rel_l2(P.grid("psi"), g["psi1"]) < TOL and O.rel_l2(P.grid("E"), g["E1"]) < TOL
    P.step(4)
    for b in ("x", "v", "a"):
        assert O.rel_l2(P.bodies(b), g[b + "5"]) < TOL, b


@pytest.mark.parametrize("deposit_all", [False, True])
def test_pm_vs_restatement_with_wrapping_bodies(fb, deposit_all):
    n, N = 16, 700
    rng = np.random.default_rng(5)
    x = rng.uniform(-10, 10, (N, 3))
    x[:5] = np.array([[9.99, -10.0, 9.5], [-10.0, 9.999, -10.0], [9.9, 9.9, 9.9], [-10, -10, -10], [1.25, 2.5, 9.9999999]])
    v = rng.uniform(-300, 300, (N, 3))
    mass = rng.uniform(0.2, 1.7, N)
    Or = restatement(n, x, v, mass, dt=0.002, deposit_all=deposit_all)
    P = fb.NBodyPM(n=n, dt=0.002, deposit_all=deposit_all, **BOX)
    P.set_bodies(x, v, mass)
    P.calc_a_pm()                      # calc_a_pm alone leaves x, v untouched
    assert np.array_equal(P.bodies("x"), x) and np.array_equal(P.bodies("v"), v)
    Or.calc_a_pm()
    assert O.rel_l2(P.bodies("a"), Or.a) < TOL
    Or.step(3)
    P.step(3)
    for b in ("x", "v", "a"):
        assert O.rel_l2(P.bodies(b), getattr(Or, b)) < TOL, b
    xs = P.bodies("x")
    assert xs.min() >= -10.0 and xs.max() < 10.0 and np.abs(xs - x).max() > 1.0


def test_pm_large_properties(fb):
    """n = 128, N = 10^6 (no CPU run): the deposit conserves mass (the cloud-in-cell weights of a body sum to 1, so
    with every body deposited sum f = sum m - n^3 * mass / l^3, the mean term of :296-302 being a density per CELL),
    the Poisson solve returns a zero-mean potential, and a second handle fed the same bodies in reverse order agrees
    to round-off (the atomics' order does not matter beyond that)."""
    n, N = 128, 1_000_000
    rng = np.random.default_rng(9)
    x = rng.uniform(-10, 10, (N, 3))
    v = np.zeros((N, 3))
    mass = rng.uniform(0.2, 1.7, N)
    P = fb.NBodyPM(n=n, deposit_all=True, **BOX)
    P.set_bodies(x, v, mass)
    P.calc_a_pm()
    f = P.grid("f")
    want = mass.sum() * (1.0 - n ** 3 / BOX["l"] ** 3)
    assert abs(f.sum() - want) < 1e-9 * abs(want)
    assert abs(P.grid("psi").mean()) < 1e-12 * np.abs(P.grid("psi")).max()
    Q = fb.NBodyPM(n=n, deposit_all=True, **BOX)
    Q.set_bodies(x[::-1], v, mass[::-1])
    Q.calc_a_pm()
    assert O.rel_l2(Q.grid("f"), f) < TOL
    assert O.rel_l2(Q.bodies("a")[::-1], P.bodies("a")) < 1e-10


def test_pm_errors(fb):
    with pytest.raises(fb.FdmB200Error):
        fb.NBodyPM(n=24, **BOX)                         # periodic axes need n = 2^k (src/fft.cpp:67)
    P = fb.NBodyPM(n=16, **BOX)
    with pytest.raises(fb.FdmB200Error):
        P.step(1)                                       # no bodies yet
    with pytest.raises(fb.FdmB200Error):
        P.set_bodies(np.array([[10.0, 0.0, 0.0]]), np.zeros((1, 3)), np.ones(1))     # x = origin + l is outside
