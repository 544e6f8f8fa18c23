"""Size-independent properties of the path, checked on the numpy restatement with randomised sizes and inputs
(hypothesis): the GPU tier asserts the same properties at BASELINE's full sizes, where no CPU run is affordable."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import fdm_oracle as O

POW2 = st.sampled_from([4, 8, 16, 32])


@settings(max_examples=25, deadline=None)
@given(N=POW2, seed=st.integers(0, 2 ** 16), dx=st.floats(0.01, 3.0))
def test_transform_round_trips(N, seed, dx):
    # ut/ut_fft.cpp:53-78,191-228: sFFT o sFFT (2/N) = id, pFFT o pFFT_1 (2/N) = id
    rng = np.random.default_rng(seed)
    s = rng.uniform(-1, 1, N - 1)
    assert np.allclose(O.sFFT(O.sFFT(s, dx), 2.0 / N / dx), s, rtol=0, atol=1e-13)
    p = rng.uniform(-1, 1, N)
    assert np.allclose(O.pFFT(O.pFFT_1(p, dx), 2.0 / N / dx), p, rtol=0, atol=1e-13)


@settings(max_examples=15, deadline=None)
@given(nz=POW2, ny=POW2, nx=POW2, seed=st.integers(0, 2 ** 16), a=st.floats(-3, 3), b=st.floats(-3, 3))
def test_lapl_cube_is_linear_and_inverts_the_stencil(nz, ny, nx, seed, a, b):
    nz, ny, nx = nz - 1, ny - 1, nx - 1                     # Dirichlet: n + 1 = 2^k
    dx, dy, dz = 1.0 / nx, 0.7 / ny, 1.3 / nz
    S = O.LaplCube(dx, dy, dz, 1.0 + dx, 0.7 + dy, 1.3 + dz, nx, ny, nz)
    rng = np.random.default_rng(seed)
    f, g = rng.uniform(-1, 1, (2, nz, ny, nx))
    uf, ug = S.solve(f).reshape(nz, ny, nx), S.solve(g).reshape(nz, ny, nx)
    assert O.rel_l2(S.solve(a * f + b * g).reshape(nz, ny, nx), a * uf + b * ug) < 1e-12 or abs(a) + abs(b) < 1e-6
    # the 7-point Laplacian with zero ghosts applied to the answer gives the right-hand side back
    u = np.pad(uf, 1)
    lap = ((u[1:-1, 1:-1, 2:] - 2 * uf + u[1:-1, 1:-1, :-2]) / dx ** 2 + (u[1:-1, 2:, 1:-1] - 2 * uf + u[1:-1, :-2, 1:-1]) / dy ** 2
           + (u[2:, 1:-1, 1:-1] - 2 * uf + u[:-2, 1:-1, 1:-1]) / dz ** 2)
    # (when two point counts coincide the reference aliases their eigenvalue tables although the spacings differ,
    #  src/lapl_cube.cpp:162,171: the solver then inverts a different operator -- reproduced, not fixed)
    if len({nx, ny, nz}) == 3:
        assert O.rel_l2(lap, f) < 1e-9


@settings(max_examples=10, deadline=None)
@given(nr=st.integers(3, 20), nz=POW2, nphi=POW2, seed=st.integers(0, 2 ** 16))
def test_lapl_cyl_inverts_the_cylindrical_stencil(nr, nz, nphi, seed):
    nz -= 1
    R0, R1 = math.pi / 2, math.pi
    dr, dz, dphi = (R1 - R0) / nr, 10.0 / nz, 2 * math.pi / nphi
    S = O.LaplCyl3FFT2(dr, dz, R0 - dr / 2, R1 - R0 + dr, 10.0 + dz, nr, nz, nphi)
    f = np.random.default_rng(seed).uniform(-1, 1, (nphi, nz, nr))
    u = S.solve(f).reshape(nphi, nz, nr)
    r = (R0 - dr / 2 + dr * np.arange(1, nr + 1))[None, None, :]           # lapl_cyl.cpp:151-159
    up = np.pad(u, ((0, 0), (1, 1), (1, 1)))
    c = up[:, 1:-1, 1:-1]
    lap = (((r + dr / 2) / r * up[:, 1:-1, 2:] - 2 * c + (r - dr / 2) / r * up[:, 1:-1, :-2]) / dr ** 2
           + (up[:, 2:, 1:-1] - 2 * c + up[:, :-2, 1:-1]) / dz ** 2
           + (np.roll(c, -1, 0) - 2 * c + np.roll(c, 1, 0)) / dphi ** 2 / r ** 2)
    assert O.rel_l2(lap, f) < 1e-9


@settings(max_examples=10, deadline=None)
@given(n=st.sampled_from([7, 15]), steps=st.integers(1, 4), Re=st.floats(10, 500))
def test_ns_cube_projection_leaves_a_divergence_free_interior(n, steps, Re):
    """After update_uvwp the discrete divergence vanishes wherever all six faces of a cell were corrected, i.e. away
    from the walls (src/ns_cube.cpp:246-270 leaves the wall-normal faces untouched)."""
    P = O.NSCube(nx=n, nz=n, Re=Re, dt=0.005)
    for _ in range(steps):
        P.step()
    u, v, w = P.u, P.v, P.w
    I = K = J = (2, n - 1)
    sh = lambda t, I, K, J: t.v(I, K, J)      # noqa: E731
    div = ((sh(u, I, K, J) - sh(u, I, K, (1, n - 2))) / P.dx + (sh(v, I, K, J) - sh(v, I, (1, n - 2), J)) / P.dy
           + (sh(w, I, K, J) - sh(w, (1, n - 2), K, J)) / P.dz)
    scale = max(np.abs(u.a).max() / P.dx, 1e-30)
    assert np.abs(div).max() < 1e-9 * scale
