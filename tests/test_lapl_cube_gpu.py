"""GPU parity tests for LaplCube and the batched 1-D transforms, through the C ABI.
Bar: relative L2 <= 1e-12 (fp64) against the oracle / the compiled reference."""
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


@pytest.mark.parametrize("N", [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048])
def test_fft_batch_vs_oracle(fb, N):
    rng = np.random.default_rng(N)
    batch = 37   # ragged: not a multiple of the rows-per-CTA tile
    x = rng.uniform(-1, 1, (batch, N - 1))
    assert O.rel_l2(fb.fft_batch("sFFT", N, x, 0.37), O.sFFT(x, 0.37)) < 1e-13
    x = rng.uniform(-1, 1, (batch, N))
    assert O.rel_l2(fb.fft_batch("pFFT_1", N, x, 0.37), O.pFFT_1(x, 0.37)) < 1e-14
    assert O.rel_l2(fb.fft_batch("pFFT", N, x, 0.37), O.pFFT(x, 0.37)) < 1e-14


def test_fft_golden(fb, golden):
    for N in (32, 128):
        s = golden[f"sFFT_{N}_in"]
        assert O.rel_l2(fb.fft_batch("sFFT", N, s[1:N], 0.37), golden[f"sFFT_{N}_out"][1:N]) < 1e-13
        s = golden[f"pFFT_1_{N}_in"]
        assert O.rel_l2(fb.fft_batch("pFFT_1", N, s[:N], 0.37), golden[f"pFFT_1_{N}_out"][:N]) < 1e-13
        assert O.rel_l2(fb.fft_batch("pFFT", N, s[:N], 0.37), golden[f"pFFT_{N}_out"][:N]) < 1e-13


def test_fft_roundtrips(fb):
    # ut/ut_fft.cpp:53-78,191-228
    N = 1024
    rng = np.random.default_rng(2)
    x = rng.uniform(-1, 1, (5, N - 1))
    assert O.rel_l2(fb.fft_batch("sFFT", N, fb.fft_batch("sFFT", N, x, 1.0), 2.0 / N), x) < 1e-14
    x = rng.uniform(-1, 1, (5, N))
    assert O.rel_l2(fb.fft_batch("pFFT", N, fb.fft_batch("pFFT_1", N, x, 2.0 / N), 1.0), x) < 1e-14


def test_fft_empty_batch(fb):
    out = fb.fft_batch("sFFT", 32, np.zeros((0, 31)))
    assert out.shape == (0, 31)


def test_cube_golden(fb, golden):
    n = 15; dx = 1.0 / n; l = 1 + dx
    a = fb.LaplCube(dx, dx, dx, l, l, l, n, n, n).solve(golden["cube_d15_rhs"])
    assert O.rel_l2(a, golden["cube_d15_ans"]) < TOL
    a = fb.LaplCube(0.1, 0.2, 0.3, 1.6, 3.2, 4.8, n, n, n).solve(golden["cube_d15_rhs"])
    assert O.rel_l2(a, golden["cube_aniso15_ans"]) < TOL          # aliasing quirk, lapl_cube.cpp:162,171
    a = fb.LaplCube(0.1, 0.2, 0.3, 3.2, 3.2, 2.4, 31, 15, 7).solve(golden["cube_ragged_rhs"])
    assert O.rel_l2(a, golden["cube_ragged_ans"]) < TOL
    n = 16; dx = 2 * math.pi / n; l = 2 * math.pi
    a = fb.LaplCube(dx, dx, dx, l, l, l, n, n, n, True).solve(golden["cube_p16_rhs"])
    assert O.rel_l2(a, golden["cube_p16_ans"]) < TOL


@pytest.mark.parametrize("n", [3, 7, 31, 63, 127, 255])
def test_cube_dirichlet_vs_oracle(fb, n):
    dx = 1.0 / n; l = 1 + dx
    rhs = O.synthetic_rhs((n, n, n), seed=n)
    a = fb.LaplCube(dx, dx, dx, l, l, l, n, n, n).solve(rhs)
    assert O.rel_l2(a, O.LaplCube(dx, dx, dx, l, l, l, n, n, n).solve(rhs)) < TOL


@pytest.mark.parametrize("shape", [(7, 15, 31), (31, 7, 63), (127, 31, 15), (3, 255, 7)])
def test_cube_ragged_vs_oracle(fb, shape):
    nz, ny, nx = shape
    rhs = O.synthetic_rhs(shape, seed=sum(shape))
    args = (0.1, 0.2, 0.3, 0.1 * (nx + 1), 0.2 * (ny + 1), 0.3 * (nz + 1), nx, ny, nz)
    assert O.rel_l2(fb.LaplCube(*args).solve(rhs), O.LaplCube(*args).solve(rhs)) < TOL


@pytest.mark.parametrize("n", [4, 32, 64, 128])
def test_cube_periodic_vs_oracle(fb, n):
    dx = 2 * math.pi / n; l = 2 * math.pi
    rhs = O.synthetic_rhs((n, n, n), seed=n + 1); rhs -= rhs.mean()
    a = fb.LaplCube(dx, dx, dx, l, l, l, n, n, n, True).solve(rhs)
    assert O.rel_l2(a, O.LaplCube(dx, dx, dx, l, l, l, n, n, n, True).solve(rhs)) < TOL


def test_cube_vs_compiled_reference(fb, ref):
    n = 127; dx = 1.0 / n; l = 1 + dx
    rhs = O.synthetic_rhs((n, n, n), seed=1234)
    a = fb.LaplCube(dx, dx, dx, l, l, l, n, n, n).solve(rhs)
    assert O.rel_l2(a, ref.LaplCube(dx, dx, dx, l, l, l, n, n, n).solve(rhs)) < TOL


def test_cube_analytic(fb):
    # ut/ut_lapl_cube.cpp:39-125
    n = 31; d = 1.0 / (n + 1)
    c = np.arange(0, n + 2) * d
    Z, Y, X = np.meshgrid(c, c, c, indexing="ij")
    u = np.sin(X) ** 2 + np.cos(Y) ** 2 + np.sin(Z) ** 2
    f = 2 * np.cos(2 * X) - 2 * np.cos(2 * Y) + 2 * np.cos(2 * Z)
    rhs = f[1:-1, 1:-1, 1:-1].copy(); d2 = d * d
    rhs[0] -= u[0, 1:-1, 1:-1] / d2; rhs[-1] -= u[-1, 1:-1, 1:-1] / d2
    rhs[:, 0] -= u[1:-1, 0, 1:-1] / d2; rhs[:, -1] -= u[1:-1, -1, 1:-1] / d2
    rhs[:, :, 0] -= u[1:-1, 1:-1, 0] / d2; rhs[:, :, -1] -= u[1:-1, 1:-1, -1] / d2
    a = fb.LaplCube(d, d, d, 1.0, 1.0, 1.0, n, n, n).solve(rhs)
    assert np.max(np.abs(a - u[1:-1, 1:-1, 1:-1])) / np.max(np.abs(u)) < 1e-4


def test_cube_full_size_properties(fb):
    """255^3 (BASELINE config 3 solver size): linearity and the discrete-operator residual,
    size-independent checks that need no CPU oracle run."""
    n = 255; d = 1.0 / n; l = 1 + d
    S = fb.LaplCube(d, d, d, l, l, l, n, n, n)
    r1 = O.synthetic_rhs((n, n, n), seed=1); r2 = O.synthetic_rhs((n, n, n), seed=2)
    a1 = S.solve(r1); a2 = S.solve(r2); a12 = S.solve(r1 + 2.0 * r2)
    assert O.rel_l2(a12, a1 + 2.0 * a2) < 1e-13
    # apply the 7-point operator with homogeneous Dirichlet ghosts: must reproduce rhs
    p = np.zeros((n + 2, n + 2, n + 2)); p[1:-1, 1:-1, 1:-1] = a1
    lap = (p[2:, 1:-1, 1:-1] + p[:-2, 1:-1, 1:-1] + p[1:-1, 2:, 1:-1] + p[1:-1, :-2, 1:-1]
           + p[1:-1, 1:-1, 2:] + p[1:-1, 1:-1, :-2] - 6 * a1) / (d * d)
    assert O.rel_l2(lap, r1) < 1e-10


def test_errors(fb):
    with pytest.raises(fb.FdmB200Error):
        fb.LaplCube(0.1, 0.1, 0.1, 3.3, 3.3, 3.3, 32, 32, 32)       # Dirichlet needs 2^k - 1
    with pytest.raises(fb.FdmB200Error):
        fb.LaplCube(0.1, 0.1, 0.1, 3.1, 3.1, 3.1, 31, 31, 31, True)  # periodic needs 2^k
    with pytest.raises(ValueError):
        fb.LaplCube(0.1, 0.1, 0.1, 1.6, 1.6, 1.6, 15, 15, 15).solve(np.zeros(10))


@pytest.mark.parametrize("shape", [(31, 31, 31), (63, 63, 63), (127, 127, 127), (63, 31, 127), (31, 127, 63),
                                   (255, 31, 63)])
def test_cube_blocked_work_layout(fb, ref, monkeypatch, shape):
    """The blocked work-array layout ([yb][z][yi][x], used by default for grids >= 511^2 per plane) forced on
    at sizes the compiled reference finishes in seconds; ragged shapes cover partial last blocks."""
    nz, ny, nx = shape
    args = (0.1, 0.2, 0.3, 0.1 * (nx + 1), 0.2 * (ny + 1), 0.3 * (nz + 1), nx, ny, nz)
    rhs = O.synthetic_rhs(shape, seed=sum(shape))
    want = ref.LaplCube(*args).solve(rhs)
    monkeypatch.setenv("FDMB_BLOCKED", "1")
    S = fb.LaplCube(*args)
    got = S.solve(rhs)
    assert O.rel_l2(got, want) < TOL
    assert O.rel_l2(S.solve(rhs), want) < TOL        # a second solve: padding rows untouched


@pytest.mark.parametrize("n,count", [(63, 1), (63, 2), (127, 5)])
def test_cube_solve_batch_equals_single_solves(fb, n, count):
    """fdmb_lapl_cube_solve_batch pipelines uploads, solves and downloads over two staging pairs; the answers are
    bit-identical to one solve() per right-hand side (page-locked and pageable host memory)."""
    import torch
    dx = 1.0 / n; l = 1 + dx
    S = fb.LaplCube(dx, dx, dx, l, l, l, n, n, n)
    rhs = [O.synthetic_rhs((n, n, n), seed=100 + i) for i in range(count)]
    want = [S.solve(r) for r in rhs]
    # pageable numpy arrays
    ans = [np.full((n, n, n), np.nan) for _ in range(count)]
    S.solve_batch([a.ctypes.data for a in ans], [r.ctypes.data for r in rhs])
    for a, w in zip(ans, want):
        assert np.array_equal(a, w)
    # page-locked buffers, answers rotating through two arrays like bench.py does
    prhs = [torch.from_numpy(r).pin_memory() for r in rhs]
    pans = [torch.empty(n ** 3, dtype=torch.float64).pin_memory() for _ in range(min(count, 2))]
    S.solve_batch([pans[i % 2].data_ptr() for i in range(count)], [t.data_ptr() for t in prhs])
    for i in range(max(0, count - 2), count):           # the last answers written to each rotating buffer survive
        assert np.array_equal(pans[i % 2].numpy().reshape(n, n, n), want[i])
    # a plain solve afterwards still works on the shared staging buffers
    assert np.array_equal(S.solve(rhs[0]), want[0])


@pytest.mark.parametrize("shape", [(255, 255, 255), (200 * 0 + 127, 255, 511), (63, 511, 255)], ids=lambda s: "x".join(map(str, s)))
def test_cube_host_streaming_equals_plain(fb, monkeypatch, shape):
    """The host-pointer entry point streams large grids through the sweeps in z chunks (upload of chunk c+1 beside the
    x / y sweeps of chunk c, downloads beside the inverse sweeps).  Same kernels on the same data: the answer is
    bit-identical to the unchunked path for every chunk count, odd plane counts and ragged last chunks included."""
    nz, ny, nx = shape
    args = (0.1, 0.2, 0.3, 0.1 * (nx + 1), 0.2 * (ny + 1), 0.3 * (nz + 1), nx, ny, nz)
    rhs = O.synthetic_rhs(shape, seed=sum(shape))
    S = fb.LaplCube(*args)
    monkeypatch.setenv("FDMB_HOST_CHUNKS", "1")
    plain = S.solve(rhs)
    for chunks in ("2", "3", "8", "16"):
        monkeypatch.setenv("FDMB_HOST_CHUNKS", chunks)
        assert np.array_equal(S.solve(rhs), plain), chunks
    monkeypatch.delenv("FDMB_HOST_CHUNKS")
    assert np.array_equal(S.solve(rhs), plain)
    if shape == (255, 255, 255):
        assert O.rel_l2(plain, O.LaplCube(*args).solve(rhs)) < TOL
