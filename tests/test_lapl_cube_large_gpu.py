"""GPU parity tests at the sizes bench.py measures (BASELINE configs[4]: LaplCube Dirichlet 1023^3) and for every
instantiation of the persistent sweep kernels those sizes use (k_rows_pipe / k_cols_pipe<512|1024|2048>, 8-column
tiles with the block swizzle, 32 threads per sequence, computed twiddle powers), through the C ABI.

  * the 1-D transforms through the persistent contiguous-axis sweep (fdmb_fft_batch_impl, impl = pipe) against the
    oracle for every length, and the DCT-I (cFFT) against the oracle and the golden vectors;
  * solves with one long axis (1023 / 511 / 2047 along x, y or z) against the oracle: under 2.1 M points each;
  * the full 1023^3 solve against (a) an eigenvector known answer built on the device -- no CPU solve needed -- and
    (b) a strided sample, one row, the norm and the sum of the answer of the UNMODIFIED reference, computed once by
    tests/golden/make_golden_cube1023.py (27.8 s on 8 cores) and committed as golden_cube1023_v1.npz.
Bar: relative L2 <= 1e-12 (fp64)."""
import os

import numpy as np
import pytest

from oracle import fdm_oracle as O
from tests.golden import cube1023 as G

pytestmark = pytest.mark.gpu
TOL = 1e-12
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fb():
    import fdm_b200
    assert fdm_b200.lib().fdmb_device_count() > 0, "GPU tests need a CUDA device"
    return fdm_b200


@pytest.mark.parametrize("N", [32, 64, 128, 256, 512, 1024])
@pytest.mark.parametrize("batch", [1, 37, 300])
def test_fft_batch_pipe_vs_oracle(fb, N, batch):
    """The persistent rows sweep (what LaplCube's x sweeps run), ragged batches: partial last tile, odd tails."""
    rng = np.random.default_rng(N + batch)
    x = rng.uniform(-1, 1, (batch, N - 1))
    got = fb.fft_batch("sFFT", N, x, 0.37, impl="pipe")
    assert O.rel_l2(got, O.sFFT(x, 0.37)) < 1e-13
    assert O.rel_l2(got, fb.fft_batch("sFFT", N, x, 0.37, impl="plain")) < 1e-13
    x = rng.uniform(-1, 1, (batch, N))
    assert O.rel_l2(fb.fft_batch("pFFT_1", N, x, 0.37, impl="pipe"), O.pFFT_1(x, 0.37)) < 1e-14
    assert O.rel_l2(fb.fft_batch("pFFT", N, x, 0.37, impl="pipe"), O.pFFT(x, 0.37)) < 1e-14


def test_fft_batch_pipe_rejects_unsupported_lengths(fb):
    """N = 2048 rows do not fit the persistent contiguous-axis sweep (staging + tile > one SM's shared memory): auto
    falls back to the plain kernel, an explicit request is an error -- never a failed launch."""
    x = np.random.default_rng(1).uniform(-1, 1, (5, 2047))
    assert O.rel_l2(fb.fft_batch("sFFT", 2048, x, 0.37), O.sFFT(x, 0.37)) < 1e-13
    with pytest.raises(fb.FdmB200Error):
        fb.fft_batch("sFFT", 2048, x, 0.37, impl="pipe")
    with pytest.raises(fb.FdmB200Error):
        fb.fft_batch("sFFT", 16, x[:, :15], 0.37, impl="pipe")


@pytest.mark.parametrize("N", [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048])
def test_cfft_vs_oracle(fb, N):
    """FFT<T>::cFFT (src/fft.cpp:368-445): DCT-I with halved end points over N + 1 values."""
    rng = np.random.default_rng(3 * N)
    x = rng.uniform(-1, 1, (37, N + 1))
    assert O.rel_l2(fb.fft_batch("cFFT", N, x, 0.37), O.cFFT(x, 0.37)) < 1e-13
    # O(N^2) definition of src/asp_fft.cpp:404-418 on one row
    j = np.arange(N + 1)
    w = np.ones(N + 1); w[0] = w[N] = 0.5
    want = 0.37 * (np.cos(np.pi * (np.outer(j, j) % (2 * N)) / N) @ (w * x[0]))     # exact argument reduction
    assert O.rel_l2(fb.fft_batch("cFFT", N, x[0], 0.37), want) < 5e-13               # (an O(N^2) fp64 sum is itself noisy)


def test_cfft_golden(fb, golden):
    for N in (32, 128):
        s = golden[f"cFFT_{N}_in"]
        assert O.rel_l2(fb.fft_batch("cFFT", N, s[:N + 1], 0.37), golden[f"cFFT_{N}_out"][:N + 1]) < 1e-13


LONG = [(31, 31, 1023), (31, 1023, 31), (1023, 31, 31), (15, 1023, 127), (127, 15, 1023),
        (31, 31, 511), (31, 511, 31), (511, 31, 31), (63, 511, 63),
        (31, 31, 2047), (31, 2047, 31), (2047, 31, 31)]


@pytest.mark.parametrize("shape", LONG, ids=lambda s: "x".join(map(str, s)))
def test_cube_long_axis_vs_oracle(fb, shape):
    """(nz, ny, nx) with one axis of the benchmarked length: every k_rows_pipe / k_cols_pipe<512|1024|2048>
    instantiation of the single-GPU solve, ragged column tiles (31 is not a multiple of the tile width)."""
    nz, ny, nx = shape
    rhs = O.synthetic_rhs(shape, seed=sum(shape))
    args = (0.1, 0.2, 0.3, 0.1 * (nx + 1), 0.2 * (ny + 1), 0.3 * (nz + 1), nx, ny, nz)
    S = fb.LaplCube(*args)
    want = O.LaplCube(*args).solve(rhs)
    assert O.rel_l2(S.solve(rhs), want) < TOL
    assert O.rel_l2(S.solve(rhs), want) < TOL        # second solve on the same handle


@pytest.mark.parametrize("shape", [(31, 31, 1023), (31, 1023, 31), (1023, 31, 31)], ids=lambda s: "x".join(map(str, s)))
def test_cube_long_axis_vs_compiled_reference(fb, ref, shape):
    nz, ny, nx = shape
    rhs = O.synthetic_rhs(shape, seed=7 + sum(shape))
    d = 1.0 / 1023
    args = (d, d, d, d * (nx + 1), d * (ny + 1), d * (nz + 1), nx, ny, nz)
    assert O.rel_l2(fb.LaplCube(*args).solve(rhs), ref.LaplCube(*args).solve(rhs)) < TOL


@pytest.mark.parametrize("n", [255, 511, 1023])
def test_cube_eigenvector_kat_device(fb, n):
    """rhs = a few discrete sine products (low, middle, highest modes); the answer is known in closed form."""
    import torch
    d, l = G.geometry(n)
    dev = torch.device("cuda", 0)
    rhs, want = G.kat_device(torch, n, d, 0, n, dev)
    ans = torch.full_like(rhs, float("nan"))
    S = fb.LaplCube(d, d, d, l, l, l, n, n, n)
    torch.cuda.synchronize()
    S.solve_device(ans.data_ptr(), rhs.data_ptr())
    fb.capi.check(fb.lib().fdmb_device_synchronize(), "sync")
    assert G.rel_l2_device(torch, ans, want) < TOL
    # the solve does not modify rhs and is repeatable bit for bit
    ans2 = torch.full_like(rhs, float("nan"))
    S.solve_device(ans2.data_ptr(), rhs.data_ptr())
    fb.capi.check(fb.lib().fdmb_device_synchronize(), "sync")
    assert torch.equal(ans, ans2)
    S.close()


def test_cube1023_with_concurrent_memory_traffic(fb):
    """Solves on the handle's stream while another stream saturates DRAM (tile loads arrive late and out of order): the
    answer must stay bit-identical.  Regression test for the tile-ring hand-off (a consumer group two barrier phases
    ahead of a buffer whose previous tile had not landed yet passed its parity wait on the stale phase)."""
    import torch
    n = 1023
    d, l = G.geometry(n)
    dev = torch.device("cuda", 0)
    rhs = torch.rand(n ** 3, dtype=torch.float64, device=dev) - 0.5
    quiet = torch.empty_like(rhs)
    S = fb.LaplCube(d, d, d, l, l, l, n, n, n)
    torch.cuda.synchronize()
    S.solve_device(quiet.data_ptr(), rhs.data_ptr())
    fb.capi.check(fb.lib().fdmb_device_synchronize(), "sync")
    noise_a = torch.empty(1 << 29, dtype=torch.float64, device=dev)      # 4 GB each way
    noise_b = torch.zeros(1 << 29, dtype=torch.float64, device=dev)
    side = torch.cuda.Stream(device=dev)
    for rep in range(3):
        busy = torch.full_like(rhs, float("nan"))
        with torch.cuda.stream(side):
            for _ in range(6):
                noise_a.copy_(noise_b)
        S.solve_device(busy.data_ptr(), rhs.data_ptr())                 # no synchronisation in between
        fb.capi.check(fb.lib().fdmb_device_synchronize(), "sync")
        assert torch.equal(busy, quiet), f"repetition {rep}"
    S.close()


def test_cube1023_vs_reference_sample(fb):
    """The benchmarked configuration against the unmodified reference's own 1023^3 answer (sample, row, norm, sum),
    through the host-pointer entry point the reference's callers use."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_cube1023_v1.npz"))
    n = int(g["n"]); st = int(g["stride"])
    assert n == G.N and int(g["seed"]) == G.SEED
    d, l = G.geometry(n)
    rhs = G.rhs_planes(n, 0, n)
    assert abs(float(np.linalg.norm(rhs.ravel())) / float(g["rhs_norm"]) - 1) < 1e-14      # same inputs
    ans = fb.LaplCube(d, d, d, l, l, l, n, n, n).solve(rhs)
    assert O.rel_l2(ans[::st, ::st, ::st], g["sample"]) < TOL
    assert O.rel_l2(ans[n // 2, n // 3, :], g["row"]) < TOL
    assert abs(float(np.linalg.norm(ans.ravel())) / float(g["ans_norm"]) - 1) < TOL
    assert abs(float(ans.sum(dtype=np.longdouble)) - float(g["ans_sum"])) < 1e-9 * abs(float(g["ans_sum"])) + 1e-6
