"""GPU parity tests for the z-slab decomposed LaplCube solve through the C ABI.

Needs >= 2 visible B200s (`gpurun --gpus 2|4|8`); skipped on a single-GPU box.  Two ways of wiring
the ranks are covered: all ranks in this process (attach_local) and one process per GPU launched
with torch.distributed.run (IPC handles exchanged through the process group).
Bar: relative L2 <= 1e-12 (fp64) of the gathered slabs against the oracle's full solve."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import fdm_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fb():
    import fdm_b200
    if fdm_b200.lib().fdmb_device_count() < 2:
        pytest.skip("the sharded solve needs at least 2 GPUs")
    return fdm_b200


def solve_in_process(fb, args, rhs, P, periodic=False, repeats=1):
    import torch
    L = fb.lib()
    solvers, bufs = [], []
    for r in range(P):
        fb.capi.check(L.fdmb_set_device(r), "set_device")
        s = fb.LaplCubeSharded(*args, rank=r, nranks=P, periodic=periodic)
        solvers.append(s)
        with torch.cuda.device(r):
            slab = np.ascontiguousarray(rhs[s.z_first:s.z_first + s.nz_local])
            d_rhs = torch.from_numpy(slab).cuda(r)
            d_ans = torch.full_like(d_rhs, float("nan"))
        bufs.append((d_rhs, d_ans))
    fb.LaplCubeSharded.connect_local(solvers)
    for _ in range(repeats):
        for s, (d_rhs, d_ans) in zip(solvers, bufs):
            s.solve_device(d_ans.data_ptr(), d_rhs.data_ptr())      # asynchronous on the handle's stream
    for r in range(P):
        fb.capi.check(L.fdmb_set_device(r), "set_device")
        fb.capi.check(L.fdmb_device_synchronize(), "sync")
    out = np.concatenate([b[1].cpu().numpy() for b in bufs], axis=0)
    for s in solvers:
        s.close()
    fb.capi.check(L.fdmb_set_device(0), "set_device")
    return out


def ranks_available(fb):
    n = fb.lib().fdmb_device_count()
    return [p for p in (2, 4, 8) if p <= n]


@pytest.mark.parametrize("n", [31, 63, 127])
def test_sharded_dirichlet_vs_oracle(fb, n):
    dx = 1.0 / n; l = 1 + dx
    args = (dx, dx, dx, l, l, l, n, n, n)
    rhs = O.synthetic_rhs((n, n, n), seed=n)
    want = O.LaplCube(*args).solve(rhs)
    for P in ranks_available(fb):
        got = solve_in_process(fb, args, rhs, P, repeats=3)       # repeats exercise the barrier epochs
        assert got.shape == want.shape
        assert O.rel_l2(got, want) < TOL, f"P={P}"


def test_sharded_ragged_vs_oracle(fb):
    nz, ny, nx = 63, 31, 127
    rhs = O.synthetic_rhs((nz, ny, nx), seed=5)
    args = (0.1, 0.2, 0.3, 0.1 * (nx + 1), 0.2 * (ny + 1), 0.3 * (nz + 1), nx, ny, nz)
    want = O.LaplCube(*args).solve(rhs)
    for P in ranks_available(fb):
        assert O.rel_l2(solve_in_process(fb, args, rhs, P), want) < TOL, f"P={P}"


@pytest.mark.parametrize("n", [32, 64])
def test_sharded_periodic_vs_oracle(fb, n):
    dx = 2 * math.pi / n; l = 2 * math.pi
    args = (dx, dx, dx, l, l, l, n, n, n)
    rhs = O.synthetic_rhs((n, n, n), seed=n + 1)
    want = O.LaplCube(*args, periodic=True).solve(rhs)
    for P in ranks_available(fb):
        assert O.rel_l2(solve_in_process(fb, args, rhs, P, periodic=True), want) < TOL, f"P={P}"


def test_sharded_matches_single_gpu_255(fb):
    # full-size property: the sharded solve reproduces the single-GPU solve
    n = 255; dx = 1.0 / n; l = 1 + dx
    args = (dx, dx, dx, l, l, l, n, n, n)
    rhs = O.synthetic_rhs((n, n, n), seed=9)
    single = fb.LaplCube(*args).solve(rhs)
    for P in ranks_available(fb):
        assert O.rel_l2(solve_in_process(fb, args, rhs, P), single) < 1e-13, f"P={P}"


@pytest.mark.parametrize("shape", [(1023, 63, 31), (63, 1023, 31), (127, 1023, 63), (1023, 127, 15), (511, 63, 31), (63, 511, 127)],
                         ids=lambda s: "x".join(map(str, s)))
def test_sharded_long_axis_vs_oracle(fb, shape):
    """Ny or Nz = 1024 / 512: the wide-tile transposing sweeps of the 1023^3 sharded solve (PipeCfg<1024, WIDE>:
    16 columns, 512 threads) and their <512> siblings, ragged in the other axes."""
    nz, ny, nx = shape
    rhs = O.synthetic_rhs(shape, seed=sum(shape))
    args = (0.1, 0.2, 0.3, 0.1 * (nx + 1), 0.2 * (ny + 1), 0.3 * (nz + 1), nx, ny, nz)
    want = O.LaplCube(*args).solve(rhs)
    for P in ranks_available(fb):
        assert O.rel_l2(solve_in_process(fb, args, rhs, P, repeats=2), want) < TOL, f"P={P}"


@pytest.mark.parametrize("n", [255, 1023])
def test_sharded_eigenvector_kat_device(fb, n):
    """The benchmarked sharded configuration (1023^3, z-slabs) against the closed-form eigenvector answer, built slab
    by slab on each device (tests/golden/cube1023.py); no CPU solve needed."""
    import torch
    from tests.golden import cube1023 as G
    L = fb.lib()
    d, l = G.geometry(n)
    for P in ranks_available(fb):
        solvers, bufs = [], []
        for r in range(P):
            fb.capi.check(L.fdmb_set_device(r), "set_device")
            s = fb.LaplCubeSharded(d, d, d, l, l, l, n, n, n, rank=r, nranks=P)
            solvers.append(s)
            with torch.cuda.device(r):
                rhs, want = G.kat_device(torch, n, d, s.z_first, s.nz_local, torch.device("cuda", r))
                bufs.append((rhs, want, torch.full_like(rhs, float("nan"))))
                torch.cuda.synchronize()
        fb.LaplCubeSharded.connect_local(solvers)
        for _ in range(2):
            for s, (rhs, want, ans) in zip(solvers, bufs):
                s.solve_device(ans.data_ptr(), rhs.data_ptr())
        num = den = 0.0
        for r in range(P):
            fb.capi.check(L.fdmb_set_device(r), "set_device")
            fb.capi.check(L.fdmb_device_synchronize(), "sync")
            rhs, want, ans = bufs[r]
            with torch.cuda.device(r):
                num += float(((ans - want) ** 2).sum()); den += float((want ** 2).sum())
        for s in solvers:
            s.close()
        del bufs
        fb.capi.check(L.fdmb_set_device(0), "set_device")
        assert (num / den) ** 0.5 < TOL, f"P={P}"


@pytest.mark.parametrize("shape", [(63, 63, 63), (127, 255, 31), (1023, 63, 31), (63, 1023, 31)], ids=lambda s: "x".join(map(str, s)))
def test_sharded_tma_tile_stores(fb, monkeypatch, shape):
    """FDMB_MG_TMA_STORE=1 (off by default, see lapl_cube.cu): the transposing sweeps leave their tiles through TMA tile
    stores into the peers' buffers -- odd / even slot views per destination rank -- instead of per-value peer stores."""
    monkeypatch.setenv("FDMB_MG_TMA_STORE", "1")
    nz, ny, nx = shape
    rhs = O.synthetic_rhs(shape, seed=sum(shape) + 1)
    args = (0.1, 0.2, 0.3, 0.1 * (nx + 1), 0.2 * (ny + 1), 0.3 * (nz + 1), nx, ny, nz)
    want = O.LaplCube(*args).solve(rhs)
    for P in ranks_available(fb):
        assert O.rel_l2(solve_in_process(fb, args, rhs, P, repeats=2), want) < TOL, f"P={P}"


def test_sharded_rejects_bad_split(fb):
    with pytest.raises(fb.FdmB200Error):
        fb.LaplCubeSharded(1, 1, 1, 16, 16, 16, 15, 15, 15, rank=0, nranks=2)     # transform length 16 < 32
    with pytest.raises(fb.FdmB200Error):
        fb.LaplCubeSharded(1, 1, 1, 64, 64, 64, 63, 63, 63, rank=0, nranks=3)
    s = fb.LaplCubeSharded(1, 1, 1, 64, 64, 64, 63, 63, 63, rank=0, nranks=2)
    with pytest.raises(fb.FdmB200Error):                                           # not connected yet
        s.solve(np.zeros(s.shape))


def test_sharded_one_process_per_gpu(fb, tmp_path):
    """torch.distributed.run, one rank per GPU, IPC handles exchanged through the process group."""
    P = 2
    out = tmp_path / "res.txt"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={P}",
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "mp", "sharded_worker.py"), "--size", "127", "--out", str(out)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    err = float(out.read_text())
    assert err < TOL
