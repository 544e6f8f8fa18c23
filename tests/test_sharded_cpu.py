"""CPU tests of the multi-GPU host logic: the slab plan exported by the C ABI and a world_size-2
(gloo) rehearsal of the sharded LaplCube data path against the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_range_partitions_every_axis():
    import fdm_b200
    for periodic in (False, True):
        for N in (4, 32, 128, 1024):
            n = N if periodic else N - 1
            for P in (1, 2, 4, 8):
                if N // P < 2:
                    with pytest.raises(fdm_b200.FdmB200Error):
                        fdm_b200.slab_range(n, periodic, P, 0)
                    continue
                parts = [fdm_b200.slab_range(n, periodic, P, r) for r in range(P)]
                pos = 0
                for first, count in parts:
                    assert first == pos and count > 0
                    pos += count
                assert pos == n
                # every rank but the first owns exactly N/P slots; rank 0 loses the Dirichlet boundary slot
                assert all(c == N // P for _, c in parts[1:])
                assert parts[0][1] == N // P - (0 if periodic else 1)


def test_slab_range_rejects_bad_arguments():
    import fdm_b200
    for bad in [(30, False, 2, 0), (31, False, 3, 0), (31, False, 2, 2), (31, False, 2, -1), (0, False, 1, 0)]:
        with pytest.raises(fdm_b200.FdmB200Error):
            fdm_b200.slab_range(*bad)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_data_path_gloo(tmp_path, world):
    out = tmp_path / "err.txt"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29650 + world),
           os.path.join(ROOT, "tests", "mp", "sharded_emul_worker.py"), "--out", str(out)]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert float(out.read_text()) < 1e-13


def test_ns_owned_planes_tile_the_fields():
    import fdm_b200
    nz = 63
    for f, (lo, hi) in {"u": (0, nz + 1), "w": (-1, nz + 1), "x": (1, nz), "H": (0, nz)}.items():
        for P in (1, 2, 4, 8):
            pos = lo
            for r in range(P):
                z0, n = fdm_b200.owned_planes(nz, f, r, P)
                assert z0 == pos and n > 0
                pos += n
            assert pos == hi + 1
    with pytest.raises(fdm_b200.FdmB200Error):
        fdm_b200.owned_planes(7, "u", 0, 4)          # fewer than 4 planes per rank


@pytest.mark.parametrize("world", [2, 4])
def test_ns_halo_plan_gloo(tmp_path, world):
    out = tmp_path / "plan.txt"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29660 + world),
           os.path.join(ROOT, "tests", "mp", "ns_plan_worker.py"), "--out", str(out)]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert out.read_text() == "ok"
