"""GPU: the thin drivers under examples/ (the reference drivers' command lines over the drop-in classes, state kept
on the device between plot intervals) reproduce the compiled reference: fields to 1e-12, VTK files to the printed
digits.  Sorted last on purpose: it exercises everything else."""
import math
import subprocess

import numpy as np
import pytest

from oracle import fdm_oracle as O
from tests import cxx_build

pytestmark = pytest.mark.gpu


def vtk_vectors(path):
    lines = open(path).read().splitlines()
    k = lines.index("VECTORS u double")
    return lines[:k + 1], np.array([[float(x) for x in ln.split()] for ln in lines[k + 1:]])


def test_fdm_ns_cube_example(tmp_path, ref):
    exe = cxx_build.build_example("fdm_ns_cube", str(tmp_path / "fdm_ns_cube"))
    n, steps = 31, 30
    r = subprocess.run([exe, f"--ns:nx={n}", f"--ns:nz={n}", "--ns:Re=250", "--ns:dt=0.01", f"--ns:steps={steps}",
                        "--plot:interval=10", "--plot:png=1", "--plot:vtk=1", "--out:prefix=run"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    assert "It took me" in r.stdout
    R = ref.NSCube(nx=n, nz=n, Re=250.0, dt=0.01)
    d = 2 * math.pi / n
    P = ref.VelocityPlotter(d, d, d, n, n, n, -math.pi, math.pi, -math.pi, math.pi, -math.pi, math.pi)
    for step in (0, 10, 20, 30):
        if step:
            R.step(10)
        P.update(*(R.field(f) for f in "uvw"))
        P.vtk_out(tmp_path / "ref.vtk", step)
        hg, vg = vtk_vectors(tmp_path / f"step_{step:07d}.vtk")
        hr, vr = vtk_vectors(tmp_path / "ref.vtk")
        assert hg == hr and vg.shape == vr.shape and np.max(np.abs(vg - vr)) < 1.5e-6
        assert (tmp_path / f"step_{step:07d}.ppm").read_bytes().startswith(b"P6\n")
    got = np.concatenate([np.fromfile(tmp_path / f"run_{f}.bin") for f in "uvwp"])
    want = np.concatenate([R.field(f) for f in "uvwp"])
    assert O.rel_l2(got, want) < 1e-12


@pytest.mark.parametrize("zperiod,nz", [(1, 32), (0, 31)])
def test_fdm_ns_cyl_example(tmp_path, ref, zperiod, nz):
    exe = cxx_build.build_example("fdm_ns_cyl", str(tmp_path / "fdm_ns_cyl"))
    nr, nphi, steps = 32, 32, 20
    r = subprocess.run([exe, f"--ns:nr={nr}", f"--ns:nz={nz}", f"--ns:nphi={nphi}", "--ns:Re=200", "--ns:dt=0.01",
                        f"--ns:steps={steps}", f"--ns:zperiod={zperiod}", "--plot:interval=10", "--plot:png=1",
                        f"--plot:vtk={zperiod}", "--out:prefix=run"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    R = ref.NSCyl(zperiodic=bool(zperiod), nr=nr, nz=nz, nphi=nphi, Re=200.0, dt=0.01)
    R.step(steps)
    got = np.concatenate([np.fromfile(tmp_path / f"run_{f}.bin") for f in "uvwp"])
    want = np.concatenate([R.field(f) for f in "uvwp"])
    assert O.rel_l2(got, want) < 1e-12
    assert (tmp_path / f"step_{steps:07d}.ppm").read_bytes().startswith(b"P6\n")
    if zperiod:
        r0, R0, h1, h2 = math.pi / 2, math.pi, 0.0, 10.0
        P = ref.VelocityPlotter((R0 - r0) / nr, (h2 - h1) / nz, 2 * math.pi / nphi, nr, nz, nphi, r0, R0, h1, h2, 0.0,
                                2 * math.pi, cyl=True, zperiodic=True, yperiodic=True)
        P.update(*(R.field(f) for f in "uvw"))
        P.vtk_out(tmp_path / "ref.vtk", steps)
        hg, vg = vtk_vectors(tmp_path / f"step_{steps:07d}.vtk")
        hr, vr = vtk_vectors(tmp_path / "ref.vtk")
        assert hg == hr and vg.shape == vr.shape
        assert np.array_equal(np.isnan(vg), np.isnan(vr)) and np.nanmax(np.abs(vg - vr)) < 1.5e-6


def test_fdm_ns_cyl_refuses_stabilisation(tmp_path):
    exe = cxx_build.build_example("fdm_ns_cyl", str(tmp_path / "fdm_ns_cyl"))
    r = subprocess.run([exe, "--st:enable=1"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 2 and "not supported" in r.stderr


def test_fdm_nbody_example(tmp_path, ref, capfd):
    exe = cxx_build.build_example("fdm_nbody", str(tmp_path / "fdm_nbody"))
    n, N, steps = 32, 3000, 12
    r = subprocess.run([exe, f"--nbody:n={n}", f"--nbody:N={N}", f"--nbody:steps={steps}", "--plot:interval=5",
                        "--out:prefix=nb"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    assert r.stdout.count("step=") == 3 and "total:" in r.stdout
    R = ref.NBody(n=n, N=N)
    R.step(steps)
    capfd.readouterr()
    for f in "xva":
        assert O.rel_l2(np.fromfile(tmp_path / f"nb_{f}.bin").reshape(N, 3), R.bodies(f)) < 1e-12, f
    r = subprocess.run([exe, "--nbody:local=1"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 2 and "not supported" in r.stderr
