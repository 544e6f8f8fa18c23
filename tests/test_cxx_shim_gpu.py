"""GPU: the drop-in C++ classes (fdm_b200/cxx) called like the reference's callers, checked
against the oracle.  Bar 1e-12 rel-L2 (fp64); 1e-5 for the float instantiation."""
import math
import subprocess

import numpy as np
import pytest

from oracle import fdm_oracle as O
from tests import cxx_build

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    return cxx_build.build(str(tmp_path_factory.mktemp("cxx") / "shim_check"))


def run_cube(exe, tmp_path, mode, n, rhs):
    rhs.tofile(tmp_path / "rhs.bin")
    subprocess.run([exe, mode, str(n), str(tmp_path / "rhs.bin"), str(tmp_path / "ans.bin")], check=True)
    return np.fromfile(tmp_path / "ans.bin").reshape(n, n, n)


def test_cxx_lapl_cube_dirichlet(exe, tmp_path):
    n = 31; dx = 1.0 / n; l = 1 + dx
    rhs = O.synthetic_rhs((n, n, n), seed=5)
    a = run_cube(exe, tmp_path, "cube", n, rhs)
    assert O.rel_l2(a, O.LaplCube(dx, dx, dx, l, l, l, n, n, n).solve(rhs)) < 1e-12


def test_cxx_lapl_cube_periodic(exe, tmp_path):
    n = 32; dx = 2 * math.pi / n; l = 2 * math.pi
    rhs = O.synthetic_rhs((n, n, n), seed=6); rhs -= rhs.mean()
    a = run_cube(exe, tmp_path, "cubep", n, rhs)
    assert O.rel_l2(a, O.LaplCube(dx, dx, dx, l, l, l, n, n, n, True).solve(rhs)) < 1e-12


def test_cxx_lapl_cube_float(exe, tmp_path):
    n = 15; dx = 1.0 / n; l = 1 + dx
    rhs = O.synthetic_rhs((n, n, n), seed=7).astype(np.float32).astype(np.float64)
    a = run_cube(exe, tmp_path, "cubef", n, rhs)
    assert O.rel_l2(a, O.LaplCube(dx, dx, dx, l, l, l, n, n, n).solve(rhs)) < 1e-5


def test_cxx_ns_cube(exe, tmp_path):
    n, steps = 15, 8
    r = subprocess.run([exe, "ns", str(n), str(steps), str(tmp_path / "ns"), "--ns:Re=250", "--ns:dt=0.01"],
                       check=True, capture_output=True, text=True)
    assert f"time_index {steps}" in r.stdout
    po = O.NSCube(nx=n, nz=n, Re=250.0, dt=0.01)
    for _ in range(steps):
        po.step()
    got = np.concatenate([np.fromfile(tmp_path / f"ns_{f}.bin") for f in "uvwp"])
    want = np.concatenate([po.fields()[f].ravel() for f in "uvwp"])
    assert O.rel_l2(got, want) < 1e-12


@pytest.mark.parametrize("mode", ["rect", "rectfft2"])
def test_cxx_lapl_rect_with_public_scales(exe, tmp_path, ref, mode):
    # src/velocity_plot.h:113-127: the caller overwrites lm_y_scale / L_scale / U_scale after construction
    nx, ny = 63, 31
    dx = (math.pi / 2) / nx; dy = 10.0 / ny
    rhs = O.synthetic_rhs((ny, nx), seed=21)
    rhs.tofile(tmp_path / "rhs.bin")
    subprocess.run([exe, mode, str(nx), str(ny), str(tmp_path / "rhs.bin"), str(tmp_path / "ans.bin")], check=True)
    got = np.fromfile(tmp_path / "ans.bin").reshape(ny, nx)
    j = np.arange(nx + 1, dtype=np.float64)
    r = math.pi / 2 + j * dx - dx / 2
    lm = 1.0 / r / r; U = (r + dx / 2) / r; L = (r - dx / 2) / r
    lm[0] = U[0] = L[0] = 1.0
    R = ref.LaplRect("rect" if mode == "rect" else "fft2", dx, dy, math.pi / 2 + dx, 10.0 + dy, nx, ny, 0)
    R.set_scales(lm, L, U)
    assert O.rel_l2(got, R.solve(rhs)) < 1e-12


def test_cxx_lapl_cyl(exe, tmp_path, ref):
    nr, nz, nphi = 32, 31, 32
    rhs = O.synthetic_rhs((nphi, nz, nr), seed=22)
    rhs.tofile(tmp_path / "rhs.bin")
    subprocess.run([exe, "cyl", str(nr), str(nz), str(nphi), str(tmp_path / "rhs.bin"), str(tmp_path / "ans.bin")], check=True)
    got = np.fromfile(tmp_path / "ans.bin").reshape(nphi, nz, nr)
    R, r0, h = math.pi, math.pi / 2, 10.0
    dr = (R - r0) / nr; dz = h / nz
    want = ref.LaplCyl3FFT2(dr, dz, r0 - dr / 2, R - r0 + dr, h + dz, nr, nz, nphi).solve(rhs)
    assert O.rel_l2(got, want) < 1e-12


@pytest.mark.parametrize("lsteps", [0, 3])
def test_cxx_ns_cyl(exe, tmp_path, ref, lsteps):
    # README Taylor-vortex run (32 x 31 x 32, Re = 200, dt = 0.01), then optionally L_step about the current state
    steps = 5
    r = subprocess.run([exe, "nscyl", str(steps), str(lsteps), str(tmp_path / "nc"), "--ns:Re=200", "--ns:dt=0.01"],
                       check=True, capture_output=True, text=True)
    assert f"time_index {steps + lsteps}" in r.stdout
    po = ref.NSCyl(nr=32, nz=31, nphi=32, Re=200.0, dt=0.01)
    po.step(steps)
    if lsteps:
        for f in "uvw":
            po.set_field(f + "0", po.field(f))
        po.step(lsteps, linear=True)
    got = np.concatenate([np.fromfile(tmp_path / f"nc_{f}.bin") for f in "uvwp"])
    want = np.concatenate([po.field(f).ravel() for f in "uvwp"])
    assert O.rel_l2(got, want) < 1e-12


def test_cxx_ns_cyl_spectral_driver_pattern(exe, tmp_path, ref):
    """test/test_ns_cyl_spectral.cpp:72-104 against the drop-in NSCyl with NO B200-specific call: the caller writes
    ns.w0 through the public tensor, sets the public member U0 = 0, assigns ns.u/v/w/p from its own storage, runs
    L_step() and copies the state out.  The header notices the host writes (mirror digests) and the changed U0."""
    nr, nz, nphi, lsteps = 16, 16, 16, 3
    kv = dict(nr=nr, nz=nz, nphi=nphi, Re=150.0, dt=0.01)
    # extents of the driver's interior tensors: u [nphi][nz][1..nr-1], v, w, p [nphi][nz][1..nr]
    sizes = [nphi * nz * (nr - 1)] + [nphi * nz * nr] * 3
    x = O.synthetic_rhs((sum(sizes),), seed=31) * 1e-2
    x.tofile(tmp_path / "x.bin")
    subprocess.run([exe, "nscylspec", str(lsteps), str(tmp_path / "sp"), str(tmp_path / "x.bin")]
                   + [f"--ns:{k}={v}" for k, v in kv.items()], check=True, capture_output=True, text=True)
    # the same sequence on the compiled reference through its public members
    R = ref.NSCyl(zperiodic=True, **kv)
    R0, Rr = math.pi, math.pi / 2
    dr = (R0 - Rr) / nr

    def embed(name, block, jlo, jhi):
        """Range-intersection assignment ns.f = f (tensor.h:103-111): interior block into the full field."""
        full = R.field(name)
        n_r = {"u": nr + 3}.get(name, nr + 2); r_lo = {"u": -1}.get(name, 0)
        nzf = full.size // (nphi * n_r)
        z_lo = -1 if (name == "v" and nzf == nz + 1) else 0
        f = full.reshape(nphi, nzf, n_r)
        f[:, 0 - z_lo:nz - z_lo, jlo - r_lo:jhi - r_lo + 1] = block
        R.set_field(name, f)

    def extract(name, jlo, jhi):
        full = R.field(name)
        n_r = {"u": nr + 3}.get(name, nr + 2); r_lo = {"u": -1}.get(name, 0)
        nzf = full.size // (nphi * n_r)
        z_lo = -1 if (name == "v" and nzf == nz + 1) else 0
        return full.reshape(nphi, nzf, n_r)[:, 0 - z_lo:nz - z_lo, jlo - r_lo:jhi - r_lo + 1].copy()

    w0 = R.field("w0").reshape(nphi, -1, nr + 2)
    j = np.arange(nr + 1)
    r = Rr + dr * j + dr / 2
    w0[:, :nz, :nr + 1] = (-1.0 * Rr ** 2 / (R0 ** 2 - Rr ** 2) + 1.0 * Rr ** 2 * R0 ** 2 / (R0 ** 2 - Rr ** 2) / r / r)[None, None, :]
    R.set_field("w0", w0)
    R.set_u0(0.0)
    xin = x.copy()
    for it in range(2):
        off = 0
        for name, sz, (jlo, jhi) in zip("uvwp", sizes, [(1, nr - 1)] + [(1, nr)] * 3):
            embed(name, xin[off:off + sz].reshape(nphi, nz, jhi - jlo + 1), jlo, jhi); off += sz
        R.step(lsteps, linear=True)
        want = np.concatenate([extract(name, jlo, jhi).ravel() for name, (jlo, jhi) in zip("uvwp", [(1, nr - 1)] + [(1, nr)] * 3)])
        got = np.fromfile(tmp_path / f"sp_y{it}.bin")
        assert O.rel_l2(got, want) < 1e-12, f"iteration {it}"
        xin = want


def test_lapl_rect_rejects_non_dominant_scales():
    """LaplRect's device recurrence does not pivot; scales for which LAPACK gtsv (src/lapl_rect.cpp:90) would are
    refused instead of answered differently."""
    import fdm_b200
    nx, ny = 31, 15
    S = fdm_b200.LaplRect(0.1, 0.1, 3.2, 1.6, nx, ny)
    ok = np.ones(nx + 1)
    S.set_scales(ok, ok, ok)
    bad = ok.copy(); bad[5] = 1.5
    with pytest.raises(fdm_b200.FdmB200Error):
        S.set_scales(ok, bad, ok)                    # |L| + |U| = 2.5 > 2
    neg = ok.copy(); neg[7] = -0.5
    with pytest.raises(fdm_b200.FdmB200Error):
        S.set_scales(neg, None, None)
    rhs = O.synthetic_rhs((ny, nx), seed=3)
    a = S.solve(rhs)                                  # the rejected calls left the handle's scales untouched
    assert O.rel_l2(a, O.LaplRect(0.1, 0.1, 3.2, 1.6, nx, ny).solve(rhs)) < 1e-12


def test_reference_driver_fdm_ns_cube_on_the_gpu_path(tmp_path, ref):
    """The reference's own driver (test/test_ns_cube.cpp, unmodified; built by __graft_entry__.build() where the
    reference tree exists) runs the README cavity case on the B200 path and writes its VTK files through the
    unmodified velocity_plotter::vtk_out; the cell-centred velocities must equal the compiled reference's."""
    import os
    exe = cxx_build.DRIVER
    if not os.path.exists(exe):
        if not os.path.isdir("/root/reference/src"):
            pytest.skip("tests/cxx/_build/fdm_ns_cube not built (needs /root/reference at build time)")
        cxx_build.build_reference_driver(cxx_build.make_overlay(str(tmp_path / "overlay")))
    n, steps = 31, 20
    r = subprocess.run([exe, f"--ns:nx={n}", f"--ns:nz={n}", "--ns:Re=250", "--ns:dt=0.01", f"--ns:steps={steps}",
                        "--plot:png=0", "--plot:vtk=1", "--plot:interval=10"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    assert "It took me" in r.stdout
    R = ref.NSCube(nx=n, nz=n, Re=250.0, dt=0.01)
    for step in (0, 10, 20):
        if step:
            R.step(10)
        lines = (tmp_path / f"step_{step:07d}.vtk").read_text().splitlines()
        assert lines[1] == f"step {step}" and lines[3] == "DATASET STRUCTURED_POINTS"
        k = lines.index("VECTORS u double")
        got = np.array([[float(x) for x in ln.split()] for ln in lines[k + 1:k + 1 + n ** 3]])
        u = R.field("u").reshape(n + 2, n + 2, n + 3); v = R.field("v").reshape(n + 2, n + 3, n + 2)
        w = R.field("w").reshape(n + 3, n + 2, n + 2)
        # src/velocity_plot.cpp:205-215: i,k,j = 1..n; u index j -> j+1, v index k -> k+1, w index i -> i+1
        uc = 0.5 * (u[1:n + 1, 1:n + 1, 2:n + 2] + u[1:n + 1, 1:n + 1, 1:n + 1])
        vc = 0.5 * (v[1:n + 1, 2:n + 2, 1:n + 1] + v[1:n + 1, 1:n + 1, 1:n + 1])
        wc = 0.5 * (w[2:n + 2, 1:n + 1, 1:n + 1] + w[1:n + 1, 1:n + 1, 1:n + 1])
        want = np.stack([uc.ravel(), vc.ravel(), wc.ravel()], axis=1)
        assert np.max(np.abs(got - want)) < 1.5e-6        # "%f": six decimals


def test_cxx_velocity_plotter_on_ns_cube(exe, tmp_path, ref):
    """The drop-in velocity_plotter used like test/test_ns_cube.cpp:24-50: host use(u,v,w), device use(ns) and the
    float instantiation give the reference plotter's stream functions on the same fields; VTK files byte-identical."""
    n, steps = 31, 12
    r = subprocess.run([exe, "vplot", str(n), str(steps), str(tmp_path / "vp"), "--ns:Re=250", "--ns:dt=0.01"],
                       check=True, capture_output=True, text=True)
    assert f"time_index {steps}" in r.stdout
    u, v, w = (np.fromfile(tmp_path / f"vp_{f}.bin") for f in "uvw")
    d = 2 * math.pi / n
    R = ref.VelocityPlotter(d, d, d, n, n, n, -math.pi, math.pi, -math.pi, math.pi, -math.pi, math.pi)
    R.update(u, v, w)
    for s in ("psi_x", "psi_y", "psi_z"):
        want = R.slice(s)
        assert np.abs(want).max() > 0 or s != "psi_y"
        for kind in ("host", "dev"):
            assert O.rel_l2(np.fromfile(tmp_path / f"vp_{kind}_{s}.bin"), want) < 1e-12, (s, kind)
    assert O.rel_l2(np.fromfile(tmp_path / "vp_flt_psi_y.bin"), R.slice("psi_y")) < 1e-5
    R.vtk_out(tmp_path / "ref.vtk", steps)
    want = (tmp_path / "ref.vtk").read_bytes()
    assert (tmp_path / "vp_host.vtk").read_bytes() == want
    assert (tmp_path / "vp_dev.vtk").read_bytes() == want
    ppm = (tmp_path / "vp_dev.ppm").read_bytes()
    assert ppm.startswith(b"P6\n")


def test_cxx_velocity_plotter_on_ns_cyl(exe, tmp_path, ref):
    """test/test_ns_cyl.cpp:53-92 (periodic z): cylindrical column scales, hexahedral VTK."""
    nr, nz, nphi, steps = 32, 32, 32, 10
    r = subprocess.run([exe, "vplotcyl", str(steps), "0", str(tmp_path / "vc"), f"--ns:nr={nr}", f"--ns:nz={nz}",
                        f"--ns:nphi={nphi}", "--ns:Re=200", "--ns:dt=0.01"], check=True, capture_output=True, text=True)
    assert f"time_index {steps}" in r.stdout
    u, v, w = (np.fromfile(tmp_path / f"vc_{f}.bin") for f in "uvw")
    r0, R0, h1, h2 = math.pi / 2, math.pi, 0.0, 10.0
    R = ref.VelocityPlotter((R0 - r0) / nr, (h2 - h1) / nz, 2 * math.pi / nphi, nr, nz, nphi, r0, R0, h1, h2, 0.0,
                            2 * math.pi, cyl=True, zperiodic=True, yperiodic=True)
    R.update(u, v, w)
    for s in ("psi_x", "psi_y", "psi_z"):
        assert O.rel_l2(np.fromfile(tmp_path / f"vc_dev_{s}.bin"), R.slice(s)) < 1e-12, s
    R.vtk_out(tmp_path / "ref.vtk", steps)
    la = (tmp_path / "ref.vtk").read_text().splitlines()
    lb = (tmp_path / "vc_dev.vtk").read_text().splitlines()
    k = la.index("VECTORS u double")
    assert la[:k + 1] == lb[:k + 1] and len(la) == len(lb)


def test_reference_driver_with_the_native_plotter(tmp_path, ref):
    """The unmodified test/test_ns_cube.cpp built with velocity_plot.h replaced as well (no src/velocity_plot.cpp, no
    plplot): its VTK files equal the compiled reference's byte for byte, and plot() leaves its panels as .ppm."""
    import os
    exe = cxx_build.DRIVER_NATIVE_PLOT
    if not os.path.exists(exe):
        if not os.path.isdir("/root/reference/src"):
            pytest.skip("tests/cxx/_build/fdm_ns_cube_native_plot not built (needs /root/reference at build time)")
        cxx_build.build_reference_driver_native_plot(
            cxx_build.make_overlay(str(tmp_path / "overlay"), native_plotter=True))
    n, steps = 31, 20
    r = subprocess.run([exe, f"--ns:nx={n}", f"--ns:nz={n}", "--ns:Re=250", "--ns:dt=0.01", f"--ns:steps={steps}",
                        "--plot:png=1", "--plot:vtk=1", "--plot:interval=10"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    Rn = ref.NSCube(nx=n, nz=n, Re=250.0, dt=0.01)
    d = 2 * math.pi / n
    P = ref.VelocityPlotter(d, d, d, n, n, n, -math.pi, math.pi, -math.pi, math.pi, -math.pi, math.pi)
    for step in (0, 10, 20):
        if step:
            Rn.step(10)
        P.update(*(Rn.field(f) for f in "uvw"))
        P.vtk_out(tmp_path / "ref.vtk", step)
        got = (tmp_path / f"step_{step:07d}.vtk").read_text().splitlines()
        want = (tmp_path / "ref.vtk").read_text().splitlines()
        k = want.index("VECTORS u double")
        assert got[:k + 1] == want[:k + 1] and len(got) == len(want)
        a = np.array([[float(x) for x in ln.split()] for ln in got[k + 1:]])
        b = np.array([[float(x) for x in ln.split()] for ln in want[k + 1:]])
        assert np.max(np.abs(a - b)) < 1.5e-6          # the two NS runs agree to 1e-12, "%f" prints six decimals
        assert (tmp_path / f"step_{step:07d}.ppm").read_bytes().startswith(b"P6\n")
