"""GPU parity tests for the device-side velocity_plotter through the C ABI.
Bar: slices and right-hand sides bit-exact (their arithmetic has no multiply-add pair), stream functions
fp64 rel-L2 <= 1e-12, box VTK files byte-identical to the unmodified reference's."""
import os

import numpy as np
import pytest

from oracle import fdm_oracle as O
from tests.test_velocity_plot_cpu import CASES, PSI, SLICES, fields, parse_vtk_vectors

pytestmark = pytest.mark.gpu
TOL = 1e-12
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fb():
    import fdm_b200
    assert fdm_b200.lib().fdmb_device_count() > 0, "GPU tests need a CUDA device"
    return fdm_b200


def check_slices(P, want):
    for s in SLICES:
        assert np.array_equal(P.slice(s).ravel(), np.ravel(want(s))), s
    for s in PSI:
        assert O.rel_l2(P.slice(s).ravel(), np.ravel(want(s))) < TOL, s


@pytest.mark.parametrize("name", sorted(CASES))
def test_update_vs_numpy_restatement_and_golden(fb, name):
    args, kw = CASES[name]
    u, v, w = fields(name)
    P = fb.VelocityPlotter(*args, **kw)
    P.use(u, v, w)
    P.update()
    want = O.VelocityPlotter(*args, **kw).update(u, v, w)
    check_slices(P, lambda s: want[s])
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_vplot_v1.npz"))
    check_slices(P, lambda s: g[f"{name}/{s}"])
    assert np.array_equal(P.cell_velocity(), O.VelocityPlotter(*args, **kw).cell_velocity(u, v, w))


@pytest.mark.parametrize("name", sorted(CASES))
def test_update_and_vtk_vs_compiled_reference(fb, ref, name, tmp_path):
    args, kw = CASES[name]
    u, v, w = fields(name, seed=11)
    R = ref.VelocityPlotter(*args, **kw)
    R.update(u, v, w)
    P = fb.VelocityPlotter(*args, **kw)
    P.use(u, v, w)
    P.update()
    check_slices(P, R.slice)
    if kw.get("cyl") and not kw.get("zperiodic"):
        return
    R.vtk_out(tmp_path / "ref.vtk", 40)
    P.vtk_out(tmp_path / "gpu.vtk", 40)
    a, b = (tmp_path / "ref.vtk").read_bytes(), (tmp_path / "gpu.vtk").read_bytes()
    if not kw.get("cyl"):
        assert a == b                 # exact face averages, same printf
    else:
        # the cylinder branch rotates and normalises on the host (libm sin/cos/sqrt, contraction differs between
        # builds): identical structure, vectors equal to the six printed decimals
        la, lb = a.decode().splitlines(), b.decode().splitlines()
        k = la.index("VECTORS u double")
        assert la[:k + 1] == lb[:k + 1]
        va = np.array([[float(x) for x in ln.split()] for ln in la[k + 1:]])
        vb = np.array([[float(x) for x in ln.split()] for ln in lb[k + 1:]])
        assert va.shape == vb.shape and np.max(np.abs(va - vb)) < 1.5e-6


def test_update_before_use_is_an_error(fb):
    args, kw = CASES["box31"]
    P = fb.VelocityPlotter(*args, **kw)
    with pytest.raises(fb.FdmB200Error):
        P.update()


def test_invalid_sizes_and_flags(fb):
    # the reference aborts in FFT (src/fft.cpp:67) for ny+1 != 2^k; periodic y without periodic z is not instantiated
    with pytest.raises(fb.FdmB200Error):
        fb.VelocityPlotter(0.1, 0.1, 0.1, 31, 30, 31, 0, 3.1, 0, 3.0, 0, 3.1)
    with pytest.raises(fb.FdmB200Error):
        fb.VelocityPlotter(0.1, 0.1, 0.1, 31, 32, 31, 0, 3.1, 0, 3.2, 0, 3.1, yperiodic=True)


def test_plotter_reads_ns_cube_state_on_the_device(fb, ref, tmp_path):
    """test/test_ns_cube.cpp:24-50 with the NS state left in HBM: the slices equal the reference plotter's on the same
    fields, and the VTK file equals the one the reference writes from them."""
    n = 31
    ns = fb.NSCube(nx=n, nz=n, Re=250.0, dt=0.01)
    P = fb.VelocityPlotter.for_ns_cube(ns)
    for steps in (0, 10, 10):
        ns.step(steps)
        P.update()
        u, v, w = (ns.field(f) for f in "uvw")
        p = ns.params
        d = (p.x2 - p.x1) / n
        R = ref.VelocityPlotter(d, d, d, n, n, n, p.x1, p.x2, p.y1, p.y2, p.z1, p.z2)
        R.update(u, v, w)
        check_slices(P, R.slice)
        R.vtk_out(tmp_path / "ref.vtk", ns.time_index)
        P.vtk_out(tmp_path / "gpu.vtk", ns.time_index)
        assert (tmp_path / "ref.vtk").read_bytes() == (tmp_path / "gpu.vtk").read_bytes()


@pytest.mark.parametrize("zperiodic", [False, True])
def test_plotter_reads_ns_cyl_state_on_the_device(fb, ref, zperiodic):
    """test/test_ns_cyl.cpp:53-92: (r, z, phi) plotter with the cylindrical column scales."""
    import math
    nr, nz, nphi = 32, (32 if zperiodic else 31), 32
    ns = fb.NSCyl(nr=nr, nz=nz, nphi=nphi, Re=200.0, dt=0.01, vrandom=1, zperiodic=zperiodic)
    ns.step(20)
    P = fb.VelocityPlotter.for_ns_cyl(ns)
    P.update()
    p = ns.params
    R = ref.VelocityPlotter((p.R - p.r) / nr, (p.h2 - p.h1) / nz, 2 * math.pi / nphi, nr, nz, nphi, p.r, p.R, p.h1, p.h2,
                            0.0, 2 * math.pi, cyl=True, zperiodic=True, yperiodic=zperiodic)
    R.update(*(ns.field(f) for f in "uvw"))
    check_slices(P, R.slice)
    assert np.array_equal(P.cell_velocity(),
                          O.VelocityPlotter((p.R - p.r) / nr, (p.h2 - p.h1) / nz, 2 * math.pi / nphi, nr, nz, nphi, p.r,
                                            p.R, p.h1, p.h2, 0.0, 2 * math.pi, cyl=True, zperiodic=True,
                                            yperiodic=zperiodic).cell_velocity(*(ns.field(f) for f in "uvw")))


def test_full_size_cavity_properties(fb):
    """BASELINE configs[2] size (255^3): size-independent properties instead of a CPU run -- the stream-function
    right-hand sides are linear in the velocity (doubling u,v,w doubles RHS and psi), and a state at rest gives 0."""
    n = 255
    ns = fb.NSCube(nx=n, nz=n, Re=1000.0, dt=0.005)
    ns.step(5)
    P = fb.VelocityPlotter.for_ns_cube(ns)
    P.update()
    psi1 = {s: P.slice(s) for s in PSI + ("RHS_x", "RHS_y", "RHS_z")}
    assert max(np.abs(psi1[s]).max() for s in PSI) > 0
    for f in "uvw":
        ns.set_field(f, 2.0 * ns.field(f))
    P.update()
    for s, a in psi1.items():
        b = P.slice(s)
        if s.startswith("RHS"):
            assert np.array_equal(b, 2.0 * a), s       # scaling by 2 is exact in binary floating point
        else:
            assert O.rel_l2(b, 2.0 * a) < TOL, s
    for f in "uvw":
        ns.set_field(f, np.zeros(ns.field_size(f)))
    P.update()
    assert all(np.abs(P.slice(s)).max() == 0 for s in PSI)
