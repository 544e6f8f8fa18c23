"""velocity_plotter: the numpy restatement against the compiled, unmodified reference
(src/velocity_plot.cpp) and against the committed golden vectors.  No GPU."""
import os

import numpy as np
import pytest

from oracle import fdm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SLICES = ("vx", "wx", "uy", "wy", "uz", "vz", "RHS_x", "RHS_y", "RHS_z")
PSI = ("psi_x", "psi_y", "psi_z")

# name -> (constructor arguments, keywords); sizes keep every transformed axis a power of two
CASES = {
    "box31": ((0.2, 0.2, 0.2, 31, 31, 31, -3.1, 3.1, -3.1, 3.1, -3.1, 3.1), {}),
    "box_15_31_63": ((0.1, 0.05, 0.025, 15, 31, 63, 0.0, 1.5, 0.0, 1.55, 0.0, 1.575), {}),
    "cyl_dirichlet_z": ((0.05, 0.3, 0.19634954084936207, 32, 31, 32, 1.5, 3.1, 0.0, 9.3, 0.0, 6.283185307179586),
                        dict(cyl=True, zperiodic=True)),
    "cyl_periodic_z": ((0.05, 0.3, 0.19634954084936207, 24, 32, 32, 1.5, 2.7, 0.0, 9.6, 0.0, 6.283185307179586),
                       dict(cyl=True, zperiodic=True, yperiodic=True)),
}


def fields(name, seed=7):
    args, kw = CASES[name]
    shapes = O.VelocityPlotter(*args, **kw).shapes()
    rng = np.random.default_rng(seed)
    return [rng.uniform(-1, 1, s) for s in shapes]


def parse_vtk_vectors(path, ncell):
    lines = open(path).read().splitlines()
    k = lines.index("VECTORS u double")
    return lines[:k + 1], np.array([[float(x) for x in ln.split()] for ln in lines[k + 1:k + 1 + ncell]])


@pytest.mark.parametrize("name", sorted(CASES))
def test_restatement_vs_compiled_reference(ref, name):
    args, kw = CASES[name]
    u, v, w = fields(name)
    R = ref.VelocityPlotter(*args, **kw)
    R.update(u, v, w)
    got = O.VelocityPlotter(*args, **kw).update(u, v, w)
    for s in SLICES:     # no multiply-add pair in the slice arithmetic: bit-exact
        assert np.array_equal(got[s].ravel(), R.slice(s)), s
    for s in PSI:
        assert O.rel_l2(got[s].ravel(), R.slice(s)) < 1e-12, s


@pytest.mark.parametrize("name", ["box31", "box_15_31_63"])
def test_cell_velocity_vs_reference_vtk(ref, name, tmp_path):
    args, kw = CASES[name]
    u, v, w = fields(name)
    R = ref.VelocityPlotter(*args, **kw)
    R.update(u, v, w)
    R.vtk_out(tmp_path / "r.vtk", 3)
    want = O.VelocityPlotter(*args, **kw).cell_velocity(u, v, w)
    head, got = parse_vtk_vectors(tmp_path / "r.vtk", want.shape[0])
    assert head[1] == "step 3" and head[3] == "DATASET STRUCTURED_POINTS"
    assert np.max(np.abs(got - want)) < 1.5e-6       # "%f": six decimals


@pytest.mark.parametrize("name", sorted(CASES))
def test_restatement_vs_golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_vplot_v1.npz"))
    args, kw = CASES[name]
    u, v, w = fields(name)
    got = O.VelocityPlotter(*args, **kw).update(u, v, w)
    for s in SLICES:
        assert np.array_equal(got[s].ravel(), g[f"{name}/{s}"]), s
    for s in PSI:
        assert O.rel_l2(got[s].ravel(), g[f"{name}/{s}"]) < 1e-12, s


# ---- the device arithmetic itself, emulated on the host ---------------------------------------------------
@pytest.fixture(scope="module")
def emul():
    import ctypes as C
    import subprocess
    here = os.path.join(ROOT, "tests", "host_emul")
    src, lib = os.path.join(here, "vplot_host.cpp"), os.path.join(here, "libvplot_host.so")
    hdr = os.path.join(ROOT, "fdm_b200", "csrc", "vplot_math.h")
    if (not os.path.exists(lib)) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        # -ffp-contract=off: what the device code does is irrelevant here, the arithmetic has no multiply-add pair
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", src, "-o", lib], check=True)
    L = C.CDLL(lib)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.emul_vplot_field_elems.restype = C.c_longlong
    L.emul_vplot_field_elems.argtypes = [C.c_int] * 6
    L.emul_vplot_dims.argtypes = [C.c_int] * 5 + [ip]
    L.emul_vplot_update.argtypes = [C.c_int] * 5 + [C.c_double] * 3 + [dp] * 12
    L.emul_vplot_cells.restype = C.c_longlong
    L.emul_vplot_cells.argtypes = [C.c_int] * 5 + [dp] * 4
    return L


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_arithmetic_emulated_on_the_host(emul, name):
    """fdm_b200/csrc/vplot_math.h (the functions the CUDA kernels call) against the restatement, bit for bit."""
    import ctypes as C
    args, kw = CASES[name]
    dx, dy, dz, nx, ny, nz = args[:6]
    zp, yp = int(kw.get("zperiodic", False)), int(kw.get("yperiodic", False))
    u, v, w = fields(name, seed=3)
    Or = O.VelocityPlotter(*args, **kw)
    assert [emul.emul_vplot_field_elems(nx, ny, nz, zp, yp, f) for f in range(3)] == [a.size for a in (u, v, w)]
    dims = (C.c_int * 18)()
    emul.emul_vplot_dims(nx, ny, nz, zp, yp, dims)
    want = Or.update(u, v, w)
    outs = []
    for s, key in enumerate(SLICES):
        assert (dims[2 * s], dims[2 * s + 1]) == want[key].shape, key
        outs.append(np.full(want[key].shape, np.nan))
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))      # noqa: E731
    emul.emul_vplot_update(nx, ny, nz, zp, yp, dx, dy, dz, p(u), p(v), p(w), *[p(a) for a in outs])
    for key, a in zip(SLICES, outs):
        assert np.array_equal(a, want[key]), key
    cells = Or.cell_velocity(u, v, w)
    assert emul.emul_vplot_cells(nx, ny, nz, zp, yp, p(u), p(v), p(w), None) == cells.shape[0]
    got = np.full(cells.shape, np.nan)
    emul.emul_vplot_cells(nx, ny, nz, zp, yp, p(u), p(v), p(w), p(got))
    assert np.array_equal(got, cells)
