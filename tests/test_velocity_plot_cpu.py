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
