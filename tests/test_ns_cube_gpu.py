"""GPU parity tests for NSCube through the C ABI.  The reference has no NSCube tests
(SURVEY 4), so parity is pinned by the compiled reference (oracle/_ref) and the oracle
restatement.  Bar: relative L2 <= 1e-12 on the concatenated state and per field (fields
that are exactly zero in the reference must be exactly zero here)."""
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


def compare(ns, fields_ref, names="uvwp", tol=TOL):
    cat_a, cat_b = [], []
    for f in names:
        a = ns.field(f); b = np.asarray(fields_ref[f]).ravel()
        assert a.size == b.size, f
        nb = np.linalg.norm(b)
        if nb > 0 and nb > 1e-3 * max(np.linalg.norm(np.asarray(fields_ref[g]).ravel()) for g in names):
            assert O.rel_l2(a, b) < tol, (f, O.rel_l2(a, b))
        cat_a.append(a); cat_b.append(b)
    err = O.rel_l2(np.concatenate(cat_a), np.concatenate(cat_b))
    assert err < tol, err
    return err


def test_field_sizes(fb):
    n = 7
    ns = fb.NSCube(nx=n, nz=n, Re=10.0, dt=0.01)
    # SURVEY appendix C / ns_cube.h:66-75
    assert ns.field_size("u") == (n + 2) ** 2 * (n + 3)
    assert ns.field_size("p") == (n + 2) ** 3
    assert ns.field_size("x") == n ** 3
    assert ns.field_size("F") == n * n * (n + 1)
    assert np.all(ns.field("u") == 0.0)


@pytest.mark.parametrize("n,steps", [(7, 5), (15, 10), (31, 20)])
def test_ns_vs_oracle(fb, n, steps):
    kw = dict(nx=n, nz=n, Re=250.0, dt=0.01)
    ns = fb.NSCube(**kw); po = O.NSCube(**kw)
    ns.step(steps)
    for _ in range(steps):
        po.step()
    compare(ns, po.fields(), "uvwpxFGH")
    compare(ns, po.fields(), ["RHS"])
    assert ns.time_index == steps


def test_ns_first_step_zero_fields(fb):
    # after step 1 the uniform lid gives zero divergence: v, w, p stay exactly 0
    ns = fb.NSCube(nx=15, nz=15, Re=250.0, dt=0.01)
    ns.step(1)
    for f in "vwp":
        assert np.all(ns.field(f) == 0.0), f
    assert np.any(ns.field("u") != 0.0)


@pytest.mark.parametrize("steps", [1, 10, 100])
def test_ns31_vs_compiled_reference(fb, ref, steps):
    # BASELINE configs[0] (nx=32 aborts in the reference; 31 is the runnable size)
    kw = dict(nx=31, nz=31, Re=250.0, dt=0.01)
    ns = fb.NSCube(**kw); r = ref.NSCube(**kw)
    ns.step(steps); r.step(steps)
    compare(ns, {f: r.field(f) for f in "uvwp"})


def test_ns63_re1000_vs_compiled_reference(fb, ref):
    kw = dict(nx=63, nz=63, Re=1000.0, dt=0.005)
    ns = fb.NSCube(**kw); r = ref.NSCube(**kw)
    ns.step(20); r.step(20)
    compare(ns, {f: r.field(f) for f in "uvwp"})


def test_ns_set_field_roundtrip_and_restart(fb, ref):
    """State exported from the reference, imported here, stepped on both sides."""
    kw = dict(nx=15, nz=15, Re=100.0, dt=0.01)
    r = ref.NSCube(**kw); r.step(7)
    ns = fb.NSCube(**kw)
    for f in "uvwp":
        ns.set_field(f, r.field(f))
        assert np.array_equal(ns.field(f), r.field(f))
    ns.step(5); r.step(5)
    compare(ns, {f: r.field(f) for f in "uvwp"})


@pytest.mark.parametrize("nx,nz", [(15, 7), (7, 15)])
def test_ns_anisotropic_box(fb, ref, nx, nz):
    """nz != nx: the reference's lid loop bound (j = -1..nz+1, ns_cube.cpp:68) stops short of / runs past the x range;
    k_bound_all's substitution of the lid values into the u mirror follows it.  Against the restatement and, where
    the reference stays inside its arrays (nz <= nx: with nz > nx its lid loop writes past the rows of u), ghosts
    included against the compiled reference on every field."""
    kw = dict(nx=nx, nz=nz, Re=50.0, dt=0.005, x1=0.0, x2=1.0, y1=0.0, y2=2.0, z1=-1.0, z2=0.5, u0=0.7)
    ns = fb.NSCube(**kw); po = O.NSCube(**kw)
    ns.step(6)
    for _ in range(6):
        po.step()
    compare(ns, po.fields())
    if nz <= nx:
        R = ref.NSCube(**kw)
        R.step(6)
        names = ("u", "v", "w", "p", "x", "F", "G", "H", "RHS")
        compare(ns, {f: R.field(f) for f in names}, names)


def test_ns255_one_step_properties(fb):
    """BASELINE configs[2] size: after one step from rest, u equals the closed-form predictor
    (only the lid ghost row is non-zero), and divergence-free fields stay exactly zero."""
    n = 255
    ns = fb.NSCube(nx=n, nz=n, Re=1000.0, dt=0.005)
    ns.step(1)
    u = ns.field("u").reshape(n + 2, n + 2, n + 3)
    dz = 2 * np.pi / n
    expect = 0.005 * (2.0 / 1000.0 / dz / dz)     # F = dt * nu * (u[nz+1] - 0)/dz2 with u[nz+1] = 2 U0
    top = u[n, 1:n + 1, 2:n + 1]                  # i = nz, k = 1..ny, j = 1..nx-1
    assert np.allclose(top, expect, rtol=1e-13, atol=0)
    assert np.all(u[1:n, :, :] == 0.0)
    for f in "vwp":
        assert np.all(ns.field(f) == 0.0)


def test_ns_errors(fb):
    with pytest.raises(fb.FdmB200Error):
        fb.NSCube(nx=32, nz=32)    # the README size: reference aborts in FFTTable (fft.cpp:67)


@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("n,steps", [(15, 5), (31, 20), (63, 6), (127, 3)])
def test_fused_fgh_rhs_matches(fb, ref, monkeypatch, n, steps, fused):
    """Both FGH + divergence paths against the compiled reference on all nine public fields (F, G, H, RHS included):
    the default register-marching single sweep (k_fgh_div, ns_cube.cu) and the two-kernel fallback FDMB_FGH_FUSED=0
    (k_fgh + k_rhs, the reference's own order of evaluation)."""
    monkeypatch.setenv("FDMB_FGH_FUSED", fused)
    ns = fb.NSCube(nx=n, nz=n, Re=400.0, dt=0.005)
    R = ref.NSCube(nx=n, nz=n, Re=400.0, dt=0.005)
    ns.step(steps); R.step(steps)
    names = ("u", "v", "w", "p", "x", "F", "G", "H", "RHS")
    compare(ns, {f: R.field(f) for f in names}, names)


@pytest.mark.parametrize("steps", [1, 10])
def test_ns255_multi_step_vs_compiled_reference(fb, ref, steps):
    """BASELINE configs[2] (NSCube 255^3, Re = 1000, dt = 0.005) after 1 and 10 steps against the unmodified reference
    (SURVEY 8d C3); per field and on the concatenated state, guarded against the exactly-zero fields of step 1."""
    n = 255
    ns = fb.NSCube(nx=n, nz=n, Re=1000.0, dt=0.005)
    R = ref.NSCube(nx=n, nz=n, Re=1000.0, dt=0.005)
    ns.step(steps); R.step(steps)
    compare(ns, {f: R.field(f) for f in "uvwp"})
    if steps == 1:      # uniform lid, zero divergence: v, w, p stay exactly zero
        for f in "vwp":
            assert not np.any(R.field(f)) and not np.any(ns.field(f)), f
