"""GPU parity tests for NSCyl through the C ABI.  The reference has no NSCyl tests, so
parity is pinned by the compiled reference (oracle/_ref) and by the golden vectors generated from it
(tests/golden/make_golden.py).  Bar: relative L2 <= 1e-12 on the concatenated state and per field
(fields that are negligible against the state are only checked through the concatenation).

The reference aborts through verify() when the wall invariants of init_bound are violated
(src/ns_cyl.cpp:136-163), so perturbed states only touch entries that init_bound does not pin."""
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


def compare(ns, fields_ref, names=("u", "v", "w", "p"), tol=TOL):
    cat_a, cat_b = [], []
    big = max(np.linalg.norm(np.asarray(fields_ref[g]).ravel()) for g in names)
    for f in names:
        a = ns.field(f); b = np.asarray(fields_ref[f]).ravel()
        assert a.size == b.size, f
        nb = np.linalg.norm(b)
        if nb > 1e-3 * big:
            assert O.rel_l2(a, b) < tol, (f, O.rel_l2(a, b))
        cat_a.append(a); cat_b.append(b)
    err = O.rel_l2(np.concatenate(cat_a), np.concatenate(cat_b))
    assert err < tol, err
    return err


def perturb(ns_list, nr, nz, nphi, zperiodic, seed, amp=1e-2):
    """Same interior perturbation of u, v, w on every solver in ns_list (wall entries untouched)."""
    rng = np.random.default_rng(seed)
    zn = nz - 1 if zperiodic else nz
    z1 = 0 if zperiodic else 1
    shapes = {"u": (nphi, (nz if zperiodic else nz + 2), nr + 3), "v": (nphi, (nz if zperiodic else nz + 3), nr + 2),
              "w": (nphi, (nz if zperiodic else nz + 2), nr + 2)}
    lows = {"u": (0, -1), "v": (0 if zperiodic else -1, 0), "w": (0, 0)}
    for f in "uvw":
        a = ns_list[0].field(f).reshape(shapes[f])
        lz, lr = lows[f]
        kmax = zn if f != "v" else (zn if zperiodic else nz - 1)
        jmax = nr - 1 if f == "u" else nr
        a[:, z1 - lz:kmax - lz + 1, 1 - lr:jmax - lr + 1] += amp * rng.uniform(-1, 1, (nphi, kmax - z1 + 1, jmax))
        for ns in ns_list:
            ns.set_field(f, a)


def test_field_sizes(fb):
    nr, nz, nphi = 8, 7, 8
    ns = fb.NSCyl(nr=nr, nz=nz, nphi=nphi, Re=10.0, dt=0.01)
    # SURVEY appendix C / ns_cyl.h:80-93
    assert ns.field_size("u") == nphi * (nz + 2) * (nr + 3)
    assert ns.field_size("v") == nphi * (nz + 3) * (nr + 2)
    assert ns.field_size("w") == nphi * (nz + 2) * (nr + 2)
    assert ns.field_size("p") == nphi * (nz + 2) * (nr + 2)
    assert ns.field_size("x") == nphi * nz * nr
    assert ns.field_size("F") == nphi * nz * (nr + 1)
    assert ns.field_size("G") == nphi * (nz + 1) * nr
    assert ns.size() == sum(ns.field_size(f) for f in "uvwp")
    nsp = fb.NSCyl(nr=nr, nz=8, nphi=nphi, Re=10.0, dt=0.01, zperiodic=True)
    assert nsp.field_size("u") == nphi * 8 * (nr + 3)
    assert nsp.field_size("v") == nphi * 8 * (nr + 2)
    assert np.all(ns.field("w") == 0.0)


def test_ns_cyl_golden(fb, golden):
    ns = fb.NSCyl(nr=16, nz=15, nphi=16, Re=200.0, dt=0.01)
    done = 0
    for steps in (1, 5):
        ns.step(steps - done); done = steps
        compare(ns, {f: golden[f"nscyl_s{steps}_{f}"] for f in "uvwp"})
    assert ns.time_index == 5


@pytest.mark.parametrize("steps", [1, 10, 100])
def test_ns_cyl_readme_vs_compiled_reference(fb, ref, steps):
    # README Taylor-vortex run (32 x 31 x 32, Re=200, dt=0.01) plus a small 3-D perturbation.  (The
    # reference's own vrandom=1 noise trips its verify() at ns_cyl.cpp:160 when z is Dirichlet.)
    kw = dict(nr=32, nz=31, nphi=32, Re=200.0, dt=0.01)
    ns = fb.NSCyl(**kw); r = ref.NSCyl(False, **kw)
    perturb([ns, r], 32, 31, 32, False, seed=1, amp=1e-3)
    ns.step(steps); r.step(steps)
    compare(ns, {f: r.field(f) for f in "uvwp"})
    compare(ns, {f: r.field(f) for f in ("x", "F", "G", "H", "RHS")}, names=("x", "F", "G", "H", "RHS"))


@pytest.mark.parametrize("zperiodic,nz,nr", [(False, 15, 16), (True, 16, 24)])
def test_ns_cyl_perturbed_state(fb, ref, zperiodic, nz, nr):
    """phi-dependent state: exercises every stencil term, the periodic wraps and (Dirichlet z) the
    v mirror whose r index runs over the z range (ns_cyl.cpp:107-112): with nz+1 = nr the ghost
    v[.][-1][nr+1] is never written and stays 0; a larger nr would trip the reference's own
    verify() at ns_cyl.cpp:162."""
    kw = dict(nr=nr, nz=nz, nphi=16, Re=120.0, dt=0.005, u0=0.8, R=2.5, r=1.0, h1=-1.0, h2=3.0)
    ns = fb.NSCyl(zperiodic=zperiodic, **kw); r = ref.NSCyl(zperiodic, **kw)
    ns.step(3); r.step(3)
    perturb([ns, r], nr, nz, 16, zperiodic, seed=5)
    for f in "uvw":
        assert np.array_equal(ns.field(f), r.field(f))
    for steps in (1, 9):
        ns.step(steps); r.step(steps)
        compare(ns, {f: r.field(f) for f in "uvwp"})


def test_ns_cyl_vrandom_periodic_z(fb, ref):
    # ns_cyl.h:99-108: default-seeded std::default_random_engine noise in v (usable with periodic z)
    kw = dict(nr=16, nz=16, nphi=16, Re=200.0, dt=0.01, vrandom=1)
    ns = fb.NSCyl(zperiodic=True, **kw); r = ref.NSCyl(True, **kw)
    # same engine and draw order; the last bit may differ (the reference build contracts a*b+c to fma)
    assert O.rel_l2(ns.field("v"), r.field("v")) < 1e-15
    ns.step(10); r.step(10)
    compare(ns, {f: r.field(f) for f in "uvwp"})


def test_ns_cyl_tall_grid_mirror_quirk(fb):
    # nz+1 > nr+1: the v mirror loop (ns_cyl.cpp:108) would run past the r extent; both sides stop at nr+1
    kw = dict(nr=8, nz=15, nphi=8, Re=50.0, dt=0.005)
    ns = fb.NSCyl(**kw)
    perturb([ns], 8, 15, 8, False, seed=2, amp=1e-3)
    ns.step(4)
    assert np.all(np.isfinite(ns.field("v")))


@pytest.mark.parametrize("zperiodic,nz", [(False, 15), (True, 16)])
def test_ns_cyl_linearised_step(fb, ref, zperiodic, nz):
    # L_step (ns_cyl.cpp:66-78, 280-405): base flow u0,v0,w0 = a developed state, perturbation stepped linearly
    kw = dict(nr=16, nz=nz, nphi=16, Re=150.0, dt=0.01)
    base = ref.NSCyl(zperiodic, **kw)
    perturb([base], 16, nz, 16, zperiodic, seed=3, amp=1e-3)
    base.step(20)
    ns2 = fb.NSCyl(zperiodic=zperiodic, u0=0.0, **kw); r2 = ref.NSCyl(zperiodic, u0=0.0, **kw)
    for f in "uvw":
        ns2.set_field(f + "0", base.field(f)); r2.set_field(f + "0", base.field(f))
    perturb([ns2, r2], 16, nz, 16, zperiodic, seed=8, amp=1e-3)
    ns2.L_step(7); r2.step(7, linear=True)
    compare(ns2, {f: r2.field(f) for f in "uvwp"})
    assert ns2.time_index == 7


def test_ns_cyl_config4_two_steps(fb, ref):
    # BASELINE configs[3]: nr=128, nz=127, nphi=128, Re=200
    kw = dict(nr=128, nz=127, nphi=128, Re=200.0, dt=0.01)
    ns = fb.NSCyl(**kw); r = ref.NSCyl(False, **kw)
    perturb([ns, r], 128, 127, 128, False, seed=4, amp=1e-3)
    ns.step(2); r.step(2)
    compare(ns, {f: r.field(f) for f in "uvwp"})


def test_ns_cyl_config4_1_10_100_steps(fb, ref):
    """BASELINE configs[3] (nr=128, nz=127, nphi=128, Re=200, dt=0.01, vrandom=0: SURVEY 8d C4) from the reference's own
    initial state, compared after 1, 10 and 100 steps with the unmodified reference (per field where the field is not
    negligible, and on the concatenated state)."""
    kw = dict(nr=128, nz=127, nphi=128, Re=200.0, dt=0.01)
    ns = fb.NSCyl(**kw); r = ref.NSCyl(False, **kw)
    done = 0
    for steps in (1, 10, 100):
        ns.step(steps - done); r.step(steps - done); done = steps
        compare(ns, {f: r.field(f) for f in "uvwp"})


def test_ns_cyl_errors(fb):
    with pytest.raises(fb.FdmB200Error):
        fb.NSCyl(nr=32, nz=32, nphi=32)      # Dirichlet z needs nz+1 = 2^k (reference aborts, src/fft.cpp:67)
    with pytest.raises(fb.FdmB200Error):
        fb.NSCyl(nr=32, nz=31, nphi=24)
