"""The numpy restatement of NSCyl (oracle/fdm_oracle.py: init_bound, FGH, L_FGH, poisson, update_uvwp following
src/ns_cyl.cpp line by line) pinned against the compiled, unmodified reference and the committed golden vectors."""
import numpy as np
import pytest

from oracle import fdm_oracle as O

TOL = 1e-12


def perturb(ns_list, nr, nz, nphi, zperiodic, seed, amp=1e-2):
    """Same interior perturbation of u, v, w on every solver (wall entries untouched: the reference asserts them)."""
    rng = np.random.default_rng(seed)
    zn, z1 = (nz - 1, 0) if zperiodic else (nz, 1)
    shapes = {"u": (nphi, (nz if zperiodic else nz + 2), nr + 3), "v": (nphi, (nz if zperiodic else nz + 3), nr + 2),
              "w": (nphi, (nz if zperiodic else nz + 2), nr + 2)}
    lows = {"u": (0, -1), "v": (0 if zperiodic else -1, 0), "w": (0, 0)}
    for f in "uvw":
        a = np.array(ns_list[0].field(f)).reshape(shapes[f])
        lz, lr = lows[f]
        kmax = zn if f != "v" else (zn if zperiodic else nz - 1)
        jmax = nr - 1 if f == "u" else nr
        a[:, z1 - lz:kmax - lz + 1, 1 - lr:jmax - lr + 1] += amp * rng.uniform(-1, 1, (nphi, kmax - z1 + 1, jmax))
        for ns in ns_list:
            ns.set_field(f, a)


def test_field_extents():
    # SURVEY appendix C / ns_cyl.h:80-93
    nr, nz, nphi = 8, 7, 8
    P = O.NSCyl(nr=nr, nz=nz, nphi=nphi)
    f = P.fields()
    assert f["u"].shape == (nphi, nz + 2, nr + 3) and f["v"].shape == (nphi, nz + 3, nr + 2)
    assert f["w"].shape == f["p"].shape == (nphi, nz + 2, nr + 2)
    assert f["x"].shape == f["RHS"].shape == f["H"].shape == (nphi, nz, nr)
    assert f["F"].shape == (nphi, nz, nr + 1) and f["G"].shape == (nphi, nz + 1, nr)
    Q = O.NSCyl(nr=nr, nz=8, nphi=nphi, zperiodic=True)
    assert all(a.shape[1] == 8 for a in Q.fields().values())       # ns_cyl.h:70-74: every field z 0..nz-1


def test_restatement_vs_golden(golden):
    # tests/golden/make_golden.py: NSCyl 16 x 15 x 16, Re = 200, states of the compiled reference after 1 and 5 steps
    P = O.NSCyl(nr=16, nz=15, nphi=16, Re=200.0, dt=0.01)
    P.step(1)
    for f in "uvwp":
        assert O.rel_l2(P.field(f), golden[f"nscyl_s1_{f}"]) < TOL, f
    P.step(4)
    for f in "uvwp":
        assert O.rel_l2(P.field(f), golden[f"nscyl_s5_{f}"]) < TOL, f


@pytest.mark.parametrize("zp,nr,nz,nphi", [(False, 16, 15, 16), (True, 24, 16, 16), (False, 32, 31, 32)])
def test_restatement_vs_compiled_reference(ref, zp, nr, nz, nphi):
    kw = dict(nr=nr, nz=nz, nphi=nphi, Re=200.0, dt=0.01)
    R, P = ref.NSCyl(zperiodic=zp, **kw), O.NSCyl(zperiodic=zp, **kw)
    R.step(3)
    P.step(3)
    perturb([R, P], nr, nz, nphi, zp, seed=4)         # a fully three-dimensional state: every term of FGH is live
    R.step(1)
    P.step(1)
    for f in ("u", "v", "w", "p", "F", "G", "H", "RHS", "x"):
        assert O.rel_l2(P.field(f), R.field(f)) < TOL, f
    R.step(9)
    P.step(9)
    for f in "uvwp":
        assert O.rel_l2(P.field(f), R.field(f)) < TOL, f
    # L_step: linearise about the current state (test/test_ns_cyl_spectral.cpp:47-58), perturb, step
    for f in "uvw":
        R.set_field(f + "0", R.field(f))
        P.set_field(f + "0", P.field(f))
    perturb([R, P], nr, nz, nphi, zp, seed=5, amp=1e-3)
    R.step(1, linear=True)
    P.step(1, linear=True)
    for f in ("u", "v", "w", "p", "F", "G", "H"):
        assert O.rel_l2(P.field(f), R.field(f)) < TOL, f
    R.step(5, linear=True)
    P.step(5, linear=True)
    for f in "uvwp":
        assert O.rel_l2(P.field(f), R.field(f)) < TOL, f
