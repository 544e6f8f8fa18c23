"""The arithmetic of the fused FGH + divergence sweep (fdm_b200/csrc/ns_cube.cu: k_fgh_div, fgh_combine) restated in
numpy and pinned against the restatement of the reference (oracle.fdm_oracle.NSCube.FGH / poisson, which follow
src/ns_cube.cpp:126-238 term by term, themselves pinned against the compiled reference):

* dt folded into the coefficients and the eleven shared products of face sums give F, G, H within a few ulps of the
  reference's order of evaluation;
* the k-shifted tap list with which a thread evaluates G[k-1] itself is the G stencil at (i, k-1, j), exactly;
* the divergence formed from F[j-1] (the lane to the left), that G[k-1] and H[i-1] (the previous plane) plus the
  ghost-pressure corrections is the reference's right-hand side.

The GPU parity tests (tests/test_ns_cube_gpu.py) compare the kernel itself with the compiled reference; this test keeps
the algebra checkable without a GPU."""
import numpy as np
import pytest

from oracle import fdm_oracle as O


def _state(nx, nz, seed):
    ns = O.NSCube(nx=nx, nz=nz, Re=37.0, dt=0.004, u0=0.8, x1=0.0, x2=1.0, y1=-0.5, y2=0.7, z1=0.2, z2=1.9)
    rng = np.random.default_rng(seed)
    for f in (ns.u, ns.v, ns.w, ns.p):
        f.a[...] = rng.standard_normal(f.a.shape)
    ns.init_bound()
    return ns


def _combine(c, sx, sy, sz, a, b, qa, d1, q1, d2, q2, k):
    """fgh_combine: nine multiply-adds (the device fuses them; numpy rounds twice, a fraction of an ulp each)."""
    acc = k["dcc"] * c + c
    acc = k["dcx"] * sx + acc
    acc = k["dcy"] * sy + acc
    acc = k["dcz"] * sz + acc
    t = a * a
    t = -b * b + t
    acc = -qa * t + acc
    acc = -q1 * d1 + acc
    return -q2 * d2 + acc


@pytest.mark.parametrize("nx,nz,seed", [(7, 7, 1), (15, 7, 2), (7, 15, 3), (15, 15, 4)])
def test_fgh_div_algebra(nx, nz, seed):
    ns = _state(nx, nz, seed)
    ny = ns.ny
    ns.FGH()
    dt, Re = ns.dt, ns.Re
    k = dict(dcx=dt * (1.0 / Re / ns.dx2), dcy=dt * (1.0 / Re / ns.dy2), dcz=dt * (1.0 / Re / ns.dz2))
    k["dcc"] = -2.0 * (k["dcx"] + k["dcy"] + k["dcz"])
    qx, qy, qz = 0.25 * dt / ns.dx, 0.25 * dt / ns.dy, 0.25 * dt / ns.dz

    def taps(I, K, J):
        """tap(field, di, dk, dj) over the inclusive index box I x K x J"""
        def tap(t, di=0, dk=0, dj=0):
            return t.v((I[0] + di, I[1] + di), (K[0] + dk, K[1] + dk), (J[0] + dj, J[1] + dj))
        return tap

    def stencils(tap, want):
        u, v, w = ns.u, ns.v, ns.w
        u000, v000, w000 = tap(u), tap(v), tap(w)
        out = {}
        if "F" in want or "G" in want:
            Puv1 = (u000 + tap(u, dk=1)) * (tap(v, dj=1) + v000)
        if "F" in want or "H" in want:
            Puw1 = (u000 + tap(u, di=1)) * (tap(w, dj=1) + w000)
        if "G" in want or "H" in want:
            Pwv1 = (w000 + tap(w, dk=1)) * (v000 + tap(v, di=1))
        if "F" in want:
            Puv2 = (tap(u, dk=-1) + u000) * (tap(v, dk=-1, dj=1) + tap(v, dk=-1))
            Puw2 = (tap(u, di=-1) + u000) * (tap(w, di=-1, dj=1) + tap(w, di=-1))
            out["F"] = _combine(u000, tap(u, dj=1) + tap(u, dj=-1), tap(u, dk=1) + tap(u, dk=-1),
                                tap(u, di=1) + tap(u, di=-1), u000 + tap(u, dj=1), tap(u, dj=-1) + u000, qx,
                                Puv1 - Puv2, qy, Puw1 - Puw2, qz, k)
        if "G" in want:
            Pg2 = (tap(u, dj=-1) + tap(u, dk=1, dj=-1)) * (v000 + tap(v, dj=-1))
            Pg4 = (tap(w, di=-1) + tap(w, di=-1, dk=1)) * (tap(v, di=-1) + v000)
            out["G"] = _combine(v000, tap(v, dj=1) + tap(v, dj=-1), tap(v, dk=1) + tap(v, dk=-1),
                                tap(v, di=1) + tap(v, di=-1), v000 + tap(v, dk=1), tap(v, dk=-1) + v000, qy,
                                Puv1 - Pg2, qx, Pwv1 - Pg4, qz, k)
        if "H" in want:
            Ph2 = (tap(u, di=1, dj=-1) + tap(u, dj=-1)) * (w000 + tap(w, dj=-1))
            Pwv2 = (tap(w, dk=-1) + w000) * (tap(v, dk=-1) + tap(v, di=1, dk=-1))
            out["H"] = _combine(w000, tap(w, dj=1) + tap(w, dj=-1), tap(w, dk=1) + tap(w, dk=-1),
                                tap(w, di=1) + tap(w, di=-1), tap(w, di=1) + w000, tap(w, di=-1) + w000, qz,
                                Puw1 - Ph2, qx, Pwv1 - Pwv2, qy, k)
        return out

    F = stencils(taps((1, nz), (1, ny), (0, nx)), "F")["F"]
    G = stencils(taps((1, nz), (0, ny), (1, nx)), "G")["G"]
    H = stencils(taps((0, nz), (1, ny), (1, nx)), "H")["H"]
    for name, mine, ref in (("F", F, ns.F.a), ("G", G, ns.G.a), ("H", H, ns.H.a)):
        assert O.rel_l2(mine, ref) < 5e-15, (name, O.rel_l2(mine, ref))

    # G[k-1] as the thread evaluates it: the G stencil with every tap shifted by dk = -1 (k_fgh_div's second g-stencil:
    # v0m0, v0mp, v0mm, v000, v0M0, vpm0, vmm0 | u0m0, u000, u0mm, u00m | w0m0, w000, wmm0, wm00)
    core = taps((1, nz), (1, ny), (1, nx))
    shifted = lambda t, di=0, dk=0, dj=0: core(t, di, dk - 1, dj)      # noqa: E731
    gm = stencils(shifted, "G")["G"]
    assert np.array_equal(gm, G[:, 0:ny, :])            # rows k-1 = 0..ny-1 of G, bit for bit
    # the divergence and the ghost pressures (ns_cube.cpp:204-233) from the thread's own values
    R = ((F[:, :, 1:] - F[:, :, :-1]) * (1.0 / ns.dx) + (G[:, 1:, :] - gm) * (1.0 / ns.dy) +
         (H[1:, :, :] - H[:-1, :, :]) * (1.0 / ns.dz)) * (1.0 / dt)
    p = ns.p
    I, K, J = (1, nz), (1, ny), (1, nx)
    R[0, :, :] -= p.v(0, K, J)[0] * (1.0 / ns.dz2)
    R[:, 0, :] -= p.v(I, 0, J)[:, 0, :] * (1.0 / ns.dy2)
    R[:, :, 0] -= p.v(I, K, 0)[:, :, 0] * (1.0 / ns.dx2)
    R[:, :, -1] -= p.v(I, K, nx + 1)[:, :, 0] * (1.0 / ns.dx2)
    R[:, -1, :] -= p.v(I, ny + 1, J)[:, 0, :] * (1.0 / ns.dy2)
    R[-1, :, :] -= p.v(nz + 1, K, J)[0] * (1.0 / ns.dz2)
    # the reference's right-hand side from ITS F, G, H
    Fo, Go, Ho = ns.F, ns.G, ns.H
    Ro = ((Fo.v(I, K, J) - Fo.v(I, K, (0, nx - 1))) / ns.dx + (Go.v(I, K, J) - Go.v(I, (0, ny - 1), J)) / ns.dy +
          (Ho.v(I, K, J) - Ho.v((0, nz - 1), K, J)) / ns.dz) / dt
    Ro[0, :, :] -= p.v(0, K, J)[0] / ns.dz2
    Ro[:, 0, :] -= p.v(I, 0, J)[:, 0, :] / ns.dy2
    Ro[:, :, 0] -= p.v(I, K, 0)[:, :, 0] / ns.dx2
    Ro[:, :, -1] -= p.v(I, K, nx + 1)[:, :, 0] / ns.dx2
    Ro[:, -1, :] -= p.v(I, ny + 1, J)[:, 0, :] / ns.dy2
    Ro[-1, :, :] -= p.v(nz + 1, K, J)[0] / ns.dz2
    assert O.rel_l2(R, Ro) < 1e-12, O.rel_l2(R, Ro)
