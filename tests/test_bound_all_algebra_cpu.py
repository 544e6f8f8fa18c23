"""k_bound_all (fdm_b200/csrc/ns_cube.cu) fills all ghosts of init_bound (src/ns_cube.cpp:65-122) in ONE launch: its
seven roles run concurrently, so every value an ordered fill would have read from an earlier fill is formed by the
reading thread itself (lid values inside the u mirror, mirrored ghosts inside the pressure ghosts).  This test restates
that substitution in numpy -- every role reads ONLY the state from before the launch -- and requires the result to be
bit-identical to the ordered fills of the reference restatement, for cubes and for nz != nx both ways (the lid loop's
j = -1..nz+1 bound stops short of / runs past the x range)."""
import copy

import numpy as np
import pytest

from oracle import fdm_oracle as O


def bound_all(ns):
    """One-launch init_bound: returns new (u, v, w, p) arrays computed from the OLD state only."""
    nx, ny, nz, U0, Re = ns.nx, ns.ny, ns.nz, ns.U0, ns.Re
    old = {n: copy.deepcopy(getattr(ns, n)) for n in "uvwp"}          # what every role reads
    new = {n: copy.deepcopy(getattr(ns, n)) for n in "uvwp"}          # what the roles write
    uo, vo, wo, po = old["u"], old["v"], old["w"], old["p"]
    u, v, w, p = new["u"], new["v"], new["w"], new["p"]
    jmax = min(nz + 1, nx + 1)
    K2 = (0, ny + 1)
    # role 0, lid: j = 0 .. min(jmax, nx)   (j = -1 and nx+1 of that plane belong to the u mirror)
    jl = min(jmax, nx)
    u.v(nz + 1, K2, (0, jl))[...] = 2 * U0 - uo.v(nz, K2, (0, jl))
    # role 1, u mirror on planes 0..nz+1; on the lid plane the operands are the lid values where the lid loop reaches
    for i in range(0, nz + 2):
        lid = i == nz + 1
        a = (2 * U0 - uo.v(nz, K2, 1)) if (lid and 1 <= jmax) else uo.v(i, K2, 1)
        b = (2 * U0 - uo.v(nz, K2, nx - 1)) if (lid and nx - 1 <= jmax) else uo.v(i, K2, nx - 1)
        u.v(i, K2, -1)[...] = a
        u.v(i, K2, nx + 1)[...] = b
    # roles 2, 3: v and w mirrors
    I2, J2 = (0, nz + 1), (0, nx + 1)
    v.v(I2, -1, J2)[...] = vo.v(I2, 1, J2)
    v.v(I2, ny + 1, J2)[...] = vo.v(I2, ny - 1, J2)
    w.v(-1, K2, J2)[...] = wo.v(1, K2, J2)
    w.v(nz + 1, K2, J2)[...] = wo.v(nz - 1, K2, J2)
    # roles 4..6: pressure ghosts with the mirrored ghosts substituted (u[-1] = u[1], u[nx+1] = u[nx-1], ...)
    I, K, J = (1, nz), (1, ny), (1, nx)
    dx, dy, dz = ns.dx, ns.dy, ns.dz
    p.v(I, K, 0)[...] = po.v(I, K, 1) - (uo.v(I, K, 1) - 2 * uo.v(I, K, 0) + uo.v(I, K, 1)) / Re / dx
    p.v(I, K, nx + 1)[...] = po.v(I, K, nx) - (uo.v(I, K, nx - 1) - 2 * uo.v(I, K, nx) + uo.v(I, K, nx - 1)) / Re / dx
    p.v(I, 0, J)[...] = po.v(I, 1, J) - (vo.v(I, 1, J) - 2 * vo.v(I, 0, J) + vo.v(I, 1, J)) / Re / dy
    p.v(I, ny + 1, J)[...] = po.v(I, ny, J) - (vo.v(I, ny - 1, J) - 2 * vo.v(I, ny, J) + vo.v(I, ny - 1, J)) / Re / dy
    p.v(0, K, J)[...] = po.v(1, K, J) - (wo.v(1, K, J) - 2 * wo.v(0, K, J) + wo.v(1, K, J)) / Re / dz
    p.v(nz + 1, K, J)[...] = po.v(nz, K, J) - (wo.v(nz - 1, K, J) - 2 * wo.v(nz, K, J) + wo.v(nz - 1, K, J)) / Re / dz
    return new


@pytest.mark.parametrize("nx,nz", [(7, 7), (15, 7), (7, 15), (15, 3), (3, 15), (31, 31)])
def test_one_launch_init_bound_equals_ordered_fills(nx, nz):
    ns = O.NSCube(nx=nx, nz=nz, Re=13.0, dt=0.01, u0=0.9, x1=0.0, x2=1.0, y1=0.0, y2=1.5, z1=-1.0, z2=0.25)
    rng = np.random.default_rng(100 * nx + nz)
    for f in (ns.u, ns.v, ns.w, ns.p):
        f.a[...] = rng.standard_normal(f.a.shape)
    merged = bound_all(ns)
    ns.init_bound()                      # the reference's ordered fills
    for n in "uvwp":
        assert np.array_equal(merged[n].a, getattr(ns, n).a), n
