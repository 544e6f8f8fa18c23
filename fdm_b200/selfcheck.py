"""Closed-form self-check of the LaplCube solve: eigenvector known answers (no CPU solve of any kind).

The discrete sine products phi_k(j) = sin(pi k j / (n+1)) are the eigenvectors of the 7-point Dirichlet Laplacian
(eigenvalue -(lx + ly + lz), l = 4/d^2 sin^2(k pi / (2(n+1))), reference src/lapl_cube.cpp:145-172), so for
rhs = sum_i a_i phi_kz(z) phi_ky(y) phi_kx(x) the solve must return sum_i a_i / -(lz + ly + lx) phi phi phi.
Used by bench.py (the "check" object of every cube line, at every GPU count) and by the full-size GPU tests.
"""
import math

import numpy as np

KAT_MODES = [  # (amplitude, kz, ky, kx) in units of "fraction of n" so that any size uses low, middle and high modes
    (1.0, 1, 1, 1),
    (-0.7, 2, 5, 3),
    (0.45, 0.5, 0.25, 0.75),
    (0.3, 1.0, 1.0, 1.0),
    (-0.2, 0.999, 0.002, 0.5),
]


def kat_modes(n):
    out = []
    for a, kz, ky, kx in KAT_MODES:
        ks = [k if isinstance(k, int) else max(1, min(n, int(round(k * n)))) for k in (kz, ky, kx)]
        out.append((a, *ks))
    return out


def kat_factors(n, d, modes=None):
    """Per mode: (amplitude, sz, sy, sx, -1/(lz+ly+lx)) with the 1-D sine vectors over the interior points."""
    j = np.arange(1, n + 1, dtype=np.float64)
    res = []
    for a, kz, ky, kx in (modes or kat_modes(n)):
        s = [np.sin(math.pi * k * j / (n + 1)) for k in (kz, ky, kx)]
        lam = sum(4.0 / (d * d) * math.sin(k * math.pi * 0.5 / (n + 1)) ** 2 for k in (kz, ky, kx))
        res.append((a, s[0], s[1], s[2], -1.0 / lam))
    return res


def kat_device(torch, n, d, z0, nzl, device, modes=None, chunk=32):
    """(rhs, want) for the planes [z0, z0+nzl) as device tensors, built chunk by chunk (no full-size temporaries)."""
    fac = kat_factors(n, d, modes)
    rhs = torch.zeros((nzl, n, n), dtype=torch.float64, device=device)
    want = torch.zeros((nzl, n, n), dtype=torch.float64, device=device)
    planes = []
    for a, sz, sy, sx, inv in fac:
        p = torch.from_numpy(np.outer(sy, sx)).to(device)
        planes.append((a, torch.from_numpy(sz[z0:z0 + nzl].copy()).to(device), p, inv))
    for c0 in range(0, nzl, chunk):
        c1 = min(nzl, c0 + chunk)
        for a, sz, p, inv in planes:
            t = sz[c0:c1, None, None] * p[None]
            rhs[c0:c1].add_(t, alpha=a)
            want[c0:c1].add_(t, alpha=a * inv)
    return rhs, want


def rel_l2_device(torch, got, want):
    return float((got - want).norm() / want.norm())
