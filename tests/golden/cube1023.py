"""Inputs of the full-size LaplCube 1023^3 known-answer tests (BASELINE configs[4]).

Shared by tests/golden/make_golden_cube1023.py (which runs the unmodified reference once and stores a sample of
its answer), the GPU tests and bench.py's in-line check.  Nothing here touches the oracle or the product.
"""
import math

import numpy as np

N = 1023
SEED = 20261018
STRIDE = 31          # sample = ans[::31, ::31, ::31] (33^3 values)


def geometry(n=N):
    """Unit-cube convention of ut/ut_lapl_cube.cpp:56-61: (d, l)."""
    d = 1.0 / n
    return d, 1.0 + d


def rhs_planes(n, z0, nz, seed=SEED):
    """Planes [z0, z0+nz) of the synthetic right-hand side, uniform(-0.5, 0.5); plane z is drawn from its own
    counter-based stream Philox(key=(seed, z)), so every slab decomposition sees identical data."""
    out = np.empty((nz, n, n), dtype=np.float64)
    for i in range(nz):
        rng = np.random.Generator(np.random.Philox(key=[seed, z0 + i]))
        rng.random(out=out[i])
    out -= 0.5
    return out


# the eigenvector known answer lives in the package (bench.py prints it as its in-line check)
from fdm_b200.selfcheck import KAT_MODES, kat_device, kat_factors, kat_modes, rel_l2_device  # noqa: E402,F401
