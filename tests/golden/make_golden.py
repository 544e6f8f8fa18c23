"""Generate tests/golden/golden_v1.npz from the UNMODIFIED reference.

Run in the dev container (needs /root/reference):  python tests/golden/make_golden.py
Inputs come from a fixed-seed Philox generator; outputs are whatever
oracle/_ref/libfdm_ref.so (the reference's own translation units) returns.
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402


def rnd(shape, seed):
    rng = np.random.Generator(np.random.Philox(key=seed))
    return rng.random(shape) * 2.0 - 1.0


def main():
    assert ref.build(quiet=False), "reference build failed"
    g = {}
    # 1-D transforms, N = 32 and 128 (ut/ut_fft.cpp style inputs)
    for N in (32, 128):
        s = np.zeros(N + 1); s[1:N] = rnd(N - 1, 10 + N)
        g[f"sFFT_{N}_in"] = s; g[f"sFFT_{N}_out"] = ref.fft1d("sFFT", s, 0.37)
        s = rnd(N + 1, 20 + N)
        g[f"cFFT_{N}_in"] = s; g[f"cFFT_{N}_out"] = ref.fft1d("cFFT", s, 0.37)
        s = np.zeros(N + 1); s[:N] = rnd(N, 30 + N)
        g[f"pFFT_1_{N}_in"] = s; g[f"pFFT_1_{N}_out"] = ref.fft1d("pFFT_1", s, 0.37)
        g[f"pFFT_{N}_in"] = s; g[f"pFFT_{N}_out"] = ref.fft1d("pFFT", s, 0.37)
    # LaplCube Dirichlet 15^3 (unit cube convention of ut/ut_lapl_cube.cpp:56-61)
    n = 15; dx = 1.0 / n; l = 1 + dx
    rhs = rnd((n, n, n), 1)
    g["cube_d15_rhs"] = rhs
    g["cube_d15_ans"] = ref.LaplCube(dx, dx, dx, l, l, l, n, n, n).solve(rhs)
    # anisotropic spacing with equal point counts: the lm aliasing quirk (lapl_cube.cpp:162,171)
    g["cube_aniso15_ans"] = ref.LaplCube(0.1, 0.2, 0.3, 1.6, 3.2, 4.8, n, n, n).solve(rhs)
    # ragged sizes nz=7, ny=15, nx=31
    rhs = rnd((7, 15, 31), 2)
    g["cube_ragged_rhs"] = rhs
    g["cube_ragged_ans"] = ref.LaplCube(0.1, 0.2, 0.3, 3.2, 3.2, 2.4, 31, 15, 7).solve(rhs)
    # periodic 16^3 on [0,2pi]^3 (ut/ut_lapl_cube.cpp:157-243), mean removed
    n = 16; dx = 2 * math.pi / n; l = 2 * math.pi
    rhs = rnd((n, n, n), 3); rhs -= rhs.mean()
    g["cube_p16_rhs"] = rhs
    g["cube_p16_ans"] = ref.LaplCube(dx, dx, dx, l, l, l, n, n, n, True).solve(rhs)
    # NSCube 15^3, Re=100, dt=0.01: state after 1, 2 and 10 steps
    ns = ref.NSCube(nx=15, nz=15, Re=100.0, dt=0.01)
    done = 0
    for steps in (1, 2, 10):
        ns.step(steps - done); done = steps
        for f in ("u", "v", "w", "p"):
            g[f"nscube15_s{steps}_{f}"] = ns.field(f)
    # LaplCyl3FFT2 16 x 15 x 16 (Dirichlet z) and 16 x 16 x 16 (periodic z)
    for zp in (False, True):
        nr, nz, nphi = 16, (16 if zp else 15), 16
        R0, R1 = math.pi / 2, math.pi
        dr = (R1 - R0) / nr; dz = 10.0 / nz
        rhs = rnd((nphi, nz, nr), 4 + zp)
        tag = "p" if zp else "d"
        g[f"cyl_{tag}_rhs"] = rhs
        g[f"cyl_{tag}_ans"] = ref.LaplCyl3FFT2(dr, dz, R0 - dr / 2, R1 - R0 + dr, 10.0 if zp else 10.0 + dz,
                                               nr, nz, nphi, zp).solve(rhs)
    # LaplRect (gtsv) and LaplRectFFT2, 31 x 15 Dirichlet
    nx, ny = 31, 15; dx, dy = 0.1, 0.05
    rhs = rnd((ny, nx), 6)
    g["rect_rhs"] = rhs
    g["rect_ans"] = ref.LaplRect("rect", dx, dy, dx * (nx + 1), dy * (ny + 1), nx, ny, 0).solve(rhs)
    g["rectfft2_ans"] = ref.LaplRect("fft2", dx, dy, dx * (nx + 1), dy * (ny + 1), nx, ny, 0).solve(rhs)
    # NSCyl 16 x 15 x 16, Re=200: state after 1 and 5 steps
    ns = ref.NSCyl(False, nr=16, nz=15, nphi=16, Re=200.0, dt=0.01)
    done = 0
    for steps in (1, 5):
        ns.step(steps - done); done = steps
        for f in ("u", "v", "w", "p"):
            g[f"nscyl_s{steps}_{f}"] = ns.field(f)
    out = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
