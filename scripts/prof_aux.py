"""A few invocations of the paths bench.py does not cover -- LaplRect / LaplRectFFT2 (the plotter's 2-D solvers, on-the-fly
tridiagonals), the device-side velocity_plotter and the particle-mesh N-body step -- for an ncu launch list / capture.
Usage (under gpurun):  ncu ... python scripts/prof_aux.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fdm_b200  # noqa: E402

rng = np.random.default_rng(1)
# the reference's own 511 x 511 LaplRect case (ut/ut_lapl_rect.cpp:384-455) and a 2047 x 255 one
for nx, ny in ((511, 511), (2047, 255)):
    dx, dy = 1.0 / nx, 1.0 / ny
    rhs = rng.uniform(-1, 1, (ny, nx))
    for cls in (fdm_b200.LaplRect, fdm_b200.LaplRectFFT2):
        S = cls(dx, dy, 1 + dx, 1 + dy, nx, ny)
        for _ in range(3):
            S.solve(rhs)
# velocity_plotter on a 255^3 cavity state
ns = fdm_b200.NSCube(nx=255, nz=255, Re=1000.0, dt=0.005)
ns.step(3)
P = fdm_b200.VelocityPlotter.for_ns_cube(ns)
for _ in range(3):
    P.update()
# PM N-body step, n = 128, N = 1e6 (uniform random bodies in the box)
N = 1000000
B = fdm_b200.NBodyPM(n=128, x0=-10.0, y0=-10.0, z0=-10.0, l=20.0)
B.set_bodies(rng.uniform(-10, 10, (N, 3)), rng.normal(0, 0.1, (N, 3)), rng.uniform(0.5, 1.5, N))
B.step(3)
print("done")
