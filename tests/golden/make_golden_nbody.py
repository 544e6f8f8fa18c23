"""Generates tests/golden/golden_nbody_v1.npz from the compiled, unmodified reference program test/nbody.cpp
(oracle/_ref): the bodies its default_random_engine seeds (n = 16, N = 300), psi and E after the first step, bodies
after five steps.  Run in the dev container:  python tests/golden/make_golden_nbody.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref as R                      # noqa: E402

assert R.build(), "needs /root/reference"
n, N = 16, 300
B = R.NBody(n=n, N=N, x0=-10.0, y0=-10.0, z0=-10.0, l=20.0)
out = {"n": n, "x0": B.bodies("x"), "v0": B.bodies("v"), "mass": B.bodies("mass")}
B.step(1)
out["psi1"], out["E1"] = B.grid("psi"), B.grid("E")
B.step(4)
for b in ("x", "v", "a"):
    out[b + "5"] = B.bodies(b)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_nbody_v1.npz"), **out)
print("wrote", len(out), "arrays")
