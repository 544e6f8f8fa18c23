"""Generates tests/golden/golden_vplot_v1.npz from the compiled, unmodified reference (oracle/_ref):
velocity_plotter::update() slices, right-hand sides and stream functions for the seeded inputs of
tests/test_velocity_plot_cpu.py (the inputs are regenerated from the seed, only outputs are stored).
Run in the dev container:  python tests/golden/make_golden_vplot.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref as R                      # noqa: E402
from tests.test_velocity_plot_cpu import CASES, PSI, SLICES, fields   # noqa: E402

assert R.build(), "needs /root/reference"
out = {}
for name, (args, kw) in CASES.items():
    u, v, w = fields(name)
    P = R.VelocityPlotter(*args, **kw)
    P.update(u, v, w)
    for s in SLICES + PSI:
        out[f"{name}/{s}"] = P.slice(s)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_vplot_v1.npz"), **out)
print("wrote", len(out), "arrays")
