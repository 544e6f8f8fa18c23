"""Times the particle-mesh N-body step on the GPU next to the unmodified reference program on the host cores.
The program's default problem (test/nbody.cpp:600-611: n = 32, N = 100000) and a larger one.
Usage (under gpurun):  python scripts/gpu_pm.py > gpurun_out/pm.json"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fdm_b200  # noqa: E402
from oracle import fdm_oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402

BOX = dict(x0=-10.0, y0=-10.0, z0=-10.0, l=20.0)
res = []
for n, N, ref_steps in ((32, 100000, 5), (128, 1000000, 2)):
    row = {"n": n, "N": N}
    if R.available():
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)                       # the reference prints a line per step
        try:
            B = R.NBody(n=n, N=N, **BOX)
            x, v, m = B.bodies("x"), B.bodies("v"), B.bodies("mass")
            B.step(1)
            t0 = time.perf_counter(); B.step(ref_steps); row["reference_ms_per_step"] = 1e3 * (time.perf_counter() - t0) / ref_steps
            row["reference_threads"] = R.num_threads()
        finally:
            os.dup2(saved, 1)
    else:
        rng = np.random.default_rng(1)
        x, v, m = rng.uniform(-10, 10, (N, 3)), np.zeros((N, 3)), rng.uniform(0.2, 1.7, N)
    P = fdm_b200.NBodyPM(n=n, **BOX)
    P.set_bodies(x, v, m)
    P.step(1 + (ref_steps if R.available() else 0))
    if R.available():
        row["x_rel_l2_vs_reference"] = O.rel_l2(P.bodies("x"), B.bodies("x"))
        row["a_rel_l2_vs_reference"] = O.rel_l2(P.bodies("a"), B.bodies("a"))
    L = fdm_b200.lib()
    steps = 200
    P.step(10)
    L.fdmb_profile_begin()
    t0 = time.perf_counter(); P.step(steps); row["ms_per_step"] = 1e3 * (time.perf_counter() - t0) / steps
    out = C.create_string_buffer(65536)
    L.fdmb_profile_end(out, 65536)
    row["kernels_us"] = {ln.split()[0]: round(1e3 * float(ln.split()[2]) / int(ln.split()[1]), 2) for ln in out.value.decode().splitlines()}
    res.append(row)
print(json.dumps(res))
