"""Times the device-side velocity_plotter next to the NS step it follows (255^3 cavity, BASELINE configs[2]) and the
path a host-side plotter needs (download u,v,w + the unmodified reference plotter on the host cores).
Usage (under gpurun):  python scripts/gpu_vplot.py [n] > gpurun_out/vplot.json"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fdm_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 255
ns = fdm_b200.NSCube(nx=n, nz=n, Re=1000.0, dt=0.005)
ns.step(20)
ns.synchronize()
t0 = time.perf_counter(); ns.step(100); ns.synchronize(); t_step = (time.perf_counter() - t0) / 100

P = fdm_b200.VelocityPlotter.for_ns_cube(ns)
P.update()
L = fdm_b200.lib()
import ctypes as C  # noqa: E402
L.fdmb_profile_begin()
t0 = time.perf_counter()
for _ in range(20):
    P.update()
t_update = (time.perf_counter() - t0) / 20
out = C.create_string_buffer(65536)
L.fdmb_profile_end(out, 65536)
kern = {}
for ln in out.value.decode().splitlines():
    tag, cnt, ms = ln.split()
    kern[tag] = {"launches": int(cnt) / 20, "us_per_launch": 1e3 * float(ms) / int(cnt)}

t0 = time.perf_counter(); c = P.cell_velocity(); t_cells = time.perf_counter() - t0
with tempfile.TemporaryDirectory() as tmp:
    t0 = time.perf_counter(); P.vtk_out(os.path.join(tmp, "a.vtk"), 1); t_vtk = time.perf_counter() - t0
    vtk_bytes = os.path.getsize(os.path.join(tmp, "a.vtk"))

res = {"n": n, "ns_step_ms": 1e3 * t_step, "update_ms": 1e3 * t_update, "update_kernels": kern,
       "cell_velocity_ms": 1e3 * t_cells, "cell_velocity_bytes": int(c.nbytes), "vtk_out_s": t_vtk, "vtk_bytes": vtk_bytes}

# what a host-side plotter costs: D2H of the three fields + the reference's update() / vtk_out on the host cores
t0 = time.perf_counter(); u, v, w = (ns.field(f) for f in "uvw"); t_d2h = time.perf_counter() - t0
res["download_uvw_ms"] = 1e3 * t_d2h
try:
    from oracle import ref as R
    if R.available():
        p = ns.params
        d = (p.x2 - p.x1) / n
        RP = R.VelocityPlotter(d, d, d, n, n, n, p.x1, p.x2, p.y1, p.y2, p.z1, p.z2)
        t0 = time.perf_counter(); RP.update(u, v, w); res["reference_update_ms"] = 1e3 * (time.perf_counter() - t0)
        with tempfile.TemporaryDirectory() as tmp:
            t0 = time.perf_counter(); RP.vtk_out(os.path.join(tmp, "r.vtk"), 1)
            res["reference_vtk_out_s"] = time.perf_counter() - t0
        res["reference_threads"] = R.num_threads()
        from oracle import fdm_oracle as O
        res["psi_y_rel_l2_vs_reference"] = O.rel_l2(P.slice("psi_y").ravel(), RP.slice("psi_y"))
except Exception as e:  # the timing of our own path above stands on its own
    res["reference_error"] = repr(e)
print(json.dumps(res))
