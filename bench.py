#!/usr/bin/env python
"""bench.py -- headline benchmark of the fdm_b200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input:
  cube1023 (default) one LaplCube Dirichlet solve, 1023^3 fp64  (BASELINE.json configs[4]: the
                     configuration the Gpts/s metric is quoted on "at 1/2/4/8 B200"; it fits one GPU)
  cube127            one LaplCube Dirichlet solve, 127^3 fp64   (configs[1])
  cube255 / cube511  one LaplCube Dirichlet solve, 255^3 / 511^3 fp64
  nscube255          one NSCube lid-driven-cavity step, 255^3   (configs[2])
  nscube31           one NSCube step, 31^3                      (configs[0])
N>1: ONE solve / ONE time step of the same grid, z-slab decomposed over the N ranks ("scaling": "strong"); the
two slab<->pencil transposes of the solve are peer stores over NVLink fused into the y and z sweeps, the NS
stencils read their halo planes from the neighbours' memory.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` goes through the
host-pointer C-ABI entry point with pinned host buffers (H2D + solve + D2H inside the timed
region); `roofline` is for the dominant kernel, timed live with CUDA events; `cpu_baseline` is the
UNMODIFIED reference (oracle/_ref) on this box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALGO_BYTES_PER_SWEEP = 16.0      # one fp64 read + one fp64 write per grid point
SOLVE_BYTES_PER_PT = 48.0        # 3 sweeps
NS_BYTES_PER_PT = 168.0

WORKLOADS = {
    "cube63": dict(kind="cube", n=63, label="LaplCube Dirichlet 63^3 fp64 solve"),
    "cube127": dict(kind="cube", n=127, label="LaplCube Dirichlet 127^3 fp64 solve (BASELINE configs[1])"),
    "cube255": dict(kind="cube", n=255, label="LaplCube Dirichlet 255^3 fp64 solve"),
    "cube511": dict(kind="cube", n=511, label="LaplCube Dirichlet 511^3 fp64 solve"),
    "cube1023": dict(kind="cube", n=1023, label="LaplCube Dirichlet 1023^3 fp64 solve, 1024 slots per axis "
                                                "(BASELINE configs[4]; z-slab decomposed over the ranks when N>1)"),
    "nscube31": dict(kind="ns", n=31, Re=250.0, dt=0.01, label="NSCube 31^3 Re=250 dt=0.01 step (configs[0])"),
    "nscube255": dict(kind="ns", n=255, Re=1000.0, dt=0.005, label="NSCube 255^3 Re=1000 dt=0.005 step (configs[2])"),
    # cylindrical workloads: nr x nz x nphi (phi-slabs when N > 1)
    "nscyl128": dict(kind="nscyl", n=128, nr=128, nz=127, nphi=128, Re=200.0, dt=0.01,
                     label="NSCyl Taylor-Couette nr=128 nz=127 nphi=128 Re=200 dt=0.01 step (configs[3])"),
    "cyl128": dict(kind="cyl", n=128, nr=128, nz=127, nphi=128,
                   label="LaplCyl3FFT2 Dirichlet-z solve nr=128 nz=127 nphi=128 (the pressure solve of configs[3])"),
}


def cyl_geometry(wl):
    import math
    # the solver NSCyl builds for itself: src/ns_cyl.h:94-97 with the README box R = pi, r = pi/2, h = 10
    R, r0, h = math.pi, math.pi / 2, 10.0
    dr = (R - r0) / wl["nr"]; dz = h / wl["nz"]
    return (dr, dz, r0 - dr / 2, R - r0 + dr, h + dz, wl["nr"], wl["nz"], wl["nphi"])


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cube_geometry(n):
    # unit-cube convention of ut/ut_lapl_cube.cpp:56-61
    dx = 1.0 / n
    return dict(dx=dx, l=1.0 + dx)


def golden_full_size():
    """The committed one-off run of the unmodified reference at the full 1023^3 size (tests/golden/make_golden_cube1023.py)."""
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "golden_cube1023_v1.npz"))
        sec, thr, n = float(g["ref_seconds"]), int(g["ref_threads"]), int(g["n"])
        return {"n": n, "seconds": sec, "threads": thr, "gpts_per_s": n ** 3 / 1e9 / sec,
                "where": "dev container, tests/golden/make_golden_cube1023.py (committed fixture golden_cube1023_v1.npz)"}
    except Exception:
        return None


def reference_full_size(n, cores):
    """ONE real solve of the full-size workload by the unmodified reference on this box (needs ~27 GB of host memory
    and 10-30 s); skipped when the host is short of memory or FDMB_BENCH_FULLREF=0."""
    if os.environ.get("FDMB_BENCH_FULLREF", "1") == "0":
        return {"skipped": "FDMB_BENCH_FULLREF=0"}
    try:
        import psutil
        need = 3.3 * 8 * n ** 3            # rhs, ans, the reference's internal copy + slack
        avail = psutil.virtual_memory().available
        if avail < need + 8e9:
            return {"skipped": f"host has {avail / 1e9:.0f} GB available, the {n}^3 reference solve needs {need / 1e9:.0f} GB"}
        from oracle import ref
        g = cube_geometry(n)
        rhs = np.empty((n, n, n))
        rng = np.random.Generator(np.random.Philox(key=99))
        plane = rng.random((n, n)) - 0.5
        for z in range(n):                      # cheap synthetic field: one random plane, modulated along z
            np.multiply(plane, np.cos(0.37 * z) + 0.1, out=rhs[z])
        S = ref.LaplCube(g["dx"], g["dx"], g["dx"], g["l"], g["l"], g["l"], n, n, n)
        t0 = time.perf_counter()
        ans = S.solve(rhs)
        dt = time.perf_counter() - t0
        ok = bool(np.isfinite(ans[::97, ::89, ::83]).all())
        del ans, rhs, S
        return {"n": n, "seconds": dt, "threads": cores, "gpts_per_s": n ** 3 / 1e9 / dt, "finite": ok,
                "where": "this run, this box: one full-size solve of the unmodified reference"}
    except Exception as e:
        return {"skipped": f"{type(e).__name__}: {e}"}


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation (oracle/_ref) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref, fdm_oracle as O
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libfdm_ref.so not built"}))
        return
    n = wl["n"]
    cores = ref.use_all_cores()       # torch.distributed.run exports OMP_NUM_THREADS=1: use every core we may run on
    extra_cfg = {}
    full = None
    if wl["kind"] == "cube":
        nn = min(n, 255)      # bounded sample: the reference needs ~27 GB and 10-30 s per 1023^3 solve
        g = cube_geometry(nn)
        S = ref.LaplCube(g["dx"], g["dx"], g["dx"], g["l"], g["l"], g["l"], nn, nn, nn)
        rhs = O.synthetic_rhs((nn, nn, nn), seed=1234)
        step = lambda: S.solve(rhs)
        units = nn ** 3 / 1e9
        metric, unit = "poisson_solve_gpts_per_s", "Gpts/s"
        sample = f"full {nn}^3 solve per step" + ("" if nn == n else f" (bounded sample of the {n}^3 workload; Gpts/s is size-normalised)")
        if nn != n:
            extra_cfg["reference_sample"] = (f"each step is one full {nn}^3 LaplCube solve of the unmodified reference; Gpts/s is "
                                             f"size-normalised; cpu_baseline.full_size holds real {n}^3 solves")
    elif wl["kind"] == "cyl":
        S = ref.LaplCyl3FFT2(*cyl_geometry(wl))
        shape = (wl["nphi"], wl["nz"], wl["nr"])
        rhs = O.synthetic_rhs(shape, seed=1234)
        step = lambda: S.solve(rhs)
        units = shape[0] * shape[1] * shape[2] / 1e9
        metric, unit = "poisson_solve_gpts_per_s", "Gpts/s"
        sample = "full solve per step"
    elif wl["kind"] == "nscyl":
        ns = ref.NSCyl(nr=wl["nr"], nz=wl["nz"], nphi=wl["nphi"], Re=wl["Re"], dt=wl["dt"])
        step = lambda: ns.step(1)
        units = 1.0
        metric, unit = "ns_steps_per_s", "steps/s"
        sample = "full step per step"
    else:
        ns = ref.NSCube(nx=n, nz=n, Re=wl["Re"], dt=wl["dt"])
        step = lambda: ns.step(1)
        units = 1.0
        metric, unit = "ns_steps_per_s", "steps/s"
        sample = f"full {n}^3 step per step"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = units * args.steps / dt
    cb = {"value": value, "unit": unit, "cores": cores, "kind": "reference", "sample": sample}
    if wl["kind"] == "cube" and n > 255:
        cb["full_size_committed"] = golden_full_size()
        if args.gpus == 1:
            cb["full_size"] = reference_full_size(n, cores)
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong" if wl["kind"] == "cube" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], **extra_cfg},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


CPU_SAMPLE_SECONDS = 10.0     # bounded sample of CPU work per cpu_baseline (the task statement asks for about 10-30 s)


def timed_sample(fn, min_seconds=None, max_reps=1000):
    """Repeats fn() until min_seconds of work have accumulated; returns (reps, mean seconds, best seconds)."""
    min_seconds = CPU_SAMPLE_SECONDS if min_seconds is None else min_seconds
    total, best, reps = 0.0, 1e30, 0
    while reps < max_reps and (total < min_seconds or reps < 3):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        total += dt
        best = min(best, dt)
        reps += 1
    return reps, total / reps, best


def cpu_baseline(wl):
    """The unmodified reference on the host cores, bounded sample (rank 0, N=1 only); value = mean over the sample."""
    try:
        from oracle import ref, fdm_oracle as O
        if not ref.available():
            return {"value": None, "unit": "", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
        n = wl["n"]
        cores = ref.use_all_cores()
        if wl["kind"] == "cube":
            nn = min(n, 255)
            g = cube_geometry(nn)
            S = ref.LaplCube(g["dx"], g["dx"], g["dx"], g["l"], g["l"], g["l"], nn, nn, nn)
            rhs = O.synthetic_rhs((nn, nn, nn), seed=1234)
            S.solve(rhs)
            reps, mean, best = timed_sample(lambda: S.solve(rhs))
            out = {"value": nn ** 3 / 1e9 / mean, "unit": "Gpts/s", "cores": cores, "kind": "reference",
                   "sample": f"{reps} full {nn}^3 solves, {reps * mean:.1f} s of work: mean {mean * 1e3:.1f} ms, best "
                             f"{best * 1e3:.1f} ms" + ("" if nn == n else f" (Gpts/s is size-normalised; the {n}^3 solve needs 27 GB on the host)")}
            if nn != n:
                out["full_size_committed"] = golden_full_size()
            return out
        if wl["kind"] == "cyl":
            S = ref.LaplCyl3FFT2(*cyl_geometry(wl))
            shape = (wl["nphi"], wl["nz"], wl["nr"])
            rhs = O.synthetic_rhs(shape, seed=1234)
            S.solve(rhs)
            reps, mean, best = timed_sample(lambda: S.solve(rhs))
            return {"value": rhs.size / 1e9 / mean, "unit": "Gpts/s", "cores": cores, "kind": "reference",
                    "sample": f"{reps} full solves, {reps * mean:.1f} s of work: mean {mean * 1e3:.1f} ms, best {best * 1e3:.1f} ms"}
        if wl["kind"] == "nscyl":
            ns = ref.NSCyl(nr=wl["nr"], nz=wl["nz"], nphi=wl["nphi"], Re=wl["Re"], dt=wl["dt"])
            ns.step(1)
            reps, mean, best = timed_sample(lambda: ns.step(1))
            return {"value": 1.0 / mean, "unit": "steps/s", "cores": cores, "kind": "reference",
                    "sample": f"{reps} full steps, {reps * mean:.1f} s of work: mean {mean * 1e3:.1f} ms, best {best * 1e3:.1f} ms"}
        nn = min(n, 127)
        ns = ref.NSCube(nx=nn, nz=nn, Re=wl["Re"], dt=wl["dt"])
        ns.step(1)
        reps, mean, best = timed_sample(lambda: ns.step(1))
        val = 1.0 / mean
        note = f"{reps} steps at {nn}^3, {reps * mean:.1f} s of work: mean {mean * 1e3:.1f} ms, best {best * 1e3:.1f} ms"
        if nn != n:
            val *= (nn / n) ** 3
            note += f"; scaled by ({nn}/{n})^3 to the {n}^3 workload"
        return {"value": val, "unit": "steps/s", "cores": cores, "kind": "reference", "sample": note}
    except Exception as e:  # never let the baseline kill the bench line
        return {"value": None, "unit": "", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}


def bind_to_gpu_numa_node(torch, local):
    """One process per GPU: run (and therefore first-touch the pinned host buffers of the e2e leg) on the NUMA node
    the GPU hangs off, so that the eight ranks' PCIe traffic does not cross the socket interconnect.  Pure host
    placement; a no-op wherever the topology is not exposed (sysfs numa_node = -1, containers, VMs)."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return {"gpu": bdf, "node": node, "bound": False}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"gpu": bdf, "node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"gpu": bdf, "node": node, "bound": True, "cpus": len(cpus)}
    except Exception as e:        # placement is an optimisation, never a reason to fail the run
        return {"bound": False, "error": type(e).__name__}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="cube1023", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-batch", action="store_true", help="skip the pipelined multi-solve e2e measurement")
    ap.add_argument("--no-extra", action="store_true", help="skip the NS steps/s lines added to the default workload")
    ap.add_argument("--no-check", action="store_true", help="skip the closed-form known-answer check of the cube solve")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return

    import torch
    import torch.distributed as dist
    import fdm_b200
    from fdm_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    affinity0 = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    L = fdm_b200.lib()
    capi.check(L.fdmb_set_device(local), "set_device")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    # a non-default stream: the C ABI treats a NULL stream as "the handle's own stream"
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0
    W = max(args.warmup, 3)
    K = args.steps
    n = wl["n"]
    peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sharded = world > 1 and wl["kind"] == "cube"       # (the NS branch decides for itself below)
    e2e_batch = None
    if wl["kind"] == "cube":
        g = cube_geometry(n)
        if sharded:
            S = fdm_b200.LaplCubeSharded(g["dx"], g["dx"], g["dx"], g["l"], g["l"], g["l"], n, n, n, rank=rank, nranks=world)
            S.connect()
            nz_local = S.nz_local
        else:
            S = fdm_b200.LaplCube(g["dx"], g["dx"], g["dx"], g["l"], g["l"], g["l"], n, n, n)
            nz_local = n
        pts = n ** 3                        # whole job
        lpts = n * n * nz_local             # this rank's slab
        pair_bytes = 2 * 8 * lpts
        nbuf = max(2, int(np.ceil(2.2 * 126e6 / pair_bytes)))       # rotate > 2x L2 worth of rhs/ans pairs
        rhs = [torch.rand(lpts, dtype=torch.float64, device=dev) - 0.5 for _ in range(nbuf)]
        ans = [torch.empty(lpts, dtype=torch.float64, device=dev) for _ in range(nbuf)]
        l2_policy = f"rotating {nbuf} rhs/ans pairs ({nbuf * pair_bytes / 1e6:.0f} MB > 126 MB L2)"

        def step(i):
            b = i % nbuf
            S.solve_device(ans[b].data_ptr(), rhs[b].data_ptr(), sptr)
        units_per_step = pts / 1e9 / (world if sharded else 1)       # x world below: one job over all ranks
        metric, unit = "poisson_solve_gpts_per_s", "Gpts/s"
        algo_bytes_step = SOLVE_BYTES_PER_PT * lpts
        # e2e: host-pointer entry point, pinned buffers (this rank's slab)
        h_rhs = torch.rand(lpts, dtype=torch.float64).pin_memory()
        h_ans = torch.empty(lpts, dtype=torch.float64).pin_memory()
        dp = C.POINTER(C.c_double)

        def e2e_step():
            capi.check(L.fdmb_lapl_cube_solve(S._h, C.cast(h_ans.data_ptr(), dp), C.cast(h_rhs.data_ptr(), dp)), "solve")
        h2d = d2h = 8 * lpts
        pts_kernel = lpts
        if not sharded:
            # the pipelined multi-solve entry point (uploads / downloads of neighbouring solves overlap the solve)
            def e2e_batch(count):
                h_ans1 = getattr(e2e_batch, "h_ans1", None)
                if h_ans1 is None:
                    h_ans1 = e2e_batch.h_ans1 = torch.empty(lpts, dtype=torch.float64).pin_memory()
                outs = [h_ans.data_ptr(), h_ans1.data_ptr()]
                S.solve_batch([outs[i % 2] for i in range(count)], [h_rhs.data_ptr()] * count)
    elif wl["kind"] in ("cyl", "nscyl"):
        sharded = world > 1
        shard_kw = dict(rank=rank, nranks=world) if sharded else {}
        pts = wl["nr"] * wl["nz"] * wl["nphi"]
        lpts = pts // world if sharded else pts                 # phi-slabs
        if wl["kind"] == "cyl":
            S = fdm_b200.LaplCyl3FFT2(*cyl_geometry(wl), **shard_kw)
            S.connect()
            pair_bytes = 2 * 8 * lpts
            nbuf = max(2, int(np.ceil(2.2 * 126e6 / pair_bytes)))
            rhs = [torch.rand(lpts, dtype=torch.float64, device=dev) - 0.5 for _ in range(nbuf)]
            ans = [torch.empty(lpts, dtype=torch.float64, device=dev) for _ in range(nbuf)]
            l2_policy = f"rotating {nbuf} rhs/ans pairs ({nbuf * pair_bytes / 1e6:.0f} MB > 126 MB L2)"

            def step(i):
                S.solve_device(ans[i % nbuf].data_ptr(), rhs[i % nbuf].data_ptr(), sptr)
            units_per_step = pts / 1e9 / (world if sharded else 1)
            metric, unit = "poisson_solve_gpts_per_s", "Gpts/s"
            algo_bytes_step = SOLVE_BYTES_PER_PT * lpts
            h_rhs = torch.rand(lpts, dtype=torch.float64).pin_memory()
            h_ans = torch.empty(lpts, dtype=torch.float64).pin_memory()
            dp = C.POINTER(C.c_double)

            def e2e_step():
                capi.check(L.fdmb_lapl_cyl_solve(S._h, C.cast(h_ans.data_ptr(), dp), C.cast(h_rhs.data_ptr(), dp)), "solve")
            h2d = d2h = 8 * lpts
        else:
            kw = dict(nr=wl["nr"], nz=wl["nz"], nphi=wl["nphi"], Re=wl["Re"], dt=wl["dt"])
            ns = fdm_b200.NSCyl(**kw, **shard_kw)
            ns.connect()
            l2_policy = f"state of 13 arrays = {13 * 8 * lpts / 1e6:.0f} MB per GPU" + (" > 126 MB L2" if 13 * 8 * lpts > 126e6 else " (fits L2)")

            def step(i):
                ns.step_device(1, sptr)
            units_per_step = 1.0 / (world if sharded else 1)
            metric, unit = "ns_steps_per_s", "steps/s"
            algo_bytes_step = NS_BYTES_PER_PT * lpts
            ns2 = fdm_b200.NSCyl(**kw, **shard_kw)
            ns2.connect()

            def e2e_step():
                ns2.step_host_roundtrip()
            h2d = d2h = ns2.state_bytes()
        pts_kernel = lpts
    else:
        if not hasattr(fdm_b200, "NSCube"):
            raise SystemExit("NSCube workload not built")
        sharded = world > 1 and (n + 1) // world >= 4
        shard_kw = dict(rank=rank, nranks=world) if sharded else {}
        ns = fdm_b200.NSCube(nx=n, nz=n, Re=wl["Re"], dt=wl["dt"], **shard_kw)
        ns.connect()
        pts = n ** 3
        lpts = n * n * (ns.local_planes("x")[1])
        l2_policy = f"state of 13 arrays = {13 * 8 * lpts / 1e6:.0f} MB per GPU" + (" > 126 MB L2" if 13 * 8 * lpts > 126e6 else " (fits L2; launch-bound size)")

        def step(i):
            ns.step_device(1, sptr)
        units_per_step = 1.0 / (world if sharded else 1)      # x world below: one time step of the whole grid
        metric, unit = "ns_steps_per_s", "steps/s"
        algo_bytes_step = NS_BYTES_PER_PT * lpts
        ns2 = fdm_b200.NSCube(nx=n, nz=n, Re=wl["Re"], dt=wl["dt"], **shard_kw)
        ns2.connect()

        def e2e_step():
            ns2.step_host_roundtrip()
        h2d = d2h = ns2.state_bytes()
        pts_kernel = lpts

    # ---- device-resident timing --------------------------------------------------
    for i in range(W):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.fdmb_launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(K):
        step(W + i)
    e1.record(stream)
    barrier()
    launches = L.fdmb_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # nvidia-smi samples every 100 ms: a timed region shorter than ~0.5 s is followed by an UNTIMED tail of the same
    # steps (every rank runs it: the sharded steps are collective) so that the clocks are sampled under this load
    tail = 0
    if ms < 500.0:
        tail = int(500.0 / max(ms / K, 1e-3)) + 1
        for i in range(tail):
            step(W + K + i)
        barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed region" if tail == 0 else f"timed region + {tail} identical untimed steps"
    value = units_per_step * K * world / (ms * 1e-3)

    # ---- per-kernel live timing for the roofline (separate pass, same steps) -----
    L.fdmb_profile_begin.restype = C.c_int
    L.fdmb_profile_end.argtypes = [C.c_char_p, C.c_int]
    capi.check(L.fdmb_profile_begin(), "profile_begin")
    for i in range(K):
        step(W + i)
    buf = C.create_string_buffer(1 << 16)
    capi.check(L.fdmb_profile_end(buf, len(buf)), "profile_end")
    kern = []
    for line in buf.value.decode().splitlines():
        tag, cnt, tot = line.split()
        kern.append((tag, int(cnt), float(tot)))
    tot_kernel_ms = sum(k[2] for k in kern) or 1.0
    top = max(kern, key=lambda k: k[2])
    per_launch_ms = top[2] / top[1]
    # algorithmic bytes of one launch of the dominant kernel: one read+write sweep over the grid
    algo_launch = ALGO_BYTES_PER_SWEEP * pts_kernel * kernel_sweeps(top[0])
    achieved = algo_launch / (per_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": load_traffic(top[0], args.workload), "peak_source": peak_src,
                "kernel_share_of_step": top[2] / tot_kernel_ms,
                "step_achieved_gbs": algo_bytes_step / (ms / K * 1e-3) / 1e9,
                "step_frac": algo_bytes_step / (ms / K * 1e-3) / 1e9 / peak,
                "kernels": {k[0]: {"launches_per_step": k[1] / K, "ms_per_launch": k[2] / k[1]} for k in kern}}

    if sharded:
        # per GPU 48 n^3 / P bytes of HBM traffic plus two slab<->pencil transposes, each sending (and
        # receiving) 8 (n^3 / P) (P - 1) / P bytes over NVLink (900 GB/s per direction per GPU, nominal)
        nvl = 900.0
        t_hbm = (SOLVE_BYTES_PER_PT if wl["kind"] in ("cube", "cyl") else NS_BYTES_PER_PT) * pts / world / (peak * 1e9) * 1e3
        xpose_bytes = 8.0 * pts / world * (world - 1) / world
        t_nvl = 2 * xpose_bytes / (nvl * 1e9) * 1e3
        t_step = ms / K
        roofline["sharded"] = {
            "hbm_ms": t_hbm, "nvlink_ms": t_nvl, "nvlink_peak_gbs_per_dir": nvl,
            "nvlink_bytes_per_gpu_per_transpose": xpose_bytes,
            "t_roof_ms_no_overlap": t_hbm + t_nvl, "t_roof_ms_overlap": max(t_hbm, t_nvl),
            "frac_no_overlap": (t_hbm + t_nvl) / t_step, "frac_overlap": max(t_hbm, t_nvl) / t_step,
            "note": "step_frac above is the per-GPU HBM fraction alone; frac_no_overlap is the aggregate "
                    "HBM + NVLink roofline of BASELINE.json's 8-GPU target (>= 0.5)"}

    # ---- e2e through the host-pointer C ABI -----------------------------------------
    Ke = max(3, min(K, 50 if pts < 5e8 else 10))
    for _ in range(2 if pts >= 5e8 else 3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    t = torch.tensor([te], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    te = float(t.item())
    e2e = {"value": units_per_step * Ke * world / te, "unit": unit, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": Ke, "ms_per_step": 1e3 * te / Ke,
           "api": "one synchronous host-pointer call per step (the reference's solve(ans, rhs) / step())"}
    if numa is not None:
        e2e["host_placement"] = numa      # rank 0's NUMA binding (see bind_to_gpu_numa_node)
    if e2e_batch is not None and not args.no_e2e_batch:
        # same copies per solve, but neighbouring solves' transfers overlap (fdmb_lapl_cube_solve_batch)
        e2e_batch(2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_batch(Ke)
        tb = time.perf_counter() - t0
        e2e["pipelined"] = {"value": units_per_step * Ke / tb, "unit": unit, "steps": Ke, "ms_per_step": 1e3 * tb / Ke,
                            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                            "api": "fdmb_lapl_cube_solve_batch: independent solves, transfers overlapped"}

    # ---- correctness of what was just timed: closed-form eigenvector known answer (fdm_b200/selfcheck.py) ----
    check = None
    if wl["kind"] == "cube" and not args.no_check:
        from fdm_b200 import selfcheck
        g = cube_geometry(n)
        z0 = S.z_first if sharded else 0
        k_rhs, k_want = selfcheck.kat_device(torch, n, g["dx"], z0, nz_local, dev)
        k_ans = ans[0].view(nz_local, n, n)
        k_ans.fill_(float("nan"))
        torch.cuda.synchronize()
        S.solve_device(k_ans.data_ptr(), k_rhs.data_ptr(), sptr)
        torch.cuda.synchronize()
        nd = torch.stack([((k_ans - k_want) ** 2).sum(), (k_want ** 2).sum()])
        if world > 1:
            dist.all_reduce(nd)
        check = {"kind": "eigenvector known answer: rhs = 5 discrete sine products (modes 1 .. n), ans = rhs / -(lx+ly+lz) "
                         "in closed form; whole grid, all ranks (fdm_b200/selfcheck.py)",
                 "rel_l2": float((nd[0] / nd[1]).sqrt().item()), "tolerance": 1e-12,
                 "modes": selfcheck.kat_modes(n)}
        check["ok"] = bool(check["rel_l2"] <= check["tolerance"])
        del k_rhs, k_want

    # ---- the other half of the metric: NS steps/s on the same GPUs (default workload only) -------------------
    extra = None
    if args.workload == "cube1023" and not args.no_extra:
        extra = {}
        for name in ("nscube255", "nscyl128"):
            try:
                extra[name] = measure_ns(name, torch, dist, fdm_b200, world, rank, dev, stream, peak)
            except Exception as e:          # never let an extra kill the headline line
                extra[name] = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        out = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if (sharded or world == 1) else "weak",
            "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["label"], "l2_policy": l2_policy,
                       "parallelism": ("single GPU" if world == 1 else
                                       f"{'phi' if 'cyl' in wl['kind'] else 'z'}-slabs over {world} GPUs, slab<->pencil transposes as peer stores over NVLink"
                                       + ("" if wl["kind"] in ("cube", "cyl") else ", stencil halo planes pulled from the neighbours")
                                       if sharded else "independent replicas per rank")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        }
        if check is not None:
            out["check"] = check
        if extra is not None:
            out["extra"] = extra
        if not args.no_cpu_baseline:
            try:
                os.sched_setaffinity(0, affinity0)      # undo the NUMA binding of the e2e leg: the baseline gets every core
            except Exception:
                pass
            out["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_ns(name, torch, dist, fdm_b200, world, rank, dev, stream, peak, K=100, W=5):
    """Device-resident NS steps/s of one of the NS workloads (sharded over the ranks when they are several)."""
    wl = WORKLOADS[name]
    sptr = stream.cuda_stream
    if wl["kind"] == "nscyl":
        sharded = world > 1
        kw = dict(rank=rank, nranks=world) if sharded else {}
        ns = fdm_b200.NSCyl(nr=wl["nr"], nz=wl["nz"], nphi=wl["nphi"], Re=wl["Re"], dt=wl["dt"], **kw)
        pts = wl["nr"] * wl["nz"] * wl["nphi"]
    else:
        n = wl["n"]
        sharded = world > 1 and (n + 1) // world >= 4
        kw = dict(rank=rank, nranks=world) if sharded else {}
        ns = fdm_b200.NSCube(nx=n, nz=n, Re=wl["Re"], dt=wl["dt"], **kw)
        pts = n ** 3
    ns.connect()
    for _ in range(W):
        ns.step_device(1, sptr)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        ns.step_device(1, sptr)
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    ns.close()
    lpts = pts / world if sharded else pts
    return {"workload": wl["label"], "metric": "ns_steps_per_s", "value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms,
            "steps": K, "warmup": W, "sharded_over": world if sharded else 1,
            "step_frac": NS_BYTES_PER_PT * lpts / (ms * 1e-3) / 1e9 / peak,
            "note": "fraction of the 168 B/pt HBM roofline per GPU (SURVEY 8d); device-resident, CUDA events, max over ranks"}


def kernel_sweeps(tag):
    """How many read+write sweeps over the grid one launch of this kernel stands for."""
    return {"ns_fgh": 3.0, "ns_rhs": 2.0, "ns_fgh_rhs": 3.5, "ns_update": 4.0,
            "nscyl_fgh": 3.0, "nscyl_lfgh": 4.5, "nscyl_rhs": 2.0, "nscyl_update": 4.0}.get(tag, 1.0)


def load_traffic(tag, workload):
    """dram bytes per launch from the committed ncu --set full capture, or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(path)).get(workload, {}).get(tag)
    except Exception:
        return None


if __name__ == "__main__":
    main()
