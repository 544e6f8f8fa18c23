"""Host-side mirror of fdm::NSCyl (reference src/ns_cyl.h:17-132, src/ns_cyl.cpp:23-484).

Constructor keywords are the reference's ``[ns]`` config keys (src/ns_cyl.h:57-68; note the inner
radius key is ``r``); ``zperiodic`` selects the ``zflag`` template argument.  ``step()`` / ``L_step()``
advance the projection scheme on the device; ``field(name)`` returns the same raw storage the
reference exposes as ``ns.u.vec`` etc. ([phi][z][r], ghosts included, src/ns_cyl.h:80-93).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import capi

FIELD_IDS = {"u": 0, "v": 1, "w": 2, "p": 3, "x": 4, "F": 5, "G": 6, "H": 7, "RHS": 8, "u0": 9, "v0": 10, "w0": 11}


class NSCylParams(C.Structure):
    _fields_ = [("R", C.c_double), ("r", C.c_double), ("h1", C.c_double), ("h2", C.c_double),
                ("u0", C.c_double), ("Re", C.c_double), ("dt", C.c_double),
                ("nr", C.c_int), ("nz", C.c_int), ("nphi", C.c_int),
                ("verbose", C.c_int), ("vrandom", C.c_int), ("zperiodic", C.c_int)]


def _bind(L):
    if getattr(L, "_ns_cyl_bound", False):
        return
    P = C.POINTER(NSCylParams)
    L.fdmb_ns_cyl_default_params.argtypes = [P]
    L.fdmb_ns_cyl_create.argtypes = [C.POINTER(C.c_void_p), P]
    L.fdmb_ns_cyl_step.argtypes = [C.c_void_p, C.c_int]
    L.fdmb_ns_cyl_lstep.argtypes = [C.c_void_p, C.c_int]
    L.fdmb_ns_cyl_step_async.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.fdmb_ns_cyl_field_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
    L.fdmb_ns_cyl_get_field.argtypes = [C.c_void_p, C.c_int, capi.dp]
    L.fdmb_ns_cyl_set_field.argtypes = [C.c_void_p, C.c_int, capi.dp]
    L.fdmb_ns_cyl_field_device_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.fdmb_ns_cyl_time_index.argtypes = [C.c_void_p]
    L.fdmb_ns_cyl_time_index.restype = C.c_longlong
    L.fdmb_ns_cyl_destroy.argtypes = [C.c_void_p]
    L.fdmb_ns_cyl_set_u0.argtypes = [C.c_void_p, C.c_double]
    L.fdmb_ns_cyl_create_sharded.argtypes = [C.POINTER(C.c_void_p), P, C.c_int, C.c_int]
    L.fdmb_ns_cyl_local_slab.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.fdmb_ns_cyl_export_ipc.argtypes = [C.c_void_p, C.c_void_p]
    L.fdmb_ns_cyl_attach_ipc.argtypes = [C.c_void_p, C.c_void_p]
    L.fdmb_ns_cyl_attach_local.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.fdmb_ns_cyl_synchronize.argtypes = [C.c_void_p]
    L._ns_cyl_bound = True


class NSCyl:
    """Taylor-Couette flow; defaults are the reference's (src/ns_cyl.h:57-68)."""

    def __init__(self, nr=32, nz=31, nphi=32, Re=1.0, dt=0.001, u0=1.0, R=math.pi, r=math.pi / 2,
                 h1=0.0, h2=10.0, verbose=0, vrandom=0, zperiodic=False, rank=0, nranks=1):
        L = capi.lib()
        _bind(L)
        self.params = NSCylParams(R, r, h1, h2, u0, Re, dt, int(nr), int(nz), int(nphi), int(verbose),
                                  int(vrandom), int(bool(zperiodic)))
        self.nr, self.nz, self.nphi = int(nr), int(nz), int(nphi)
        self.zperiodic = bool(zperiodic)
        self.dt = float(dt)
        self.rank, self.nranks = int(rank), int(nranks)
        self._h = C.c_void_p()
        if self.nranks > 1:
            capi.check(L.fdmb_ns_cyl_create_sharded(C.byref(self._h), C.byref(self.params), self.rank, self.nranks),
                       "NSCyl create_sharded")
        else:
            capi.check(L.fdmb_ns_cyl_create(C.byref(self._h), C.byref(self.params)), "NSCyl create")
        a, b = C.c_int(), C.c_int()
        capi.check(L.fdmb_ns_cyl_local_slab(self._h, C.byref(a), C.byref(b)), "local_slab")
        self.phi_first, self.nphi_local = a.value, b.value      # this rank's phi planes (all of them on one GPU)

    # ---- several GPUs: phi-slabs; field()/set_field() act on this rank's planes ------------------------
    def connect(self, group=None):
        """One process per GPU: exchange the IPC handles over ``torch.distributed`` and attach the peers."""
        if self.nranks == 1:
            return
        import torch.distributed as dist
        from .lapl_cube import IPC_HANDLE_BYTES
        buf = C.create_string_buffer(2 * IPC_HANDLE_BYTES)
        capi.check(capi.lib().fdmb_ns_cyl_export_ipc(self._h, buf), "export_ipc")
        gathered = [None] * self.nranks
        dist.all_gather_object(gathered, buf.raw, group=group)
        blob = b"".join(gathered)
        capi.check(capi.lib().fdmb_ns_cyl_attach_ipc(self._h, C.create_string_buffer(blob, len(blob))), "attach_ipc")
        dist.barrier(group=group)

    @staticmethod
    def connect_local(parts):
        """All ranks live in this process (one handle per device)."""
        arr = (C.c_void_p * len(parts))(*[s._h for s in parts])
        for s in parts:
            capi.check(capi.lib().fdmb_ns_cyl_attach_local(s._h, arr), "attach_local")

    def synchronize(self):
        capi.check(capi.lib().fdmb_ns_cyl_synchronize(self._h), "synchronize")

    # ---- reference API -----------------------------------------------------------------
    def step(self, nsteps=1):
        capi.check(capi.lib().fdmb_ns_cyl_step(self._h, int(nsteps)), "NSCyl step")

    def L_step(self, nsteps=1):
        capi.check(capi.lib().fdmb_ns_cyl_lstep(self._h, int(nsteps)), "NSCyl L_step")

    def set_u0(self, u0):
        """The public, non-const member ``U0`` of the reference class (src/ns_cyl.h:23): wall speed used from the
        next step on (test/test_ns_cyl_spectral.cpp sets it to 0 for the perturbation problem)."""
        capi.check(capi.lib().fdmb_ns_cyl_set_u0(self._h, float(u0)), "NSCyl set_u0")

    def size(self):
        """u.size + v.size + w.size + p.size (src/ns_cyl.h:114-116)."""
        return sum(self.field_size(f) for f in "uvwp")

    @property
    def time_index(self):
        return capi.lib().fdmb_ns_cyl_time_index(self._h)

    def field_size(self, name):
        n = C.c_longlong()
        capi.check(capi.lib().fdmb_ns_cyl_field_size(self._h, FIELD_IDS[name], C.byref(n)), "field_size")
        return n.value

    def field(self, name, out=None):
        if out is None:
            out = np.empty(self.field_size(name), dtype=np.float64)
        capi.check(capi.lib().fdmb_ns_cyl_get_field(self._h, FIELD_IDS[name], capi.as_dp(out)), "get_field")
        return out

    def set_field(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        if a.size != self.field_size(name):
            raise ValueError(f"field {name} has {self.field_size(name)} elements, got {a.size}")
        capi.check(capi.lib().fdmb_ns_cyl_set_field(self._h, FIELD_IDS[name], capi.as_dp(a)), "set_field")

    # ---- device-side extras --------------------------------------------------------------
    def step_device(self, nsteps=1, stream=0, linear=False):
        capi.check(capi.lib().fdmb_ns_cyl_step_async(self._h, int(nsteps), int(linear), C.c_void_p(stream)),
                   "NSCyl step_async")

    def field_device_ptr(self, name):
        p = C.c_void_p()
        capi.check(capi.lib().fdmb_ns_cyl_field_device_ptr(self._h, FIELD_IDS[name], C.byref(p)), "device_ptr")
        return p.value

    def state_bytes(self):
        return 8 * self.size()

    def step_host_roundtrip(self):
        """One step the way a host-resident caller sees it: upload u,v,w,p from pinned host memory, step,
        download u,v,w,p (what the reference keeps in ns.u.vec ...)."""
        import torch
        if getattr(self, "_pinned", None) is None:
            self._pinned = {f: torch.from_numpy(self.field(f)).pin_memory() for f in "uvwp"}
        L = capi.lib()
        for f, t in self._pinned.items():
            capi.check(L.fdmb_ns_cyl_set_field(self._h, FIELD_IDS[f], C.cast(t.data_ptr(), capi.dp)), "set_field")
        self.step(1)
        for f, t in self._pinned.items():
            capi.check(L.fdmb_ns_cyl_get_field(self._h, FIELD_IDS[f], C.cast(t.data_ptr(), capi.dp)), "get_field")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            capi.lib().fdmb_ns_cyl_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
