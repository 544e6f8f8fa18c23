"""Host-side mirror of fdm::NSCube (reference src/ns_cube.h:13-92, src/ns_cube.cpp:27-277).

Constructor keywords are the reference's ``[ns]`` config keys (src/ns_cube.h:47-61);
``step()`` advances the projection scheme on the device; ``field(name)`` returns the same
raw storage the reference exposes as ``ns.u.vec`` etc. (ghosts included, src/ns_cube.h:66-75).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import capi

FIELD_IDS = {"u": 0, "v": 1, "w": 2, "p": 3, "x": 4, "F": 5, "G": 6, "H": 7, "RHS": 8}


class NSCubeParams(C.Structure):
    _fields_ = [("x1", C.c_double), ("y1", C.c_double), ("z1", C.c_double),
                ("x2", C.c_double), ("y2", C.c_double), ("z2", C.c_double),
                ("u0", C.c_double), ("Re", C.c_double), ("dt", C.c_double),
                ("nx", C.c_int), ("nz", C.c_int), ("verbose", C.c_int)]


def _bind(L):
    if getattr(L, "_ns_cube_bound", False):
        return
    P = C.POINTER(NSCubeParams)
    L.fdmb_ns_cube_default_params.argtypes = [P]
    L.fdmb_ns_cube_create.argtypes = [C.POINTER(C.c_void_p), P]
    L.fdmb_ns_cube_step.argtypes = [C.c_void_p, C.c_int]
    L.fdmb_ns_cube_step_async.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.fdmb_ns_cube_field_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
    L.fdmb_ns_cube_get_field.argtypes = [C.c_void_p, C.c_int, capi.dp]
    L.fdmb_ns_cube_set_field.argtypes = [C.c_void_p, C.c_int, capi.dp]
    L.fdmb_ns_cube_field_device_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.fdmb_ns_cube_time_index.argtypes = [C.c_void_p]
    L.fdmb_ns_cube_time_index.restype = C.c_longlong
    L.fdmb_ns_cube_destroy.argtypes = [C.c_void_p]
    L.fdmb_ns_cube_create_sharded.argtypes = [C.POINTER(C.c_void_p), P, C.c_int, C.c_int]
    ip = C.POINTER(C.c_int)
    L.fdmb_ns_cube_owned_planes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip, ip]
    L.fdmb_ns_cube_local_planes.argtypes = [C.c_void_p, C.c_int, ip, ip]
    L.fdmb_ns_cube_export_ipc.argtypes = [C.c_void_p, C.c_void_p]
    L.fdmb_ns_cube_attach_ipc.argtypes = [C.c_void_p, C.c_void_p]
    L.fdmb_ns_cube_attach_local.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.fdmb_ns_cube_synchronize.argtypes = [C.c_void_p]
    fp = C.POINTER(C.c_float)
    L.fdmb_ns_cube_f32_create.argtypes = [C.POINTER(C.c_void_p), P]
    L.fdmb_ns_cube_f32_step.argtypes = [C.c_void_p, C.c_int]
    L.fdmb_ns_cube_f32_field_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
    L.fdmb_ns_cube_f32_get_field.argtypes = [C.c_void_p, C.c_int, fp]
    L.fdmb_ns_cube_f32_set_field.argtypes = [C.c_void_p, C.c_int, fp]
    L.fdmb_ns_cube_f32_time_index.argtypes = [C.c_void_p]
    L.fdmb_ns_cube_f32_time_index.restype = C.c_longlong
    L.fdmb_ns_cube_f32_destroy.argtypes = [C.c_void_p]
    L._ns_cube_bound = True


def owned_planes(nz, field, rank, nranks):
    """(first global z index, number of planes) of ``field`` that ``rank`` reports through ``field()``; the ranks'
    ranges tile the reference array's z range (pure function, no device needed)."""
    L = capi.lib()
    _bind(L)
    a, b = C.c_int(), C.c_int()
    capi.check(L.fdmb_ns_cube_owned_planes(int(nz), FIELD_IDS[field], int(rank), int(nranks), C.byref(a), C.byref(b)),
               "owned_planes")
    return a.value, b.value


class NSCube:
    """Lid-driven cavity on a staggered grid; ``ny`` is taken from ``nx`` like the reference
    (src/ns_cube.h:58).  Defaults are the reference's (src/ns_cube.h:47-61)."""

    def __init__(self, nx=32, nz=32, Re=1.0, dt=0.001, u0=1.0,
                 x1=-math.pi, y1=-math.pi, z1=-math.pi, x2=math.pi, y2=math.pi, z2=math.pi, verbose=0,
                 rank=0, nranks=1):
        L = capi.lib()
        _bind(L)
        self.params = NSCubeParams(x1, y1, z1, x2, y2, z2, u0, Re, dt, int(nx), int(nz), int(verbose))
        self.nx, self.ny, self.nz = int(nx), int(nx), int(nz)
        self.dt = float(dt)
        self.rank, self.nranks = int(rank), int(nranks)
        self._h = C.c_void_p()
        if self.nranks > 1:
            capi.check(L.fdmb_ns_cube_create_sharded(C.byref(self._h), C.byref(self.params), self.rank, self.nranks),
                       "NSCube create_sharded")
        else:
            capi.check(L.fdmb_ns_cube_create(C.byref(self._h), C.byref(self.params)), "NSCube create")
        self._pinned = None

    # ---- several GPUs: z-slabs, every rank holds and reports its own planes ----------------------
    def local_planes(self, name):
        a, b = C.c_int(), C.c_int()
        capi.check(capi.lib().fdmb_ns_cube_local_planes(self._h, FIELD_IDS[name], C.byref(a), C.byref(b)), "local_planes")
        return a.value, b.value

    def connect(self, group=None):
        """One process per GPU: exchange the IPC handles over ``torch.distributed`` and attach the peers."""
        if self.nranks == 1:
            return
        import torch.distributed as dist
        from .lapl_cube import IPC_HANDLE_BYTES
        buf = C.create_string_buffer(2 * IPC_HANDLE_BYTES)
        capi.check(capi.lib().fdmb_ns_cube_export_ipc(self._h, buf), "export_ipc")
        gathered = [None] * self.nranks
        dist.all_gather_object(gathered, buf.raw, group=group)
        blob = b"".join(gathered)
        capi.check(capi.lib().fdmb_ns_cube_attach_ipc(self._h, C.create_string_buffer(blob, len(blob))), "attach_ipc")
        dist.barrier(group=group)

    @staticmethod
    def connect_local(parts):
        """All ranks live in this process (one handle per device)."""
        arr = (C.c_void_p * len(parts))(*[s._h for s in parts])
        for s in parts:
            capi.check(capi.lib().fdmb_ns_cube_attach_local(s._h, arr), "attach_local")

    def synchronize(self):
        capi.check(capi.lib().fdmb_ns_cube_synchronize(self._h), "synchronize")

    # ---- reference API -----------------------------------------------------------------
    def step(self, nsteps=1):
        capi.check(capi.lib().fdmb_ns_cube_step(self._h, int(nsteps)), "NSCube step")

    @property
    def time_index(self):
        return capi.lib().fdmb_ns_cube_time_index(self._h)

    def field_size(self, name):
        n = C.c_longlong()
        capi.check(capi.lib().fdmb_ns_cube_field_size(self._h, FIELD_IDS[name], C.byref(n)), "field_size")
        return n.value

    def field(self, name, out=None):
        if out is None:
            out = np.empty(self.field_size(name), dtype=np.float64)
        capi.check(capi.lib().fdmb_ns_cube_get_field(self._h, FIELD_IDS[name], capi.as_dp(out)), "get_field")
        return out

    def set_field(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        if a.size != self.field_size(name):
            raise ValueError(f"field {name} has {self.field_size(name)} elements, got {a.size}")
        capi.check(capi.lib().fdmb_ns_cube_set_field(self._h, FIELD_IDS[name], capi.as_dp(a)), "set_field")

    # ---- device-side extras --------------------------------------------------------------
    def step_device(self, nsteps=1, stream=0):
        """Asynchronous steps on ``stream`` (raw cudaStream_t as int; 0 = the handle's own stream)."""
        capi.check(capi.lib().fdmb_ns_cube_step_async(self._h, int(nsteps), C.c_void_p(stream)), "NSCube step_async")

    def field_device_ptr(self, name):
        p = C.c_void_p()
        capi.check(capi.lib().fdmb_ns_cube_field_device_ptr(self._h, FIELD_IDS[name], C.byref(p)), "device_ptr")
        return p.value

    def state_bytes(self):
        return 8 * sum(self.field_size(f) for f in "uvwp")

    def step_host_roundtrip(self):
        """One step the way a host-resident caller sees it: upload u,v,w,p from pinned host
        memory, step, download u,v,w,p (what the reference keeps in ns.u.vec ...)."""
        import torch
        if self._pinned is None:
            self._pinned = {f: torch.from_numpy(self.field(f)).pin_memory() for f in "uvwp"}
        L = capi.lib()
        for f, t in self._pinned.items():
            capi.check(L.fdmb_ns_cube_set_field(self._h, FIELD_IDS[f], C.cast(t.data_ptr(), capi.dp)), "set_field")
        self.step(1)
        for f, t in self._pinned.items():
            capi.check(L.fdmb_ns_cube_get_field(self._h, FIELD_IDS[f], C.cast(t.data_ptr(), capi.dp)), "get_field")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            capi.lib().fdmb_ns_cube_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NSCubeF32:
    """``fdm::NSCube<float,check>`` (src/ns_cube.cpp:281-282): float32 fields and pressure solve, same keywords,
    field names and extents as :class:`NSCube` (single GPU)."""

    def __init__(self, nx=32, nz=32, Re=1.0, dt=0.001, u0=1.0,
                 x1=-math.pi, y1=-math.pi, z1=-math.pi, x2=math.pi, y2=math.pi, z2=math.pi, verbose=0):
        L = capi.lib()
        _bind(L)
        self.params = NSCubeParams(x1, y1, z1, x2, y2, z2, u0, Re, dt, int(nx), int(nz), int(verbose))
        self.nx, self.ny, self.nz = int(nx), int(nx), int(nz)
        self._h = C.c_void_p()
        capi.check(L.fdmb_ns_cube_f32_create(C.byref(self._h), C.byref(self.params)), "NSCube<float> create")

    def step(self, nsteps=1):
        capi.check(capi.lib().fdmb_ns_cube_f32_step(self._h, int(nsteps)), "NSCube<float> step")

    @property
    def time_index(self):
        return capi.lib().fdmb_ns_cube_f32_time_index(self._h)

    def field_size(self, name):
        n = C.c_longlong()
        capi.check(capi.lib().fdmb_ns_cube_f32_field_size(self._h, FIELD_IDS[name], C.byref(n)), "field_size")
        return n.value

    def field(self, name):
        out = np.empty(self.field_size(name), dtype=np.float32)
        capi.check(capi.lib().fdmb_ns_cube_f32_get_field(self._h, FIELD_IDS[name], out.ctypes.data_as(C.POINTER(C.c_float))),
                   "get_field")
        return out

    def set_field(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float32).ravel()
        if a.size != self.field_size(name):
            raise ValueError(f"field {name} has {self.field_size(name)} elements, got {a.size}")
        capi.check(capi.lib().fdmb_ns_cube_f32_set_field(self._h, FIELD_IDS[name], a.ctypes.data_as(C.POINTER(C.c_float))),
                   "set_field")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            capi.lib().fdmb_ns_cube_f32_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
