"""Host-side mirror of fdm::LaplCube (reference src/lapl_cube.h:9-106).

Same constructor arguments and ``solve(ans, rhs)`` meaning as the reference class;
the body is the CUDA path behind the C ABI (include/fdm_b200.h).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class LaplCube:
    """3-D Poisson solve, all-Dirichlet (``periodic=False``) or all-periodic.

    Mirrors ``fdm::LaplCube<double,check,F>(dx,dy,dz,lx,ly,lz,nx,ny,nz)``
    (src/lapl_cube.h:58-100); ``periodic`` selects between the two instantiated
    flag sets of src/lapl_cube.cpp:174-182.  Invalid sizes raise instead of the
    reference's ``verify`` abort (src/fft.cpp:67).
    """

    def __init__(self, dx, dy, dz, lx, ly, lz, nx, ny, nz, periodic=False):
        self.dx, self.dy, self.dz = float(dx), float(dy), float(dz)
        self.lx, self.ly, self.lz = float(lx), float(ly), float(lz)
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.periodic = bool(periodic)
        self._h = C.c_void_p()
        L = capi.lib()
        capi.check(L.fdmb_lapl_cube_create(C.byref(self._h), self.dx, self.dy, self.dz, self.lx, self.ly, self.lz,
                                           self.nx, self.ny, self.nz, int(self.periodic)), "LaplCube create")

    @property
    def shape(self):
        return (self.nz, self.ny, self.nx)

    def solve(self, ans, rhs=None):
        """``solve(ans, rhs)`` like the reference (host arrays, interior points only).

        ``solve(rhs)`` with one argument allocates and returns ``ans``.
        """
        if rhs is None:
            rhs, ans = ans, None
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        if rhs.size != self.nx * self.ny * self.nz:
            raise ValueError(f"rhs has {rhs.size} elements, expected {self.nx * self.ny * self.nz}")
        if ans is None:
            ans = np.empty(self.shape, dtype=np.float64)
        if not (isinstance(ans, np.ndarray) and ans.dtype == np.float64 and ans.flags.c_contiguous
                and ans.size == rhs.size):
            raise ValueError("ans must be a C-contiguous float64 array of the same size as rhs")
        capi.check(capi.lib().fdmb_lapl_cube_solve(self._h, capi.as_dp(ans), capi.as_dp(rhs)), "LaplCube solve")
        return ans

    def solve_batch(self, ans_ptrs, rhs_ptrs):
        """Pipelined independent solves: raw HOST pointers (ints; page-locked memory lets the copies overlap the
        solves).  Same results as one ``solve`` per pair."""
        n = len(rhs_ptrs)
        if len(ans_ptrs) != n:
            raise ValueError("one ans per rhs")
        a = (C.c_void_p * n)(*ans_ptrs)
        r = (C.c_void_p * n)(*rhs_ptrs)
        capi.check(capi.lib().fdmb_lapl_cube_solve_batch(self._h, n, a, r), "LaplCube solve_batch")

    def solve_device(self, d_ans, d_rhs, stream=0):
        """Device-resident solve: raw device pointers (ints), asynchronous on ``stream``."""
        capi.check(capi.lib().fdmb_lapl_cube_solve_device(self._h, C.c_void_p(d_ans), C.c_void_p(d_rhs),
                                                         C.c_void_p(stream)), "LaplCube solve_device")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            capi.lib().fdmb_lapl_cube_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LaplCubeF32:
    """``fdm::LaplCube<float,check,F>`` (src/lapl_cube.cpp:176-177,181-182): same constructor and ``solve`` as
    :class:`LaplCube` with float32 arrays; the device arrays, tables and butterflies are float32 (single GPU)."""

    def __init__(self, dx, dy, dz, lx, ly, lz, nx, ny, nz, periodic=False):
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.periodic = bool(periodic)
        self._h = C.c_void_p()
        capi.check(capi.lib().fdmb_lapl_cube_f32_create(C.byref(self._h), float(dx), float(dy), float(dz), float(lx),
                                                        float(ly), float(lz), self.nx, self.ny, self.nz,
                                                        int(self.periodic)), "LaplCube<float> create")

    @property
    def shape(self):
        return (self.nz, self.ny, self.nx)

    def solve(self, ans, rhs=None):
        if rhs is None:
            rhs, ans = ans, None
        rhs = np.ascontiguousarray(rhs, dtype=np.float32)
        if rhs.size != self.nx * self.ny * self.nz:
            raise ValueError(f"rhs has {rhs.size} elements, expected {self.nx * self.ny * self.nz}")
        if ans is None:
            ans = np.empty(self.shape, dtype=np.float32)
        if not (isinstance(ans, np.ndarray) and ans.dtype == np.float32 and ans.flags.c_contiguous and ans.size == rhs.size):
            raise ValueError("ans must be a C-contiguous float32 array of the same size as rhs")
        fp = C.POINTER(C.c_float)
        capi.check(capi.lib().fdmb_lapl_cube_f32_solve(self._h, ans.ctypes.data_as(fp), rhs.ctypes.data_as(fp)),
                   "LaplCube<float> solve")
        return ans

    def solve_device(self, d_ans, d_rhs, stream=0):
        capi.check(capi.lib().fdmb_lapl_cube_f32_solve_device(self._h, C.c_void_p(d_ans), C.c_void_p(d_rhs),
                                                             C.c_void_p(stream)), "LaplCube<float> solve_device")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            capi.lib().fdmb_lapl_cube_f32_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


IPC_HANDLE_BYTES = 64


def slab_range(n, periodic, nranks, rank):
    """(first, count) of the 0-based interior entries of an axis of ``n`` points that ``rank`` owns
    (``fdmb_slab_range``; host-only arithmetic, usable without a GPU)."""
    first, count = C.c_int(), C.c_int()
    capi.check(capi.lib().fdmb_slab_range(int(n), int(bool(periodic)), int(nranks), int(rank),
                                          C.byref(first), C.byref(count)), "slab_range")
    return first.value, count.value


class LaplCubeSharded(LaplCube):
    """One rank of the z-slab decomposed LaplCube solve (include/fdm_b200.h).

    Same constructor as :class:`LaplCube` plus ``rank``/``nranks``; ``solve`` takes and returns this
    rank's slab ``[nz_local][ny][nx]``.  Before the first solve the ranks must be connected:
    ``connect(group)`` (one process per GPU, handles exchanged through ``torch.distributed``) or
    ``LaplCubeSharded.connect_local(solvers)`` (all ranks in one process).
    """

    def __init__(self, dx, dy, dz, lx, ly, lz, nx, ny, nz, rank, nranks, periodic=False):
        self.dx, self.dy, self.dz = float(dx), float(dy), float(dz)
        self.lx, self.ly, self.lz = float(lx), float(ly), float(lz)
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.periodic = bool(periodic)
        self.rank, self.nranks = int(rank), int(nranks)
        self._h = C.c_void_p()
        L = capi.lib()
        capi.check(L.fdmb_lapl_cube_create_sharded(C.byref(self._h), self.dx, self.dy, self.dz, self.lx, self.ly,
                                                   self.lz, self.nx, self.ny, self.nz, int(self.periodic),
                                                   self.rank, self.nranks), "LaplCube create_sharded")
        zf, nzl = C.c_int(), C.c_int()
        capi.check(L.fdmb_lapl_cube_local_slab(self._h, C.byref(zf), C.byref(nzl)), "local_slab")
        self.z_first, self.nz_local = zf.value, nzl.value

    @property
    def shape(self):
        return (self.nz_local, self.ny, self.nx)

    def export_ipc(self) -> bytes:
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        capi.check(capi.lib().fdmb_lapl_cube_export_ipc(self._h, buf), "export_ipc")
        return buf.raw

    def attach_ipc(self, handles):
        """``handles``: the exported handles of all ranks, ordered by rank."""
        blob = b"".join(handles)
        if len(blob) != IPC_HANDLE_BYTES * self.nranks:
            raise ValueError("need one IPC handle per rank")
        capi.check(capi.lib().fdmb_lapl_cube_attach_ipc(self._h, C.create_string_buffer(blob, len(blob))), "attach_ipc")

    def connect(self, group=None):
        """Exchange the IPC handles over ``torch.distributed`` (any backend) and attach the peers."""
        if self.nranks == 1:
            return
        import torch
        import torch.distributed as dist
        mine = self.export_ipc()
        gathered = [None] * self.nranks
        dist.all_gather_object(gathered, mine, group=group)
        self.attach_ipc(gathered)
        dist.barrier(group=group)
        del torch

    @staticmethod
    def connect_local(solvers):
        """All ranks live in this process (one handle per device): enable peer access both ways."""
        arr = (C.c_void_p * len(solvers))(*[s._h for s in solvers])
        for s in solvers:
            capi.check(capi.lib().fdmb_lapl_cube_attach_local(s._h, arr), "attach_local")

    def solve(self, ans, rhs=None):
        if rhs is None:
            rhs, ans = ans, None
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        want = self.nx * self.ny * self.nz_local
        if rhs.size != want:
            raise ValueError(f"rhs slab has {rhs.size} elements, expected {want}")
        if ans is None:
            ans = np.empty(self.shape, dtype=np.float64)
        capi.check(capi.lib().fdmb_lapl_cube_solve(self._h, capi.as_dp(ans), capi.as_dp(rhs)), "LaplCube solve (slab)")
        return ans
