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
