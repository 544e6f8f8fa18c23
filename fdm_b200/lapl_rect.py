"""Host-side mirrors of fdm::LaplRect and fdm::LaplRectFFT2 (reference src/lapl_rect.h, src/lapl_rect.cpp)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def _bind(L):
    if getattr(L, "_lapl_rect_bound", False):
        return
    L.fdmb_lapl_rect_create.argtypes = [C.POINTER(C.c_void_p)] + [C.c_int] * 3 + [C.c_double] * 4 + [C.c_int] * 2
    L.fdmb_lapl_rect_set_scales.argtypes = [C.c_void_p, capi.dp, capi.dp, capi.dp]
    L.fdmb_lapl_rect_solve.argtypes = [C.c_void_p, capi.dp, capi.dp]
    L.fdmb_lapl_rect_solve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fdmb_lapl_rect_destroy.argtypes = [C.c_void_p]
    L._lapl_rect_bound = True


class LaplRect:
    """y transform + tridiagonal solves along x; arrays ``[ny][nx]`` (x fastest).

    Mirrors ``fdm::LaplRect<double,check,F>(dx,dy,lx,ly,nx,ny)`` (src/lapl_rect.h:40-68);
    ``yperiodic`` selects ``F = tensor_flags<tensor_flag::periodic>``.  x is always Dirichlet
    (src/lapl_rect.cpp:47).
    """

    _KIND = 0

    def __init__(self, dx, dy, lx, ly, nx, ny, yperiodic=False, xperiodic=False):
        L = capi.lib()
        _bind(L)
        self.nx, self.ny = int(nx), int(ny)
        self._h = C.c_void_p()
        capi.check(L.fdmb_lapl_rect_create(C.byref(self._h), self._KIND, int(bool(yperiodic)), int(bool(xperiodic)),
                                           float(dx), float(dy), float(lx), float(ly), self.nx, self.ny),
                   f"{type(self).__name__} create")

    def set_scales(self, lm_y_scale=None, L_scale=None, U_scale=None):
        """Overwrite the per-column scales (src/lapl_rect.h:57-59); each has nx+1 entries, entry j for column j."""
        keep = []

        def arg(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.size != self.nx + 1:
                raise ValueError(f"scale arrays hold nx+1 = {self.nx + 1} entries")
            keep.append(a)
            return capi.as_dp(a)
        capi.check(capi.lib().fdmb_lapl_rect_set_scales(self._h, arg(lm_y_scale), arg(L_scale), arg(U_scale)),
                   "LaplRect set_scales")

    def solve(self, ans, rhs=None):
        if rhs is None:
            rhs, ans = ans, None
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        n = self.nx * self.ny
        if rhs.size != n:
            raise ValueError(f"rhs has {rhs.size} elements, expected {n}")
        if ans is None:
            ans = np.empty((self.ny, self.nx), dtype=np.float64)
        if not (isinstance(ans, np.ndarray) and ans.dtype == np.float64 and ans.flags.c_contiguous and ans.size == n):
            raise ValueError("ans must be a C-contiguous float64 array of the same size as rhs")
        capi.check(capi.lib().fdmb_lapl_rect_solve(self._h, capi.as_dp(ans), capi.as_dp(rhs)), "LaplRect solve")
        return ans

    def solve_device(self, d_ans, d_rhs, stream=0):
        capi.check(capi.lib().fdmb_lapl_rect_solve_device(self._h, C.c_void_p(d_ans), C.c_void_p(d_rhs),
                                                         C.c_void_p(stream)), "LaplRect solve_device")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            capi.lib().fdmb_lapl_rect_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LaplRectFFT2(LaplRect):
    """Transforms on both axes (``fdm::LaplRectFFT2``, src/lapl_rect.h:89-105, src/lapl_rect.cpp:113-207)."""

    _KIND = 1
