"""Host-side mirror of the particle-mesh N-body step of the reference's test/nbody.cpp (class NBody, :24-596,
local = 0): cloud-in-cell deposit, periodic LaplCube, 4-point field differencing, gather, velocity Verlet."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

_BODY = {"x": 0, "v": 1, "a": 2, "aprev": 3, "mass": 4}
_GRID = {"f": 0, "rhs": 1, "psi": 2, "E": 3}


class PMParams(C.Structure):
    _fields_ = [("x0", C.c_double), ("y0", C.c_double), ("z0", C.c_double), ("l", C.c_double),
                ("dt", C.c_double), ("G", C.c_double), ("n", C.c_int), ("deposit_all", C.c_int)]


def _bind(L):
    if getattr(L, "_pm_bound", False):
        return
    L.fdmb_pm_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(PMParams)]
    L.fdmb_pm_set_bodies.argtypes = [C.c_void_p, C.c_longlong, capi.dp, capi.dp, capi.dp]
    L.fdmb_pm_count.argtypes = [C.c_void_p]
    L.fdmb_pm_count.restype = C.c_longlong
    L.fdmb_pm_calc_accel.argtypes = [C.c_void_p]
    L.fdmb_pm_step.argtypes = [C.c_void_p, C.c_int]
    L.fdmb_pm_get_bodies.argtypes = [C.c_void_p, C.c_int, capi.dp]
    L.fdmb_pm_get_grid.argtypes = [C.c_void_p, C.c_int, capi.dp]
    L.fdmb_pm_destroy.argtypes = [C.c_void_p]
    L._pm_bound = True


class NBodyPM:
    """``NBody(x0, y0, z0, l, n, ..., dt, G)`` (test/nbody.cpp:104); defaults are the program's (:600-611).
    ``deposit_all=False`` reproduces the reference's cell walk, which deposits only the bodies whose cell indices are
    all even or all odd (:257-272); ``True`` deposits every body."""

    def __init__(self, x0=-10.0, y0=-10.0, z0=-10.0, l=20.0, n=32, dt=0.001, G=1.0, deposit_all=False):
        L = capi.lib()
        _bind(L)
        self.n = int(n)
        self.params = PMParams(x0, y0, z0, l, dt, G, self.n, int(bool(deposit_all)))
        self._h = C.c_void_p()
        capi.check(L.fdmb_pm_create(C.byref(self._h), C.byref(self.params)), "NBody create")

    def set_bodies(self, x, v, mass):
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        mass = np.ascontiguousarray(mass, dtype=np.float64)
        N = mass.size
        if x.shape != (N, 3) or v.shape != (N, 3):
            raise ValueError("x and v must have shape (N, 3), mass (N,)")
        capi.check(capi.lib().fdmb_pm_set_bodies(self._h, N, capi.as_dp(x), capi.as_dp(v), capi.as_dp(mass)),
                   "NBody set_bodies")

    @property
    def N(self):
        return capi.lib().fdmb_pm_count(self._h)

    def calc_a_pm(self):
        capi.check(capi.lib().fdmb_pm_calc_accel(self._h), "NBody calc_a_pm")

    def step(self, nsteps=1):
        capi.check(capi.lib().fdmb_pm_step(self._h, int(nsteps)), "NBody step")

    def bodies(self, name):
        out = np.empty(self.N if name == "mass" else (self.N, 3), dtype=np.float64)
        capi.check(capi.lib().fdmb_pm_get_bodies(self._h, _BODY[name], capi.as_dp(out)), "NBody get_bodies")
        return out

    def grid(self, name):
        n = self.n
        out = np.empty((n, n, n, 3) if name == "E" else (n, n, n), dtype=np.float64)
        capi.check(capi.lib().fdmb_pm_get_grid(self._h, _GRID[name], capi.as_dp(out)), "NBody get_grid")
        return out

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            capi.lib().fdmb_pm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
