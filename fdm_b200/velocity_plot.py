"""Host-side mirror of fdm::velocity_plotter (reference src/velocity_plot.h:11-143, src/velocity_plot.cpp:10-222):
mid-plane slices of a staggered velocity field, their stream functions and the ASCII VTK writer, computed on the
device from the NS state where it lives."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

SLICE_IDS = {"vx": 0, "wx": 1, "uy": 2, "wy": 3, "uz": 4, "vz": 5, "RHS_x": 6, "RHS_y": 7, "RHS_z": 8,
             "psi_x": 9, "psi_y": 10, "psi_z": 11}


class VPlotParams(C.Structure):
    _fields_ = [("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
                ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("xx1", C.c_double), ("xx2", C.c_double), ("yy1", C.c_double), ("yy2", C.c_double),
                ("zz1", C.c_double), ("zz2", C.c_double),
                ("cyl", C.c_int), ("zperiodic", C.c_int), ("yperiodic", C.c_int)]


def _bind(L):
    if getattr(L, "_vplot_bound", False):
        return
    L.fdmb_vplot_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(VPlotParams)]
    L.fdmb_vplot_field_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
    L.fdmb_vplot_use_host.argtypes = [C.c_void_p, capi.dp, capi.dp, capi.dp]
    L.fdmb_vplot_use_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fdmb_vplot_use_ns_cube.argtypes = [C.c_void_p, C.c_void_p]
    L.fdmb_vplot_use_ns_cyl.argtypes = [C.c_void_p, C.c_void_p]
    L.fdmb_vplot_update.argtypes = [C.c_void_p]
    L.fdmb_vplot_slice_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.fdmb_vplot_get_slice.argtypes = [C.c_void_p, C.c_int, capi.dp]
    L.fdmb_vplot_cell_velocity.argtypes = [C.c_void_p, capi.dp]
    L.fdmb_vplot_vtk_out.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.fdmb_vplot_destroy.argtypes = [C.c_void_p]
    L._vplot_bound = True


class VelocityPlotter:
    """``velocity_plotter<double,check,F>(dx,dy,dz, nx,ny,nz, xx1,xx2, yy1,yy2, zz1,zz2, cyl)``
    (src/velocity_plot.h:59-71); ``zperiodic`` / ``yperiodic`` select F (src/velocity_plot.cpp:222-235).
    Axis names are the reference's: z slowest, x fastest; for cylinders (phi, z, r)."""

    def __init__(self, dx, dy, dz, nx, ny, nz, xx1, xx2, yy1, yy2, zz1, zz2, cyl=False, zperiodic=False,
                 yperiodic=False):
        L = capi.lib()
        _bind(L)
        self.params = VPlotParams(dx, dy, dz, int(nx), int(ny), int(nz), xx1, xx2, yy1, yy2, zz1, zz2,
                                  int(bool(cyl)), int(bool(zperiodic)), int(bool(yperiodic)))
        self._h = C.c_void_p()
        self._keep = None
        capi.check(L.fdmb_vplot_create(C.byref(self._h), C.byref(self.params)), "velocity_plotter create")

    @classmethod
    def for_ns_cube(cls, ns):
        """The plotter test/test_ns_cube.cpp:24-30 builds, reading the NSCube state on the device."""
        p = ns.params
        nx, ny, nz = ns.nx, ns.ny, ns.nz
        self = cls((p.x2 - p.x1) / nx, (p.y2 - p.y1) / ny, (p.z2 - p.z1) / nz, nx, ny, nz,
                   p.x1, p.x2, p.y1, p.y2, p.z1, p.z2)
        self.use(ns)
        return self

    @classmethod
    def for_ns_cyl(cls, ns):
        """The plotter test/test_ns_cyl.cpp:53-58 builds (dr, dz, dphi; nr, nz, nphi; r0, R; h1, h2; 0, 2 pi; cyl)."""
        import math
        p = ns.params
        dr, dz, dphi = (p.R - p.r) / p.nr, (p.h2 - p.h1) / p.nz, 2 * math.pi / p.nphi
        self = cls(dr, dz, dphi, p.nr, p.nz, p.nphi, p.r, p.R, p.h1, p.h2, 0.0, 2 * math.pi, cyl=True,
                   zperiodic=True, yperiodic=bool(p.zperiodic))
        self.use(ns)
        return self

    def field_size(self, name):
        n = C.c_longlong()
        capi.check(capi.lib().fdmb_vplot_field_size(self._h, "uvw".index(name), C.byref(n)), "field_size")
        return n.value

    def use(self, u, v=None, w=None):
        """``use(u, v, w)`` with host arrays of the reference's extents (re-read at every ``update()``), or
        ``use(ns)`` with an ``NSCube`` / ``NSCyl`` whose device state is read in place."""
        L = capi.lib()
        if v is None:
            from .ns_cube import NSCube
            from .ns_cyl import NSCyl
            if isinstance(u, NSCube):
                capi.check(L.fdmb_vplot_use_ns_cube(self._h, u._h), "use(NSCube)")
            elif isinstance(u, NSCyl):
                capi.check(L.fdmb_vplot_use_ns_cyl(self._h, u._h), "use(NSCyl)")
            else:
                raise TypeError("use(ns) needs an NSCube or NSCyl")
            self._keep = u
            return
        arrs = []
        for name, a in zip("uvw", (u, v, w)):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous):
                raise ValueError("use() keeps pointers like the reference: pass C-contiguous float64 arrays")
            if a.size != self.field_size(name):
                raise ValueError(f"{name} has {a.size} elements, expected {self.field_size(name)}")
            arrs.append(a)
        self._keep = arrs
        capi.check(L.fdmb_vplot_use_host(self._h, *[capi.as_dp(a) for a in arrs]), "use")

    def use_device(self, d_u, d_v, d_w):
        capi.check(capi.lib().fdmb_vplot_use_device(self._h, C.c_void_p(d_u), C.c_void_p(d_v), C.c_void_p(d_w)),
                   "use_device")

    def update(self):
        capi.check(capi.lib().fdmb_vplot_update(self._h), "velocity_plotter update")

    def slice(self, name):
        sid = SLICE_IDS[name]
        r, c = C.c_int(), C.c_int()
        capi.check(capi.lib().fdmb_vplot_slice_dims(self._h, sid, C.byref(r), C.byref(c)), "slice_dims")
        out = np.empty((r.value, c.value), dtype=np.float64)
        capi.check(capi.lib().fdmb_vplot_get_slice(self._h, sid, capi.as_dp(out)), "get_slice")
        return out

    def cell_velocity(self):
        """The VECTORS block of vtk_out before formatting: (cells, 3), i = z1..zn, k = y1..yn, j = 1..nx."""
        r, c = C.c_int(), C.c_int()
        capi.check(capi.lib().fdmb_vplot_slice_dims(self._h, SLICE_IDS["RHS_x"], C.byref(r), C.byref(c)), "slice_dims")
        out = np.empty((r.value * c.value * self.params.nx, 3), dtype=np.float64)
        capi.check(capi.lib().fdmb_vplot_cell_velocity(self._h, capi.as_dp(out)), "cell_velocity")
        return out

    def vtk_out(self, name, time_index):
        capi.check(capi.lib().fdmb_vplot_vtk_out(self._h, str(name).encode(), int(time_index)), "vtk_out")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            capi.lib().fdmb_vplot_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
