"""ctypes binding of oracle/_ref/libfdm_ref.so -- TEST INFRASTRUCTURE ONLY.

The library is the UNMODIFIED reference (resetius/fdm) compiled by
oracle/Makefile plus the C shim oracle/ref_shim.cpp.  It is built in the dev
container (where /root/reference exists) and travels to the GPU box as a
prebuilt, git-ignored file.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libfdm_ref.so")
_lib = None

_dp = C.POINTER(C.c_double)


def build(reference="/root/reference", quiet=True):
    """Compile the reference from where it lies; no-op when /root/reference is absent."""
    if not os.path.isdir(os.path.join(reference, "src")):
        return os.path.exists(LIB_PATH)
    subprocess.run(["make", "-C", _HERE, f"REF={reference}", "-j8"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)
    return os.path.exists(LIB_PATH)


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_num_threads.restype = C.c_int
        L.ref_fft1d.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_double]
        L.ref_lapl_cube_create.restype = C.c_void_p
        L.ref_lapl_cube_create.argtypes = [C.c_double] * 6 + [C.c_int] * 4
        L.ref_lapl_cube_solve.argtypes = [C.c_void_p, _dp, _dp]
        L.ref_lapl_cube_destroy.argtypes = [C.c_void_p]
        L.ref_lapl_rect_create.restype = C.c_void_p
        L.ref_lapl_rect_create.argtypes = [C.c_int, C.c_int] + [C.c_double] * 4 + [C.c_int] * 2
        L.ref_lapl_rect_set_scales.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int]
        L.ref_lapl_rect_solve.argtypes = [C.c_void_p, _dp, _dp]
        L.ref_lapl_rect_destroy.argtypes = [C.c_void_p]
        L.ref_lapl_cyl_create.restype = C.c_void_p
        L.ref_lapl_cyl_create.argtypes = [C.c_double] * 5 + [C.c_int] * 4
        L.ref_lapl_cyl_solve.argtypes = [C.c_void_p, _dp, _dp]
        L.ref_lapl_cyl_destroy.argtypes = [C.c_void_p]
        L.ref_ns_cube_create.restype = C.c_void_p
        L.ref_ns_cube_create.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        L.ref_ns_cube_step.argtypes = [C.c_void_p, C.c_int]
        L.ref_ns_cube_field_size.argtypes = [C.c_void_p, C.c_int]
        L.ref_ns_cube_get_field.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_ns_cube_set_field.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_ns_cube_destroy.argtypes = [C.c_void_p]
        L.ref_ns_cyl_create.restype = C.c_void_p
        L.ref_ns_cyl_create.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int]
        L.ref_ns_cyl_step.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_ns_cyl_field_size.argtypes = [C.c_void_p, C.c_int]
        L.ref_ns_cyl_get_field.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_ns_cyl_set_field.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_ns_cyl_destroy.argtypes = [C.c_void_p]
        L.ref_ns_cyl_set_u0.argtypes = [C.c_void_p, C.c_double]
        fp = C.POINTER(C.c_float)
        L.ref_lapl_cube_f32_create.restype = C.c_void_p
        L.ref_lapl_cube_f32_create.argtypes = [C.c_double] * 6 + [C.c_int] * 4
        L.ref_lapl_cube_f32_solve.argtypes = [C.c_void_p, fp, fp]
        L.ref_lapl_cube_f32_destroy.argtypes = [C.c_void_p]
        L.ref_ns_cube_f32_create.restype = C.c_void_p
        L.ref_ns_cube_f32_create.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        L.ref_ns_cube_f32_step.argtypes = [C.c_void_p, C.c_int]
        L.ref_ns_cube_f32_field_size.argtypes = [C.c_void_p, C.c_int]
        L.ref_ns_cube_f32_get_field.argtypes = [C.c_void_p, C.c_int, fp]
        L.ref_ns_cube_f32_destroy.argtypes = [C.c_void_p]
        L.ref_nbody_create.restype = C.c_void_p
        L.ref_nbody_create.argtypes = [C.c_double] * 4 + [C.c_int] * 3 + [C.c_double] * 3 + [C.c_int] * 2
        L.ref_nbody_count.argtypes = [C.c_void_p]
        L.ref_nbody_total_mass.argtypes = [C.c_void_p]
        L.ref_nbody_total_mass.restype = C.c_double
        L.ref_nbody_get.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_nbody_step.argtypes = [C.c_void_p, C.c_int]
        L.ref_nbody_get_grid.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_nbody_destroy.argtypes = [C.c_void_p]
        L.ref_vplot_create.restype = C.c_void_p
        L.ref_vplot_create.argtypes = [C.c_int] + [C.c_double] * 3 + [C.c_int] * 3 + [C.c_double] * 6 + [C.c_int]
        L.ref_vplot_update.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.ref_vplot_get_slice.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_vplot_vtk_out.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.ref_vplot_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def num_threads():
    return lib().ref_num_threads()


def use_all_cores():
    """OpenMP threads = the cores this process may run on, whatever OMP_NUM_THREADS says (torch.distributed.run sets
    it to 1 for its workers).  Call BEFORE constructing a solver; returns the thread count now in force."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    L = lib()
    L.ref_set_num_threads.argtypes = [C.c_int]
    L.ref_set_num_threads.restype = C.c_int
    return L.ref_set_num_threads(int(n))


FFT_KINDS = {"sFFT": 0, "cFFT": 1, "pFFT_1": 2, "pFFT": 3}


def fft1d(kind, s, dx):
    """s: N+1 doubles (index 0..N as the reference's scratch rows).  Returns N+1 doubles."""
    s = np.ascontiguousarray(s, dtype=np.float64)
    N = s.size - 1
    out = np.zeros(N + 1)
    lib().ref_fft1d(FFT_KINDS[kind], N, _p(s), _p(out), float(dx))
    return out


class LaplCube:
    def __init__(self, dx, dy, dz, lx, ly, lz, nx, ny, nz, periodic=False):
        self.shape = (nz, ny, nx)
        self.h = lib().ref_lapl_cube_create(dx, dy, dz, lx, ly, lz, nx, ny, nz, int(periodic))

    def solve(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64).reshape(self.shape)
        ans = np.empty_like(rhs)
        lib().ref_lapl_cube_solve(self.h, _p(ans), _p(rhs))
        return ans

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_lapl_cube_destroy(self.h)
            self.h = None


class LaplCubeF32:
    """The unmodified fdm::LaplCube<float,false,F> (src/lapl_cube.cpp:176-177,181-182)."""

    def __init__(self, dx, dy, dz, lx, ly, lz, nx, ny, nz, periodic=False):
        self.shape = (nz, ny, nx)
        self.h = lib().ref_lapl_cube_f32_create(dx, dy, dz, lx, ly, lz, nx, ny, nz, int(periodic))

    def solve(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=np.float32).reshape(self.shape).copy()
        ans = np.empty_like(rhs)
        fp = C.POINTER(C.c_float)
        lib().ref_lapl_cube_f32_solve(self.h, ans.ctypes.data_as(fp), rhs.ctypes.data_as(fp))
        return ans

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_lapl_cube_f32_destroy(self.h)
            self.h = None


class NSCubeF32:
    """The unmodified fdm::NSCube<float,false> (src/ns_cube.cpp:281-282)."""

    def __init__(self, **params):
        n, arr, self._keep = _kv(params)
        self.h = lib().ref_ns_cube_f32_create(n, arr)

    def step(self, nsteps=1):
        lib().ref_ns_cube_f32_step(self.h, nsteps)

    def field(self, name):
        fid = FIELD_IDS[name]
        n = lib().ref_ns_cube_f32_field_size(self.h, fid)
        out = np.empty(n, dtype=np.float32)
        lib().ref_ns_cube_f32_get_field(self.h, fid, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_ns_cube_f32_destroy(self.h)
            self.h = None


class LaplRect:
    """kind: 'rect' (y transform + tridiagonal x) or 'fft2'.  flags: 0 D/D, 1 y periodic, 3 both."""

    def __init__(self, kind, dx, dy, lx, ly, nx, ny, flags=0):
        self.nx, self.ny = nx, ny
        rows = ny if flags & 1 else ny
        cols = nx
        self.shape = (rows, cols)
        self.h = lib().ref_lapl_rect_create(0 if kind == "rect" else 1, flags, dx, dy, lx, ly, nx, ny)

    def set_scales(self, lm_y_scale=None, L_scale=None, U_scale=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64)
                for a in (lm_y_scale, L_scale, U_scale)]
        n = self.nx + 1
        lib().ref_lapl_rect_set_scales(self.h, *[None if a is None else _p(a) for a in arrs], n)

    def solve(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64).reshape(self.shape)
        ans = np.empty_like(rhs)
        lib().ref_lapl_rect_solve(self.h, _p(ans), _p(rhs))
        return ans

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_lapl_rect_destroy(self.h)
            self.h = None


class LaplCyl3FFT2:
    def __init__(self, dr, dz, r0, lr, lz, nr, nz, nphi, zperiodic=False):
        self.shape = (nphi, nz, nr)
        self.h = lib().ref_lapl_cyl_create(dr, dz, r0, lr, lz, nr, nz, nphi, int(zperiodic))

    def solve(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64).reshape(self.shape)
        ans = np.empty_like(rhs)
        lib().ref_lapl_cyl_solve(self.h, _p(ans), _p(rhs))
        return ans

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_lapl_cyl_destroy(self.h)
            self.h = None


FIELD_IDS = {"u": 0, "v": 1, "w": 2, "p": 3, "x": 4, "F": 5, "G": 6, "H": 7, "RHS": 8,
             "u0": 9, "v0": 10, "w0": 11}


def _kv(params):
    items = [f"--ns:{k}={v!r}".encode() if isinstance(v, float) else f"--ns:{k}={v}".encode()
             for k, v in params.items()]
    arr = (C.c_char_p * len(items))(*items)
    return len(items), arr, items


class NSCube:
    """params: reference config keys of section [ns] (nx, nz, Re, dt, u0, x1..z2)."""

    def __init__(self, **params):
        n, arr, self._keep = _kv(params)
        self.h = lib().ref_ns_cube_create(n, arr)

    def step(self, nsteps=1):
        lib().ref_ns_cube_step(self.h, nsteps)

    def field(self, name):
        fid = FIELD_IDS[name]
        n = lib().ref_ns_cube_field_size(self.h, fid)
        out = np.empty(n)
        lib().ref_ns_cube_get_field(self.h, fid, _p(out))
        return out

    def set_field(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        fid = FIELD_IDS[name]
        assert a.size == lib().ref_ns_cube_field_size(self.h, fid)
        lib().ref_ns_cube_set_field(self.h, fid, _p(a))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_ns_cube_destroy(self.h)
            self.h = None


class NSCyl:
    def __init__(self, zperiodic=False, **params):
        n, arr, self._keep = _kv(params)
        self.h = lib().ref_ns_cyl_create(n, arr, int(zperiodic))

    def step(self, nsteps=1, linear=False):
        lib().ref_ns_cyl_step(self.h, nsteps, int(linear))

    def field(self, name):
        fid = FIELD_IDS[name]
        n = lib().ref_ns_cyl_field_size(self.h, fid)
        out = np.empty(n)
        lib().ref_ns_cyl_get_field(self.h, fid, _p(out))
        return out

    def set_field(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        fid = FIELD_IDS[name]
        assert a.size == lib().ref_ns_cyl_field_size(self.h, fid)
        lib().ref_ns_cyl_set_field(self.h, fid, _p(a))

    def set_u0(self, u0):
        """The public member U0 (src/ns_cyl.h:23)."""
        lib().ref_ns_cyl_set_u0(self.h, float(u0))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_ns_cyl_destroy(self.h)
            self.h = None


SLICE_IDS = {"vx": 0, "wx": 1, "uy": 2, "wy": 3, "uz": 4, "vz": 5, "RHS_x": 6, "RHS_y": 7, "RHS_z": 8,
             "psi_x": 9, "psi_y": 10, "psi_z": 11}


class VelocityPlotter:
    """The unmodified fdm::velocity_plotter<double,false,F> (src/velocity_plot.h, src/velocity_plot.cpp)."""

    def __init__(self, dx, dy, dz, nx, ny, nz, xx1, xx2, yy1, yy2, zz1, zz2, cyl=False, zperiodic=False,
                 yperiodic=False):
        flags = 3 if (zperiodic and yperiodic) else (1 if zperiodic else 0)
        assert not (yperiodic and not zperiodic)
        self.h = lib().ref_vplot_create(flags, dx, dy, dz, nx, ny, nz, xx1, xx2, yy1, yy2, zz1, zz2, int(bool(cyl)))
        self._keep = None

    def update(self, u, v, w):
        """use(u, v, w); update()"""
        self._keep = [np.ascontiguousarray(a, dtype=np.float64).ravel() for a in (u, v, w)]
        lib().ref_vplot_update(self.h, *[_p(a) for a in self._keep])

    def slice(self, name):
        n = lib().ref_vplot_get_slice(self.h, SLICE_IDS[name], None)
        out = np.empty(n)
        lib().ref_vplot_get_slice(self.h, SLICE_IDS[name], _p(out))
        return out

    def vtk_out(self, name, time_index):
        lib().ref_vplot_vtk_out(self.h, str(name).encode(), int(time_index))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_vplot_destroy(self.h)
            self.h = None


class NBody:
    """The unmodified NBody<double,false,CIC3<double>> of test/nbody.cpp (local = 0: particle-mesh forces only)."""

    _BODY = {"x": 0, "v": 1, "a": 2, "aprev": 3, "mass": 4}
    _GRID = {"f": 0, "rhs": 1, "psi": 2, "E": 3}

    def __init__(self, x0=-10.0, y0=-10.0, z0=-10.0, l=20.0, n=32, npp=64, N=1000, dt=0.001, G=1.0, vel=4.0, sgn=-1,
                 solar=0):
        self.n = n
        self.h = lib().ref_nbody_create(x0, y0, z0, l, n, npp, N, dt, G, vel, sgn, solar)
        self.N = lib().ref_nbody_count(self.h)

    def bodies(self, name):
        out = np.empty(self.N if name == "mass" else (self.N, 3))
        lib().ref_nbody_get(self.h, self._BODY[name], _p(out))
        return out

    def total_mass(self):
        return lib().ref_nbody_total_mass(self.h)

    def step(self, nsteps=1):
        lib().ref_nbody_step(self.h, nsteps)      # prints one line per step like the reference (:505-507)

    def grid(self, name):
        n = self.n
        out = np.empty((n, n, n, 3) if name == "E" else (n, n, n))
        lib().ref_nbody_get_grid(self.h, self._GRID[name], _p(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_nbody_destroy(self.h)
            self.h = None
