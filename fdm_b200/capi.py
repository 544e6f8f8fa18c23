"""ctypes binding of libfdm_b200.so (the C ABI declared in include/fdm_b200.h).

There is no CPU fallback: if the shared library is missing or a CUDA call
fails the caller gets an exception, never a silently different code path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfdm_b200.so")
_lib = None

dp = C.POINTER(C.c_double)


class FdmB200Error(RuntimeError):
    pass


def lib():
    """Load the library once.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FdmB200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C fdm_b200/csrc`.  fdm_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.fdmb_last_error.restype = C.c_char_p
    L.fdmb_version.restype = C.c_int
    L.fdmb_device_count.restype = C.c_int
    L.fdmb_set_device.argtypes = [C.c_int]
    L.fdmb_launch_count.restype = C.c_ulonglong
    L.fdmb_malloc.argtypes = [C.POINTER(C.c_void_p), C.c_ulonglong]
    L.fdmb_free.argtypes = [C.c_void_p]
    L.fdmb_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong]
    L.fdmb_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong]
    L.fdmb_fft_batch.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_double, dp, dp]
    L.fdmb_fft_batch_impl.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_double, dp, dp, C.c_int]
    L.fdmb_lapl_cube_create.argtypes = [C.POINTER(C.c_void_p)] + [C.c_double] * 6 + [C.c_int] * 4
    L.fdmb_lapl_cube_solve.argtypes = [C.c_void_p, dp, dp]
    L.fdmb_lapl_cube_solve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fdmb_lapl_cube_solve_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.fdmb_lapl_cube_destroy.argtypes = [C.c_void_p]
    ip = C.POINTER(C.c_int)
    L.fdmb_slab_range.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip, ip]
    L.fdmb_lapl_cube_create_sharded.argtypes = [C.POINTER(C.c_void_p)] + [C.c_double] * 6 + [C.c_int] * 6
    L.fdmb_lapl_cube_local_slab.argtypes = [C.c_void_p, ip, ip]
    L.fdmb_lapl_cube_export_ipc.argtypes = [C.c_void_p, C.c_void_p]
    L.fdmb_lapl_cube_attach_ipc.argtypes = [C.c_void_p, C.c_void_p]
    L.fdmb_lapl_cube_attach_local.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    fp = C.POINTER(C.c_float)
    L.fdmb_lapl_cube_f32_create.argtypes = [C.POINTER(C.c_void_p)] + [C.c_double] * 6 + [C.c_int] * 4
    L.fdmb_lapl_cube_f32_solve.argtypes = [C.c_void_p, fp, fp]
    L.fdmb_lapl_cube_f32_solve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fdmb_lapl_cube_f32_destroy.argtypes = [C.c_void_p]
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().fdmb_last_error()
        raise FdmB200Error(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def as_dp(a):
    return a.ctypes.data_as(dp)
