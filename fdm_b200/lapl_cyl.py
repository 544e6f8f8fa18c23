"""Host-side mirror of fdm::LaplCyl3FFT2 (reference src/lapl_cyl.h:172-249, src/lapl_cyl.cpp:11-170)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def _bind(L):
    if getattr(L, "_lapl_cyl_bound", False):
        return
    L.fdmb_lapl_cyl_create.argtypes = [C.POINTER(C.c_void_p)] + [C.c_double] * 5 + [C.c_int] * 4
    L.fdmb_lapl_cyl_solve.argtypes = [C.c_void_p, capi.dp, capi.dp]
    L.fdmb_lapl_cyl_solve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fdmb_lapl_cyl_destroy.argtypes = [C.c_void_p]
    L.fdmb_lapl_cyl_create_sharded.argtypes = [C.POINTER(C.c_void_p)] + [C.c_double] * 5 + [C.c_int] * 6
    L.fdmb_lapl_cyl_local_slab.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.fdmb_lapl_cyl_export_ipc.argtypes = [C.c_void_p, C.c_void_p]
    L.fdmb_lapl_cyl_attach_ipc.argtypes = [C.c_void_p, C.c_void_p]
    L.fdmb_lapl_cyl_attach_local.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L._lapl_cyl_bound = True


class LaplCyl3FFT2:
    """Cylindrical Poisson solve, arrays ``[nphi][nz][nr]`` (r fastest).

    Mirrors ``fdm::LaplCyl3FFT2<double,check,zflag>(dr,dz,r0,lr,lz,nr,nz,nphi)``;
    ``zperiodic`` selects ``zflag = tensor_flag::periodic`` (src/lapl_cyl.cpp:172-180).
    """

    def __init__(self, dr, dz, r0, lr, lz, nr, nz, nphi, zperiodic=False, rank=0, nranks=1):
        L = capi.lib()
        _bind(L)
        self.nr, self.nz, self.nphi = int(nr), int(nz), int(nphi)
        self.rank, self.nranks = int(rank), int(nranks)
        self._h = C.c_void_p()
        geo = (float(dr), float(dz), float(r0), float(lr), float(lz), self.nr, self.nz, self.nphi, int(bool(zperiodic)))
        if self.nranks > 1:
            capi.check(L.fdmb_lapl_cyl_create_sharded(C.byref(self._h), *geo, self.rank, self.nranks),
                       "LaplCyl3FFT2 create_sharded")
        else:
            capi.check(L.fdmb_lapl_cyl_create(C.byref(self._h), *geo), "LaplCyl3FFT2 create")
        a, b = C.c_int(), C.c_int()
        capi.check(L.fdmb_lapl_cyl_local_slab(self._h, C.byref(a), C.byref(b)), "local_slab")
        self.phi_first, self.nphi_local = a.value, b.value      # this rank's phi-slab (everything on one GPU)

    @property
    def shape(self):
        return (self.nphi_local, self.nz, self.nr)

    # ---- several GPUs: phi-slabs; solve() takes and returns this rank's slab ------------------------
    def connect(self, group=None):
        """One process per GPU: exchange the IPC handles over ``torch.distributed`` and attach the peers."""
        if self.nranks == 1:
            return
        import torch.distributed as dist
        from .lapl_cube import IPC_HANDLE_BYTES
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        capi.check(capi.lib().fdmb_lapl_cyl_export_ipc(self._h, buf), "export_ipc")
        gathered = [None] * self.nranks
        dist.all_gather_object(gathered, buf.raw, group=group)
        blob = b"".join(gathered)
        capi.check(capi.lib().fdmb_lapl_cyl_attach_ipc(self._h, C.create_string_buffer(blob, len(blob))), "attach_ipc")
        dist.barrier(group=group)

    @staticmethod
    def connect_local(parts):
        """All ranks live in this process (one handle per device)."""
        arr = (C.c_void_p * len(parts))(*[s._h for s in parts])
        for s in parts:
            capi.check(capi.lib().fdmb_lapl_cyl_attach_local(s._h, arr), "attach_local")

    def solve(self, ans, rhs=None):
        if rhs is None:
            rhs, ans = ans, None
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        n = self.nr * self.nz * self.nphi_local
        if rhs.size != n:
            raise ValueError(f"rhs has {rhs.size} elements, expected {n}")
        if ans is None:
            ans = np.empty(self.shape, dtype=np.float64)
        if not (isinstance(ans, np.ndarray) and ans.dtype == np.float64 and ans.flags.c_contiguous and ans.size == n):
            raise ValueError("ans must be a C-contiguous float64 array of the same size as rhs")
        capi.check(capi.lib().fdmb_lapl_cyl_solve(self._h, capi.as_dp(ans), capi.as_dp(rhs)), "LaplCyl3FFT2 solve")
        return ans

    def solve_device(self, d_ans, d_rhs, stream=0):
        capi.check(capi.lib().fdmb_lapl_cyl_solve_device(self._h, C.c_void_p(d_ans), C.c_void_p(d_rhs),
                                                        C.c_void_p(stream)), "LaplCyl3FFT2 solve_device")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            capi.lib().fdmb_lapl_cyl_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
