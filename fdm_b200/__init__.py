"""fdm_b200 -- B200-native fast Poisson / Navier-Stokes projection path of resetius/fdm.

The package is a thin host-side mirror of the reference's class interfaces over the
C ABI in ``include/fdm_b200.h``; all arithmetic runs in hand-written sm_100a kernels
(``fdm_b200/csrc``).  Importing the package does not load the CUDA library; the first
object construction does, and fails loudly if it is missing.
"""
from .capi import FdmB200Error, lib  # noqa: F401
from .lapl_cube import LaplCube, LaplCubeSharded, slab_range  # noqa: F401
from .ns_cube import NSCube, owned_planes  # noqa: F401
from .lapl_cyl import LaplCyl3FFT2  # noqa: F401
from .lapl_rect import LaplRect, LaplRectFFT2  # noqa: F401
from .ns_cyl import NSCyl  # noqa: F401
from .velocity_plot import VelocityPlotter  # noqa: F401
from .nbody import NBodyPM  # noqa: F401


def fft_batch(kind, N, data, dx=1.0):
    """Batched fdm::FFT<double> transforms.  kind: 'sFFT' | 'pFFT_1' | 'pFFT'."""
    import numpy as np
    from . import capi
    k = {"sFFT": 0, "pFFT_1": 1, "pFFT": 2}[kind]
    a = np.ascontiguousarray(data, dtype=np.float64)
    nvalid = N - 1 if k == 0 else N
    if a.ndim == 1:
        a = a[None, :]
    if a.shape[-1] != nvalid:
        raise ValueError(f"rows must hold {nvalid} entries for {kind} with N={N}")
    out = np.empty_like(a)
    batch = a.size // nvalid if nvalid else 0
    capi.check(capi.lib().fdmb_fft_batch(k, int(N), batch, float(dx), capi.as_dp(a), capi.as_dp(out)), "fft_batch")
    return out.reshape(np.shape(data))
