"""fdm_b200 -- B200-native fast Poisson / Navier-Stokes projection path of resetius/fdm.

The package is a thin host-side mirror of the reference's class interfaces over the
C ABI in ``include/fdm_b200.h``; all arithmetic runs in hand-written sm_100a kernels
(``fdm_b200/csrc``).  Importing the package does not load the CUDA library; the first
object construction does, and fails loudly if it is missing.
"""
from .capi import FdmB200Error, lib  # noqa: F401
from .lapl_cube import LaplCube, LaplCubeF32, LaplCubeSharded, slab_range  # noqa: F401
from .ns_cube import NSCube, NSCubeF32, owned_planes  # noqa: F401
from .lapl_cyl import LaplCyl3FFT2  # noqa: F401
from .lapl_rect import LaplRect, LaplRectFFT2  # noqa: F401
from .ns_cyl import NSCyl  # noqa: F401
from .velocity_plot import VelocityPlotter  # noqa: F401
from .nbody import NBodyPM  # noqa: F401


def fft_batch(kind, N, data, dx=1.0, impl="auto"):
    """Batched fdm::FFT<double> transforms.  kind: 'sFFT' | 'pFFT_1' | 'pFFT' | 'cFFT' (rows of N-1, N, N, N+1
    values).  impl: 'auto' (the kernel the solvers use for this length), 'plain' or 'pipe'."""
    import numpy as np
    from . import capi
    k = {"sFFT": 0, "pFFT_1": 1, "pFFT": 2, "cFFT": 3}[kind]
    a = np.ascontiguousarray(data, dtype=np.float64)
    nvalid = {0: N - 1, 3: N + 1}.get(k, N)
    if a.ndim == 1:
        a = a[None, :]
    if a.shape[-1] != nvalid:
        raise ValueError(f"rows must hold {nvalid} entries for {kind} with N={N}")
    out = np.empty_like(a)
    batch = a.size // nvalid if nvalid else 0
    capi.check(capi.lib().fdmb_fft_batch_impl(k, int(N), batch, float(dx), capi.as_dp(a), capi.as_dp(out),
                                              {"auto": 0, "plain": 1, "pipe": 2}[impl]), "fft_batch")
    return out.reshape(np.shape(data))
