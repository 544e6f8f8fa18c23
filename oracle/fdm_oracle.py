"""CPU restatement of the resetius/fdm hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product
(``fdm_b200``) never does; it fails loudly when its CUDA library is missing.

Parity status: PINNED.  Every function here is checked in
``tests/test_oracle_cpu.py`` (transforms, LaplCube, LaplRect*, LaplCyl3FFT2, NSCube),
``tests/test_oracle_nscyl_cpu.py`` (NSCyl step / L_step), ``tests/test_velocity_plot_cpu.py``
(velocity_plotter), ``tests/test_nbody_cpu.py`` (particle-mesh N-body step) and
``tests/test_reference_kat.py`` against (a) the unmodified reference compiled into
``oracle/_ref/libfdm_ref.so`` (when present), (b) the golden vectors under
``tests/golden`` that were generated from that library by
``tests/golden/make_golden.py``, and (c) the reference's own known-answer
constructions (ut/ut_fft.cpp round trips and O(N^2) definitions,
ut/ut_lapl_cube.cpp analytic solution).

All arithmetic is float64 numpy/scipy.  The 1-D transforms use the closed forms
that the reference's FFTW backend defines unambiguously (src/fft_fftw3.cpp:8-64)
instead of re-tracing the Samarskii-Nikolaev butterflies of src/fft.cpp; the two
agree to ~1e-15 (probed, and asserted by the tests).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.fft as sfft
import scipy.linalg

# --------------------------------------------------------------------------------------
# 1-D transforms (reference: src/fft.cpp, definitions src/asp_fft.cpp:308-319,404-418)
# --------------------------------------------------------------------------------------


def sFFT(s: np.ndarray, dx: float, axis: int = -1) -> np.ndarray:
    """DST-I.  ``s`` holds the N-1 interior values s[1..N-1] along ``axis``.

    S[k] = dx * sum_{j=1}^{N-1} s[j] sin(pi k j / N), k = 1..N-1
    (src/fft.cpp:294-365; FFTW RODFT00 * 0.5, src/fft_fftw3.cpp:45-53).
    """
    return sfft.dst(s, type=1, axis=axis) * (0.5 * dx)


def cFFT(s: np.ndarray, dx: float, axis: int = -1) -> np.ndarray:
    """DCT-I with halved end points over N+1 values s[0..N] (src/fft.cpp:368-445).

    S[k] = dx * (s[0]/2 + sum_{j=1}^{N-1} s[j] cos(pi k j / N) + (-1)^k s[N]/2)
    (src/asp_fft.cpp:404-418; FFTW REDFT00 * 0.5, src/fft_fftw3.cpp:55-64).
    """
    return sfft.dct(s, type=1, axis=axis) * (0.5 * dx)


def pFFT_1(s: np.ndarray, dx: float, axis: int = -1) -> np.ndarray:
    """Periodic forward transform, values -> coefficients (src/fft.cpp:109-193).

    S[k]   = dx * sum_j s[j] cos(2 pi k j / N), k = 0..N/2
    S[N-k] = dx * sum_j s[j] sin(2 pi k j / N), k = 1..N/2-1
    (src/fft_fftw3.cpp:8-22).
    """
    s = np.moveaxis(np.asarray(s, dtype=np.float64), axis, -1)
    N = s.shape[-1]
    X = np.fft.rfft(s, axis=-1)
    out = np.empty_like(s)
    out[..., : N // 2 + 1] = X.real
    if N > 2:
        out[..., N // 2 + 1 :] = (-X.imag[..., 1 : N // 2])[..., ::-1]
    return np.moveaxis(out * dx, -1, axis)


def pFFT(s: np.ndarray, dx: float, axis: int = -1) -> np.ndarray:
    """Periodic inverse transform, coefficients -> values (src/fft.cpp:196-212).

    S[j] = dx * (s[0]/2 + sum_{k=1}^{N/2-1} (s[k] cos(2 pi j k/N) + s[N-k] sin(2 pi j k/N))
                 + (-1)^j s[N/2]/2)            (src/fft_fftw3.cpp:25-42)
    """
    s = np.moveaxis(np.asarray(s, dtype=np.float64), axis, -1)
    N = s.shape[-1]
    X = np.zeros(s.shape[:-1] + (N // 2 + 1,), dtype=np.complex128)
    X.real[...] = s[..., : N // 2 + 1]
    if N > 2:
        X.imag[..., 1 : N // 2] = -(s[..., N // 2 + 1 :][..., ::-1])
    # irfft computes (1/N)(X0 + 2 sum Re(Xk e^{+i...}) + (-1)^j X_{N/2});  we want half of N*that.
    out = np.fft.irfft(X, n=N, axis=-1) * (0.5 * N)
    return np.moveaxis(out * dx, -1, axis)


# --------------------------------------------------------------------------------------
# LaplCube (reference: src/lapl_cube.h:58-100, src/lapl_cube.cpp:9-172)
# --------------------------------------------------------------------------------------


class LaplCube:
    """3-D Poisson solve; all-Dirichlet (periodic=False) or all-periodic."""

    def __init__(self, dx, dy, dz, lx, ly, lz, nx, ny, nz, periodic=False):
        self.dx, self.dy, self.dz = dx, dy, dz
        self.lx, self.ly, self.lz = lx, ly, lz
        self.nx, self.ny, self.nz = nx, ny, nz
        self.periodic = bool(periodic)
        self.slx, self.sly, self.slz = (math.sqrt(2.0 / l) for l in (lx, ly, lz))
        # lapl_cube.h:66-80 (x1/xn/xpoints test zflag, harmless: only all-D / all-P exist)
        if periodic:
            self.z1 = self.y1 = self.x1 = 0
            self.zn, self.yn, self.xn = nz - 1, ny - 1, nx - 1
            self.zpoints, self.ypoints, self.xpoints = nz, ny, nx
        else:
            self.z1 = self.y1 = self.x1 = 1
            self.zn, self.yn, self.xn = nz, ny, nx
            self.zpoints, self.ypoints, self.xpoints = nz + 1, ny + 1, nx + 1
        self._init_lm()

    def _init_lm(self):
        # lapl_cube.cpp:145-172, including the aliasing quirk (:162,:171):
        # lm_x := lm_y when xpoints == ypoints, lm_z := lm_y when zpoints == ypoints,
        # regardless of the spacings.
        def lm(n, d2, lo, hi):
            k = np.arange(lo, hi + 1, dtype=np.float64)
            if self.periodic:
                return 4.0 / d2 * np.sin(k * math.pi / n) ** 2
            return 4.0 / d2 * np.sin(k * math.pi * 0.5 / (n + 1)) ** 2

        self.lm_y = lm(self.ny, self.dy * self.dy, self.y1, self.yn)
        lm_x = lm(self.nx, self.dx * self.dx, self.x1, self.xn)
        lm_z = lm(self.nz, self.dz * self.dz, self.z1, self.zn)
        self.lm_x = self.lm_y if self.xpoints == self.ypoints else lm_x
        self.lm_z = self.lm_y if self.zpoints == self.ypoints else lm_z

    def solve(self, rhs: np.ndarray) -> np.ndarray:
        """rhs: [nz][ny][nx] interior-only array (last index fastest).  Returns ans."""
        a = np.asarray(rhs, dtype=np.float64).reshape(self.nz, self.ny, self.nx)
        fwd = pFFT_1 if self.periodic else sFFT
        inv = pFFT if self.periodic else sFFT
        # forward z, y, x  (lapl_cube.cpp:13-68)
        a = fwd(a, self.dz * self.slz, axis=0)
        a = fwd(a, self.dy * self.sly, axis=1)
        a = fwd(a, self.dx * self.slx, axis=2)
        # divide by -(lm_z + lm_y + lm_x)  (lapl_cube.cpp:70-80)
        k2 = self.lm_z[:, None, None] + self.lm_y[None, :, None] + self.lm_x[None, None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            a = a / (-k2)
        if self.periodic:
            a[0, 0, 0] = 0.0  # lapl_cube.cpp:82-84
        # inverse x, y, z  (lapl_cube.cpp:86-142)
        a = inv(a, self.slx, axis=2)
        a = inv(a, self.sly, axis=1)
        a = inv(a, self.slz, axis=0)
        return np.ascontiguousarray(a)


# --------------------------------------------------------------------------------------
# Tridiagonal solve (reference boundary: LAPACK gtsv / gttrf+gttrs, src/blas.h:65-91)
# --------------------------------------------------------------------------------------


def tridiag_solve(L, D, U, b):
    """Solve the tridiagonal system along the last axis.

    L[..., j] multiplies x[j-1] in row j (L[..., 0] ignored), U[..., j] multiplies
    x[j+1] (U[..., n-1] ignored).  Thomas algorithm without pivoting -- the reference's
    matrices are diagonally dominant (lapl_cyl.cpp:151-159, lapl_rect.cpp:44-60) so
    LAPACK's partial pivoting never interchanges; ut/ut_tdiag.cpp:27-170 asserts
    Thomas == gtsv == gttrs to 1e-15.
    """
    D = np.array(np.broadcast_to(D, b.shape), dtype=np.float64)
    L = np.broadcast_to(L, b.shape)
    U = np.broadcast_to(U, b.shape)
    x = np.array(b, dtype=np.float64)
    n = b.shape[-1]
    for j in range(1, n):
        f = L[..., j] / D[..., j - 1]
        D[..., j] = D[..., j] - f * U[..., j - 1]
        x[..., j] = x[..., j] - f * x[..., j - 1]
    x[..., n - 1] = x[..., n - 1] / D[..., n - 1]
    for j in range(n - 2, -1, -1):
        x[..., j] = (x[..., j] - U[..., j] * x[..., j + 1]) / D[..., j]
    return x


# --------------------------------------------------------------------------------------
# LaplRect / LaplRectFFT2 (reference: src/lapl_rect.h, src/lapl_rect.cpp)
# --------------------------------------------------------------------------------------


class LaplRect:
    """y-transform + tridiagonal in x (lapl_rect.cpp:63-110).  x is always Dirichlet."""

    def __init__(self, dx, dy, lx, ly, nx, ny, yperiodic=False):
        self.dx, self.dy, self.lx, self.ly, self.nx, self.ny = dx, dy, lx, ly, nx, ny
        self.yperiodic = bool(yperiodic)
        self.slx, self.sly = math.sqrt(2.0 / lx), math.sqrt(2.0 / ly)
        self.y1, self.yn = (0, ny - 1) if yperiodic else (1, ny)
        self.ypoints = ny if yperiodic else ny + 1
        k = np.arange(self.y1, self.yn + 1, dtype=np.float64)
        if yperiodic:
            self.lm_y = 4.0 / (dy * dy) * np.sin(k * math.pi / ny) ** 2
        else:
            self.lm_y = 4.0 / (dy * dy) * np.sin(k * math.pi * 0.5 / (ny + 1)) ** 2
        # index 0 unused, 1..nx (lapl_rect.h:57-59)
        self.lm_y_scale = np.ones(nx + 1)
        self.L_scale = np.ones(nx + 1)
        self.U_scale = np.ones(nx + 1)

    def solve(self, rhs):
        rows = self.yn - self.y1 + 1
        a = np.asarray(rhs, dtype=np.float64).reshape(rows, self.nx)
        fwd = pFFT_1 if self.yperiodic else sFFT
        inv = pFFT if self.yperiodic else sFFT
        a = fwd(a, self.dy * self.sly, axis=0)
        dx2 = self.dx * self.dx
        # init_Mat (lapl_rect.cpp:44-60)
        D = -2.0 / dx2 - self.lm_y[:, None] * self.lm_y_scale[None, 1:]
        L = np.broadcast_to(self.L_scale[None, 1:] / dx2, D.shape)
        U = np.broadcast_to(self.U_scale[None, 1:] / dx2, D.shape)
        a = tridiag_solve(L, D, U, a)
        a = inv(a, self.sly, axis=0)
        return np.ascontiguousarray(a)


class LaplRectFFT2(LaplRect):
    """Transforms on both axes (lapl_rect.cpp:113-207)."""

    def __init__(self, dx, dy, lx, ly, nx, ny, yperiodic=False, xperiodic=False):
        super().__init__(dx, dy, lx, ly, nx, ny, yperiodic)
        self.xperiodic = bool(xperiodic)
        self.x1, self.xn = (0, nx - 1) if xperiodic else (1, nx)
        self.xpoints = nx if xperiodic else nx + 1
        j = np.arange(self.x1, self.xn + 1, dtype=np.float64)
        if xperiodic:
            lm_x = 4.0 / (dx * dx) * np.sin(j * math.pi / self.xpoints) ** 2
        else:
            lm_x = 4.0 / (dx * dx) * np.sin(j * math.pi * 0.5 / self.xpoints) ** 2
        # aliasing only in the all-Dirichlet instantiation (lapl_rect.cpp:36-40)
        if not yperiodic and not xperiodic and self.xpoints == self.ypoints:
            self.lm_x = self.lm_y
        else:
            self.lm_x = lm_x

    def solve(self, rhs):
        rows = self.yn - self.y1 + 1
        cols = self.xn - self.x1 + 1
        a = np.asarray(rhs, dtype=np.float64).reshape(rows, cols)
        fy, iy = (pFFT_1, pFFT) if self.yperiodic else (sFFT, sFFT)
        fx, ix = (pFFT_1, pFFT) if self.xperiodic else (sFFT, sFFT)
        a = fy(a, self.dy * self.sly, axis=0)
        a = fx(a, self.dx * self.slx, axis=1)
        scale = self.lm_y_scale[self.x1 : self.xn + 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            a = a / (-self.lm_y[:, None] * scale[None, :] - self.lm_x[None, :])
        if self.y1 == 0 and self.x1 == 0:
            a[0, 0] = 1.0  # "hack for double-period", lapl_rect.cpp:169-172
        a = ix(a, self.slx, axis=1)
        a = iy(a, self.sly, axis=0)
        return np.ascontiguousarray(a)


# --------------------------------------------------------------------------------------
# LaplCyl3FFT2 (reference: src/lapl_cyl.h:12-37,171-249, src/lapl_cyl.cpp:11-170)
# --------------------------------------------------------------------------------------

SQRT_M_1_PI = 0.56418958354775629  # lapl_cyl.h:174


class LaplCyl3FFT2:
    def __init__(self, dr, dz, r0, lr, lz, nr, nz, nphi, zperiodic=False):
        self.dr, self.dz, self.r0, self.lr, self.lz = dr, dz, r0, lr, lz
        self.nr, self.nz, self.nphi = nr, nz, nphi
        self.zperiodic = bool(zperiodic)
        self.dphi = 2 * math.pi / nphi
        self.slz = math.sqrt(2.0 / lz)
        self.zpoints = nz if zperiodic else nz + 1
        self.z1, self.zn = (0, nz - 1) if zperiodic else (1, nz)
        dphi2, dz2 = self.dphi * self.dphi, dz * dz
        i = np.arange(nphi, dtype=np.float64)
        self.lm_phi = 4.0 / dphi2 * np.sin(i * self.dphi * 0.5) ** 2  # lapl_cyl.cpp:132-134
        k = np.arange(self.zpoints, dtype=np.float64)
        if zperiodic:
            self.lm_z = 4.0 / dz2 * np.sin(k * math.pi / self.zpoints) ** 2
        else:
            self.lm_z = 4.0 / dz2 * np.sin(k * math.pi * 0.5 / self.zpoints) ** 2

    def solve(self, rhs):
        nzr = self.zn - self.z1 + 1
        a = np.asarray(rhs, dtype=np.float64).reshape(self.nphi, nzr, self.nr)
        a = pFFT_1(a, self.dphi * SQRT_M_1_PI, axis=0)  # lapl_cyl.cpp:14-29
        if self.zperiodic:
            a = pFFT_1(a, self.dz * self.slz, axis=1)
        else:
            a = sFFT(a, self.dz * self.slz, axis=1)  # lapl_cyl.cpp:31-50
        # tridiagonal in r, lapl_cyl.cpp:151-159
        dr, dr2 = self.dr, self.dr * self.dr
        j = np.arange(1, self.nr + 1, dtype=np.float64)
        r = self.r0 + j * dr
        lmz = self.lm_z[self.z1 : self.zn + 1]
        D = -2.0 / dr2 - self.lm_phi[:, None, None] / r[None, None, :] / r[None, None, :] - lmz[None, :, None]
        # LAPACK layout: L[j-1] is the sub-diagonal of row j (j>1); restated row-wise here
        Lrow = (r - 0.5 * dr) / dr2 / r
        Urow = (r + 0.5 * dr) / dr2 / r
        a = tridiag_solve(Lrow[None, None, :], D, Urow[None, None, :], a)
        if self.zperiodic:
            a = pFFT(a, self.slz, axis=1)
        else:
            a = sFFT(a, self.slz, axis=1)
        a = pFFT(a, SQRT_M_1_PI, axis=0)
        return np.ascontiguousarray(a)


# --------------------------------------------------------------------------------------
# Offset-indexed helper (reference: src/tensor.h -- row-major, last index fastest)
# --------------------------------------------------------------------------------------


class OT:
    """numpy array with per-axis inclusive [lo,hi] index ranges, like fdm::tensor."""

    def __init__(self, ranges):
        self.lo = [r[0] for r in ranges]
        self.hi = [r[1] for r in ranges]
        self.a = np.zeros([h - l + 1 for l, h in ranges], dtype=np.float64)

    def v(self, *rng):
        """View for inclusive reference-index ranges; an int selects one index (kept as len-1)."""
        sl = []
        for ax, r in enumerate(rng):
            if isinstance(r, int):
                r = (r, r)
            sl.append(slice(r[0] - self.lo[ax], r[1] - self.lo[ax] + 1))
        return self.a[tuple(sl)]

    @property
    def size(self):
        return self.a.size


def _sq(x):
    return x * x


# --------------------------------------------------------------------------------------
# NSCube (reference: src/ns_cube.h:46-78, src/ns_cube.cpp:27-277)
# --------------------------------------------------------------------------------------


class NSCube:
    def __init__(self, nx=32, nz=32, Re=1.0, dt=0.001, u0=1.0,
                 x1=-math.pi, y1=-math.pi, z1=-math.pi, x2=math.pi, y2=math.pi, z2=math.pi):
        self.x1, self.y1, self.z1, self.x2, self.y2, self.z2 = x1, y1, z1, x2, y2, z2
        self.U0, self.Re, self.dt = u0, Re, dt
        self.nx = nx
        self.ny = nx  # ns_cube.h:58 -- ny is read from key "nx"
        self.nz = nz
        nx, ny, nz = self.nx, self.ny, self.nz
        self.dx, self.dy, self.dz = (x2 - x1) / nx, (y2 - y1) / ny, (z2 - z1) / nz
        self.dx2, self.dy2, self.dz2 = self.dx**2, self.dy**2, self.dz**2
        self.u = OT([(0, nz + 1), (0, ny + 1), (-1, nx + 1)])
        self.v = OT([(0, nz + 1), (-1, ny + 1), (0, nx + 1)])
        self.w = OT([(-1, nz + 1), (0, ny + 1), (0, nx + 1)])
        self.p = OT([(0, nz + 1), (0, ny + 1), (0, nx + 1)])
        self.x = OT([(1, nz), (1, ny), (1, nx)])
        self.F = OT([(1, nz), (1, ny), (0, nx)])
        self.G = OT([(1, nz), (0, ny), (1, nx)])
        self.H = OT([(0, nz), (1, ny), (1, nx)])
        self.RHS = OT([(1, nz), (1, ny), (1, nx)])
        self.lapl = LaplCube(self.dx, self.dy, self.dz,
                             x2 - x1 + self.dx, y2 - y1 + self.dy, z2 - z1 + self.dz, nx, ny, nz)
        self.time_index = 0

    def fields(self):
        return {"u": self.u.a, "v": self.v.a, "w": self.w.a, "p": self.p.a, "x": self.x.a,
                "F": self.F.a, "G": self.G.a, "H": self.H.a, "RHS": self.RHS.a}

    def step(self):
        self.init_bound()
        self.FGH()
        self.poisson()
        self.update_uvwp()
        self.time_index += 1

    def init_bound(self):  # ns_cube.cpp:65-122, statement order preserved
        nx, ny, nz, U0, Re = self.nx, self.ny, self.nz, self.U0, self.Re
        u, v, w, p = self.u, self.v, self.w, self.p
        # lid: k=0..ny+1, j=-1..nz+1 (sic: nz used for the x extent, ns_cube.cpp:68)
        u.v(nz + 1, (0, ny + 1), (-1, nz + 1))[...] = 2 * U0 - u.v(nz, (0, ny + 1), (-1, nz + 1))
        u.v((0, nz + 1), (0, ny + 1), -1)[...] = u.v((0, nz + 1), (0, ny + 1), 1)
        u.v((0, nz + 1), (0, ny + 1), nx + 1)[...] = u.v((0, nz + 1), (0, ny + 1), nx - 1)
        v.v((0, nz + 1), -1, (0, nx + 1))[...] = v.v((0, nz + 1), 1, (0, nx + 1))
        v.v((0, nz + 1), ny + 1, (0, nx + 1))[...] = v.v((0, nz + 1), ny - 1, (0, nx + 1))
        w.v(-1, (0, ny + 1), (0, nx + 1))[...] = w.v(1, (0, ny + 1), (0, nx + 1))
        w.v(nz + 1, (0, ny + 1), (0, nx + 1))[...] = w.v(nz - 1, (0, ny + 1), (0, nx + 1))
        I, K, J = (1, nz), (1, ny), (1, nx)
        dx, dy, dz = self.dx, self.dy, self.dz
        p.v(I, K, 0)[...] = p.v(I, K, 1) - (u.v(I, K, 1) - 2 * u.v(I, K, 0) + u.v(I, K, -1)) / Re / dx
        p.v(I, K, nx + 1)[...] = p.v(I, K, nx) - (u.v(I, K, nx + 1) - 2 * u.v(I, K, nx) + u.v(I, K, nx - 1)) / Re / dx
        p.v(I, 0, J)[...] = p.v(I, 1, J) - (v.v(I, 1, J) - 2 * v.v(I, 0, J) + v.v(I, -1, J)) / Re / dy
        p.v(I, ny + 1, J)[...] = p.v(I, ny, J) - (v.v(I, ny + 1, J) - 2 * v.v(I, ny, J) + v.v(I, ny - 1, J)) / Re / dy
        p.v(0, K, J)[...] = p.v(1, K, J) - (w.v(1, K, J) - 2 * w.v(0, K, J) + w.v(-1, K, J)) / Re / dz
        p.v(nz + 1, K, J)[...] = p.v(nz, K, J) - (w.v(nz + 1, K, J) - 2 * w.v(nz, K, J) + w.v(nz - 1, K, J)) / Re / dz

    def FGH(self):  # ns_cube.cpp:126-200
        nx, ny, nz, Re, dt = self.nx, self.ny, self.nz, self.Re, self.dt
        dx, dy, dz, dx2, dy2, dz2 = self.dx, self.dy, self.dz, self.dx2, self.dy2, self.dz2
        u, v, w = self.u, self.v, self.w

        def sh(t, I, K, J, di=0, dk=0, dj=0):
            return t.v((I[0] + di, I[1] + di), (K[0] + dk, K[1] + dk), (J[0] + dj, J[1] + dj))

        # F: i=1..nz, k=1..ny, j=0..nx
        I, K, J = (1, nz), (1, ny), (0, nx)
        U = lambda di=0, dk=0, dj=0: sh(u, I, K, J, di, dk, dj)
        V = lambda di=0, dk=0, dj=0: sh(v, I, K, J, di, dk, dj)
        W = lambda di=0, dk=0, dj=0: sh(w, I, K, J, di, dk, dj)
        self.F.a[...] = U() + dt * (
            (U(dj=1) - 2 * U() + U(dj=-1)) / Re / dx2 +
            (U(dk=1) - 2 * U() + U(dk=-1)) / Re / dy2 +
            (U(di=1) - 2 * U() + U(di=-1)) / Re / dz2 -
            (_sq(0.5 * (U() + U(dj=1))) - _sq(0.5 * (U(dj=-1) + U()))) / dx -
            0.25 * ((U() + U(dk=1)) * (V(dj=1) + V()) -
                    (U(dk=-1) + U()) * (V(dk=-1, dj=1) + V(dk=-1))) / dy -
            0.25 * ((U() + U(di=1)) * (W(dj=1) + W()) -
                    (U(di=-1) + U()) * (W(di=-1, dj=1) + W(di=-1))) / dz)
        # G: i=1..nz, k=0..ny, j=1..nx
        I, K, J = (1, nz), (0, ny), (1, nx)
        self.G.a[...] = V() + dt * (
            (V(dj=1) - 2 * V() + V(dj=-1)) / Re / dx2 +
            (V(dk=1) - 2 * V() + V(dk=-1)) / Re / dy2 +
            (V(di=1) - 2 * V() + V(di=-1)) / Re / dz2 -
            (_sq(0.5 * (V() + V(dk=1))) - _sq(0.5 * (V(dk=-1) + V()))) / dy -
            0.25 * ((U() + U(dk=1)) * (V(dj=1) + V()) -
                    (U(dj=-1) + U(dk=1, dj=-1)) * (V() + V(dj=-1))) / dx -
            0.25 * ((W() + W(dk=1)) * (V() + V(di=1)) -
                    (W(di=-1) + W(di=-1, dk=1)) * (V(di=-1) + V())) / dz)
        # H: i=0..nz, k=1..ny, j=1..nx
        I, K, J = (0, nz), (1, ny), (1, nx)
        self.H.a[...] = W() + dt * (
            (W(dj=1) - 2 * W() + W(dj=-1)) / Re / dx2 +
            (W(dk=1) - 2 * W() + W(dk=-1)) / Re / dy2 +
            (W(di=1) - 2 * W() + W(di=-1)) / Re / dz2 -
            (_sq(0.5 * (W(di=1) + W())) - _sq(0.5 * (W(di=-1) + W()))) / dz -
            0.25 * ((U(di=1) + U()) * (W(dj=1) + W()) -
                    (U(di=1, dj=-1) + U(dj=-1)) * (W() + W(dj=-1))) / dx -
            0.25 * ((W() + W(dk=1)) * (V() + V(di=1)) -
                    (W(dk=-1) + W()) * (V(dk=-1) + V(di=1, dk=-1))) / dy)

    def poisson(self):  # ns_cube.cpp:204-238
        nx, ny, nz, dt = self.nx, self.ny, self.nz, self.dt
        dx, dy, dz, dx2, dy2, dz2 = self.dx, self.dy, self.dz, self.dx2, self.dy2, self.dz2
        F, G, H, p, R = self.F, self.G, self.H, self.p, self.RHS
        I, K, J = (1, nz), (1, ny), (1, nx)
        R.a[...] = ((F.v(I, K, J) - F.v(I, K, (0, nx - 1))) / dx +
                    (G.v(I, K, J) - G.v(I, (0, ny - 1), J)) / dy +
                    (H.v(I, K, J) - H.v((0, nz - 1), K, J)) / dz) / dt
        # the reference applies the six corrections in this order per point (:213-233)
        R.v(1, K, J)[...] -= p.v(0, K, J) / dz2
        R.v(I, 1, J)[...] -= p.v(I, 0, J) / dy2
        R.v(I, K, 1)[...] -= p.v(I, K, 0) / dx2
        R.v(I, K, nx)[...] -= p.v(I, K, nx + 1) / dx2
        R.v(I, ny, J)[...] -= p.v(I, ny + 1, J) / dy2
        R.v(nz, K, J)[...] -= p.v(nz + 1, K, J) / dz2
        self.x.a[...] = self.lapl.solve(R.a)

    def update_uvwp(self):  # ns_cube.cpp:241-277
        nx, ny, nz, dt = self.nx, self.ny, self.nz, self.dt
        dx, dy, dz = self.dx, self.dy, self.dz
        u, v, w, p, x, F, G, H = self.u, self.v, self.w, self.p, self.x, self.F, self.G, self.H
        I, K, J = (1, nz), (1, ny), (1, nx)
        Jm, Km, Im = (1, nx - 1), (1, ny - 1), (1, nz - 1)
        u.v(I, K, Jm)[...] = F.v(I, K, Jm) - dt / dx * (x.v(I, K, (2, nx)) - x.v(I, K, Jm))
        v.v(I, Km, J)[...] = G.v(I, Km, J) - dt / dy * (x.v(I, (2, ny), J) - x.v(I, Km, J))
        w.v(Im, K, J)[...] = H.v(Im, K, J) - dt / dz * (x.v((2, nz), K, J) - x.v(Im, K, J))
        p.v(I, K, J)[...] = x.a  # tensor::operator= copies the index-range intersection


# --------------------------------------------------------------------------------------
# NSCyl (reference: src/ns_cyl.h:56-110, src/ns_cyl.cpp:23-484)
# --------------------------------------------------------------------------------------


class PT:
    """[phi][z][r] array with inclusive z / r index ranges; phi always wraps, z wraps when ``zper``
    (fdm::tensor with tensor_flags<periodic, zflag>, src/ns_cyl.h:21-22, src/tensor.h:119-123)."""

    def __init__(self, nphi, zr, rr, zper):
        self.nphi, self.lz, self.lr, self.zper = nphi, zr[0], rr[0], zper
        self.a = np.zeros((nphi, zr[1] - zr[0] + 1, rr[1] - rr[0] + 1), dtype=np.float64)

    def g(self, K, J, di=0, dk=0, dj=0):
        """Values at (i + di, k + dk, j + dj) for all i, k in K = (k0, k1), j in J = (j0, j1)."""
        I = (np.arange(self.nphi) + di) % self.nphi
        Kk = np.arange(K[0], K[1] + 1) + dk - self.lz
        if self.zper:
            Kk %= self.a.shape[1]
        Jj = np.arange(J[0], J[1] + 1) + dj - self.lr
        return self.a[np.ix_(I, Kk, Jj)]

    def put(self, K, J, val):
        self.a[:, K[0] - self.lz:K[1] - self.lz + 1, J[0] - self.lr:J[1] - self.lr + 1] = val


class NSCyl:
    """Flow between two coaxial cylinders, inner one rotating; fields [phi][z][r] with the reference's extents.
    The verify() wall invariants inside init_bound (ns_cyl.cpp:136-163) are not re-checked here."""

    def __init__(self, nr=32, nz=31, nphi=32, Re=1.0, dt=0.001, u0=1.0, R=math.pi, r=math.pi / 2, h1=0.0, h2=10.0,
                 zperiodic=False):
        self.R, self.r0, self.h1, self.h2, self.U0, self.Re, self.dt = R, r, h1, h2, u0, Re, dt
        self.nr, self.nz, self.nphi, self.zp = nr, nz, nphi, bool(zperiodic)
        zp = self.zp
        # ns_cyl.h:70-74
        self.z_, self.z0, self.z1, self.zn, self.znn = (0, 0, 0, nz - 1, nz - 1) if zp else (-1, 0, 1, nz, nz + 1)
        self.dr, self.dz, self.dphi = (R - r) / nr, (h2 - h1) / nz, 2 * math.pi / nphi
        self.dr2, self.dz2, self.dphi2 = self.dr ** 2, self.dz ** 2, self.dphi ** 2
        z_, z0, z1, zn, znn = self.z_, self.z0, self.z1, self.zn, self.znn
        mk = lambda zr, rr: PT(nphi, zr, rr, zp)      # noqa: E731
        # ns_cyl.h:80-93
        self.u, self.v, self.w = mk((z0, znn), (-1, nr + 1)), mk((z_, znn), (0, nr + 1)), mk((z0, znn), (0, nr + 1))
        self.p = mk((z0, znn), (0, nr + 1))
        self.u0, self.v0, self.w0 = mk((z0, znn), (-1, nr + 1)), mk((z_, znn), (0, nr + 1)), mk((z0, znn), (0, nr + 1))
        self.x, self.RHS = mk((z1, zn), (1, nr)), mk((z1, zn), (1, nr))
        self.F, self.G, self.H = mk((z1, zn), (0, nr)), mk((z0, zn), (1, nr)), mk((z1, zn), (1, nr))
        # ns_cyl.h:95-97
        self.lapl = LaplCyl3FFT2(self.dr, self.dz, r - self.dr / 2, R - r + self.dr,
                                 h2 - h1 if zp else h2 - h1 + self.dz, nr, nz, nphi, zperiodic=zp)
        self.time_index = 0

    def fields(self):
        return {k: getattr(self, k).a for k in ("u", "v", "w", "p", "x", "F", "G", "H", "RHS", "u0", "v0", "w0")}

    def field(self, name):
        return self.fields()[name].ravel()

    def set_field(self, name, a):
        t = getattr(self, name).a
        t[...] = np.asarray(a, dtype=np.float64).reshape(t.shape)

    def step(self, nsteps=1, linear=False):
        for _ in range(nsteps):                       # ns_cyl.cpp:23-63 / :66-78
            self.init_bound()
            self.L_FGH() if linear else self.FGH()
            self.poisson()
            self.update_uvwp()
            self.time_index += 1

    def init_bound(self):  # ns_cyl.cpp:81-172, statement order preserved
        nr, nz, U0, Re, dr, dz, r0 = self.nr, self.nz, self.U0, self.Re, self.dr, self.dz, self.r0
        u, v, w, p = self.u, self.v, self.w, self.p
        z_, z0, z1, zn, znn = self.z_, self.z0, self.z1, self.zn, self.znn
        Kw, Kv = (z0, znn), (z_, znn)
        w.put(Kw, (0, 0), 2 * U0 - w.g(Kw, (1, 1)))                      # inner cylinder: 0.5 (w0 + w1) = U0
        w.put(Kw, (nr + 1, nr + 1), -w.g(Kw, (nr, nr)))
        v.put(Kv, (0, 0), -v.g(Kv, (1, 1)))
        v.put(Kv, (nr + 1, nr + 1), -v.g(Kv, (nr, nr)))
        u.put(Kw, (-1, -1), u.g(Kw, (1, 1)))
        u.put(Kw, (nr + 1, nr + 1), u.g(Kw, (nr - 1, nr - 1)))
        if not self.zp:
            # :107-112 loops the r index over z0..znn (sic)
            assert znn <= nr + 1, "the reference walks out of v's r range when nz > nr"
            Jq = (z0, znn)
            v.put((-1, -1), Jq, v.g((1, 1), Jq))
            v.put((nz + 1, nz + 1), Jq, v.g((nz - 1, nz - 1), Jq))
            Ju, Jw = (-1, nr + 1), (0, nr + 1)
            u.put((0, 0), Ju, -u.g((1, 1), Ju))
            u.put((nz + 1, nz + 1), Ju, -u.g((nz, nz), Ju))
            w.put((0, 0), Jw, -w.g((1, 1), Jw))
            w.put((nz + 1, nz + 1), Jw, -w.g((nz, nz), Jw))
        K = (z1, zn)
        r = r0 + 0 * dr - dr / 2
        p.put(K, (0, 0), p.g(K, (1, 1)) - ((r + 0.5 * dr) * u.g(K, (1, 1)) / r - 2 * u.g(K, (0, 0))
                                          + (r - 0.5 * dr) * u.g(K, (-1, -1)) / r) / Re / dr)
        r = r0 + nr * dr - dr / 2
        p.put(K, (nr + 1, nr + 1), p.g(K, (nr, nr)) + ((r + 0.5 * dr) * u.g(K, (nr + 1, nr + 1)) / r - 2 * u.g(K, (nr, nr))
                                                      + (r - 0.5 * dr) * u.g(K, (nr - 1, nr - 1)) / r) / Re / dr)
        if not self.zp:
            J = (1, nr)
            p.put((0, 0), J, p.g((1, 1), J) - (v.g((1, 1), J) - 2 * v.g((0, 0), J) + v.g((-1, -1), J)) / Re / dz)
            p.put((nz + 1, nz + 1), J, p.g((nz, nz), J)
                  + (v.g((nz + 1, nz + 1), J) - 2 * v.g((nz, nz), J) + v.g((nz - 1, nz - 1), J)) / Re / dz)

    def _radii(self, J, staggered):
        j = np.arange(J[0], J[1] + 1, dtype=np.float64)
        r = self.r0 + self.dr * j - (self.dr / 2 if staggered else 0.0)      # :184 / :217
        return r, (r + 0.5 * self.dr) / r, (r - 0.5 * self.dr) / r, r * r

    def FGH(self):  # ns_cyl.cpp:175-277
        nr, Re, dt, dr, dz, dphi = self.nr, self.Re, self.dt, self.dr, self.dz, self.dphi
        dr2, dz2, dphi2 = self.dr2, self.dz2, self.dphi2
        u, v, w = self.u, self.v, self.w
        z0, z1, zn = self.z0, self.z1, self.zn
        # F (r): i = 1..nphi with periodic wrap = every phi (:181), k = z1..zn, j = 0..nr
        K, J = (z1, zn), (0, nr)
        r, r2, r1, rr = self._radii(J, False)
        U = lambda di=0, dk=0, dj=0: u.g(K, J, di, dk, dj)      # noqa: E731
        V = lambda di=0, dk=0, dj=0: v.g(K, J, di, dk, dj)      # noqa: E731
        W = lambda di=0, dk=0, dj=0: w.g(K, J, di, dk, dj)      # noqa: E731
        self.F.put(K, J, U() + dt * (
            (r2 * U(dj=1) - 2 * U() + r1 * U(dj=-1)) / Re / dr2 +
            (U(dk=1) - 2 * U() + U(dk=-1)) / Re / dz2 +
            (U(di=1) - 2 * U() + U(di=-1)) / Re / dphi2 / rr -
            (r2 * _sq(0.5 * (U() + U(dj=1))) - r1 * _sq(0.5 * (U(dj=-1) + U()))) / dr -
            0.25 * ((U() + U(dk=1)) * (V(dj=1) + V()) - (U(dk=-1) + U()) * (V(dk=-1, dj=1) + V(dk=-1))) / dz -
            0.25 * ((U() + U(di=1)) * (W(dj=1) + W()) - (U(di=-1) + U()) * (W(di=-1, dj=1) + W(di=-1))) / dphi / r
            + _sq(0.5 * (W(dj=1) + W())) / r - U() / rr / Re
            - 2 * (0.5 * (W(dj=1) + W()) - 0.5 * (W(di=-1, dj=1) + W(di=-1))) / rr / dphi / Re))
        # G (z): k = z0..zn, j = 1..nr
        K, J = (z0, zn), (1, nr)
        r, r2, r1, rr = self._radii(J, True)
        self.G.put(K, J, V() + dt * (
            (r2 * V(dj=1) - 2 * V() + r1 * V(dj=-1)) / Re / dr2 +
            (V(dk=1) - 2 * V() + V(dk=-1)) / Re / dz2 +
            (V(di=1) - 2 * V() + V(di=-1)) / Re / dphi2 / rr -
            (_sq(0.5 * (V() + V(dk=1))) - _sq(0.5 * (V(dk=-1) + V()))) / dz -
            0.25 * (r2 * (U() + U(dk=1)) * (V(dj=1) + V()) - r1 * (U(dj=-1) + U(dk=1, dj=-1)) * (V() + V(dj=-1))) / dr -
            0.25 * ((W() + W(dk=1)) * (V() + V(di=1)) - (W(di=-1) + W(di=-1, dk=1)) * (V(di=-1) + V())) / dphi / r))
        # H (phi): k = z1..zn, j = 1..nr
        K, J = (z1, zn), (1, nr)
        self.H.put(K, J, W() + dt * (
            (r2 * W(dj=1) - 2 * W() + r1 * W(dj=-1)) / Re / dr2 +
            (W(dk=1) - 2 * W() + W(dk=-1)) / Re / dz2 +
            (W(di=1) - 2 * W() + W(di=-1)) / Re / dphi2 / rr -
            (_sq(0.5 * (W(di=1) + W())) - _sq(0.5 * (W(di=-1) + W()))) / dphi / r -
            0.25 * (r2 * (U(di=1) + U()) * (W(dj=1) + W()) - r1 * (U(di=1, dj=-1) + U(dj=-1)) * (W() + W(dj=-1))) / dr -
            0.25 * ((W() + W(dk=1)) * (V() + V(di=1)) - (W(dk=-1) + W()) * (V(dk=-1) + V(di=1, dk=-1))) / dz
            - W() * 0.5 * (U(di=1) + U()) / r - W() / rr / Re
            + 2 * (0.5 * (U(di=1) + U()) - 0.5 * (U() + U(di=-1))) / rr / dphi / Re))

    def L_FGH(self):  # ns_cyl.cpp:280-405: FGH linearised about u0, v0, w0
        nr, Re, dt, dr, dz, dphi = self.nr, self.Re, self.dt, self.dr, self.dz, self.dphi
        dr2, dz2, dphi2 = self.dr2, self.dz2, self.dphi2
        u, v, w, u0, v0, w0 = self.u, self.v, self.w, self.u0, self.v0, self.w0
        z0, z1, zn = self.z0, self.z1, self.zn
        K, J = (z1, zn), (0, nr)
        r, r2, r1, rr = self._radii(J, False)
        U = lambda di=0, dk=0, dj=0: u.g(K, J, di, dk, dj)        # noqa: E731
        V = lambda di=0, dk=0, dj=0: v.g(K, J, di, dk, dj)        # noqa: E731
        W = lambda di=0, dk=0, dj=0: w.g(K, J, di, dk, dj)        # noqa: E731
        U0 = lambda di=0, dk=0, dj=0: u0.g(K, J, di, dk, dj)      # noqa: E731
        V0 = lambda di=0, dk=0, dj=0: v0.g(K, J, di, dk, dj)      # noqa: E731
        W0 = lambda di=0, dk=0, dj=0: w0.g(K, J, di, dk, dj)      # noqa: E731
        self.F.put(K, J, U() + dt * (
            (r2 * U(dj=1) - 2 * U() + r1 * U(dj=-1)) / Re / dr2 +
            (U(dk=1) - 2 * U() + U(dk=-1)) / Re / dz2 +
            (U(di=1) - 2 * U() + U(di=-1)) / Re / dphi2 / rr -
            (r2 * (0.5 * U() + U(dj=1)) * (U0() + U0(dj=1)) - r1 * (0.5 * U(dj=-1) + U()) * (U0(dj=-1) + U0())) / dr -
            0.25 * ((U() + U(dk=1)) * (V0(dj=1) + V0()) - (U(dk=-1) + U()) * (V0(dk=-1, dj=1) + V0(dk=-1))) / dz -
            0.25 * ((U0() + U0(dk=1)) * (V(dj=1) + V()) - (U0(dk=-1) + U0()) * (V(dk=-1, dj=1) + V(dk=-1))) / dz -
            0.25 * ((U() + U(di=1)) * (W0(dj=1) + W0()) - (U(di=-1) + U()) * (W0(di=-1, dj=1) + W0(di=-1))) / dphi / r -
            0.25 * ((U0() + U0(di=1)) * (W(dj=1) + W()) - (U0(di=-1) + U0()) * (W(di=-1, dj=1) + W(di=-1))) / dphi / r
            + 0.5 * (W(dj=1) + W()) * (W0(dj=1) + W0()) / r
            - U() / rr / Re
            - 2 * (0.5 * (W(dj=1) + W()) - 0.5 * (W(di=-1, dj=1) + W(di=-1))) / rr / dphi / Re))
        K, J = (z0, zn), (1, nr)
        r, r2, r1, rr = self._radii(J, True)
        self.G.put(K, J, V() + dt * (
            (r2 * V(dj=1) - 2 * V() + r1 * V(dj=-1)) / Re / dr2 +
            (V(dk=1) - 2 * V() + V(dk=-1)) / Re / dz2 +
            (V(di=1) - 2 * V() + V(di=-1)) / Re / dphi2 / rr -
            (0.5 * (V() + V(dk=1)) * (V0() + V0(dk=1)) - 0.5 * (V(dk=-1) + V()) * (V0(dk=-1) + V0())) / dz -
            0.25 * (r2 * (U() + U(dk=1)) * (V0(dj=1) + V0()) - r1 * (U(dj=-1) + U(dk=1, dj=-1)) * (V0() + V0(dj=-1))) / dr -
            0.25 * (r2 * (U0() + U0(dk=1)) * (V(dj=1) + V()) - r1 * (U0(dj=-1) + U0(dk=1, dj=-1)) * (V() + V(dj=-1))) / dr -
            0.25 * ((W() + W(dk=1)) * (V0() + V0(di=1)) - (W(di=-1) + W(di=-1, dk=1)) * (V0(di=-1) + V0())) / dphi / r -
            0.25 * ((W0() + W0(dk=1)) * (V() + V(di=1)) - (W0(di=-1) + W0(di=-1, dk=1)) * (V(di=-1) + V())) / dphi / r))
        K, J = (z1, zn), (1, nr)
        self.H.put(K, J, W() + dt * (
            (r2 * W(dj=1) - 2 * W() + r1 * W(dj=-1)) / Re / dr2 +
            (W(dk=1) - 2 * W() + W(dk=-1)) / Re / dz2 +
            (W(di=1) - 2 * W() + W(di=-1)) / Re / dphi2 / rr -
            (0.5 * (W(di=1) + W()) * (W0(di=1) + W0()) - 0.5 * (W(di=-1) + W()) * (W0(di=-1) + W0())) / dphi / r -
            0.25 * (r2 * (U(di=1) + U()) * (W0(dj=1) + W0()) - r1 * (U(di=1, dj=-1) + U(dj=-1)) * (W0() + W0(dj=-1))) / dr -
            0.25 * (r2 * (U0(di=1) + U0()) * (W(dj=1) + W()) - r1 * (U0(di=1, dj=-1) + U0(dj=-1)) * (W() + W(dj=-1))) / dr -
            0.25 * ((W() + W(dk=1)) * (V0() + V0(di=1)) - (W(dk=-1) + W()) * (V0(dk=-1) + V0(di=1, dk=-1))) / dz -
            0.25 * ((W0() + W0(dk=1)) * (V() + V(di=1)) - (W0(dk=-1) + W0()) * (V(dk=-1) + V(di=1, dk=-1))) / dz
            - W0() * 0.5 * (U(di=1) + U()) / r
            - W() * 0.5 * (U0(di=1) + U0()) / r
            - W() / rr / Re
            + 2 * (0.5 * (U(di=1) + U()) - 0.5 * (U() + U(di=-1))) / rr / dphi / Re))

    def poisson(self):  # ns_cyl.cpp:408-442
        nr, nz, dt, dr, dz, dphi, dr2, dz2 = self.nr, self.nz, self.dt, self.dr, self.dz, self.dphi, self.dr2, self.dz2
        F, G, H, p, RHS = self.F, self.G, self.H, self.p, self.RHS
        K, J = (self.z1, self.zn), (1, nr)
        r, _, _, _ = self._radii(J, True)
        R = (((r + 0.5 * dr) * F.g(K, J) - (r - 0.5 * dr) * F.g(K, J, dj=-1)) / r / dr
             + (G.g(K, J) - G.g(K, J, dk=-1)) / dz
             + (H.g(K, J) - H.g(K, J, di=-1)) / dphi / r) / dt
        RHS.put(K, J, R)
        a = RHS.a                                             # [phi][k - z1][j - 1]
        if not self.zp:
            a[:, 0, :] -= p.g((0, 0), J)[:, 0, :] / dz2
        a[:, :, 0] -= (r[0] - dr / 2) / r[0] * p.g(K, (0, 0))[:, :, 0] / dr2
        a[:, :, -1] -= (r[-1] + dr / 2) / r[-1] * p.g(K, (nr + 1, nr + 1))[:, :, 0] / dr2
        if not self.zp:
            a[:, -1, :] -= p.g((nz + 1, nz + 1), J)[:, 0, :] / dz2
        self.x.a[...] = self.lapl.solve(a).reshape(a.shape)

    def update_uvwp(self):  # ns_cyl.cpp:445-484
        nr, nz, dt, dr, dz, dphi = self.nr, self.nz, self.dt, self.dr, self.dz, self.dphi
        u, v, w, p, x, F, G, H = self.u, self.v, self.w, self.p, self.x, self.F, self.G, self.H
        z1, zn = self.z1, self.zn
        K, J = (z1, zn), (1, nr - 1)
        u.put(K, J, F.g(K, J) - dt / dr * (x.g(K, J, dj=1) - x.g(K, J)))
        K, J = (z1, nz - 1), (1, nr)                           # "k < nz /*ok*/": the periodic case wraps x[k+1]
        v.put(K, J, G.g(K, J) - dt / dz * (x.g(K, J, dk=1) - x.g(K, J)))
        K, J = (z1, zn), (1, nr)
        r, _, _, _ = self._radii(J, True)
        w.put(K, J, H.g(K, J) - dt / dphi / r * (x.g(K, J, di=1) - x.g(K, J)))
        p.put(K, J, x.a)                                       # tensor::operator= copies the index-range intersection


# --------------------------------------------------------------------------------------
# Convenience
# --------------------------------------------------------------------------------------


# --------------------------------------------------------------------------------------
# velocity_plotter (reference: src/velocity_plot.h:59-129, src/velocity_plot.cpp:17-67,117-220)
# --------------------------------------------------------------------------------------
class VelocityPlotter:
    """update() and the VECTORS block of vtk_out, restated with numpy index arithmetic.
    Axes as in the reference: z slowest, x fastest; zperiodic / yperiodic <-> F."""

    def __init__(self, dx, dy, dz, nx, ny, nz, xx1, xx2, yy1, yy2, zz1, zz2, cyl=False, zperiodic=False,
                 yperiodic=False):
        self.dx, self.dy, self.dz, self.nx, self.ny, self.nz = dx, dy, dz, nx, ny, nz
        self.zp, self.yp, self.cyl = bool(zperiodic), bool(yperiodic), bool(cyl)
        yp, zp = self.yp, self.zp
        # velocity_plot.h:67-83
        ly = yy2 - yy1 if yp else yy2 - yy1 + dy
        lz = zz2 - zz1 if zp else zz2 - zz1 + dz
        self.y_, self.y1, self.yn, self.ynn = (0, 0, ny - 1, ny - 1) if yp else (-1, 1, ny, ny + 1)
        self.z_, self.z1, self.zn, self.znn = (0, 0, nz - 1, nz - 1) if zp else (-1, 1, nz, nz + 1)
        # velocity_plot.h:101-105
        self.lapl_x = LaplRectFFT2(dy, dz, ly, lz, ny, nz, yperiodic=zp, xperiodic=yp)
        self.lapl_y = LaplRect(dx, dz, xx2 - xx1 + dx, lz, nx, nz, yperiodic=zp)
        self.lapl_z = LaplRect(dx, dy, xx2 - xx1 + dx, ly, nx, ny, yperiodic=yp)
        if cyl:  # velocity_plot.h:113-127
            j = np.arange(nx + 1, dtype=np.float64)
            r = xx1 + j * dx - dx / 2
            one = np.ones(1)
            self.lapl_y.lm_y_scale = np.concatenate([one, (1. / r / r)[1:]])
            self.lapl_y.U_scale = np.concatenate([one, ((r + dx / 2) / r)[1:]])
            self.lapl_y.L_scale = np.concatenate([one, ((r - dx / 2) / r)[1:]])
            self.lapl_z.U_scale = self.lapl_y.U_scale.copy()
            self.lapl_z.L_scale = self.lapl_y.L_scale.copy()

    def shapes(self):
        """Extents of u, v, w (velocity_plot.h:97-99)."""
        Zc, Yc = self.znn + 1, self.ynn + 1
        return ((Zc, Yc, self.nx + 3), (Zc, self.ynn - self.y_ + 1, self.nx + 2),
                (self.znn - self.z_ + 1, Yc, self.nx + 2))

    def _views(self, u, v, w):
        su, sv, sw = self.shapes()
        u, v, w = np.reshape(u, su), np.reshape(v, sv), np.reshape(w, sw)
        nz, ny = self.nz, self.ny
        wz = (lambda i: (i + nz) % nz) if self.zp else (lambda i: i)
        wy = (lambda k: (k + ny) % ny) if self.yp else (lambda k: k)
        U = lambda i, k, j: u[i, k, j + 1]                   # noqa: E731
        V = lambda i, k, j: v[i, wy(k) - self.y_, j]         # noqa: E731
        W = lambda i, k, j: w[wz(i) - self.z_, k, j]         # noqa: E731
        return U, V, W, wz, wy

    def update(self, u, v, w):
        U, V, W, wz, wy = self._views(u, v, w)
        nx, ny, nz, dx, dy, dz = self.nx, self.ny, self.nz, self.dx, self.dy, self.dz
        I = np.arange(0, self.znn + 1)[:, None]
        K = np.arange(0, self.ynn + 1)
        J = np.arange(0, nx + 2)
        out = {}
        out["vx"] = 0.5 * (V(I, K[None, :] - 1, nx // 2) + V(I, K[None, :], nx // 2))        # :19-24
        out["wx"] = 0.5 * (W(I - 1, K[None, :], nx // 2) + W(I, K[None, :], nx // 2))
        out["uy"] = 0.5 * (U(I, ny // 2, J[None, :] - 1) + U(I, ny // 2, J[None, :]))        # :26-31
        out["wy"] = 0.5 * (W(I - 1, ny // 2, J[None, :]) + W(I, ny // 2, J[None, :]))
        Kc = K[:, None]
        out["uz"] = 0.5 * (U(nz // 2, Kc, J[None, :] - 1) + U(nz // 2, Kc, J[None, :]))      # :33-38
        out["vz"] = 0.5 * (V(nz // 2, Kc - 1, J[None, :]) + V(nz // 2, Kc, J[None, :]))
        Ii = np.arange(self.z1, self.zn + 1)[:, None]
        Ki = np.arange(self.y1, self.yn + 1)
        Ji = np.arange(1, nx + 1)[None, :]
        vx, wx, uy, wy_, uz, vz = (out[k] for k in ("vx", "wx", "uy", "wy", "uz", "vz"))
        out["RHS_x"] = ((wx[Ii, wy(Ki + 1)[None, :]] - wx[Ii, wy(Ki - 1)[None, :]]) / 2 / dy
                        - (vx[wz(Ii + 1), Ki[None, :]] - vx[wz(Ii - 1), Ki[None, :]]) / 2 / dz)   # :40-44
        out["RHS_y"] = ((wy_[Ii, Ji + 1] - wy_[Ii, Ji - 1]) / 2 / dx
                        - (uy[wz(Ii + 1), Ji] - uy[wz(Ii - 1), Ji]) / 2 / dz)                     # :48-52
        Kic = Ki[:, None]
        out["RHS_z"] = ((vz[Kic, Ji + 1] - vz[Kic, Ji - 1]) / 2 / dx
                        - (uz[wy(Kic + 1), Ji] - uz[wy(Kic - 1), Ji]) / 2 / dy)                   # :56-60
        out["psi_x"] = self.lapl_x.solve(out["RHS_x"])
        out["psi_y"] = self.lapl_y.solve(out["RHS_y"])
        out["psi_z"] = self.lapl_z.solve(out["RHS_z"])
        return out

    def cell_velocity(self, u, v, w):
        """(cells, 3): the face averages vtk_out prints (or rotates, for cylinders) -- :180-182, :209-213."""
        U, V, W, wz, wy = self._views(u, v, w)
        I = np.arange(self.z1, self.zn + 1)[:, None, None]
        K = np.arange(self.y1, self.yn + 1)[None, :, None]
        J = np.arange(1, self.nx + 1)[None, None, :]
        c = np.stack([0.5 * (U(I, K, J) + U(I, K, J - 1)), 0.5 * (V(I, K, J) + V(I, K - 1, J)),
                      0.5 * (W(I, K, J) + W(I - 1, K, J))], axis=-1)
        return c.reshape(-1, 3)


# --------------------------------------------------------------------------------------
# Particle-mesh N-body step (reference: test/nbody.cpp:24-596 with local = 0, src/interpolate.h:66-100)
# --------------------------------------------------------------------------------------
class NBodyPM:
    """One step = calc_a_pm() (:292-421: CIC deposit on a periodic n^3 grid, rhs = 4 pi G f / h^3, periodic LaplCube,
    4-point field differencing, CIC gather) followed by move() (:469-503, velocity Verlet with periodic wrap).
    Bodies are given by the caller (the reference seeds them from std::default_random_engine, :541-587)."""

    def __init__(self, x0, y0, z0, l, n, dt, G, deposit_all=False):
        self.origin = np.array([x0, y0, z0], dtype=np.float64)
        self.l, self.n, self.h, self.dt, self.G = float(l), int(n), l / n, float(dt), float(G)
        self.deposit_all = bool(deposit_all)      # False = the reference's behaviour
        h = self.h
        self.solver = LaplCube(h, h, h, l, l, l, n, n, n, periodic=True)

    def set_bodies(self, x, v, mass):
        self.x, self.v = np.array(x, dtype=np.float64), np.array(v, dtype=np.float64)
        self.m = np.array(mass, dtype=np.float64)
        self.a, self.aprev = np.zeros_like(self.x), np.zeros_like(self.x)
        self.mass = 0.0
        for mb in self.m:          # sequential like :583
            self.mass += mb

    def _cic(self):
        """interpolate.h:66-100: base cell (j0,k0,i0) and the 2x2x2 weights M[i][k][j]; x[0] <-> j (fastest axis)."""
        h = self.h
        rel = self.x - self.origin
        base = np.floor(rel / h).astype(np.int64)
        fr = (rel - base * h) / h
        fx, fy, fz = fr[:, 0], fr[:, 1], fr[:, 2]
        wx, wy, wz = np.stack([1 - fx, fx], 1), np.stack([1 - fy, fy], 1), np.stack([1 - fz, fz], 1)
        M = wz[:, :, None, None] * wy[:, None, :, None] * wx[:, None, None, :]      # [body][i][k][j]
        d = np.arange(2)
        n = self.n
        I = (base[:, 2, None] + d)[:, :, None, None] % n
        K = (base[:, 1, None] + d)[:, None, :, None] % n
        J = (base[:, 0, None] + d)[:, None, None, :] % n
        I, K, J = np.broadcast_arrays(I, K, J)
        return I, K, J, M

    def calc_a_pm(self):
        n, h, l = self.n, self.h, self.l
        I, K, J, M = self._cic()
        self.f = np.full((n, n, n), -self.mass / l / l / l)
        # distribute_masses (:257-272) walks the cell lists with ONE offset for all three axes
        # (i = off, off+2, ...; k = off, ...; j = off, ...; off = 0, 1): only bodies whose cell indices are all even
        # or all odd are deposited -- a quarter of them.  Reproduced, not fixed (the mean density still uses all).
        i0, k0, j0 = I[:, 0, 0, 0], K[:, 0, 0, 0], J[:, 0, 0, 0]
        dep = ((i0 % 2 == k0 % 2) & (k0 % 2 == j0 % 2)) | self.deposit_all
        np.add.at(self.f, (I[dep], K[dep], J[dep]), self.m[dep, None, None, None] * M[dep])
        self.rhs = 4 * self.G * math.pi * self.f / h / h / h
        self.psi = self.solver.solve(self.rhs).reshape(n, n, n)
        psi, beta = self.psi, 4. / 3.
        E = np.empty((n, n, n, 3))
        for m, ax in enumerate((2, 1, 0)):          # E[..][0] differences along x (last axis), :333-339
            E[..., m] = (-beta * (np.roll(psi, -1, ax) - np.roll(psi, 1, ax)) / 2 / h
                         - (1 - beta) * (np.roll(psi, -2, ax) - np.roll(psi, 2, ax)) / 4 / h)
        self.E = E
        self.a = np.einsum("bikjm,bikj->bm", E[I, K, J], M)       # :434-466 with F = 0

    def move(self):
        dt, l, o = self.dt, self.l, self.origin
        self.x = self.x + (dt * self.v + 0.5 * dt * dt * self.aprev)
        self.x = np.where(self.x < o, self.x + l, self.x)
        self.x = np.where(self.x >= o + l, self.x - l, self.x)
        self.v = self.v + 0.5 * dt * (self.a + self.aprev)
        self.aprev = self.a.copy()

    def step(self, nsteps=1):
        for _ in range(nsteps):
            self.calc_a_pm()
            self.move()


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2 with a 0/0 guard (both exactly zero -> 0)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    nb = float(np.linalg.norm(b))
    nd = float(np.linalg.norm(a - b))
    if nb == 0.0:
        return 0.0 if nd == 0.0 else float("inf")
    return nd / nb


def synthetic_rhs(shape, seed=1234):
    """Counter-based synthetic RHS, uniform(-0.5, 0.5), identical on every box."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    return rng.random(shape, dtype=np.float64) - 0.5
