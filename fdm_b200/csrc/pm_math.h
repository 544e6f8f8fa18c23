// Per-element arithmetic of the particle-mesh N-body step (pm.cu), written so that the same code compiles for the
// device and for the host: tests/host_emul/pm_host.cpp runs it on the CPU against the compiled reference.
// Reference: test/nbody.cpp:257-272 (deposit), :309-316 (right-hand side), :326-341 (field), :434-466 (gather),
// :469-503 (move); cloud-in-cell weights src/interpolate.h:66-100.
#pragma once
#include <cmath>

#include "vplot_math.h"     // FDMB_HD

namespace fdmb {

struct PMGeom {
    int n;                  // cells per axis (periodic)
    long long N;            // bodies
    double h, l;            // cell size, box size
    double ox, oy, oz;      // origin
    double dt, G;
    double rho0;            // -mass / l^3: the mean density subtracted before the deposit (:296-302)
    int deposit_all;        // 0: the reference's cell walk (see pm_deposits), 1: every body
    int lgn;                // log2(n) when n is a power of two (cell index by shifts), else -1
};

inline int pm_log2_or_neg(int n)
{
    if (n <= 0 || (n & (n - 1))) return -1;
    int lg = 0;
    while ((1 << lg) < n) lg++;
    return lg;
}

// Cloud-in-cell: base cell and per-axis weights; M[i][k][j] = (wz[i] * wy[k]) * wx[j] like interpolate.h:88-97.
struct PMCic {
    int j0, k0, i0;
    double wx[2], wy[2], wz[2];
};

FDMB_HD PMCic pm_cic(const PMGeom& g, double x, double y, double z)
{
    PMCic c;
    x -= g.ox; y -= g.oy; z -= g.oz;
    c.j0 = (int)floor(x / g.h); c.k0 = (int)floor(y / g.h); c.i0 = (int)floor(z / g.h);
    x = (x - c.j0 * g.h) / g.h; y = (y - c.k0 * g.h) / g.h; z = (z - c.i0 * g.h) / g.h;
    c.wx[0] = 1 - x; c.wx[1] = x; c.wy[0] = 1 - y; c.wy[1] = y; c.wz[0] = 1 - z; c.wz[1] = z;
    return c;
}

// fdm::tensor periodic wrap (src/tensor.h:119-123); n is a power of two wherever the periodic LaplCube accepts it,
// then the wrap is a mask (two's complement: -1 & (n-1) = n-1)
FDMB_HD int pm_wrap(int i, int n)
{
    if ((n & (n - 1)) == 0) return i & (n - 1);
    i %= n;
    return i < 0 ? i + n : i;
}

// distribute_masses (:257-272) walks the per-cell body lists with ONE offset for all three axes (i, k, j = off,
// off + 2, ...; off = 0, 1), so only bodies whose cell indices are all even or all odd are ever deposited.
FDMB_HD bool pm_deposits(const PMGeom& g, const PMCic& c)
{
    if (g.deposit_all) return true;
    const int pi = pm_wrap(c.i0, g.n) & 1, pk = pm_wrap(c.k0, g.n) & 1, pj = pm_wrap(c.j0, g.n) & 1;
    return pi == pk && pk == pj;
}

// f[i0+i][k0+k][j0+j] += m * M[i][k][j]  (:246-254); Add is an atomic add on the device, a plain += on the host
template <typename Add>
FDMB_HD void pm_deposit_body(const PMGeom& g, double x, double y, double z, double m, double* f, Add add)
{
    const PMCic c = pm_cic(g, x, y, z);
    if (!pm_deposits(g, c)) return;
    for (int i = 0; i < 2; i++)
        for (int k = 0; k < 2; k++)
            for (int j = 0; j < 2; j++) {
                const long long idx = ((long long)pm_wrap(c.i0 + i, g.n) * g.n + pm_wrap(c.k0 + k, g.n)) * g.n + pm_wrap(c.j0 + j, g.n);
                add(f + idx, m * (c.wz[i] * c.wy[k] * c.wx[j]));
            }
}

// rhs = 4 G pi f / h^3  (:312)
FDMB_HD double pm_rhs(const PMGeom& g, double f) { return 4 * g.G * M_PI * f / g.h / g.h / g.h; }

// E = -grad psi with the 4-point rule of Hockney & Eastwood 5.137 (:326-341); E is [n^3][3], component 0 along x
FDMB_HD void pm_field_elem(const PMGeom& g, long long t, const double* psi, double* E)
{
    const int n = g.n;
    int i, k, j;
    if (g.lgn >= 0) { j = (int)(t & (n - 1)); k = (int)((t >> g.lgn) & (n - 1)); i = (int)(t >> (2 * g.lgn)); }
    else { j = (int)(t % n); k = (int)((t / n) % n); i = (int)(t / ((long long)n * n)); }
    const double beta = 4. / 3., h = g.h;
    // the cell's own row / plane offsets once, then one wrapped index per tap
    const long long plane = (long long)n * n, row = (long long)i * plane + (long long)k * n;
    auto X = [&](int jj) { return psi[row + pm_wrap(jj, n)]; };
    auto Y = [&](int kk) { return psi[(long long)i * plane + (long long)pm_wrap(kk, n) * n + j]; };
    auto Z = [&](int ii) { return psi[(long long)pm_wrap(ii, n) * plane + (long long)k * n + j]; };
    E[3 * t + 0] = -beta * (X(j + 1) - X(j - 1)) / 2 / h - (1 - beta) * (X(j + 2) - X(j - 2)) / 4 / h;
    E[3 * t + 1] = -beta * (Y(k + 1) - Y(k - 1)) / 2 / h - (1 - beta) * (Y(k + 2) - Y(k - 2)) / 4 / h;
    E[3 * t + 2] = -beta * (Z(i + 1) - Z(i - 1)) / 2 / h - (1 - beta) * (Z(i + 2) - Z(i - 2)) / 4 / h;
}

// calc_accelerations (:434-466, F = 0 without the local pair forces): a = sum E[cell] * M
FDMB_HD void pm_gather_body(const PMGeom& g, double x, double y, double z, const double* E, double a[3])
{
    const PMCic c = pm_cic(g, x, y, z);
    a[0] = a[1] = a[2] = 0;
    for (int i = 0; i < 2; i++)
        for (int k = 0; k < 2; k++)
            for (int j = 0; j < 2; j++) {
                const long long idx = ((long long)pm_wrap(c.i0 + i, g.n) * g.n + pm_wrap(c.k0 + k, g.n)) * g.n + pm_wrap(c.j0 + j, g.n);
                const double M = c.wz[i] * c.wy[k] * c.wx[j];
                for (int m = 0; m < 3; m++) a[m] += E[3 * idx + m] * M;
            }
}

// move (:469-487): velocity Verlet, positions wrapped into [origin, origin + l)
FDMB_HD void pm_move_body(const PMGeom& g, double x[3], double v[3], const double a[3], double aprev[3])
{
    const double o[3] = {g.ox, g.oy, g.oz};
    for (int m = 0; m < 3; m++) {
        x[m] += g.dt * v[m] + 0.5 * g.dt * g.dt * aprev[m];
        if (x[m] < o[m]) x[m] += g.l;
        if (x[m] >= o[m] + g.l) x[m] -= g.l;
    }
    for (int m = 0; m < 3; m++) {
        v[m] += 0.5 * g.dt * (a[m] + aprev[m]);
        aprev[m] = a[m];
    }
}

}  // namespace fdmb
