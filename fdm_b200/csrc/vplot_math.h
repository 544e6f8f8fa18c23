// Per-element arithmetic of the device-side velocity_plotter (velocity_plot.cu), written so that the same code
// compiles for the device and for the host: tests/host_emul/vplot_host.cpp runs it element by element on the CPU
// and compares it with the compiled reference, which checks every index expression without a GPU.
// Reference: src/velocity_plot.h:73-99 (index ranges, array extents), src/velocity_plot.cpp:19-64 (update),
// :180-182,209-213 (cell-centred velocities of vtk_out).
#pragma once

#if defined(__CUDACC__)
#define FDMB_HD __host__ __device__ __forceinline__
#else
#define FDMB_HD inline
#endif

namespace fdmb {

struct VGeom {
    int nx, ny, nz;
    int zper, yper;
    int y_, y1, yn, ynn;    // y0 = z0 = 0
    int z_, z1, zn, znn;
    int Yc, Zc;             // points in y0..ynn, z0..znn
    int Yv;                 // points in y_..ynn (v)
    int Yi, Zi;             // points in y1..yn, z1..zn
    double dx, dy, dz;
};

// src/velocity_plot.h:73-83
inline VGeom vplot_make_geom(int nx, int ny, int nz, int zperiodic, int yperiodic, double dx, double dy, double dz)
{
    VGeom g{};
    g.nx = nx; g.ny = ny; g.nz = nz; g.zper = zperiodic ? 1 : 0; g.yper = yperiodic ? 1 : 0;
    g.dx = dx; g.dy = dy; g.dz = dz;
    g.y_ = g.yper ? 0 : -1; g.y1 = g.yper ? 0 : 1; g.yn = g.yper ? ny - 1 : ny; g.ynn = g.yper ? ny - 1 : ny + 1;
    g.z_ = g.zper ? 0 : -1; g.z1 = g.zper ? 0 : 1; g.zn = g.zper ? nz - 1 : nz; g.znn = g.zper ? nz - 1 : nz + 1;
    g.Yc = g.ynn + 1; g.Zc = g.znn + 1; g.Yv = g.ynn - g.y_ + 1;
    g.Yi = g.yn - g.y1 + 1; g.Zi = g.zn - g.z1 + 1;
    return g;
}

// elements of u, v, w: u[z0..znn][y0..ynn][-1..nx+1], v[z0..znn][y_..ynn][0..nx+1], w[z_..znn][y0..ynn][0..nx+1]
inline long long vplot_field_elems(const VGeom& g, int field)
{
    if (field == 0) return (long long)g.Zc * g.Yc * (g.nx + 3);
    if (field == 1) return (long long)g.Zc * g.Yv * (g.nx + 2);
    return (long long)(g.znn - g.z_ + 1) * g.Yc * (g.nx + 2);
}

FDMB_HD long long vplot_slice_elems(const VGeom& g) { return (long long)g.Zc * g.Yc + (long long)(g.Zc + g.Yc) * (g.nx + 2); }
FDMB_HD long long vplot_rhs_elems(const VGeom& g) { return (long long)g.Zi * g.Yi + (long long)(g.Zi + g.Yi) * g.nx; }
FDMB_HD long long vplot_cell_elems(const VGeom& g) { return (long long)g.Zi * g.Yi * g.nx; }

// periodic axes wrap like fdm::tensor (src/tensor.h:119-123)
FDMB_HD int vplot_wrap_z(const VGeom& g, int i) { return g.zper ? (i + g.nz) % g.nz : i; }
FDMB_HD int vplot_wrap_y(const VGeom& g, int k) { return g.yper ? (k + g.ny) % g.ny : k; }
FDMB_HD long long vplot_iu(const VGeom& g, int i, int k, int j) { return ((long long)i * g.Yc + k) * (g.nx + 3) + (j + 1); }
FDMB_HD long long vplot_iv(const VGeom& g, int i, int k, int j) { return ((long long)i * g.Yv + (k - g.y_)) * (g.nx + 2) + j; }
FDMB_HD long long vplot_iw(const VGeom& g, int i, int k, int j) { return ((long long)(i - g.z_) * g.Yc + k) * (g.nx + 2) + j; }

// src/velocity_plot.cpp:19-38: face averages on the planes x = nx/2, y = ny/2, z = nz/2.
// Element t of the concatenation [vx|wx pairs (Zc x Yc)] [uy|wy pairs (Zc x nx+2)] [uz|vz pairs (Yc x nx+2)].
FDMB_HD void vplot_slice_elem(const VGeom& g, long long t, const double* u, const double* v, const double* w, double* vx,
                              double* wx, double* uy, double* wy, double* uz, double* vz)
{
    const int X2 = g.nx + 2;
    const long long n0 = (long long)g.Zc * g.Yc, n1 = (long long)g.Zc * X2;
    const int jm = g.nx / 2, km = g.ny / 2, im = g.nz / 2;
    if (t < n0) {
        const int i = (int)(t / g.Yc), k = (int)(t % g.Yc);
        vx[t] = 0.5 * (v[vplot_iv(g, i, vplot_wrap_y(g, k - 1), jm)] + v[vplot_iv(g, i, k, jm)]);
        wx[t] = 0.5 * (w[vplot_iw(g, vplot_wrap_z(g, i - 1), k, jm)] + w[vplot_iw(g, i, k, jm)]);
    } else if (t < n0 + n1) {
        const long long s = t - n0;
        const int i = (int)(s / X2), j = (int)(s % X2);
        uy[s] = 0.5 * (u[vplot_iu(g, i, km, j - 1)] + u[vplot_iu(g, i, km, j)]);
        wy[s] = 0.5 * (w[vplot_iw(g, vplot_wrap_z(g, i - 1), km, j)] + w[vplot_iw(g, i, km, j)]);
    } else {
        const long long s = t - n0 - n1;
        const int k = (int)(s / X2), j = (int)(s % X2);
        uz[s] = 0.5 * (u[vplot_iu(g, im, k, j - 1)] + u[vplot_iu(g, im, k, j)]);
        vz[s] = 0.5 * (v[vplot_iv(g, im, vplot_wrap_y(g, k - 1), j)] + v[vplot_iv(g, im, k, j)]);
    }
}

// src/velocity_plot.cpp:40-64: centred differences of the slices (periodic axes wrap inside the slice).
// No multiply-add pair anywhere: the result is bit-identical to the reference whatever the contraction setting.
FDMB_HD void vplot_rhs_elem(const VGeom& g, long long t, const double* vx, const double* wx, const double* uy,
                            const double* wy, const double* uz, const double* vz, double* rx, double* ry, double* rz)
{
    const int X2 = g.nx + 2;
    const long long n0 = (long long)g.Zi * g.Yi, n1 = (long long)g.Zi * g.nx;
    if (t < n0) {
        const int i = (int)(t / g.Yi) + g.z1, k = (int)(t % g.Yi) + g.y1;
        const int kp = vplot_wrap_y(g, k + 1), kmn = vplot_wrap_y(g, k - 1);
        const int ip = vplot_wrap_z(g, i + 1), imn = vplot_wrap_z(g, i - 1);
        const double a = wx[(long long)i * g.Yc + kp] - wx[(long long)i * g.Yc + kmn];
        const double b = vx[(long long)ip * g.Yc + k] - vx[(long long)imn * g.Yc + k];
        rx[t] = a / 2 / g.dy - b / 2 / g.dz;
    } else if (t < n0 + n1) {
        const long long s = t - n0;
        const int i = (int)(s / g.nx) + g.z1, j = (int)(s % g.nx) + 1;
        const int ip = vplot_wrap_z(g, i + 1), imn = vplot_wrap_z(g, i - 1);
        const double a = wy[(long long)i * X2 + j + 1] - wy[(long long)i * X2 + j - 1];
        const double b = uy[(long long)ip * X2 + j] - uy[(long long)imn * X2 + j];
        ry[s] = a / 2 / g.dx - b / 2 / g.dz;
    } else {
        const long long s = t - n0 - n1;
        const int k = (int)(s / g.nx) + g.y1, j = (int)(s % g.nx) + 1;
        const int kp = vplot_wrap_y(g, k + 1), kmn = vplot_wrap_y(g, k - 1);
        const double a = vz[(long long)k * X2 + j + 1] - vz[(long long)k * X2 + j - 1];
        const double b = uz[(long long)kp * X2 + j] - uz[(long long)kmn * X2 + j];
        rz[s] = a / 2 / g.dx - b / 2 / g.dy;
    }
}

// src/velocity_plot.cpp:180-182,209-213: cell-centred velocity = average of the two faces, i=z1..zn, k=y1..yn, j=1..nx
FDMB_HD void vplot_cell_elem(const VGeom& g, long long t, const double* u, const double* v, const double* w, double* out)
{
    const int j = (int)(t % g.nx) + 1;
    const long long r = t / g.nx;
    const int k = (int)(r % g.Yi) + g.y1, i = (int)(r / g.Yi) + g.z1;
    out[3 * t + 0] = 0.5 * (u[vplot_iu(g, i, k, j)] + u[vplot_iu(g, i, k, j - 1)]);
    out[3 * t + 1] = 0.5 * (v[vplot_iv(g, i, k, j)] + v[vplot_iv(g, i, vplot_wrap_y(g, k - 1), j)]);
    out[3 * t + 2] = 0.5 * (w[vplot_iw(g, i, k, j)] + w[vplot_iw(g, vplot_wrap_z(g, i - 1), k, j)]);
}

}  // namespace fdmb
