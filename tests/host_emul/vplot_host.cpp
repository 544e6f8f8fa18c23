// Host emulation of the device-side velocity_plotter kernels: the very functions the CUDA kernels call
// (fdm_b200/csrc/vplot_math.h) run element by element on the CPU.  TEST INFRASTRUCTURE; no GPU, no product path.
#include "../../fdm_b200/csrc/vplot_math.h"

using namespace fdmb;

extern "C" {

long long emul_vplot_field_elems(int nx, int ny, int nz, int zper, int yper, int field)
{
    return vplot_field_elems(vplot_make_geom(nx, ny, nz, zper, yper, 1, 1, 1), field);
}

// dims[2*s], dims[2*s+1] = rows, cols of vx,wx,uy,wy,uz,vz,RHS_x,RHS_y,RHS_z (what velocity_plot.cu allocates)
void emul_vplot_dims(int nx, int ny, int nz, int zper, int yper, int* dims)
{
    const VGeom g = vplot_make_geom(nx, ny, nz, zper, yper, 1, 1, 1);
    const int X2 = nx + 2;
    const int r_[9] = {g.Zc, g.Zc, g.Zc, g.Zc, g.Yc, g.Yc, g.Zi, g.Zi, g.Yi};
    const int c_[9] = {g.Yc, g.Yc, X2, X2, X2, X2, g.Yi, nx, nx};
    for (int s = 0; s < 9; s++) { dims[2 * s] = r_[s]; dims[2 * s + 1] = c_[s]; }
}

void emul_vplot_update(int nx, int ny, int nz, int zper, int yper, double dx, double dy, double dz, const double* u,
                       const double* v, const double* w, double* vx, double* wx, double* uy, double* wy, double* uz,
                       double* vz, double* rx, double* ry, double* rz)
{
    const VGeom g = vplot_make_geom(nx, ny, nz, zper, yper, dx, dy, dz);
    for (long long t = 0; t < vplot_slice_elems(g); t++) vplot_slice_elem(g, t, u, v, w, vx, wx, uy, wy, uz, vz);
    for (long long t = 0; t < vplot_rhs_elems(g); t++) vplot_rhs_elem(g, t, vx, wx, uy, wy, uz, vz, rx, ry, rz);
}

long long emul_vplot_cells(int nx, int ny, int nz, int zper, int yper, const double* u, const double* v, const double* w,
                           double* out)
{
    const VGeom g = vplot_make_geom(nx, ny, nz, zper, yper, 1, 1, 1);
    if (out)
        for (long long t = 0; t < vplot_cell_elems(g); t++) vplot_cell_elem(g, t, u, v, w, out);
    return vplot_cell_elems(g);
}

}  // extern "C"
