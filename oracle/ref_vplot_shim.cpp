// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C-ABI veneer over the UNMODIFIED fdm::velocity_plotter<double,false,F>
// (reference src/velocity_plot.h:11-143, src/velocity_plot.cpp:10-222), compiled where it lies
// by oracle/Makefile.  The slice members are private in the reference (class default access,
// src/velocity_plot.h:12-56); this translation unit alone opens them to read them back.
// matrix_plotter (plplot) is stubbed at the bottom: plot() is never called from here.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "tensor.h"
#include "lapl_rect.h"
#include "matrix_plot.h"
#define class struct   /* default member access of velocity_plotter becomes public; layout and mangling unchanged */
#include "velocity_plot.h"
#undef class

using namespace fdm;

namespace {
using Fd = tensor_flags<>;
using Fp = tensor_flags<tensor_flag::periodic>;
using Fpp = tensor_flags<tensor_flag::periodic, tensor_flag::periodic>;

struct VPlotH {
    int flags;   // 0: <>, 1: <periodic> (z), 3: <periodic,periodic> (z and y)
    velocity_plotter<double, false, Fd>* d = nullptr;
    velocity_plotter<double, false, Fp>* p = nullptr;
    velocity_plotter<double, false, Fpp>* pp = nullptr;
};

template <typename M>
int copy_out(const M& m, double* out)
{
    if (out) std::memcpy(out, m.vec, sizeof(double) * m.size);
    return (int)m.size;
}

template <typename P>
int get_slice(P* q, int id, double* out)
{
    switch (id) {
    case 0: return copy_out(q->vx, out);
    case 1: return copy_out(q->wx, out);
    case 2: return copy_out(q->uy, out);
    case 3: return copy_out(q->wy, out);
    case 4: return copy_out(q->uz, out);
    case 5: return copy_out(q->vz, out);
    case 6: return copy_out(q->RHS_x, out);
    case 7: return copy_out(q->RHS_y, out);
    case 8: return copy_out(q->RHS_z, out);
    case 9: return copy_out(q->psi_x, out);
    case 10: return copy_out(q->psi_y, out);
    case 11: return copy_out(q->psi_z, out);
    }
    return -1;
}
}  // namespace

extern "C" {

void* ref_vplot_create(int flags, double dx, double dy, double dz, int nx, int ny, int nz, double xx1, double xx2,
                       double yy1, double yy2, double zz1, double zz2, int cyl)
{
    auto* h = new VPlotH;
    h->flags = flags;
    if (flags == 0) h->d = new velocity_plotter<double, false, Fd>(dx, dy, dz, nx, ny, nz, xx1, xx2, yy1, yy2, zz1, zz2, cyl);
    else if (flags == 1) h->p = new velocity_plotter<double, false, Fp>(dx, dy, dz, nx, ny, nz, xx1, xx2, yy1, yy2, zz1, zz2, cyl);
    else h->pp = new velocity_plotter<double, false, Fpp>(dx, dy, dz, nx, ny, nz, xx1, xx2, yy1, yy2, zz1, zz2, cyl);
    return h;
}

// use() + update() on host arrays with the reference extents (src/velocity_plot.h:103-105)
void ref_vplot_update(void* vh, double* u, double* v, double* w)
{
    auto* h = (VPlotH*)vh;
    if (h->d) { h->d->use(u, v, w); h->d->update(); }
    else if (h->p) { h->p->use(u, v, w); h->p->update(); }
    else { h->pp->use(u, v, w); h->pp->update(); }
}

// ids: 0 vx 1 wx 2 uy 3 wy 4 uz 5 vz 6 RHS_x 7 RHS_y 8 RHS_z 9 psi_x 10 psi_y 11 psi_z; returns the element count
int ref_vplot_get_slice(void* vh, int id, double* out)
{
    auto* h = (VPlotH*)vh;
    if (h->d) return get_slice(h->d, id, out);
    if (h->p) return get_slice(h->p, id, out);
    return get_slice(h->pp, id, out);
}

void ref_vplot_vtk_out(void* vh, const char* name, int time_index)
{
    auto* h = (VPlotH*)vh;
    if (h->d) h->d->vtk_out(name, time_index);
    else if (h->p) h->p->vtk_out(name, time_index);
    else h->pp->vtk_out(name, time_index);
}

void ref_vplot_destroy(void* vh)
{
    auto* h = (VPlotH*)vh;
    delete h->d; delete h->p; delete h->pp; delete h;
}

}  // extern "C"

// ---- plplot-backed matrix_plotter: link stubs only (src/matrix_plot.h:156-189) -------------------
namespace fdm {
matrix_plotter::matrix_plotter(const settings& s_) : levels(nullptr), s(s_) {}
matrix_plotter::~matrix_plotter() {}
void matrix_plotter::plot_internal(const page&) {}
void matrix_plotter::clear() {}
matrix_plotter::data::~data() { clear(); }
void matrix_plotter::data::clear() {}
}  // namespace fdm
