// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C-ABI veneer over the UNMODIFIED particle-mesh N-body program of the reference (test/nbody.cpp: class
// NBody<T,check,I>, :24-596), the second in-tree consumer of the periodic LaplCube.  The program is a single
// translation unit with its own main(); it is included where it lies (oracle/Makefile adds -I$(REF)/test) with main
// renamed, so that the class can be driven from Python.  cblas_ddot (src/blas.h:3, used once in init_points) is the
// only BLAS symbol it needs; the textbook loop below stands in for the unpinned system BLAS.
#include <cmath>
#include <cstring>

#define main fdm_reference_nbody_main
#include "nbody.cpp"
#undef main

extern "C" double cblas_ddot(int n, const double* x, int incx, const double* y, int incy)
{
    double s = 0;
    for (int i = 0; i < n; i++) s += x[(long)i * incx] * y[(long)i * incy];
    return s;
}
extern "C" float cblas_sdot(int n, const float* x, int incx, const float* y, int incy)
{
    float s = 0;
    for (int i = 0; i < n; i++) s += x[(long)i * incx] * y[(long)i * incy];
    return s;
}

namespace {
using RefNBody = NBody<double, false, CIC3<double>>;
}

extern "C" {

// constructor arguments as in test/nbody.cpp:104 (local = 0: particle-mesh forces only, the program's default :614)
void* ref_nbody_create(double x0, double y0, double z0, double l, int n, int npp, int N, double dt, double G, double vel,
                       int sgn, int solar)
{
    return new RefNBody(x0, y0, z0, l, n, npp, N, dt, G, vel, sgn, 0, solar, 1.0, 0);
}
int ref_nbody_count(void* vh) { return (int)((RefNBody*)vh)->bodies.size(); }
double ref_nbody_total_mass(void* vh) { return ((RefNBody*)vh)->mass; }
// field: 0 x, 1 v, 2 a, 3 aprev -> out[N][3]; 4 mass -> out[N]
void ref_nbody_get(void* vh, int field, double* out)
{
    auto* h = (RefNBody*)vh;
    const size_t N = h->bodies.size();
    for (size_t b = 0; b < N; b++) {
        const auto& B = h->bodies[b];
        if (field == 4) { out[b] = B.mass; continue; }
        const double* src = field == 0 ? B.x : field == 1 ? B.v : field == 2 ? B.a : B.aprev;
        for (int m = 0; m < 3; m++) out[3 * b + m] = src[m];
    }
}
void ref_nbody_step(void* vh, int nsteps)
{
    auto* h = (RefNBody*)vh;
    for (int i = 0; i < nsteps; i++) h->step();
}
// grid: 0 f (density after the deposit), 1 rhs, 2 psi -> out[n^3]; 3 E -> out[n^3][3]
void ref_nbody_get_grid(void* vh, int grid, double* out)
{
    auto* h = (RefNBody*)vh;
    const size_t n3 = (size_t)h->n * h->n * h->n;
    if (grid == 0) std::memcpy(out, h->f.vec, sizeof(double) * n3);
    else if (grid == 1) std::memcpy(out, h->rhs.vec, sizeof(double) * n3);
    else if (grid == 2) std::memcpy(out, h->psi.vec, sizeof(double) * n3);
    else std::memcpy(out, h->E.vec, sizeof(double) * 3 * n3);
}
void ref_nbody_destroy(void* vh) { delete (RefNBody*)vh; }

}  // extern "C"
