// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C-ABI shim over the UNMODIFIED reference classes so that Python tests and
// bench.py's reference arm can drive the real resetius/fdm CPU implementation.
// The reference translation units are compiled where they lie under
// /root/reference/src by oracle/Makefile; nothing from the reference is copied
// into this repository.  Output: oracle/_ref/libfdm_ref.so (git-ignored).
//
// Classes wrapped (reference file:line):
//   fdm::FFT<double>::{sFFT,cFFT,pFFT_1,pFFT}          src/fft.h:47-98, src/fft.cpp
//   fdm::LaplCube<double,false,F>                      src/lapl_cube.h:9-106
//   fdm::LaplRect / LaplRectFFT2<double,false,F>       src/lapl_rect.h:10-105
//   fdm::LaplCyl3FFT2<double,false,zflag>              src/lapl_cyl.h:171-249
//   fdm::NSCube<double,false>                          src/ns_cube.h:13-92
//   fdm::NSCyl<double,false,zflag>                     src/ns_cyl.h:17-132
#include <cmath>   // must precede lapl_cube.h (unqualified sqrt at lapl_cube.h:64)
#include <cstring>
#include <chrono>
#include <string>
#include <vector>
#include <algorithm>

#include "fft.h"
#include "lapl_cube.h"
#include "lapl_rect.h"
#include "lapl_cyl.h"
#include "ns_cube.h"
#include "ns_cyl.h"
#include "config.h"

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace fdm;

namespace {

using F3d = tensor_flags<>;
using F3p = tensor_flags<tensor_flag::periodic, tensor_flag::periodic, tensor_flag::periodic>;
using F2d = tensor_flags<>;
using F2p = tensor_flags<tensor_flag::periodic>;
using F2pp = tensor_flags<tensor_flag::periodic, tensor_flag::periodic>;

struct CubeH {
    int periodic;
    LaplCube<double, false, F3d>* d = nullptr;
    LaplCube<double, false, F3p>* p = nullptr;
};

struct RectH {
    int kind;  // 0 LaplRect, 1 LaplRectFFT2
    int flags; // 0 dirichlet, 1 periodic y, 3 periodic y+x
    LaplRect<double, false, F2d>* r0 = nullptr;
    LaplRect<double, false, F2p>* r1 = nullptr;
    LaplRectFFT2<double, false, F2d>* f0 = nullptr;
    LaplRectFFT2<double, false, F2p>* f1 = nullptr;
    LaplRectFFT2<double, false, F2pp>* f3 = nullptr;
};

struct CylH {
    int zperiodic;
    LaplCyl3FFT2<double, false, tensor_flag::none>* d = nullptr;
    LaplCyl3FFT2<double, false, tensor_flag::periodic>* p = nullptr;
};

struct NSCylH {
    int zperiodic;
    NSCyl<double, false, tensor_flag::none>* d = nullptr;
    NSCyl<double, false, tensor_flag::periodic>* p = nullptr;
};

Config make_config(int nkv, const char** kv)
{
    // kv[i] = "--section:key=value", exactly the reference CLI syntax
    // (src/config.cpp:122-150); argv[0] is skipped by rewrite().
    std::vector<char*> argv;
    argv.push_back(const_cast<char*>("ref"));
    for (int i = 0; i < nkv; i++) argv.push_back(const_cast<char*>(kv[i]));
    Config c;
    c.rewrite((int)argv.size(), argv.data());
    return c;
}

template <typename NS, typename T = double>
int ns_field(NS* ns, int id, T** ptr)
{
    switch (id) {
    case 0: *ptr = ns->u.vec; return ns->u.size;
    case 1: *ptr = ns->v.vec; return ns->v.size;
    case 2: *ptr = ns->w.vec; return ns->w.size;
    case 3: *ptr = ns->p.vec; return ns->p.size;
    case 4: *ptr = ns->x.vec; return ns->x.size;
    case 5: *ptr = ns->F.vec; return ns->F.size;
    case 6: *ptr = ns->G.vec; return ns->G.size;
    case 7: *ptr = ns->H.vec; return ns->H.size;
    case 8: *ptr = ns->RHS.vec; return ns->RHS.size;
    }
    *ptr = nullptr;
    return -1;
}

template <typename NS>
int nscyl_field(NS* ns, int id, double** ptr)
{
    switch (id) {
    case 9: *ptr = ns->u0.vec; return ns->u0.size;
    case 10: *ptr = ns->v0.vec; return ns->v0.size;
    case 11: *ptr = ns->w0.vec; return ns->w0.size;
    }
    return ns_field(ns, id, ptr);
}

} // namespace

extern "C" {

int ref_num_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; bench.py's reference arm calls this before it
// constructs a solver (FFTOmpSafe sizes its pool from omp_get_max_threads() at construction, src/fft.h:382-385)
int ref_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

// ---- 1-D transforms: kind 0 sFFT, 1 cFFT, 2 pFFT_1, 3 pFFT ------------------
// in: N+1 doubles (copied, the reference destroys its input), out: N+1 doubles.
void ref_fft1d(int kind, int N, const double* in, double* out, double dx)
{
    FFTTable<double> table(N);
    FFT<double> ft(table, N);
    std::vector<double> s(in, in + N + 1);
    std::vector<double> S(N + 1, 0.0);
    switch (kind) {
    case 0: ft.sFFT(S.data(), s.data(), dx); break;
    case 1: ft.cFFT(S.data(), s.data(), dx); break;
    case 2: ft.pFFT_1(S.data(), s.data(), dx); break;
    case 3: ft.pFFT(S.data(), s.data(), dx); break;
    }
    std::memcpy(out, S.data(), sizeof(double) * (N + 1));
}

// ---- LaplCube ----------------------------------------------------------------
void* ref_lapl_cube_create(double dx, double dy, double dz, double lx, double ly, double lz,
                           int nx, int ny, int nz, int periodic)
{
    auto* h = new CubeH;
    h->periodic = periodic;
    if (periodic) h->p = new LaplCube<double, false, F3p>(dx, dy, dz, lx, ly, lz, nx, ny, nz);
    else h->d = new LaplCube<double, false, F3d>(dx, dy, dz, lx, ly, lz, nx, ny, nz);
    return h;
}
void ref_lapl_cube_solve(void* vh, double* ans, double* rhs)
{
    auto* h = (CubeH*)vh;
    if (h->periodic) h->p->solve(ans, rhs); else h->d->solve(ans, rhs);
}
void ref_lapl_cube_destroy(void* vh)
{
    auto* h = (CubeH*)vh;
    delete h->p; delete h->d; delete h;
}

// ---- LaplRect / LaplRectFFT2 ---------------------------------------------------
void* ref_lapl_rect_create(int kind, int flags, double dx, double dy, double lx, double ly, int nx, int ny)
{
    auto* h = new RectH;
    h->kind = kind; h->flags = flags;
    if (kind == 0) {
        if (flags == 0) h->r0 = new LaplRect<double, false, F2d>(dx, dy, lx, ly, nx, ny);
        else h->r1 = new LaplRect<double, false, F2p>(dx, dy, lx, ly, nx, ny);
    } else {
        if (flags == 0) h->f0 = new LaplRectFFT2<double, false, F2d>(dx, dy, lx, ly, nx, ny);
        else if (flags == 1) h->f1 = new LaplRectFFT2<double, false, F2p>(dx, dy, lx, ly, nx, ny);
        else h->f3 = new LaplRectFFT2<double, false, F2pp>(dx, dy, lx, ly, nx, ny);
    }
    return h;
}
// scales: nx+1 entries each (index 0 unused), lapl_rect.h:57-59
void ref_lapl_rect_set_scales(void* vh, const double* lm_y_scale, const double* L_scale, const double* U_scale, int n)
{
    auto* h = (RectH*)vh;
    auto set = [&](auto* r) {
        if (!r) return;
        if (lm_y_scale) std::copy(lm_y_scale, lm_y_scale + n, r->lm_y_scale.begin());
        if (L_scale) std::copy(L_scale, L_scale + n, r->L_scale.begin());
        if (U_scale) std::copy(U_scale, U_scale + n, r->U_scale.begin());
    };
    set(h->r0); set(h->r1); set(h->f0); set(h->f1); set(h->f3);
}
void ref_lapl_rect_solve(void* vh, double* ans, double* rhs)
{
    auto* h = (RectH*)vh;
    if (h->r0) h->r0->solve(ans, rhs);
    else if (h->r1) h->r1->solve(ans, rhs);
    else if (h->f0) h->f0->solve(ans, rhs);
    else if (h->f1) h->f1->solve(ans, rhs);
    else if (h->f3) h->f3->solve(ans, rhs);
}
void ref_lapl_rect_destroy(void* vh)
{
    auto* h = (RectH*)vh;
    delete h->r0; delete h->r1; delete h->f0; delete h->f1; delete h->f3; delete h;
}

// ---- LaplCyl3FFT2 --------------------------------------------------------------
void* ref_lapl_cyl_create(double dr, double dz, double r0, double lr, double lz,
                          int nr, int nz, int nphi, int zperiodic)
{
    auto* h = new CylH;
    h->zperiodic = zperiodic;
    if (zperiodic) h->p = new LaplCyl3FFT2<double, false, tensor_flag::periodic>(dr, dz, r0, lr, lz, nr, nz, nphi);
    else h->d = new LaplCyl3FFT2<double, false, tensor_flag::none>(dr, dz, r0, lr, lz, nr, nz, nphi);
    return h;
}
void ref_lapl_cyl_solve(void* vh, double* ans, double* rhs)
{
    auto* h = (CylH*)vh;
    if (h->zperiodic) h->p->solve(ans, rhs); else h->d->solve(ans, rhs);
}
void ref_lapl_cyl_destroy(void* vh)
{
    auto* h = (CylH*)vh;
    delete h->p; delete h->d; delete h;
}

// ---- NSCube --------------------------------------------------------------------
void* ref_ns_cube_create(int nkv, const char** kv)
{
    Config c = make_config(nkv, kv);
    return new NSCube<double, false>(c);
}
void ref_ns_cube_step(void* vh, int nsteps)
{
    auto* ns = (NSCube<double, false>*)vh;
    for (int i = 0; i < nsteps; i++) ns->step();
}
// field ids: 0 u, 1 v, 2 w, 3 p, 4 x, 5 F, 6 G, 7 H, 8 RHS.  Returns element count.
int ref_ns_cube_field_size(void* vh, int id)
{
    double* p; return ns_field((NSCube<double, false>*)vh, id, &p);
}
int ref_ns_cube_get_field(void* vh, int id, double* out)
{
    double* p; int n = ns_field((NSCube<double, false>*)vh, id, &p);
    if (n > 0) std::memcpy(out, p, sizeof(double) * n);
    return n;
}
int ref_ns_cube_set_field(void* vh, int id, const double* in)
{
    double* p; int n = ns_field((NSCube<double, false>*)vh, id, &p);
    if (n > 0) std::memcpy(p, in, sizeof(double) * n);
    return n;
}
void ref_ns_cube_destroy(void* vh) { delete (NSCube<double, false>*)vh; }

// ---- single-precision instantiations (src/lapl_cube.cpp:176-177,181-182; src/ns_cube.cpp:281-282) -------------
struct CubeHF {
    int periodic;
    LaplCube<float, false, F3d>* d = nullptr;
    LaplCube<float, false, F3p>* p = nullptr;
};
void* ref_lapl_cube_f32_create(double dx, double dy, double dz, double lx, double ly, double lz,
                               int nx, int ny, int nz, int periodic)
{
    auto* h = new CubeHF;
    h->periodic = periodic;
    if (periodic) h->p = new LaplCube<float, false, F3p>(dx, dy, dz, lx, ly, lz, nx, ny, nz);
    else h->d = new LaplCube<float, false, F3d>(dx, dy, dz, lx, ly, lz, nx, ny, nz);
    return h;
}
void ref_lapl_cube_f32_solve(void* vh, float* ans, float* rhs)
{
    auto* h = (CubeHF*)vh;
    if (h->periodic) h->p->solve(ans, rhs); else h->d->solve(ans, rhs);
}
void ref_lapl_cube_f32_destroy(void* vh)
{
    auto* h = (CubeHF*)vh;
    delete h->p; delete h->d; delete h;
}
void* ref_ns_cube_f32_create(int nkv, const char** kv)
{
    Config c = make_config(nkv, kv);
    return new NSCube<float, false>(c);
}
void ref_ns_cube_f32_step(void* vh, int nsteps)
{
    auto* ns = (NSCube<float, false>*)vh;
    for (int i = 0; i < nsteps; i++) ns->step();
}
int ref_ns_cube_f32_field_size(void* vh, int id)
{
    float* p; return ns_field((NSCube<float, false>*)vh, id, &p);
}
int ref_ns_cube_f32_get_field(void* vh, int id, float* out)
{
    float* p; int n = ns_field((NSCube<float, false>*)vh, id, &p);
    if (n > 0) std::memcpy(out, p, sizeof(float) * n);
    return n;
}
void ref_ns_cube_f32_destroy(void* vh) { delete (NSCube<float, false>*)vh; }

// ---- NSCyl ---------------------------------------------------------------------
void* ref_ns_cyl_create(int nkv, const char** kv, int zperiodic)
{
    Config c = make_config(nkv, kv);
    auto* h = new NSCylH;
    h->zperiodic = zperiodic;
    if (zperiodic) h->p = new NSCyl<double, false, tensor_flag::periodic>(c);
    else h->d = new NSCyl<double, false, tensor_flag::none>(c);
    return h;
}
void ref_ns_cyl_step(void* vh, int nsteps, int linear)
{
    auto* h = (NSCylH*)vh;
    for (int i = 0; i < nsteps; i++) {
        if (h->zperiodic) { if (linear) h->p->L_step(); else h->p->step(); }
        else { if (linear) h->d->L_step(); else h->d->step(); }
    }
}
// ids as NSCube plus 9 u0, 10 v0, 11 w0
int ref_ns_cyl_field_size(void* vh, int id)
{
    auto* h = (NSCylH*)vh; double* p;
    return h->zperiodic ? nscyl_field(h->p, id, &p) : nscyl_field(h->d, id, &p);
}
int ref_ns_cyl_get_field(void* vh, int id, double* out)
{
    auto* h = (NSCylH*)vh; double* p;
    int n = h->zperiodic ? nscyl_field(h->p, id, &p) : nscyl_field(h->d, id, &p);
    if (n > 0) std::memcpy(out, p, sizeof(double) * n);
    return n;
}
int ref_ns_cyl_set_field(void* vh, int id, const double* in)
{
    auto* h = (NSCylH*)vh; double* p;
    int n = h->zperiodic ? nscyl_field(h->p, id, &p) : nscyl_field(h->d, id, &p);
    if (n > 0) std::memcpy(p, in, sizeof(double) * n);
    return n;
}
// the public member U0 (src/ns_cyl.h:23), changed by test/test_ns_cyl_spectral.cpp between construction and L_step
void ref_ns_cyl_set_u0(void* vh, double u0)
{
    auto* h = (NSCylH*)vh;
    if (h->zperiodic) h->p->U0 = u0; else h->d->U0 = u0;
}
void ref_ns_cyl_destroy(void* vh)
{
    auto* h = (NSCylH*)vh;
    delete h->p; delete h->d; delete h;
}

} // extern "C"
