// NSCube on B200: incompressible Navier-Stokes in a box on a staggered (MAC) grid,
// explicit Euler predictor + pressure projection.  Replaces fdm::NSCube<double,check>
// (reference src/ns_cube.h:13-92, src/ns_cube.cpp:27-277).  The whole step is device
// resident; the state arrays keep the reference's extents and layout (ghosts included)
// so get/set_field are plain copies of what the reference exposes as ns.u.vec etc.
//
// Multi-GPU: the same z-slabs as the sharded LaplCube.  Every rank keeps WINDOWS of the global
// arrays (its own planes plus the halo planes its stencils read), all kernels index with GLOBAL (i,k,j), and the
// halo planes are pulled from the neighbours' memory over NVLink (peer loads) after a cross-GPU barrier:
//   barrier -> pull u,v,w halos -> init_bound -> FGH (H also on the plane below the slab, recomputed instead of
//   exchanged) -> RHS -> sharded solve -> barrier -> pull the x plane above the slab -> update
//
//   step() = init_bound (ns_cube.cpp:65-122)  -> k_bound_all (single precision: k_bound_lid, k_bound_mirror, k_bound_p)
//            FGH        (ns_cube.cpp:126-200) -> k_fgh_div, which also forms the divergence (FDMB_FGH_FUSED=0: k_fgh)
//            poisson    (ns_cube.cpp:204-238) -> (k_rhs +) LaplCube solve
//            update_uvwp(ns_cube.cpp:241-277) -> k_update
#include <cmath>
#include <cstring>
#include <new>

#include "common.h"
#include "lapl_cube.h"

namespace fdmb {

// offset-indexed 3-D view, row-major, last index fastest (src/tensor.h:207-219).  T is the STORAGE type: double on the
// graded path, float for NSCube<float> (src/ns_cube.cpp:281-282).  The arithmetic of the stencil kernels is double in
// both cases -- exactly like the reference, whose float instantiation mixes its float fields with double dt, Re, dx in
// every expression (src/ns_cube.h:19-31) and so rounds to float only when it stores.
template <typename T> struct FldT {
    T* p;
    int lz, ly, lx;       // lowest index per axis
    long long sz, sy;     // strides (doubles)
    __host__ __device__ __forceinline__ T& at(int i, int k, int j) const
    {
        return p[(long long)(i - lz) * sz + (long long)(k - ly) * sy + (j - lx)];
    }
};
using Fld = FldT<double>;

struct NSGeom {
    int nx, ny, nz;
    double U0, dt;
    double cRx, cRy, cRz;   // 1/Re/dx2 ...
    double idx, idy, idz;   // 1/dx ...
    double iRdx, iRdy, iRdz; // 1/Re/dx ...
    double idx2, idy2, idz2; // 1/dx2 ...
    double idt;
    double dtdx, dtdy, dtdz; // dt/dx ...
    double dcx, dcy, dcz, dcc; // dt/Re/dx2 ..., -2 (dcx + dcy + dcz)     (k_fgh_div)
    double qx, qy, qz;         // dt/(4 dx) ...
};

// ---- init_bound ---------------------------------------------------------------------
// lid (ns_cube.cpp:67-72): u[nz+1][k][j] = 2 U0 - u[nz][k][j], k=0..ny+1, j=-1..jmax
template <typename T> __global__ void k_bound_lid(FldT<T> u, NSGeom g, int jmax)
{
    pdl_wait();
    pdl_trigger();
    int j = blockIdx.x * blockDim.x + threadIdx.x - 1;
    int k = blockIdx.y;
    if (j > jmax) return;
    u.at(g.nz + 1, k, j) = 2 * g.U0 - u.at(g.nz, k, j);
}

// mirror ghosts (ns_cube.cpp:76-95).  blockIdx.z selects the field.  zlo..zhi: the z planes of u and v held by
// this rank (0..nz+1 on one GPU; own planes + halos when sharded); wbot / wtop: this rank holds the bottom / top
// z ghost plane of w.
template <typename T> __global__ void k_bound_mirror(FldT<T> u, FldT<T> v, FldT<T> w, NSGeom g, int zlo, int zhi, int wbot, int wtop)
{
    pdl_wait();
    pdl_trigger();
    int a = blockIdx.x * blockDim.x + threadIdx.x;   // fast index of the face
    int b = blockIdx.y;                              // slow index of the face
    if (blockIdx.z == 0) {          // u: i = b in zlo..zhi, k = a in 0..ny+1
        b += zlo;
        if (b <= zhi && a <= g.ny + 1) {
            u.at(b, a, -1) = u.at(b, a, 1);
            u.at(b, a, g.nx + 1) = u.at(b, a, g.nx - 1);
        }
    } else if (blockIdx.z == 1) {   // v: i = b in zlo..zhi, j = a in 0..nx+1
        b += zlo;
        if (b <= zhi && a <= g.nx + 1) {
            v.at(b, -1, a) = v.at(b, 1, a);
            v.at(b, g.ny + 1, a) = v.at(b, g.ny - 1, a);
        }
    } else {                        // w: k = b in 0..ny+1, j = a in 0..nx+1
        if (b <= g.ny + 1 && a <= g.nx + 1) {
            if (wbot) w.at(-1, b, a) = w.at(1, b, a);
            if (wtop) w.at(g.nz + 1, b, a) = w.at(g.nz - 1, b, a);
        }
    }
}

// pressure ghosts (ns_cube.cpp:98-121).  ilo..ihi: this rank's interior planes (1..nz on one GPU).
template <typename T> __global__ void k_bound_p(FldT<T> u, FldT<T> v, FldT<T> w, FldT<T> p, NSGeom g, int ilo, int ihi, int wbot, int wtop)
{
    pdl_wait();
    pdl_trigger();
    int a = blockIdx.x * blockDim.x + threadIdx.x + 1;
    int b = blockIdx.y + 1;
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    if (blockIdx.z == 0) {          // x faces: i = b in ilo..ihi, k = a in 1..ny
        b += ilo - 1;
        if (b <= ihi && a <= ny) {
            int i = b, k = a;
            p.at(i, k, 0) = p.at(i, k, 1) - (u.at(i, k, 1) - 2 * u.at(i, k, 0) + u.at(i, k, -1)) * g.iRdx;
            p.at(i, k, nx + 1) = p.at(i, k, nx) - (u.at(i, k, nx + 1) - 2 * u.at(i, k, nx) + u.at(i, k, nx - 1)) * g.iRdx;
        }
    } else if (blockIdx.z == 1) {   // y faces: i = b in ilo..ihi, j = a in 1..nx
        b += ilo - 1;
        if (b <= ihi && a <= nx) {
            int i = b, j = a;
            p.at(i, 0, j) = p.at(i, 1, j) - (v.at(i, 1, j) - 2 * v.at(i, 0, j) + v.at(i, -1, j)) * g.iRdy;
            p.at(i, ny + 1, j) = p.at(i, ny, j) - (v.at(i, ny + 1, j) - 2 * v.at(i, ny, j) + v.at(i, ny - 1, j)) * g.iRdy;
        }
    } else {                        // z faces: k = b in 1..ny, j = a in 1..nx
        if (b <= ny && a <= nx) {
            int k = b, j = a;
            if (wbot) p.at(0, k, j) = p.at(1, k, j) - (w.at(1, k, j) - 2 * w.at(0, k, j) + w.at(-1, k, j)) * g.iRdz;
            if (wtop)
                p.at(nz + 1, k, j) = p.at(nz, k, j) - (w.at(nz + 1, k, j) - 2 * w.at(nz, k, j) + w.at(nz - 1, k, j)) * g.iRdz;
        }
    }
}

// init_bound in ONE launch (the double-precision step).  The three ordered fills above depend on each other only through
// values a thread can form itself: the mirror of u on the lid plane reads what the lid fill has just written there
// (2 U0 - u[nz], where the lid loop reaches), and the pressure ghosts read mirrored ghosts (u[-1] = u[1], u[nx+1] =
// u[nx-1], ... on planes the lid never touches).  With those substitutions every role writes a set of elements nobody
// else reads or writes: blockIdx.z = 0 lid (j = 0..min(jmax, nx); j = -1 and nx+1 of that plane belong to the u mirror),
// 1..3 mirrors of u, v, w, 4..6 pressure ghosts on the x, y, z faces.  Same expressions, same operands: bit-identical to
// the three-kernel sequence (which the single-precision step keeps).
template <typename T>
__global__ void k_bound_all(FldT<T> u, FldT<T> v, FldT<T> w, FldT<T> p, NSGeom g, int jmax, int zlo, int zhi, int ilo, int ihi,
                            int wbot, int wtop)
{
    pdl_wait();
    pdl_trigger();
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    switch (blockIdx.z) {
    case 0: {        // lid (ns_cube.cpp:67-72), top rank: k = r in 0..ny+1, j = t in 0..min(jmax, nx)
        if (!wtop || r > ny + 1 || t > jmax || t > nx) return;
        u.at(nz + 1, r, t) = 2 * g.U0 - u.at(nz, r, t);
        return;
    }
    case 1: {        // u mirror (:76-81): i = zlo + r, k = t in 0..ny+1
        const int i = zlo + r, k = t;
        if (i > zhi || k > ny + 1) return;
        const bool lid = wtop && i == nz + 1;       // that plane's interior is being rewritten by role 0
        const T a = (lid && 1 <= jmax) ? (T)(2 * g.U0 - u.at(nz, k, 1)) : u.at(i, k, 1);
        const T b = (lid && nx - 1 <= jmax) ? (T)(2 * g.U0 - u.at(nz, k, nx - 1)) : u.at(i, k, nx - 1);
        u.at(i, k, -1) = a;
        u.at(i, k, nx + 1) = b;
        return;
    }
    case 2: {        // v mirror (:83-88): i = zlo + r, j = t in 0..nx+1
        const int i = zlo + r, j = t;
        if (i > zhi || j > nx + 1) return;
        v.at(i, -1, j) = v.at(i, 1, j);
        v.at(i, ny + 1, j) = v.at(i, ny - 1, j);
        return;
    }
    case 3: {        // w mirror (:90-95): k = r in 0..ny+1, j = t in 0..nx+1
        if (r > ny + 1 || t > nx + 1) return;
        if (wbot) w.at(-1, r, t) = w.at(1, r, t);
        if (wtop) w.at(nz + 1, r, t) = w.at(nz - 1, r, t);
        return;
    }
    case 4: {        // pressure ghosts, x faces (:98-104): i = ilo + r, k = t + 1 in 1..ny; u[-1] = u[1], u[nx+1] = u[nx-1]
        const int i = ilo + r, k = t + 1;
        if (i > ihi || k > ny) return;
        p.at(i, k, 0) = p.at(i, k, 1) - (u.at(i, k, 1) - 2 * u.at(i, k, 0) + u.at(i, k, 1)) * g.iRdx;
        p.at(i, k, nx + 1) = p.at(i, k, nx) - (u.at(i, k, nx - 1) - 2 * u.at(i, k, nx) + u.at(i, k, nx - 1)) * g.iRdx;
        return;
    }
    case 5: {        // y faces (:106-112): i = ilo + r, j = t + 1 in 1..nx; v[-1] = v[1], v[ny+1] = v[ny-1]
        const int i = ilo + r, j = t + 1;
        if (i > ihi || j > nx) return;
        p.at(i, 0, j) = p.at(i, 1, j) - (v.at(i, 1, j) - 2 * v.at(i, 0, j) + v.at(i, 1, j)) * g.iRdy;
        p.at(i, ny + 1, j) = p.at(i, ny, j) - (v.at(i, ny - 1, j) - 2 * v.at(i, ny, j) + v.at(i, ny - 1, j)) * g.iRdy;
        return;
    }
    default: {       // z faces (:114-121): k = r + 1 in 1..ny, j = t + 1 in 1..nx; w[-1] = w[1], w[nz+1] = w[nz-1]
        const int k = r + 1, j = t + 1;
        if (k > ny || j > nx) return;
        if (wbot) p.at(0, k, j) = p.at(1, k, j) - (w.at(1, k, j) - 2 * w.at(0, k, j) + w.at(1, k, j)) * g.iRdz;
        if (wtop)
            p.at(nz + 1, k, j) = p.at(nz, k, j) - (w.at(nz - 1, k, j) - 2 * w.at(nz, k, j) + w.at(nz - 1, k, j)) * g.iRdz;
        return;
    }
    }
}

__device__ __forceinline__ double sq(double x) { return x * x; }

// linear element offset of (i,k,j); every field of the graded sizes has < 2^31 elements per axis product
template <typename T> __device__ __forceinline__ long long lin(const FldT<T>& f, int i, int k, int j)
{
    return (long long)(i - f.lz) * f.sz + (long long)(k - f.ly) * f.sy + (j - f.lx);
}

// ---- FGH (ns_cube.cpp:126-200) ----------------------------------------------------------
// One thread per (i,k,j) in [i0..]x[0..ny]x[0..nx]; F where i>=iFG and k>=1, G where i>=iFG and j>=1, H where
// k,j>=1 (one GPU: i0 = 0, iFG = 1; a sharded rank starts one plane below its slab and computes only H there).
// Interior threads (i,k,j >= 1) read the 27 distinct taps of the three stencils once through
// row pointers (immediate offsets, no per-tap address arithmetic) and produce F, G and H together;
// the O(n^2) edge threads take the generic path.
template <bool F_, bool G_, bool H_, typename T>
__device__ __forceinline__ void fgh_generic(const FldT<T>& u, const FldT<T>& v, const FldT<T>& w, const FldT<T>& F, const FldT<T>& G,
                                            const FldT<T>& H, const NSGeom& g, int i, int k, int j)
{
#define U(a, b, c) u.at(a, b, c)
#define V(a, b, c) v.at(a, b, c)
#define W(a, b, c) w.at(a, b, c)
    if (F_) {
        const double uc = U(i, k, j);
        F.at(i, k, j) = uc + g.dt * (
            (U(i, k, j + 1) - 2 * uc + U(i, k, j - 1)) * g.cRx +
            (U(i, k + 1, j) - 2 * uc + U(i, k - 1, j)) * g.cRy +
            (U(i + 1, k, j) - 2 * uc + U(i - 1, k, j)) * g.cRz -
            (sq(0.5 * (uc + U(i, k, j + 1))) - sq(0.5 * (U(i, k, j - 1) + uc))) * g.idx -
            0.25 * ((uc + U(i, k + 1, j)) * (V(i, k, j + 1) + V(i, k, j)) -
                    (U(i, k - 1, j) + uc) * (V(i, k - 1, j + 1) + V(i, k - 1, j))) * g.idy -
            0.25 * ((uc + U(i + 1, k, j)) * (W(i, k, j + 1) + W(i, k, j)) -
                    (U(i - 1, k, j) + uc) * (W(i - 1, k, j + 1) + W(i - 1, k, j))) * g.idz);
    }
    if (G_) {
        const double vc = V(i, k, j);
        G.at(i, k, j) = vc + g.dt * (
            (V(i, k, j + 1) - 2 * vc + V(i, k, j - 1)) * g.cRx +
            (V(i, k + 1, j) - 2 * vc + V(i, k - 1, j)) * g.cRy +
            (V(i + 1, k, j) - 2 * vc + V(i - 1, k, j)) * g.cRz -
            (sq(0.5 * (vc + V(i, k + 1, j))) - sq(0.5 * (V(i, k - 1, j) + vc))) * g.idy -
            0.25 * ((U(i, k, j) + U(i, k + 1, j)) * (V(i, k, j + 1) + vc) -
                    (U(i, k, j - 1) + U(i, k + 1, j - 1)) * (vc + V(i, k, j - 1))) * g.idx -
            0.25 * ((W(i, k, j) + W(i, k + 1, j)) * (vc + V(i + 1, k, j)) -
                    (W(i - 1, k, j) + W(i - 1, k + 1, j)) * (V(i - 1, k, j) + vc)) * g.idz);
    }
    if (H_) {
        const double wc = W(i, k, j);
        H.at(i, k, j) = wc + g.dt * (
            (W(i, k, j + 1) - 2 * wc + W(i, k, j - 1)) * g.cRx +
            (W(i, k + 1, j) - 2 * wc + W(i, k - 1, j)) * g.cRy +
            (W(i + 1, k, j) - 2 * wc + W(i - 1, k, j)) * g.cRz -
            (sq(0.5 * (W(i + 1, k, j) + wc)) - sq(0.5 * (W(i - 1, k, j) + wc))) * g.idz -
            0.25 * ((U(i + 1, k, j) + U(i, k, j)) * (W(i, k, j + 1) + wc) -
                    (U(i + 1, k, j - 1) + U(i, k, j - 1)) * (wc + W(i, k, j - 1))) * g.idx -
            0.25 * ((wc + W(i, k + 1, j)) * (V(i, k, j) + V(i + 1, k, j)) -
                    (W(i, k - 1, j) + wc) * (V(i, k - 1, j) + V(i + 1, k - 1, j))) * g.idy);
    }
#undef U
#undef V
#undef W
}

template <typename T> __global__ void __launch_bounds__(256) k_fgh(FldT<T> u, FldT<T> v, FldT<T> w, FldT<T> F, FldT<T> G, FldT<T> H, NSGeom g, int i0, int iFG)
{
    pdl_wait();
    pdl_trigger();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = blockIdx.z + i0;
    if (j > g.nx || k > g.ny) return;
    if (i >= iFG && k >= 1 && j >= 1) {
        const T* __restrict__ uc_ = u.p + lin(u, i, k, j);
        const T* __restrict__ vc_ = v.p + lin(v, i, k, j);
        const T* __restrict__ wc_ = w.p + lin(w, i, k, j);
        const T* __restrict__ ukp = uc_ + u.sy; const T* __restrict__ ukm = uc_ - u.sy;
        const T* __restrict__ uip = uc_ + u.sz; const T* __restrict__ uim = uc_ - u.sz;
        const T* __restrict__ vkp = vc_ + v.sy; const T* __restrict__ vkm = vc_ - v.sy;
        const T* __restrict__ vip = vc_ + v.sz; const T* __restrict__ vim = vc_ - v.sz;
        const T* __restrict__ vipkm = vip - v.sy;
        const T* __restrict__ wkp = wc_ + w.sy; const T* __restrict__ wkm = wc_ - w.sy;
        const T* __restrict__ wip = wc_ + w.sz; const T* __restrict__ wim = wc_ - w.sz;
        const T* __restrict__ wimkp = wim + w.sy;
        // the 27 taps, named by (di,dk,dj) with m = -1, p = +1
        const double u000 = uc_[0], u00p = uc_[1], u00m = uc_[-1], u0p0 = ukp[0], u0pm = ukp[-1], u0m0 = ukm[0];
        const double up00 = uip[0], up0m = uip[-1], um00 = uim[0];
        const double v000 = vc_[0], v00p = vc_[1], v00m = vc_[-1], v0p0 = vkp[0], v0m0 = vkm[0], v0mp = vkm[1];
        const double vp00 = vip[0], vpm0 = vipkm[0], vm00 = vim[0];
        const double w000 = wc_[0], w00p = wc_[1], w00m = wc_[-1], w0p0 = wkp[0], w0m0 = wkm[0];
        const double wp00 = wip[0], wm00 = wim[0], wm0p = wim[1], wmp0 = wimkp[0];

        F.p[lin(F, i, k, j)] = u000 + g.dt * (
            (u00p - 2 * u000 + u00m) * g.cRx +
            (u0p0 - 2 * u000 + u0m0) * g.cRy +
            (up00 - 2 * u000 + um00) * g.cRz -
            (sq(0.5 * (u000 + u00p)) - sq(0.5 * (u00m + u000))) * g.idx -
            0.25 * ((u000 + u0p0) * (v00p + v000) -
                    (u0m0 + u000) * (v0mp + v0m0)) * g.idy -
            0.25 * ((u000 + up00) * (w00p + w000) -
                    (um00 + u000) * (wm0p + wm00)) * g.idz);
        G.p[lin(G, i, k, j)] = v000 + g.dt * (
            (v00p - 2 * v000 + v00m) * g.cRx +
            (v0p0 - 2 * v000 + v0m0) * g.cRy +
            (vp00 - 2 * v000 + vm00) * g.cRz -
            (sq(0.5 * (v000 + v0p0)) - sq(0.5 * (v0m0 + v000))) * g.idy -
            0.25 * ((u000 + u0p0) * (v00p + v000) -
                    (u00m + u0pm) * (v000 + v00m)) * g.idx -
            0.25 * ((w000 + w0p0) * (v000 + vp00) -
                    (wm00 + wmp0) * (vm00 + v000)) * g.idz);
        H.p[lin(H, i, k, j)] = w000 + g.dt * (
            (w00p - 2 * w000 + w00m) * g.cRx +
            (w0p0 - 2 * w000 + w0m0) * g.cRy +
            (wp00 - 2 * w000 + wm00) * g.cRz -
            (sq(0.5 * (wp00 + w000)) - sq(0.5 * (wm00 + w000))) * g.idz -
            0.25 * ((up00 + u000) * (w00p + w000) -
                    (up0m + u00m) * (w000 + w00m)) * g.idx -
            0.25 * ((w000 + w0p0) * (v000 + vp00) -
                    (w0m0 + w000) * (v0m0 + vpm0)) * g.idy);
        return;
    }
    if (i >= iFG && k >= 1) fgh_generic<true, false, false>(u, v, w, F, G, H, g, i, k, j);
    if (i >= iFG && j >= 1) fgh_generic<false, true, false>(u, v, w, F, G, H, g, i, k, j);
    if (k >= 1 && j >= 1) fgh_generic<false, false, true>(u, v, w, F, G, H, g, i, k, j);
}

// ---- FGH + divergence in one sweep (ns_cube.cpp:126-200 and 204-235) -------------------------------------------
// z-marching with the taps in registers, warps independent of each other (no shared memory, no barrier).  A warp owns
// 31 consecutive j of one row k (lane 0 is the column to the left, where only F is produced) and marches over a chunk of
// z planes.  Of the taps of a plane's stencils, twelve are last step's registers (the five taps of plane i+1 become
// centre taps, seven centre taps become taps of plane i-1), so a step loads 20 values per point through row pointers
// that advance by one plane.
//   RHS = ((F - F[j-1])/dx + (G - G[k-1])/dy + (H - H[i-1])/dz)/dt - ghost pressures:
//   F[j-1] comes from the lane to the left (shuffle), H[i-1] is the thread's own value of the previous plane (the chunk's
//   first plane i = ia - 1 only produces it), and G[k-1] is evaluated a second time by the thread itself with the
//   k-shifted taps (three more loads and one more stencil, +28 % fp64 work): exchanging it between warps costs a
//   barrier per plane, which serialises the load and compute phases of a block (measured: 404 us against 388 us for
//   k_fgh + k_rhs at 255^3).
// Traffic: u, v, w read once, F, G, H, RHS written once = 56 B per point, against 80 B for k_fgh + k_rhs.
// The same kernel serves a z-slab of the sharded step: it starts one plane below the slab like k_fgh does.
// Addresses of taps a thread does not use stay inside the allocation: rows k-1 of u and w fall back to row k where
// k = 0 (row k-2 of v to row k-1), and v / w at j-1 = -1 (lane 0 of the first block column, F only) is the element
// before the row.
// one momentum stencil with dt folded into the coefficients:
//   c + dcx sx + dcy sy + dcz sz + dcc c  =  c + dt (second differences)/Re       (sx = the two x neighbours' sum ...)
//   - qa (a^2 - b^2)                      =  - dt ((a/2)^2 - (b/2)^2)/d           (the squared face averages)
//   - q1 d1 - q2 d2                       =  - dt/4 (product differences)/d       (the two cross terms)
// 9 fused operations instead of the 20 of the reference's order of evaluation; differs from it by a few ulps.
__device__ __forceinline__ double fgh_combine(const NSGeom& g, double c, double sx, double sy, double sz, double a, double b,
                                              double qa, double d1, double q1, double d2, double q2)
{
    double acc = fma(g.dcc, c, c);
    acc = fma(g.dcx, sx, acc);
    acc = fma(g.dcy, sy, acc);
    acc = fma(g.dcz, sz, acc);
    double t = a * a;
    t = fma(-b, b, t);
    acc = fma(-qa, t, acc);
    acc = fma(-q1, d1, acc);
    return fma(-q2, d2, acc);
}

constexpr int FGH_TK = 4, FGH_MINB = 3;      // 128-thread blocks (4 rows), 168 registers: two register sets of taps
template <int TK, int MINB, int PF>
__global__ void __launch_bounds__(32 * TK, MINB)
k_fgh_div(Fld u, Fld v, Fld w, Fld p, Fld F, Fld G, Fld H, Fld R, NSGeom g, int ia0, int ib0, int zchunk)
{
    pdl_wait();
    pdl_trigger();
    const int tj = threadIdx.x;
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    const int j_ = blockIdx.x * 31 + tj, k = blockIdx.y * TK + threadIdx.y;
    if (k > ny) return;                              // (whole warps)
    const bool inb = j_ <= nx;
    const int j = inb ? j_ : nx;                     // out-of-range lanes shadow the last column, store nothing
    const int ia = ia0 + blockIdx.z * zchunk;
    const int ib = (ia + zchunk - 1 < ib0) ? ia + zchunk - 1 : ib0;
    const bool core = inb && k >= 1 && tj >= 1;
    const bool stF = inb && k >= 1 && (tj >= 1 || j == 0);   // (lane 0 repeats the last column of the block to the left)
    const bool stG = inb && tj >= 1;                 // j >= 1; k >= 0
    const bool edge_kj = k <= 1 || j <= 1 || j >= nx || k >= ny;

    // row pointers at plane ia - 1 (the priming plane)
    const double* __restrict__ uc_ = u.p + lin(u, ia - 1, k, j);
    const double* __restrict__ vc_ = v.p + lin(v, ia - 1, k, j);
    const double* __restrict__ wc_ = w.p + lin(w, ia - 1, k, j);
    const long long usy = u.sy, vsy = v.sy, wsy = w.sy, usz = u.sz, vsz = v.sz, wsz = w.sz;
    const long long ukm_o = k >= 1 ? -usy : 0, wkm_o = k >= 1 ? -wsy : 0, vkM_o = k >= 1 ? -2 * vsy : -vsy;
    double* __restrict__ Fp = F.p + lin(F, ia - 1, k, j);
    double* __restrict__ Gp = G.p + lin(G, ia - 1, k, j);
    double* __restrict__ Hp = H.p + lin(H, ia - 1, k, j);
    double* __restrict__ Rp = R.p + lin(R, ia - 1, k, j);
    const double* __restrict__ pp = p.p + lin(p, ia - 1, k, j);     // ghost pressures of the boundary cells
    const long long psy = p.sy, psz = p.sz;

    // the twenty values a step loads: fifteen taps of its centre plane and the five taps of the plane above
    struct Ld {
        double u00p, u0p0, u0pm, u0m0, u0mm, v00p, v00m, v0p0, v0mp, v0mm, v0M0, w00p, w0p0, w0m0, w00m;
        double up00, up0m, vp00, vpm0, wp00;
    };
    auto load = [&](Ld& L, const bool pf) {     // centre plane = the plane uc_, vc_, wc_ point at; advances them
        L.u00p = uc_[1]; L.u0p0 = uc_[usy]; L.u0pm = uc_[usy - 1]; L.u0m0 = uc_[ukm_o]; L.u0mm = uc_[ukm_o - 1];
        L.v00p = vc_[1]; L.v00m = vc_[-1]; L.v0p0 = vc_[vsy]; L.v0mp = vc_[-vsy + 1]; L.v0mm = vc_[-vsy - 1]; L.v0M0 = vc_[vkM_o];
        L.w00p = wc_[1]; L.w0p0 = wc_[wsy]; L.w0m0 = wc_[wkm_o]; L.w00m = wc_[-1];
        L.up00 = uc_[usz]; L.up0m = uc_[usz - 1]; L.vp00 = vc_[vsz]; L.vpm0 = vc_[vsz - vsy]; L.wp00 = wc_[wsz];
        uc_ += usz; vc_ += vsz; wc_ += wsz;
        if (PF && pf) {   // the first-touch rows of the load after the next (three planes above this centre) on their way to L2
            asm volatile("prefetch.global.L2 [%0];" ::"l"(uc_ + 2 * usz));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vc_ + 2 * vsz));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(wc_ + 2 * wsz));
        }
    };
    // carried registers: the centre taps that were last step's plane-above taps, and the taps of the plane below that
    // were last step's centre taps.  As seen from centre plane ia - 2 only what the priming step (H alone) consumes
    // is real.
    double u000 = uc_[0], u00m = uc_[-1], v000 = vc_[0], v0m0 = vc_[-vsy], w000 = wc_[0];
    double um00 = 0.0, vm00 = 0.0, wm00 = wc_[-wsz], wm0p = 0.0, wmp0 = 0.0, vmm0 = 0.0, wmm0 = 0.0;
    double hprev = 0.0;
    auto step = [&](const int i, const Ld& L) {
        const bool fg = i >= ia;
        // ghost pressures of the boundary cells (ns_cube.cpp:216-232): loaded before the stencils, consumed after them
        const bool edge = fg && core && (edge_kj || i <= 1 || i >= nz);
        double pzm = 0.0, pzp = 0.0, pym = 0.0, pyp = 0.0, pxm = 0.0, pxp = 0.0;
        if (edge) {
            if (i <= 1) pzm = pp[-psz];
            if (i >= nz) pzp = pp[psz];
            if (k <= 1) pym = pp[-psy];
            if (k >= ny) pyp = pp[psy];
            if (j <= 1) pxm = pp[-1];
            if (j >= nx) pxp = pp[1];
            asm volatile("" ::: "memory");
        }
        // every thread evaluates all four stencils (the stores are predicated); the eleven distinct products of face sums
        // are formed once: F and G share Puv1, F and H Puw1, G and H Pwv1, G[k-1] shares Puv2 with F and Pwv2 with H
        const double Puv1 = (u000 + L.u0p0) * (L.v00p + v000), Puv2 = (L.u0m0 + u000) * (L.v0mp + v0m0);
        const double Puw1 = (u000 + L.up00) * (L.w00p + w000), Puw2 = (um00 + u000) * (wm0p + wm00);
        const double Pwv1 = (w000 + L.w0p0) * (v000 + L.vp00), Pwv2 = (L.w0m0 + w000) * (v0m0 + L.vpm0);
        const double Pg2 = (u00m + L.u0pm) * (v000 + L.v00m), Pg4 = (wm00 + wmp0) * (vm00 + v000);
        const double Pm2 = (L.u0mm + u00m) * (v0m0 + L.v0mm), Pm4 = (wmm0 + wm00) * (vmm0 + v0m0);
        const double Ph2 = (L.up0m + u00m) * (w000 + L.w00m);
        const double fv = fgh_combine(g, u000, L.u00p + u00m, L.u0p0 + L.u0m0, L.up00 + um00, u000 + L.u00p, u00m + u000, g.qx,
                                      Puv1 - Puv2, g.qy, Puw1 - Puw2, g.qz);                   // ns_cube.cpp:136-149
        const double gv = fgh_combine(g, v000, L.v00p + L.v00m, L.v0p0 + v0m0, L.vp00 + vm00, v000 + L.v0p0, v0m0 + v000, g.qy,
                                      Puv1 - Pg2, g.qx, Pwv1 - Pg4, g.qz);                     // :156-169
        const double gm = fgh_combine(g, v0m0, L.v0mp + L.v0mm, v000 + L.v0M0, L.vpm0 + vmm0, v0m0 + v000, L.v0M0 + v0m0, g.qy,
                                      Puv2 - Pm2, g.qx, Pwv2 - Pm4, g.qz);                     // the same at (i, k-1, j)
        const double hv = fgh_combine(g, w000, L.w00p + L.w00m, L.w0p0 + L.w0m0, L.wp00 + wm00, L.wp00 + w000, wm00 + w000, g.qz,
                                      Puw1 - Ph2, g.qx, Pwv1 - Pwv2, g.qy);                    // :176-189
        if (fg) {
            const double fm = __shfl_up_sync(0xffffffffu, fv, 1);
            if (stF) *Fp = fv;
            if (stG) *Gp = gv;
            if (core) {
                *Hp = hv;
                double r = ((fv - fm) * g.idx + (gv - gm) * g.idy + (hv - hprev) * g.idz) * g.idt;
                if (edge) {
                    if (i <= 1) r -= pzm * g.idz2;
                    if (k <= 1) r -= pym * g.idy2;
                    if (j <= 1) r -= pxm * g.idx2;
                    if (j >= nx) r -= pxp * g.idx2;
                    if (k >= ny) r -= pyp * g.idy2;
                    if (i >= nz) r -= pzp * g.idz2;
                }
                *Rp = r;
            }
        } else if (core && i == 0) {
            *Hp = hv;              // (plane ia - 1 belongs to the chunk below, except H[0])
        }
        hprev = hv;
        // shift: centre -> plane below, plane above -> centre
        um00 = u000; vm00 = v000; wm00 = w000; wm0p = L.w00p; wmp0 = L.w0p0; vmm0 = v0m0; wmm0 = L.w0m0;
        u000 = L.up00; u00m = L.up0m; v000 = L.vp00; v0m0 = L.vpm0; w000 = L.wp00;
        Fp += F.sz; Gp += G.sz; Hp += H.sz; Rp += R.sz; pp += psz;
    };
    // software pipeline: the loads of plane i + 1 are in flight while plane i is evaluated (two register sets)
    Ld A, B;
    load(A, ia - 1 <= ib - 2);
    int i = ia - 1;
#pragma unroll 1
    for (; i + 2 <= ib; i += 2) {      // (unconditional loads: the compiler keeps them ahead of the arithmetic)
        load(B, i + 1 <= ib - 2);
        asm volatile("" ::: "memory");
        step(i, A);
        load(A, i + 2 <= ib - 2);
        asm volatile("" ::: "memory");
        step(i + 1, B);
    }
    if (i + 1 <= ib) {
        load(B, false);
        asm volatile("" ::: "memory");
        step(i, A);
        step(i + 1, B);
    } else {
        step(i, A);
    }
}

// ---- poisson RHS (ns_cube.cpp:205-235) ---------------------------------------------------
template <typename T> __global__ void __launch_bounds__(256) k_rhs(FldT<T> F, FldT<T> G, FldT<T> H, FldT<T> p, FldT<T> R, NSGeom g, int ilo)
{
    pdl_wait();
    pdl_trigger();
    const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int k = blockIdx.y * blockDim.y + threadIdx.y + 1;
    const int i = blockIdx.z + ilo;
    if (j > g.nx || k > g.ny) return;
    const T* __restrict__ Fp = F.p + lin(F, i, k, j);
    const T* __restrict__ Gp = G.p + lin(G, i, k, j);
    const T* __restrict__ Hp = H.p + lin(H, i, k, j);
    double r = ((Fp[0] - Fp[-1]) * g.idx + (Gp[0] - Gp[-G.sy]) * g.idy + (Hp[0] - Hp[-H.sz]) * g.idz) * g.idt;
    if (i <= 1 || k <= 1 || j <= 1 || j >= g.nx || k >= g.ny || i >= g.nz) {
        if (i <= 1) r -= p.at(i - 1, k, j) * g.idz2;
        if (k <= 1) r -= p.at(i, k - 1, j) * g.idy2;
        if (j <= 1) r -= p.at(i, k, j - 1) * g.idx2;
        if (j >= g.nx) r -= p.at(i, k, j + 1) * g.idx2;
        if (k >= g.ny) r -= p.at(i, k + 1, j) * g.idy2;
        if (i >= g.nz) r -= p.at(i + 1, k, j) * g.idz2;
    }
    R.p[lin(R, i, k, j)] = r;
}

// ---- update_uvwp (ns_cube.cpp:241-277) ---------------------------------------------------
template <typename T> __global__ void __launch_bounds__(256) k_update(FldT<T> u, FldT<T> v, FldT<T> w, FldT<T> p, FldT<T> x, FldT<T> F, FldT<T> G, FldT<T> H, NSGeom g, int ilo)
{
    pdl_wait();
    pdl_trigger();
    const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int k = blockIdx.y * blockDim.y + threadIdx.y + 1;
    const int i = blockIdx.z + ilo;
    if (j > g.nx || k > g.ny) return;
    const T* __restrict__ xp = x.p + lin(x, i, k, j);
    const double xc = xp[0];
    if (j < g.nx) u.p[lin(u, i, k, j)] = F.p[lin(F, i, k, j)] - g.dtdx * (xp[1] - xc);
    if (k < g.ny) v.p[lin(v, i, k, j)] = G.p[lin(G, i, k, j)] - g.dtdy * (xp[x.sy] - xc);
    if (i < g.nz) w.p[lin(w, i, k, j)] = H.p[lin(H, i, k, j)] - g.dtdz * (xp[x.sz] - xc);
    p.p[lin(p, i, k, j)] = xc;   // p = x copies the index-range intersection (tensor.h:103-111)
}

// ---- halo pull: whole z planes copied from the neighbours' windows (peer loads over NVLink) ---------
constexpr int NS_MAX_PULLS = 8;
struct PullList {
    const double* src[NS_MAX_PULLS];
    double* dst[NS_MAX_PULLS];
    long long n[NS_MAX_PULLS];     // doubles, even (planes of the graded sizes) or odd: handled element-wise
    int count;
};
__global__ void __launch_bounds__(256) k_pull(PullList pl)
{
    const int s = blockIdx.y;
    if (s >= pl.count) return;
    const double* __restrict__ src = pl.src[s];
    double* __restrict__ dst = pl.dst[s];
    const long long n = pl.n[s];
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        dst[e] = src[e];
}

}  // namespace fdmb

using namespace fdmb;

// FGH + divergence run as ONE sweep (k_fgh_div) by default; FDMB_FGH_FUSED=0 selects k_fgh + k_rhs for handles created
// afterwards.  Measured at 255^3 (r02fgh): 274 us against 278 + 110 us, step 0.919 -> 0.810 ms.  History of the fused
// sweep: shared-memory staged tiles with cp.async 625 us (issue-bound staging); register marching with a G exchange
// through shared memory and a barrier per plane 404 us (the barrier serialises a block's load and compute phases);
// independent warps that evaluate G[k-1] themselves 346 us (half the stall samples on the first use of a plane's loads);
// the same with the next plane's loads in flight during the stencils 300 us; ghost-pressure loads hoisted and dt folded
// into the coefficients (11 shared products, 9 fused operations per stencil) 274 us.
static int fgh_chunks_per_sm()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("FDMB_FGH_CHUNKS"); v = e ? atoi(e) : 32; if (v < 1) v = 32; }
    return v;
}
static int fused_fgh_rhs_requested()
{
    const char* e = getenv("FDMB_FGH_FUSED");
    return e ? atoi(e) : 1;
}

// global z range of field `fld` (u v w p x F G H RHS), ns_cube.h:66-75
static inline void field_zrange(int fld, int nz, int* lo, int* hi)
{
    static const int L[9] = {0, 0, -1, 0, 1, 1, 1, 0, 1};
    static const int H1[9] = {1, 1, 1, 1, 0, 0, 0, 0, 0};      // hi = nz + H1
    *lo = L[fld]; *hi = nz + H1[fld];
}

// Planes of field `fld` that rank `rank` of `nranks` HOLDS (window: own planes + halos) and OWNS (the planes it
// reports through get_field; the owned ranges of all ranks tile the global range).  ilo..ihi = interior planes
// of the rank's z-slab, the same slabs as the sharded LaplCube (fdmb_slab_range).
static void ns_planes(int fld, int nz, int rank, int nranks, int* wlo, int* whi, int* olo, int* ohi)
{
    int first = 0, cnt = nz, glo, ghi;
    if (nranks > 1) slab_range(nz, 0, nranks, rank, &first, &cnt);
    const int ilo = first + 1, ihi = first + cnt;
    const bool bot = rank == 0, top = rank == nranks - 1;
    field_zrange(fld, nz, &glo, &ghi);
    *olo = bot ? glo : ilo;
    *ohi = top ? ghi : ihi;
    switch (fld) {
    case 0: case 1: case 3: *wlo = bot ? glo : ilo - 1; *whi = ihi + 1; break;            // u v p: one plane each side
    case 2: *wlo = bot ? glo : ilo - 2; *whi = ihi + 1; break;                            // w: H below the slab reads i-1
    case 4: *wlo = ilo; *whi = top ? ihi : ihi + 1; break;                                // x: update reads x[i+1]
    case 7: *wlo = bot ? glo : ilo - 1; *whi = ihi; break;                                // H: divergence reads H[i-1]
    default: *wlo = ilo; *whi = ihi; break;                                               // F G RHS
    }
}

struct NSLayout {
    int wlo[9], whi[9], olo[9], ohi[9];
    long long sy[9], sz[9], count[9];
    size_t off[9], bytes;
};
static void ns_layout(int nx, int ny, int nz, int rank, int nranks, NSLayout* L)
{
    // x/y extents: ns_cube.h:66-75
    static const int Y0[9] = {0, -1, 0, 0, 1, 1, 0, 1, 1}, X0[9] = {-1, 0, 0, 0, 1, 0, 1, 1, 1};
    const int Y1[9] = {ny + 1, ny + 1, ny + 1, ny + 1, ny, ny, ny, ny, ny};
    const int X1[9] = {nx + 1, nx + 1, nx + 1, nx + 1, nx, nx, nx, nx, nx};
    size_t o = 0;
    for (int f = 0; f < 9; f++) {
        ns_planes(f, nz, rank, nranks, &L->wlo[f], &L->whi[f], &L->olo[f], &L->ohi[f]);
        L->sy[f] = X1[f] - X0[f] + 1;
        L->sz[f] = (long long)(Y1[f] - Y0[f] + 1) * L->sy[f];
        L->count[f] = (long long)(L->whi[f] - L->wlo[f] + 1) * L->sz[f];
        L->off[f] = o;
        o += (sizeof(double) * (size_t)L->count[f] + 255) & ~(size_t)255;
    }
    L->bytes = o;
}

struct fdmb_ns_cube {
    fdmb_ns_cube_params prm{};
    int nx = 0, ny = 0, nz = 0;
    double dx = 0, dy = 0, dz = 0;
    NSGeom g{};
    Fld f[9]{};                 // u v w p x F G H RHS: windows over the global arrays, indexed globally
    NSLayout lay{};
    fdmb_lapl_cube* lapl = nullptr;
    cudaStream_t stream = nullptr;
    long long time_index = 0;
    // z-slab sharding (nranks == 1: the windows are the whole arrays)
    int rank = 0, nranks = 1, ilo = 1, ihi = 0, device = 0;
    void* block = nullptr;                       // all nine windows in one allocation (one IPC handle)
    void* peer_block[FDMB_MAX_RANKS] = {};
    bool peer_ipc[FDMB_MAX_RANKS] = {};
    bool attached = false;

    StepGraph graph;            // one time step as a replayed CUDA graph (single-GPU handles)
    int fused = 0;              // FGH + divergence in one sweep (k_fgh_div), 0 = k_fgh + k_rhs

    int init();
    int step(int nsteps, cudaStream_t st);
    int step_once(cudaStream_t st);
    int pull(const int* flds, const int* lo, const int* hi, const int* from, int n, cudaStream_t st);
    double* owned_ptr(int fld) const { return f[fld].p + (long long)(lay.olo[fld] - lay.wlo[fld]) * lay.sz[fld]; }
    long long owned_count(int fld) const { return (long long)(lay.ohi[fld] - lay.olo[fld] + 1) * lay.sz[fld]; }
    ~fdmb_ns_cube();
};

int fdmb_ns_cube::init()
{
    nx = prm.nx; ny = prm.nx /* ns_cube.h:58: ny is read from key "nx" */; nz = prm.nz;
    fused = fused_fgh_rhs_requested();
    if (nx < 3 || nz < 3) { set_error("NSCube: nx, nz must be >= 3"); return FDMB_ERR_INVALID; }
    if (nranks > 1 && (nz + 1) / nranks < 4) {
        set_error("NSCube: the sharded step needs at least 4 z planes per rank (nz=%d, %d ranks)", nz, nranks);
        return FDMB_ERR_INVALID;
    }
    dx = (prm.x2 - prm.x1) / nx; dy = (prm.y2 - prm.y1) / ny; dz = (prm.z2 - prm.z1) / nz;
    const double dx2 = dx * dx, dy2 = dy * dy, dz2 = dz * dz;
    int rc;
    if (nranks > 1)
        rc = fdmb_lapl_cube_create_sharded(&lapl, dx, dy, dz, prm.x2 - prm.x1 + dx, prm.y2 - prm.y1 + dy,
                                           prm.z2 - prm.z1 + dz, nx, ny, nz, 0, rank, nranks);
    else
        rc = fdmb_lapl_cube_create(&lapl, dx, dy, dz, prm.x2 - prm.x1 + dx, prm.y2 - prm.y1 + dy,
                                   prm.z2 - prm.z1 + dz, nx, ny, nz, 0);   // ns_cube.h:77
    if (rc) return rc;
    FDMB_CUDA(cudaGetDevice(&device));
    FDMB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    {
        int first = 0, cnt = nz;
        if (nranks > 1) slab_range(nz, 0, nranks, rank, &first, &cnt);
        ilo = first + 1; ihi = first + cnt;
    }
    {   // Load this translation unit's kernels NOW.  With lazy module loading the first launch of a kernel may need a
        // context synchronisation; a step starts with a cross-GPU barrier kernel that spins until the peers arrive,
        // so a first-use load behind it deadlocks when several ranks are driven by one host thread.
        cudaFuncAttributes fa;
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_pull));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_fgh<double>));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_rhs<double>));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_fgh_div<FGH_TK, FGH_MINB, 1>));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_update<double>));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_bound_all<double>));
    }
    ns_layout(nx, ny, nz, rank, nranks, &lay);
    FDMB_CUDA(cudaMalloc(&block, lay.bytes));
    FDMB_CUDA(cudaMemset(block, 0, lay.bytes));
    // cudaMemset on device memory is asynchronous to the host and runs on the legacy default stream, which the
    // handle's non-blocking streams do not order against: finish it before the handle is handed out
    FDMB_CUDA(cudaDeviceSynchronize());
    static const int Y0[9] = {0, -1, 0, 0, 1, 1, 0, 1, 1}, X0[9] = {-1, 0, 0, 0, 1, 0, 1, 1, 1};
    for (int k = 0; k < 9; k++) {
        f[k].p = reinterpret_cast<double*>(static_cast<char*>(block) + lay.off[k]);
        f[k].lz = lay.wlo[k]; f[k].ly = Y0[k]; f[k].lx = X0[k];
        f[k].sy = lay.sy[k]; f[k].sz = lay.sz[k];
    }
    peer_block[rank] = block;
    g.nx = nx; g.ny = ny; g.nz = nz; g.U0 = prm.u0; g.dt = prm.dt;
    const double Re = prm.Re;
    g.cRx = 1.0 / Re / dx2; g.cRy = 1.0 / Re / dy2; g.cRz = 1.0 / Re / dz2;
    g.idx = 1.0 / dx; g.idy = 1.0 / dy; g.idz = 1.0 / dz;
    g.iRdx = 1.0 / Re / dx; g.iRdy = 1.0 / Re / dy; g.iRdz = 1.0 / Re / dz;
    g.idx2 = 1.0 / dx2; g.idy2 = 1.0 / dy2; g.idz2 = 1.0 / dz2;
    g.idt = 1.0 / prm.dt;
    g.dtdx = prm.dt / dx; g.dtdy = prm.dt / dy; g.dtdz = prm.dt / dz;
    g.dcx = prm.dt * g.cRx; g.dcy = prm.dt * g.cRy; g.dcz = prm.dt * g.cRz; g.dcc = -2.0 * (g.dcx + g.dcy + g.dcz);
    g.qx = 0.25 * prm.dt * g.idx; g.qy = 0.25 * prm.dt * g.idy; g.qz = 0.25 * prm.dt * g.idz;
    return FDMB_OK;
}

fdmb_ns_cube::~fdmb_ns_cube()
{
    for (int q = 0; q < nranks; q++)
        if (q != rank && peer_ipc[q] && peer_block[q]) cudaIpcCloseMemHandle(peer_block[q]);
    cudaFree(block);
    if (lapl) fdmb_lapl_cube_destroy(lapl);
    if (stream) cudaStreamDestroy(stream);
}

// copy planes lo[s]..hi[s] of field flds[s] from rank from[s]'s window into this rank's window
int fdmb_ns_cube::pull(const int* flds, const int* lo, const int* hi, const int* from, int n, cudaStream_t st)
{
    PullList pl{};
    long long nmax = 0;
    for (int s = 0; s < n; s++) {
        const int k = flds[s];
        NSLayout q;
        ns_layout(nx, ny, nz, from[s], nranks, &q);
        pl.src[s] = reinterpret_cast<const double*>(static_cast<const char*>(peer_block[from[s]]) + q.off[k]) +
                    (long long)(lo[s] - q.wlo[k]) * q.sz[k];
        pl.dst[s] = f[k].p + (long long)(lo[s] - lay.wlo[k]) * lay.sz[k];
        pl.n[s] = (long long)(hi[s] - lo[s] + 1) * lay.sz[k];
        if (pl.n[s] > nmax) nmax = pl.n[s];
    }
    pl.count = n;
    LaunchScope sc("ns_halo_pull", st);
    long long bx = (nmax + 256 * 4 - 1) / (256 * 4);
    if (bx > 1024) bx = 1024;
    if (bx < 1) bx = 1;
    k_pull<<<dim3((unsigned)bx, n), 256, 0, st>>>(pl);
    FDMB_CHECK_LAUNCH();
    return FDMB_OK;
}

int fdmb_ns_cube::step(int nsteps, cudaStream_t st)
{
    if (nranks > 1 && !attached) { set_error("NSCube: sharded handle used before attach_ipc/attach_local"); return FDMB_ERR_COMM; }
    for (int s = 0; s < nsteps; s++) {
        // (sharded steps too: the cross-GPU barriers keep their epoch in device memory, so the sequence replays)
        const int rc = graph.run(st, this, nullptr, [&]() { return step_once(st); });
        if (rc) return rc;
        time_index++;
    }
    return FDMB_OK;
}

int fdmb_ns_cube::step_once(cudaStream_t st)
{
    const Fld &u = f[0], &v = f[1], &w = f[2], &p = f[3], &x = f[4], &F = f[5], &G = f[6], &H = f[7], &R = f[8];
    const int nmax = nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz);
    const bool bot = rank == 0, top = rank == nranks - 1;
    const int nzl = ihi - ilo + 1;
    PdlScope pdl(nranks == 1 && pdl_small_grid((long long)nx * ny * nz));   // launch-bound sizes (pdl.cuh)
    int rc;
    {
        if (nranks > 1) {
            // every rank has finished the previous update (or set_field): fetch the halo planes of u, v, w
            if ((rc = lapl->barrier(st))) return rc;
            int flds[6], lo[6], hi[6], from[6], n = 0;
            if (!bot) {
                flds[n] = 0; lo[n] = hi[n] = ilo - 1; from[n++] = rank - 1;
                flds[n] = 1; lo[n] = hi[n] = ilo - 1; from[n++] = rank - 1;
                flds[n] = 2; lo[n] = ilo - 2; hi[n] = ilo - 1; from[n++] = rank - 1;
            }
            if (!top) {
                for (int k = 0; k < 3; k++) { flds[n] = k; lo[n] = hi[n] = ihi + 1; from[n++] = rank + 1; }
            }
            if ((rc = pull(flds, lo, hi, from, n, st))) return rc;
        }
        {   // init_bound, one launch (k_bound_all).  The reference's lid loop runs j = -1..nz+1 (ns_cube.cpp:68): clamped
            // to the allocated x range
            LaunchScope sc("ns_bound", st);
            const int jmax = (nz + 1 < nx + 1) ? nz + 1 : nx + 1;
            const int zlo = lay.wlo[0], zhi = lay.whi[0];
            const int rows = (zhi - zlo + 1) > ny + 2 ? (zhi - zlo + 1) : ny + 2;
            dim3 grid((nmax + 2 + 127) / 128, rows, 7);
            launch_pdl(k_bound_all<double>, grid, dim3(128), 0, st, u, v, w, p, g, jmax, zlo, zhi, ilo, ihi, bot ? 1 : 0, top ? 1 : 0);
        }
        if (fused) {
            LaunchScope sc("ns_fgh_rhs", st);
            // z chunks: about 32 warps' worth of work per SM and chunk wave, not more (each chunk recomputes one plane
            // of H and loads one plane's full tap set)
            const int bx = (nx + 1 + 30) / 31, by = (ny + 1 + FGH_TK - 1) / FGH_TK;
            int nch = (fgh_chunks_per_sm() * device_sm_count() * 4 / FGH_TK + bx * by - 1) / (bx * by);
            if (nch < 1) nch = 1;
            if (nch > nzl) nch = nzl;
            const int zchunk = (nzl + nch - 1) / nch;
            dim3 block(32, FGH_TK);
            dim3 grid(bx, by, (nzl + zchunk - 1) / zchunk);
            launch_pdl(k_fgh_div<FGH_TK, FGH_MINB, 1>, grid, block, 0, st, u, v, w, p, F, G, H, R, g, ilo, ihi, zchunk);
        } else {
            {
                LaunchScope sc("ns_fgh", st);
                dim3 block(64, 4);
                const int i0 = lay.wlo[7];                 // first plane of H
                dim3 grid((nx + 1 + 63) / 64, (ny + 1 + 3) / 4, ihi - i0 + 1);
                launch_pdl(k_fgh<double>, grid, block, 0, st, u, v, w, F, G, H, g, i0, ilo);
            }
            {
                LaunchScope sc("ns_rhs", st);
                dim3 block(64, 4);
                dim3 grid((nx + 63) / 64, (ny + 3) / 4, nzl);
                launch_pdl(k_rhs<double>, grid, block, 0, st, F, G, H, p, R, g, ilo);
            }
        }
        FDMB_CHECK_LAUNCH();
        rc = lapl->solve_device(x.p, R.p, st);
        if (rc) return rc;
        if (nranks > 1) {
            // the neighbour above has written its x slab: fetch the plane the w update reads
            if ((rc = lapl->barrier(st))) return rc;
            if (!top) {
                int fld = 4, lo = ihi + 1, hi = ihi + 1, from = rank + 1;
                if ((rc = pull(&fld, &lo, &hi, &from, 1, st))) return rc;
            }
        }
        {
            LaunchScope sc("ns_update", st);
            dim3 block(64, 4);
            dim3 grid((nx + 63) / 64, (ny + 3) / 4, nzl);
            launch_pdl(k_update<double>, grid, block, 0, st, u, v, w, p, x, F, G, H, g, ilo);
        }
        FDMB_CHECK_LAUNCH();
    }
    return FDMB_OK;
}

// runs `body` with the handle's device current (several ranks may share one process)
template <typename Fn> static int on_device(fdmb_ns_cube* h, Fn body)
{
    int cur = 0;
    FDMB_CUDA(cudaGetDevice(&cur));
    if (cur != h->device) FDMB_CUDA(cudaSetDevice(h->device));
    int rc = body();
    if (cur != h->device) cudaSetDevice(cur);
    return rc;
}

extern "C" {

int fdmb_ns_cube_default_params(fdmb_ns_cube_params* p)
{
    if (!p) { set_error("null argument"); return FDMB_ERR_INVALID; }
    // defaults of ns_cube.h:47-61
    p->x1 = p->y1 = p->z1 = -M_PI;
    p->x2 = p->y2 = p->z2 = M_PI;
    p->u0 = 1.0; p->Re = 1.0; p->dt = 0.001;
    p->nx = 32; p->nz = 32; p->verbose = 0;
    return FDMB_OK;
}

static int ns_create(fdmb_ns_cube** out, const fdmb_ns_cube_params* p, int rank, int nranks)
{
    if (!out || !p) { set_error("null argument"); return FDMB_ERR_INVALID; }
    *out = nullptr;
    if (nranks != 1 && nranks != 2 && nranks != 4 && nranks != 8) {
        set_error("NSCube: nranks must be 1, 2, 4 or 8 (got %d)", nranks);
        return FDMB_ERR_INVALID;
    }
    if (rank < 0 || rank >= nranks) { set_error("NSCube: rank %d out of range", rank); return FDMB_ERR_INVALID; }
    auto* h = new (std::nothrow) fdmb_ns_cube();
    if (!h) { set_error("out of host memory"); return FDMB_ERR_NOMEM; }
    h->prm = *p; h->rank = rank; h->nranks = nranks;
    int rc = h->init();
    if (rc) { delete h; return rc; }
    *out = h;
    return FDMB_OK;
}

int fdmb_ns_cube_create(fdmb_ns_cube** out, const fdmb_ns_cube_params* p) { return ns_create(out, p, 0, 1); }

int fdmb_ns_cube_create_sharded(fdmb_ns_cube** out, const fdmb_ns_cube_params* p, int rank, int nranks)
{
    return ns_create(out, p, rank, nranks);
}

int fdmb_ns_cube_owned_planes(int nz, int field, int rank, int nranks, int* z_first, int* nplanes)
{
    if (field < 0 || field > 8 || !z_first || !nplanes || nz < 3 || nranks < 1 || rank < 0 || rank >= nranks ||
        (nranks > 1 && (!is_pow2(nz + 1) || !is_pow2(nranks) || (nz + 1) / nranks < 4))) {
        set_error("fdmb_ns_cube_owned_planes: bad argument (nz=%d field=%d rank=%d/%d)", nz, field, rank, nranks);
        return FDMB_ERR_INVALID;
    }
    int wlo, whi, olo, ohi;
    ns_planes(field, nz, rank, nranks, &wlo, &whi, &olo, &ohi);
    *z_first = olo; *nplanes = ohi - olo + 1;
    return FDMB_OK;
}

int fdmb_ns_cube_local_planes(fdmb_ns_cube* h, int field, int* z_first, int* nplanes)
{
    if (!h || field < 0 || field > 8 || !z_first || !nplanes) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    *z_first = h->lay.olo[field]; *nplanes = h->lay.ohi[field] - h->lay.olo[field] + 1;
    return FDMB_OK;
}

int fdmb_ns_cube_export_ipc(fdmb_ns_cube* h, void* handles)
{
    if (!h || !handles || h->nranks < 2) { set_error("export_ipc needs a sharded handle"); return FDMB_ERR_INVALID; }
    int rc = fdmb_lapl_cube_export_ipc(h->lapl, handles);
    if (rc) return rc;
    cudaIpcMemHandle_t ih;
    FDMB_CUDA(cudaIpcGetMemHandle(&ih, h->block));
    memcpy(static_cast<char*>(handles) + FDMB_IPC_HANDLE_BYTES, &ih, sizeof(ih));
    return FDMB_OK;
}

// handles: nranks records of 2 * FDMB_IPC_HANDLE_BYTES (solver block, field block), indexed by rank
int fdmb_ns_cube_attach_ipc(fdmb_ns_cube* h, const void* handles)
{
    if (!h || !handles || h->nranks < 2) { set_error("attach_ipc needs a sharded handle"); return FDMB_ERR_INVALID; }
    char solver[FDMB_MAX_RANKS * FDMB_IPC_HANDLE_BYTES];
    for (int q = 0; q < h->nranks; q++)
        memcpy(solver + (size_t)q * FDMB_IPC_HANDLE_BYTES, static_cast<const char*>(handles) + (size_t)q * 2 * FDMB_IPC_HANDLE_BYTES,
               FDMB_IPC_HANDLE_BYTES);
    int rc = fdmb_lapl_cube_attach_ipc(h->lapl, solver);
    if (rc) return rc;
    for (int q = 0; q < h->nranks; q++) {
        if (q == h->rank) continue;
        cudaIpcMemHandle_t ih;
        memcpy(&ih, static_cast<const char*>(handles) + (size_t)q * 2 * FDMB_IPC_HANDLE_BYTES + FDMB_IPC_HANDLE_BYTES, sizeof(ih));
        cudaError_t e = cudaIpcOpenMemHandle(&h->peer_block[q], ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("cudaIpcOpenMemHandle for rank %d failed: %s", q, cudaGetErrorString(e));
            return FDMB_ERR_COMM;
        }
        h->peer_ipc[q] = true;
    }
    h->attached = true;
    return FDMB_OK;
}

int fdmb_ns_cube_attach_local(fdmb_ns_cube* h, fdmb_ns_cube* const* all)
{
    if (!h || !all || h->nranks < 2) { set_error("attach_local needs a sharded handle"); return FDMB_ERR_INVALID; }
    fdmb_lapl_cube* solvers[FDMB_MAX_RANKS] = {};
    for (int q = 0; q < h->nranks; q++) {
        if (!all[q] || all[q]->nranks != h->nranks || all[q]->rank != q || all[q]->nz != h->nz || all[q]->nx != h->nx) {
            set_error("attach_local: handle %d does not belong to this sharded run", q);
            return FDMB_ERR_INVALID;
        }
        solvers[q] = all[q]->lapl;
    }
    int rc = fdmb_lapl_cube_attach_local(h->lapl, solvers);      // also enables peer access between the devices
    if (rc) return rc;
    for (int q = 0; q < h->nranks; q++) h->peer_block[q] = all[q]->block;
    h->attached = true;
    return FDMB_OK;
}

int fdmb_ns_cube_step(fdmb_ns_cube* h, int nsteps)
{
    if (!h || nsteps < 0) { set_error("bad argument"); return FDMB_ERR_INVALID; }
    return on_device(h, [&] {
        int rc = h->step(nsteps, h->stream);
        if (rc) return rc;
        FDMB_CUDA(cudaStreamSynchronize(h->stream));
        return (int)FDMB_OK;
    });
}

int fdmb_ns_cube_step_async(fdmb_ns_cube* h, int nsteps, void* stream)
{
    if (!h || nsteps < 0) { set_error("bad argument"); return FDMB_ERR_INVALID; }
    return on_device(h, [&] { return h->step(nsteps, stream ? (cudaStream_t)stream : h->stream); });
}

int fdmb_ns_cube_synchronize(fdmb_ns_cube* h)
{
    if (!h) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return on_device(h, [&] {
        FDMB_CUDA(cudaStreamSynchronize(h->stream));
        return (int)FDMB_OK;
    });
}

int fdmb_ns_cube_field_size(fdmb_ns_cube* h, int field, long long* count)
{
    if (!h || field < 0 || field > 8 || !count) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    *count = h->owned_count(field);
    return FDMB_OK;
}

int fdmb_ns_cube_get_field(fdmb_ns_cube* h, int field, double* host)
{
    if (!h || field < 0 || field > 8 || !host) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    return on_device(h, [&] {
        FDMB_CUDA(cudaMemcpyAsync(host, h->owned_ptr(field), sizeof(double) * h->owned_count(field), cudaMemcpyDeviceToHost,
                                  h->stream));
        FDMB_CUDA(cudaStreamSynchronize(h->stream));
        return (int)FDMB_OK;
    });
}

int fdmb_ns_cube_set_field(fdmb_ns_cube* h, int field, const double* host)
{
    if (!h || field < 0 || field > 8 || !host) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    return on_device(h, [&] {
        FDMB_CUDA(cudaMemcpyAsync(h->owned_ptr(field), host, sizeof(double) * h->owned_count(field), cudaMemcpyHostToDevice,
                                  h->stream));
        FDMB_CUDA(cudaStreamSynchronize(h->stream));
        return (int)FDMB_OK;
    });
}

int fdmb_ns_cube_field_device_ptr(fdmb_ns_cube* h, int field, void** dptr)
{
    if (!h || field < 0 || field > 8 || !dptr) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    *dptr = h->owned_ptr(field);
    return FDMB_OK;
}

long long fdmb_ns_cube_time_index(fdmb_ns_cube* h) { return h ? h->time_index : -1; }

int fdmb_ns_cube_destroy(fdmb_ns_cube* h)
{
    delete h;
    return FDMB_OK;
}

}  // extern "C"

// =====================================================================================================================
// NSCube<float> (reference instantiations src/ns_cube.cpp:281-282).  Float STORAGE for all nine fields and a float
// pressure solve (fdmb_lapl_cube_f32: the reference's member is LaplCube<T,check>, src/ns_cube.h:34), so a step moves
// half the bytes; the stencil arithmetic is double like the reference's own mixed expressions (see FldT).  Single GPU.
// =====================================================================================================================
typedef struct fdmb_lapl_cube_f32 fdmb_lapl_cube_f32;
extern "C" {
int fdmb_lapl_cube_f32_create(fdmb_lapl_cube_f32** h, double dx, double dy, double dz, double lx, double ly, double lz,
                              int nx, int ny, int nz, int periodic);
int fdmb_lapl_cube_f32_solve_device(fdmb_lapl_cube_f32* h, float* d_ans, const float* d_rhs, void* stream);
int fdmb_lapl_cube_f32_destroy(fdmb_lapl_cube_f32* h);
}

struct fdmb_ns_cube_f32 {
    fdmb_ns_cube_params prm{};
    int nx = 0, ny = 0, nz = 0;
    NSGeom g{};
    FldT<float> f[9]{};
    long long count[9] = {};
    fdmb_lapl_cube_f32* lapl = nullptr;
    cudaStream_t stream = nullptr;
    long long time_index = 0;
    void* block = nullptr;
    StepGraph graph;

    int init();
    int step_once(cudaStream_t st);
    ~fdmb_ns_cube_f32()
    {
        cudaFree(block);
        if (lapl) fdmb_lapl_cube_f32_destroy(lapl);
        if (stream) cudaStreamDestroy(stream);
    }
};

int fdmb_ns_cube_f32::init()
{
    nx = prm.nx; ny = prm.nx /* ns_cube.h:58 */; nz = prm.nz;
    if (nx < 3 || nz < 3) { set_error("NSCube<float>: nx, nz must be >= 3"); return FDMB_ERR_INVALID; }
    const double dx = (prm.x2 - prm.x1) / nx, dy = (prm.y2 - prm.y1) / ny, dz = (prm.z2 - prm.z1) / nz;
    const double dx2 = dx * dx, dy2 = dy * dy, dz2 = dz * dz;
    int rc = fdmb_lapl_cube_f32_create(&lapl, dx, dy, dz, prm.x2 - prm.x1 + dx, prm.y2 - prm.y1 + dy, prm.z2 - prm.z1 + dz,
                                       nx, ny, nz, 0);
    if (rc) return rc;
    FDMB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    NSLayout lay;
    ns_layout(nx, ny, nz, 0, 1, &lay);
    size_t off[9], o = 0;
    for (int k = 0; k < 9; k++) { off[k] = o; count[k] = lay.count[k]; o += (sizeof(float) * (size_t)lay.count[k] + 255) & ~(size_t)255; }
    FDMB_CUDA(cudaMalloc(&block, o));
    FDMB_CUDA(cudaMemset(block, 0, o));
    FDMB_CUDA(cudaDeviceSynchronize());
    static const int Y0[9] = {0, -1, 0, 0, 1, 1, 0, 1, 1}, X0[9] = {-1, 0, 0, 0, 1, 0, 1, 1, 1};
    for (int k = 0; k < 9; k++) {
        f[k].p = reinterpret_cast<float*>(static_cast<char*>(block) + off[k]);
        f[k].lz = lay.wlo[k]; f[k].ly = Y0[k]; f[k].lx = X0[k];
        f[k].sy = lay.sy[k]; f[k].sz = lay.sz[k];
    }
    g.nx = nx; g.ny = ny; g.nz = nz; g.U0 = prm.u0; g.dt = prm.dt;
    const double Re = prm.Re;
    g.cRx = 1.0 / Re / dx2; g.cRy = 1.0 / Re / dy2; g.cRz = 1.0 / Re / dz2;
    g.idx = 1.0 / dx; g.idy = 1.0 / dy; g.idz = 1.0 / dz;
    g.iRdx = 1.0 / Re / dx; g.iRdy = 1.0 / Re / dy; g.iRdz = 1.0 / Re / dz;
    g.idx2 = 1.0 / dx2; g.idy2 = 1.0 / dy2; g.idz2 = 1.0 / dz2;
    g.idt = 1.0 / prm.dt;
    g.dtdx = prm.dt / dx; g.dtdy = prm.dt / dy; g.dtdz = prm.dt / dz;
    return FDMB_OK;
}

int fdmb_ns_cube_f32::step_once(cudaStream_t st)
{
    const FldT<float> &u = f[0], &v = f[1], &w = f[2], &p = f[3], &x = f[4], &F = f[5], &G = f[6], &H = f[7], &R = f[8];
    const int nmax = nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz);
    {
        int jmax = (nz + 1 < nx + 1) ? nz + 1 : nx + 1;
        LaunchScope sc("ns32_bound_lid", st);
        k_bound_lid<float><<<dim3((jmax + 2 + 127) / 128, ny + 2), 128, 0, st>>>(u, g, jmax);
    }
    {
        LaunchScope sc("ns32_bound_mirror", st);
        const int rows = (nz + 2) > ny + 2 ? (nz + 2) : ny + 2;
        k_bound_mirror<float><<<dim3((nmax + 2 + 127) / 128, rows, 3), 128, 0, st>>>(u, v, w, g, 0, nz + 1, 1, 1);
    }
    {
        LaunchScope sc("ns32_bound_p", st);
        const int rows = nz > ny ? nz : ny;
        k_bound_p<float><<<dim3((nmax + 127) / 128, rows, 3), 128, 0, st>>>(u, v, w, p, g, 1, nz, 1, 1);
    }
    {
        LaunchScope sc("ns32_fgh", st);
        k_fgh<float><<<dim3((nx + 1 + 63) / 64, (ny + 1 + 3) / 4, nz + 1), dim3(64, 4), 0, st>>>(u, v, w, F, G, H, g, 0, 1);
    }
    {
        LaunchScope sc("ns32_rhs", st);
        k_rhs<float><<<dim3((nx + 63) / 64, (ny + 3) / 4, nz), dim3(64, 4), 0, st>>>(F, G, H, p, R, g, 1);
    }
    FDMB_CHECK_LAUNCH();
    int rc = fdmb_lapl_cube_f32_solve_device(lapl, x.p, R.p, st);
    if (rc) return rc;
    {
        LaunchScope sc("ns32_update", st);
        k_update<float><<<dim3((nx + 63) / 64, (ny + 3) / 4, nz), dim3(64, 4), 0, st>>>(u, v, w, p, x, F, G, H, g, 1);
    }
    FDMB_CHECK_LAUNCH();
    return FDMB_OK;
}

extern "C" {

int fdmb_ns_cube_f32_create(fdmb_ns_cube_f32** out, const fdmb_ns_cube_params* p)
{
    if (!out || !p) { set_error("null argument"); return FDMB_ERR_INVALID; }
    *out = nullptr;
    auto* h = new (std::nothrow) fdmb_ns_cube_f32();
    if (!h) { set_error("out of host memory"); return FDMB_ERR_NOMEM; }
    h->prm = *p;
    int rc = h->init();
    if (rc) { delete h; return rc; }
    *out = h;
    return FDMB_OK;
}

int fdmb_ns_cube_f32_step(fdmb_ns_cube_f32* h, int nsteps)
{
    if (!h || nsteps < 0) { set_error("bad argument"); return FDMB_ERR_INVALID; }
    for (int s = 0; s < nsteps; s++) {
        int rc = h->graph.run(h->stream, h, nullptr, [&]() { return h->step_once(h->stream); });
        if (rc) { cudaStreamSynchronize(h->stream); return rc; }
        h->time_index++;
    }
    FDMB_CUDA(cudaStreamSynchronize(h->stream));
    return FDMB_OK;
}

int fdmb_ns_cube_f32_field_size(fdmb_ns_cube_f32* h, int field, long long* count)
{
    if (!h || field < 0 || field > 8 || !count) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    *count = h->count[field];
    return FDMB_OK;
}

int fdmb_ns_cube_f32_get_field(fdmb_ns_cube_f32* h, int field, float* host)
{
    if (!h || field < 0 || field > 8 || !host) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    FDMB_CUDA(cudaStreamSynchronize(h->stream));
    FDMB_CUDA(cudaMemcpy(host, h->f[field].p, sizeof(float) * (size_t)h->count[field], cudaMemcpyDeviceToHost));
    return FDMB_OK;
}

int fdmb_ns_cube_f32_set_field(fdmb_ns_cube_f32* h, int field, const float* host)
{
    if (!h || field < 0 || field > 8 || !host) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    FDMB_CUDA(cudaStreamSynchronize(h->stream));
    FDMB_CUDA(cudaMemcpy(h->f[field].p, host, sizeof(float) * (size_t)h->count[field], cudaMemcpyHostToDevice));
    return FDMB_OK;
}

long long fdmb_ns_cube_f32_time_index(fdmb_ns_cube_f32* h) { return h ? h->time_index : -1; }

int fdmb_ns_cube_f32_destroy(fdmb_ns_cube_f32* h)
{
    delete h;
    return FDMB_OK;
}

}  // extern "C"
