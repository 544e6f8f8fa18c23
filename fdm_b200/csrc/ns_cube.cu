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
//   step() = init_bound (ns_cube.cpp:65-122)  -> k_bound_lid, k_bound_mirror, k_bound_p
//            FGH        (ns_cube.cpp:126-200) -> k_fgh
//            poisson    (ns_cube.cpp:204-238) -> k_rhs + LaplCube solve
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
};

// ---- init_bound ---------------------------------------------------------------------
// lid (ns_cube.cpp:67-72): u[nz+1][k][j] = 2 U0 - u[nz][k][j], k=0..ny+1, j=-1..jmax
template <typename T> __global__ void k_bound_lid(FldT<T> u, NSGeom g, int jmax)
{
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
// z-marching, shared-memory staged.  A block owns a (TK-1) x (TJ-1) core of (k, j) points and marches over a chunk of
// z planes; its TK x TJ threads cover the core plus one row below and one column to the left, where only the G / F
// values the core's divergence needs are produced.  Per plane:
//   * the (TK+2) x (TJ+2) halo tiles of u, v, w of planes i-1, i, i+1 sit in a ring of FOUR shared-memory planes per
//     field: every value is fetched from global memory once per block, and plane i+2 streams into the fourth slot with
//     asynchronous copies (cp.async, zero-filled outside the field) while plane i is computed;
//   * every thread evaluates F, G, H at its point from the staged taps and publishes F, G in shared memory;
//   * the core threads form RHS = ((F - F[j-1])/dx + (G - G[k-1])/dy + (H - H[i-1])/dz)/dt - ghost pressures, with H[i-1]
//     carried in a register from the previous plane (the chunk's first plane i = ia - 1 only produces it).
// Traffic: u, v, w read once, F, G, H, RHS written once = 56 B per point, against 80 B for k_fgh + k_rhs.
// The same kernel serves a z-slab of the sharded step: it starts one plane below the slab like k_fgh did.
template <int TJ, int TK>
__global__ void __launch_bounds__(TJ * TK, 2)
k_fgh_rhs(Fld u, Fld v, Fld w, Fld p, Fld F, Fld G, Fld H, Fld R, NSGeom g, int ia0, int ib0, int zchunk)
{
    constexpr int NT = TJ * TK;
    constexpr int PW = TJ + 2 + 1;                 // tile pitch (odd: rows start in different banks)
    constexpr int PLANE = (TK + 2) * PW;
    extern __shared__ double fgh_smem[];
    double (*su)[PLANE] = reinterpret_cast<double (*)[PLANE]>(fgh_smem);
    double (*sv)[PLANE] = su + 4;
    double (*sw)[PLANE] = sv + 4;
    double (*sF)[TJ + 1] = reinterpret_cast<double (*)[TJ + 1]>(fgh_smem + 12 * PLANE);
    double (*sG)[TJ + 1] = sF + TK;
    const int tj = threadIdx.x, tk = threadIdx.y, tid = tk * TJ + tj;
    const int j0 = 1 + blockIdx.x * (TJ - 1), k0 = 1 + blockIdx.y * (TK - 1);
    const int ia = ia0 + blockIdx.z * zchunk;
    const int ib = (ia + zchunk - 1 < ib0) ? ia + zchunk - 1 : ib0;
    const int j = j0 - 1 + tj, k = k0 - 1 + tk;
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    // global extents of the three fields (ns_cube.h:66-68); z: the planes this rank holds start at f.lz
    // one element of a halo tile: asynchronous 8-byte copy, zero-filled outside the field (z: the planes this rank
    // holds start at f.lz)
    auto stage = [&](double* dst, const Fld& f, int i, int kk, int jj) {
        const bool in = i >= f.lz && i <= nz + 1 && kk >= f.ly && kk <= ny + 1 && jj >= f.lx && jj <= nx + 1;
        const double* src = in ? &f.at(i, kk, jj) : f.p;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src),
                     "r"(in ? 8 : 0)
                     : "memory");
    };
    auto load_plane = [&](int i) {
        const int s = (i + 4) & 3;
        for (int e = tid; e < (TK + 2) * (TJ + 2); e += NT) {
            const int r = e / (TJ + 2), c = e - r * (TJ + 2);
            const int kk = k0 - 2 + r, jj = j0 - 2 + c, off = r * PW + c;
            stage(&su[s][off], u, i, kk, jj);
            stage(&sv[s][off], v, i, kk, jj);
            stage(&sw[s][off], w, i, kk, jj);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_plane(ia - 2); load_plane(ia - 1); load_plane(ia);
    const bool core = tk >= 1 && tj >= 1 && k <= ny && j <= nx;
    const bool doF = k >= 1 && k <= ny && j <= nx;            // j >= 0 always
    const bool doG = j >= 1 && j <= nx && k <= ny;            // k >= 0 always
    double hprev = 0.0;
    for (int i = ia - 1; i <= ib; i++) {
        // planes up to i + 1 have landed for everybody; plane i + 2 (needed from the next iteration on) streams into the
        // slot plane i - 2 left, which nobody reads any more (the barrier at the end of the previous iteration)
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (i + 1 <= ib) load_plane(i + 2);
        double fv = 0.0, gv = 0.0, hv = 0.0;
        const bool fg = i >= ia;
        {
            // the 27 distinct taps of the three stencils, read once (every thread's 3 x 3 x 3 neighbourhood lies inside
            // the staged tiles), named by (di,dk,dj) with m = -1, p = +1 like k_fgh
            const int c0 = (tk + 1) * PW + (tj + 1);
            const double* __restrict__ uc_ = su[(i + 4) & 3] + c0;
            const double* __restrict__ uip = su[(i + 5) & 3] + c0;
            const double* __restrict__ uim = su[(i + 3) & 3] + c0;
            const double* __restrict__ vc_ = sv[(i + 4) & 3] + c0;
            const double* __restrict__ vip = sv[(i + 5) & 3] + c0;
            const double* __restrict__ vim = sv[(i + 3) & 3] + c0;
            const double* __restrict__ wc_ = sw[(i + 4) & 3] + c0;
            const double* __restrict__ wip = sw[(i + 5) & 3] + c0;
            const double* __restrict__ wim = sw[(i + 3) & 3] + c0;
            const double u000 = uc_[0], u00p = uc_[1], u00m = uc_[-1], u0p0 = uc_[PW], u0pm = uc_[PW - 1], u0m0 = uc_[-PW];
            const double up00 = uip[0], up0m = uip[-1], um00 = uim[0];
            const double v000 = vc_[0], v00p = vc_[1], v00m = vc_[-1], v0p0 = vc_[PW], v0m0 = vc_[-PW], v0mp = vc_[-PW + 1];
            const double vp00 = vip[0], vpm0 = vip[-PW], vm00 = vim[0];
            const double w000 = wc_[0], w00p = wc_[1], w00m = wc_[-1], w0p0 = wc_[PW], w0m0 = wc_[-PW];
            const double wp00 = wip[0], wm00 = wim[0], wm0p = wim[1], wmp0 = wim[PW];
            if (fg && doF)
                fv = u000 + g.dt * (
                    (u00p - 2 * u000 + u00m) * g.cRx +
                    (u0p0 - 2 * u000 + u0m0) * g.cRy +
                    (up00 - 2 * u000 + um00) * g.cRz -
                    (sq(0.5 * (u000 + u00p)) - sq(0.5 * (u00m + u000))) * g.idx -
                    0.25 * ((u000 + u0p0) * (v00p + v000) -
                            (u0m0 + u000) * (v0mp + v0m0)) * g.idy -
                    0.25 * ((u000 + up00) * (w00p + w000) -
                            (um00 + u000) * (wm0p + wm00)) * g.idz);
            if (fg && doG)
                gv = v000 + g.dt * (
                    (v00p - 2 * v000 + v00m) * g.cRx +
                    (v0p0 - 2 * v000 + v0m0) * g.cRy +
                    (vp00 - 2 * v000 + vm00) * g.cRz -
                    (sq(0.5 * (v000 + v0p0)) - sq(0.5 * (v0m0 + v000))) * g.idy -
                    0.25 * ((u000 + u0p0) * (v00p + v000) -
                            (u00m + u0pm) * (v000 + v00m)) * g.idx -
                    0.25 * ((w000 + w0p0) * (v000 + vp00) -
                            (wm00 + wmp0) * (vm00 + v000)) * g.idz);
            if (core)
                hv = w000 + g.dt * (
                    (w00p - 2 * w000 + w00m) * g.cRx +
                    (w0p0 - 2 * w000 + w0m0) * g.cRy +
                    (wp00 - 2 * w000 + wm00) * g.cRz -
                    (sq(0.5 * (wp00 + w000)) - sq(0.5 * (wm00 + w000))) * g.idz -
                    0.25 * ((up00 + u000) * (w00p + w000) -
                            (up0m + u00m) * (w000 + w00m)) * g.idx -
                    0.25 * ((w000 + w0p0) * (v000 + vp00) -
                            (w0m0 + w000) * (v0m0 + vpm0)) * g.idy);
        }
        sF[tk][tj] = fv; sG[tk][tj] = gv;
        __syncthreads();
        if (fg) {
            if (doF && (tj >= 1 || j == 0)) F.p[lin(F, i, k, j)] = fv;
            if (doG && (tk >= 1 || k == 0)) G.p[lin(G, i, k, j)] = gv;
        }
        if (core) {
            if (fg || i == 0) H.p[lin(H, i, k, j)] = hv;         // (plane ia - 1 belongs to the chunk below, except H[0])
            if (fg) {
                double r = ((fv - sF[tk][tj - 1]) * g.idx + (gv - sG[tk - 1][tj]) * g.idy + (hv - hprev) * g.idz) * g.idt;
                if (i <= 1 || k <= 1 || j <= 1 || j >= nx || k >= ny || i >= nz) {
                    if (i <= 1) r -= p.at(i - 1, k, j) * g.idz2;
                    if (k <= 1) r -= p.at(i, k - 1, j) * g.idy2;
                    if (j <= 1) r -= p.at(i, k, j - 1) * g.idx2;
                    if (j >= nx) r -= p.at(i, k, j + 1) * g.idx2;
                    if (k >= ny) r -= p.at(i, k + 1, j) * g.idy2;
                    if (i >= nz) r -= p.at(i + 1, k, j) * g.idz2;
                }
                R.p[lin(R, i, k, j)] = r;
            }
            hprev = hv;
        }
    }
}

template <int TJ, int TK> constexpr size_t fgh_rhs_smem()
{
    return sizeof(double) * (size_t)(12 * (TK + 2) * (TJ + 3) + 2 * TK * (TJ + 1));
}

// ---- poisson RHS (ns_cube.cpp:205-235) ---------------------------------------------------
template <typename T> __global__ void __launch_bounds__(256) k_rhs(FldT<T> F, FldT<T> G, FldT<T> H, FldT<T> p, FldT<T> R, NSGeom g, int ilo)
{
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

// FDMB_FGH_FUSED=1 selects the fused, shared-memory staged FGH + divergence sweep (k_fgh_rhs) for handles created
// afterwards.  Measured at 255^3 (r02j/k): 625 us against 279 + 109 us for k_fgh + k_rhs -- both variants are bound by
// instruction issue, not by DRAM (k_fgh: 34 % of peak DRAM throughput), so the 80 -> 56 B/pt of traffic the fusion
// saves buys nothing while the per-element staging and the two barriers per plane cost issue slots.  Off by default;
// parity-tested (tests/test_ns_cube_gpu.py::test_fused_fgh_rhs_matches).
static bool fused_fgh_rhs_requested()
{
    const char* e = getenv("FDMB_FGH_FUSED");
    return e && e[0] == '1';
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
    bool fused = false;         // FGH + divergence in one sweep (k_fgh_rhs)

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
        FDMB_CUDA(cudaFuncSetAttribute(k_fgh_rhs<64, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fgh_rhs_smem<64, 8>()));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_fgh_rhs<64, 8>));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_update<double>));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_bound_lid<double>));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_bound_mirror<double>));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_bound_p<double>));
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
        if (top) {   // the reference loops j = -1..nz+1 (ns_cube.cpp:68); clamp to the allocated x range
            int jmax = (nz + 1 < nx + 1) ? nz + 1 : nx + 1;
            LaunchScope sc("ns_bound_lid", st);
            dim3 grid((jmax + 2 + 127) / 128, ny + 2);
            k_bound_lid<double><<<grid, 128, 0, st>>>(u, g, jmax);
        }
        {
            LaunchScope sc("ns_bound_mirror", st);
            const int zlo = lay.wlo[0], zhi = lay.whi[0];
            const int rows = (zhi - zlo + 1) > ny + 2 ? (zhi - zlo + 1) : ny + 2;
            dim3 grid((nmax + 2 + 127) / 128, rows, 3);
            k_bound_mirror<double><<<grid, 128, 0, st>>>(u, v, w, g, zlo, zhi, bot ? 1 : 0, top ? 1 : 0);
        }
        {
            LaunchScope sc("ns_bound_p", st);
            const int rows = nzl > ny ? nzl : ny;
            dim3 grid((nmax + 127) / 128, rows, 3);
            k_bound_p<double><<<grid, 128, 0, st>>>(u, v, w, p, g, ilo, ihi, bot ? 1 : 0, top ? 1 : 0);
        }
        if (fused) {
            LaunchScope sc("ns_fgh_rhs", st);
            constexpr int TJ = 64, TK = 8;
            // z chunks: enough blocks for two waves of two resident blocks per SM, not more (each chunk recomputes one
            // plane of H and reloads two planes)
            const int tiles = ((nx + TJ - 2) / (TJ - 1)) * ((ny + TK - 2) / (TK - 1));
            int nch = (4 * device_sm_count() + tiles - 1) / tiles;
            if (nch < 1) nch = 1;
            if (nch > nzl) nch = nzl;
            const int zchunk = (nzl + nch - 1) / nch;
            dim3 block(TJ, TK);
            dim3 grid((nx + TJ - 2) / (TJ - 1), (ny + TK - 2) / (TK - 1), (nzl + zchunk - 1) / zchunk);
            k_fgh_rhs<TJ, TK><<<grid, block, fgh_rhs_smem<TJ, TK>(), st>>>(u, v, w, p, F, G, H, R, g, ilo, ihi, zchunk);
        } else {
            {
                LaunchScope sc("ns_fgh", st);
                dim3 block(64, 4);
                const int i0 = lay.wlo[7];                 // first plane of H
                dim3 grid((nx + 1 + 63) / 64, (ny + 1 + 3) / 4, ihi - i0 + 1);
                k_fgh<double><<<grid, block, 0, st>>>(u, v, w, F, G, H, g, i0, ilo);
            }
            {
                LaunchScope sc("ns_rhs", st);
                dim3 block(64, 4);
                dim3 grid((nx + 63) / 64, (ny + 3) / 4, nzl);
                k_rhs<double><<<grid, block, 0, st>>>(F, G, H, p, R, g, ilo);
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
            k_update<double><<<grid, block, 0, st>>>(u, v, w, p, x, F, G, H, g, ilo);
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
