// NSCube on B200: incompressible Navier-Stokes in a box on a staggered (MAC) grid,
// explicit Euler predictor + pressure projection.  Replaces fdm::NSCube<double,check>
// (reference src/ns_cube.h:13-92, src/ns_cube.cpp:27-277).  The whole step is device
// resident; the state arrays keep the reference's extents and layout (ghosts included)
// so get/set_field are plain copies of what the reference exposes as ns.u.vec etc.
//
//   step() = init_bound (ns_cube.cpp:65-122)  -> k_bound_lid, k_bound_mirror, k_bound_p
//            FGH        (ns_cube.cpp:126-200) -> k_fgh
//            poisson    (ns_cube.cpp:204-238) -> k_rhs + LaplCube solve
//            update_uvwp(ns_cube.cpp:241-277) -> k_update
#include <cmath>
#include <new>

#include "common.h"
#include "lapl_cube.h"

namespace fdmb {

// offset-indexed 3-D view, row-major, last index fastest (src/tensor.h:207-219)
struct Fld {
    double* p;
    int lz, ly, lx;       // lowest index per axis
    long long sz, sy;     // strides (doubles)
    __host__ __device__ __forceinline__ double& at(int i, int k, int j) const
    {
        return p[(long long)(i - lz) * sz + (long long)(k - ly) * sy + (j - lx)];
    }
};

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
__global__ void k_bound_lid(Fld u, NSGeom g, int jmax)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x - 1;
    int k = blockIdx.y;
    if (j > jmax) return;
    u.at(g.nz + 1, k, j) = 2 * g.U0 - u.at(g.nz, k, j);
}

// mirror ghosts (ns_cube.cpp:76-95).  blockIdx.z selects the field.
__global__ void k_bound_mirror(Fld u, Fld v, Fld w, NSGeom g)
{
    int a = blockIdx.x * blockDim.x + threadIdx.x;   // fast index of the face
    int b = blockIdx.y;                              // slow index of the face
    if (blockIdx.z == 0) {          // u: i = b in 0..nz+1, k = a in 0..ny+1
        if (b <= g.nz + 1 && a <= g.ny + 1) {
            u.at(b, a, -1) = u.at(b, a, 1);
            u.at(b, a, g.nx + 1) = u.at(b, a, g.nx - 1);
        }
    } else if (blockIdx.z == 1) {   // v: i = b in 0..nz+1, j = a in 0..nx+1
        if (b <= g.nz + 1 && a <= g.nx + 1) {
            v.at(b, -1, a) = v.at(b, 1, a);
            v.at(b, g.ny + 1, a) = v.at(b, g.ny - 1, a);
        }
    } else {                        // w: k = b in 0..ny+1, j = a in 0..nx+1
        if (b <= g.ny + 1 && a <= g.nx + 1) {
            w.at(-1, b, a) = w.at(1, b, a);
            w.at(g.nz + 1, b, a) = w.at(g.nz - 1, b, a);
        }
    }
}

// pressure ghosts (ns_cube.cpp:98-121)
__global__ void k_bound_p(Fld u, Fld v, Fld w, Fld p, NSGeom g)
{
    int a = blockIdx.x * blockDim.x + threadIdx.x + 1;
    int b = blockIdx.y + 1;
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    if (blockIdx.z == 0) {          // x faces: i = b in 1..nz, k = a in 1..ny
        if (b <= nz && a <= ny) {
            int i = b, k = a;
            p.at(i, k, 0) = p.at(i, k, 1) - (u.at(i, k, 1) - 2 * u.at(i, k, 0) + u.at(i, k, -1)) * g.iRdx;
            p.at(i, k, nx + 1) = p.at(i, k, nx) - (u.at(i, k, nx + 1) - 2 * u.at(i, k, nx) + u.at(i, k, nx - 1)) * g.iRdx;
        }
    } else if (blockIdx.z == 1) {   // y faces: i = b in 1..nz, j = a in 1..nx
        if (b <= nz && a <= nx) {
            int i = b, j = a;
            p.at(i, 0, j) = p.at(i, 1, j) - (v.at(i, 1, j) - 2 * v.at(i, 0, j) + v.at(i, -1, j)) * g.iRdy;
            p.at(i, ny + 1, j) = p.at(i, ny, j) - (v.at(i, ny + 1, j) - 2 * v.at(i, ny, j) + v.at(i, ny - 1, j)) * g.iRdy;
        }
    } else {                        // z faces: k = b in 1..ny, j = a in 1..nx
        if (b <= ny && a <= nx) {
            int k = b, j = a;
            p.at(0, k, j) = p.at(1, k, j) - (w.at(1, k, j) - 2 * w.at(0, k, j) + w.at(-1, k, j)) * g.iRdz;
            p.at(nz + 1, k, j) = p.at(nz, k, j) - (w.at(nz + 1, k, j) - 2 * w.at(nz, k, j) + w.at(nz - 1, k, j)) * g.iRdz;
        }
    }
}

__device__ __forceinline__ double sq(double x) { return x * x; }

// linear element offset of (i,k,j); every field of the graded sizes has < 2^31 elements per axis product
__device__ __forceinline__ long long lin(const Fld& f, int i, int k, int j)
{
    return (long long)(i - f.lz) * f.sz + (long long)(k - f.ly) * f.sy + (j - f.lx);
}

// ---- FGH (ns_cube.cpp:126-200) ----------------------------------------------------------
// One thread per (i,k,j) in [0..nz]x[0..ny]x[0..nx]; F where i,k>=1, G where i,j>=1, H where k,j>=1.
// Interior threads (i,k,j >= 1) read the 27 distinct taps of the three stencils once through
// row pointers (immediate offsets, no per-tap address arithmetic) and produce F, G and H together;
// the O(n^2) edge threads take the generic path.
template <bool F_, bool G_, bool H_>
__device__ __forceinline__ void fgh_generic(const Fld& u, const Fld& v, const Fld& w, const Fld& F, const Fld& G,
                                            const Fld& H, const NSGeom& g, int i, int k, int j)
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

__global__ void __launch_bounds__(256) k_fgh(Fld u, Fld v, Fld w, Fld F, Fld G, Fld H, NSGeom g)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = blockIdx.z;
    if (j > g.nx || k > g.ny) return;
    if (i >= 1 && k >= 1 && j >= 1) {
        const double* __restrict__ uc_ = u.p + lin(u, i, k, j);
        const double* __restrict__ vc_ = v.p + lin(v, i, k, j);
        const double* __restrict__ wc_ = w.p + lin(w, i, k, j);
        const double* __restrict__ ukp = uc_ + u.sy; const double* __restrict__ ukm = uc_ - u.sy;
        const double* __restrict__ uip = uc_ + u.sz; const double* __restrict__ uim = uc_ - u.sz;
        const double* __restrict__ vkp = vc_ + v.sy; const double* __restrict__ vkm = vc_ - v.sy;
        const double* __restrict__ vip = vc_ + v.sz; const double* __restrict__ vim = vc_ - v.sz;
        const double* __restrict__ vipkm = vip - v.sy;
        const double* __restrict__ wkp = wc_ + w.sy; const double* __restrict__ wkm = wc_ - w.sy;
        const double* __restrict__ wip = wc_ + w.sz; const double* __restrict__ wim = wc_ - w.sz;
        const double* __restrict__ wimkp = wim + w.sy;
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
    if (i >= 1 && k >= 1) fgh_generic<true, false, false>(u, v, w, F, G, H, g, i, k, j);
    if (i >= 1 && j >= 1) fgh_generic<false, true, false>(u, v, w, F, G, H, g, i, k, j);
    if (k >= 1 && j >= 1) fgh_generic<false, false, true>(u, v, w, F, G, H, g, i, k, j);
}

// ---- poisson RHS (ns_cube.cpp:205-235) ---------------------------------------------------
__global__ void __launch_bounds__(256) k_rhs(Fld F, Fld G, Fld H, Fld p, Fld R, NSGeom g)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int k = blockIdx.y * blockDim.y + threadIdx.y + 1;
    const int i = blockIdx.z + 1;
    if (j > g.nx || k > g.ny) return;
    const double* __restrict__ Fp = F.p + lin(F, i, k, j);
    const double* __restrict__ Gp = G.p + lin(G, i, k, j);
    const double* __restrict__ Hp = H.p + lin(H, i, k, j);
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
__global__ void __launch_bounds__(256) k_update(Fld u, Fld v, Fld w, Fld p, Fld x, Fld F, Fld G, Fld H, NSGeom g)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int k = blockIdx.y * blockDim.y + threadIdx.y + 1;
    const int i = blockIdx.z + 1;
    if (j > g.nx || k > g.ny) return;
    const double* __restrict__ xp = x.p + lin(x, i, k, j);
    const double xc = xp[0];
    if (j < g.nx) u.p[lin(u, i, k, j)] = F.p[lin(F, i, k, j)] - g.dtdx * (xp[1] - xc);
    if (k < g.ny) v.p[lin(v, i, k, j)] = G.p[lin(G, i, k, j)] - g.dtdy * (xp[x.sy] - xc);
    if (i < g.nz) w.p[lin(w, i, k, j)] = H.p[lin(H, i, k, j)] - g.dtdz * (xp[x.sz] - xc);
    p.p[lin(p, i, k, j)] = xc;   // p = x copies the index-range intersection (tensor.h:103-111)
}

}  // namespace fdmb

using namespace fdmb;

struct fdmb_ns_cube {
    fdmb_ns_cube_params prm{};
    int nx = 0, ny = 0, nz = 0;
    double dx = 0, dy = 0, dz = 0;
    NSGeom g{};
    Fld f[9]{};                 // u v w p x F G H RHS
    long long count[9]{};
    fdmb_lapl_cube* lapl = nullptr;
    cudaStream_t stream = nullptr;
    long long time_index = 0;

    int init();
    int step(int nsteps, cudaStream_t st);
    ~fdmb_ns_cube();
};

static int make_field(Fld& f, long long& count, int z0, int z1, int y0, int y1, int x0, int x1)
{
    f.lz = z0; f.ly = y0; f.lx = x0;
    f.sy = x1 - x0 + 1;
    f.sz = (long long)(y1 - y0 + 1) * f.sy;
    count = (long long)(z1 - z0 + 1) * f.sz;
    FDMB_CUDA(cudaMalloc(&f.p, sizeof(double) * count));
    FDMB_CUDA(cudaMemset(f.p, 0, sizeof(double) * count));
    return FDMB_OK;
}

int fdmb_ns_cube::init()
{
    nx = prm.nx; ny = prm.nx /* ns_cube.h:58: ny is read from key "nx" */; nz = prm.nz;
    if (nx < 3 || nz < 3) { set_error("NSCube: nx, nz must be >= 3"); return FDMB_ERR_INVALID; }
    dx = (prm.x2 - prm.x1) / nx; dy = (prm.y2 - prm.y1) / ny; dz = (prm.z2 - prm.z1) / nz;
    const double dx2 = dx * dx, dy2 = dy * dy, dz2 = dz * dz;
    int rc = fdmb_lapl_cube_create(&lapl, dx, dy, dz, prm.x2 - prm.x1 + dx, prm.y2 - prm.y1 + dy,
                                   prm.z2 - prm.z1 + dz, nx, ny, nz, 0);   // ns_cube.h:77
    if (rc) return rc;
    FDMB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    // extents: ns_cube.h:66-75
    if ((rc = make_field(f[0], count[0], 0, nz + 1, 0, ny + 1, -1, nx + 1))) return rc;  // u
    if ((rc = make_field(f[1], count[1], 0, nz + 1, -1, ny + 1, 0, nx + 1))) return rc;  // v
    if ((rc = make_field(f[2], count[2], -1, nz + 1, 0, ny + 1, 0, nx + 1))) return rc;  // w
    if ((rc = make_field(f[3], count[3], 0, nz + 1, 0, ny + 1, 0, nx + 1))) return rc;   // p
    if ((rc = make_field(f[4], count[4], 1, nz, 1, ny, 1, nx))) return rc;               // x
    if ((rc = make_field(f[5], count[5], 1, nz, 1, ny, 0, nx))) return rc;               // F
    if ((rc = make_field(f[6], count[6], 1, nz, 0, ny, 1, nx))) return rc;               // G
    if ((rc = make_field(f[7], count[7], 0, nz, 1, ny, 1, nx))) return rc;               // H
    if ((rc = make_field(f[8], count[8], 1, nz, 1, ny, 1, nx))) return rc;               // RHS
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
    for (auto& a : f) cudaFree(a.p);
    if (lapl) fdmb_lapl_cube_destroy(lapl);
    if (stream) cudaStreamDestroy(stream);
}

int fdmb_ns_cube::step(int nsteps, cudaStream_t st)
{
    const Fld &u = f[0], &v = f[1], &w = f[2], &p = f[3], &x = f[4], &F = f[5], &G = f[6], &H = f[7], &R = f[8];
    const int nmax = nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz);
    for (int s = 0; s < nsteps; s++) {
        {   // the reference loops j = -1..nz+1 (ns_cube.cpp:68); clamp to the allocated x range
            int jmax = (nz + 1 < nx + 1) ? nz + 1 : nx + 1;
            LaunchScope sc("ns_bound_lid", st);
            dim3 grid((jmax + 2 + 127) / 128, ny + 2);
            k_bound_lid<<<grid, 128, 0, st>>>(u, g, jmax);
        }
        {
            LaunchScope sc("ns_bound_mirror", st);
            dim3 grid((nmax + 2 + 127) / 128, nmax + 2, 3);
            k_bound_mirror<<<grid, 128, 0, st>>>(u, v, w, g);
        }
        {
            LaunchScope sc("ns_bound_p", st);
            dim3 grid((nmax + 127) / 128, nmax, 3);
            k_bound_p<<<grid, 128, 0, st>>>(u, v, w, p, g);
        }
        {
            LaunchScope sc("ns_fgh", st);
            dim3 block(64, 4);
            dim3 grid((nx + 1 + 63) / 64, (ny + 1 + 3) / 4, nz + 1);
            k_fgh<<<grid, block, 0, st>>>(u, v, w, F, G, H, g);
        }
        {
            LaunchScope sc("ns_rhs", st);
            dim3 block(64, 4);
            dim3 grid((nx + 63) / 64, (ny + 3) / 4, nz);
            k_rhs<<<grid, block, 0, st>>>(F, G, H, p, R, g);
        }
        FDMB_CHECK_LAUNCH();
        int rc = lapl->solve_device(x.p, R.p, st);
        if (rc) return rc;
        {
            LaunchScope sc("ns_update", st);
            dim3 block(64, 4);
            dim3 grid((nx + 63) / 64, (ny + 3) / 4, nz);
            k_update<<<grid, block, 0, st>>>(u, v, w, p, x, F, G, H, g);
        }
        FDMB_CHECK_LAUNCH();
        time_index++;
    }
    return FDMB_OK;
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

int fdmb_ns_cube_create(fdmb_ns_cube** out, const fdmb_ns_cube_params* p)
{
    if (!out || !p) { set_error("null argument"); return FDMB_ERR_INVALID; }
    *out = nullptr;
    auto* h = new (std::nothrow) fdmb_ns_cube();
    if (!h) { set_error("out of host memory"); return FDMB_ERR_NOMEM; }
    h->prm = *p;
    int rc = h->init();
    if (rc) { delete h; return rc; }
    *out = h;
    return FDMB_OK;
}

int fdmb_ns_cube_step(fdmb_ns_cube* h, int nsteps)
{
    if (!h || nsteps < 0) { set_error("bad argument"); return FDMB_ERR_INVALID; }
    int rc = h->step(nsteps, h->stream);
    if (rc) return rc;
    FDMB_CUDA(cudaStreamSynchronize(h->stream));
    return FDMB_OK;
}

int fdmb_ns_cube_step_async(fdmb_ns_cube* h, int nsteps, void* stream)
{
    if (!h || nsteps < 0) { set_error("bad argument"); return FDMB_ERR_INVALID; }
    return h->step(nsteps, stream ? (cudaStream_t)stream : h->stream);
}

int fdmb_ns_cube_field_size(fdmb_ns_cube* h, int field, long long* count)
{
    if (!h || field < 0 || field > 8 || !count) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    *count = h->count[field];
    return FDMB_OK;
}

int fdmb_ns_cube_get_field(fdmb_ns_cube* h, int field, double* host)
{
    if (!h || field < 0 || field > 8 || !host) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    FDMB_CUDA(cudaMemcpyAsync(host, h->f[field].p, sizeof(double) * h->count[field], cudaMemcpyDeviceToHost, h->stream));
    FDMB_CUDA(cudaStreamSynchronize(h->stream));
    return FDMB_OK;
}

int fdmb_ns_cube_set_field(fdmb_ns_cube* h, int field, const double* host)
{
    if (!h || field < 0 || field > 8 || !host) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    FDMB_CUDA(cudaMemcpyAsync(h->f[field].p, host, sizeof(double) * h->count[field], cudaMemcpyHostToDevice, h->stream));
    FDMB_CUDA(cudaStreamSynchronize(h->stream));
    return FDMB_OK;
}

int fdmb_ns_cube_field_device_ptr(fdmb_ns_cube* h, int field, void** dptr)
{
    if (!h || field < 0 || field > 8 || !dptr) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    *dptr = h->f[field].p;
    return FDMB_OK;
}

long long fdmb_ns_cube_time_index(fdmb_ns_cube* h) { return h ? h->time_index : -1; }

int fdmb_ns_cube_destroy(fdmb_ns_cube* h)
{
    delete h;
    return FDMB_OK;
}

}  // extern "C"
