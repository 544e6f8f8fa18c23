// NSCyl on B200: incompressible Navier-Stokes between two coaxial cylinders (Taylor-Couette) on a
// staggered grid in (phi, z, r), explicit Euler predictor + pressure projection.  Replaces
// fdm::NSCyl<double,check,zflag> (reference src/ns_cyl.h:17-132, src/ns_cyl.cpp:23-484).
// The whole step is device resident; the state arrays keep the reference's extents and layout
// (ghosts included, src/ns_cyl.h:80-93) so get/set_field are plain copies of ns.u.vec etc.
//
//   step()   = init_bound (ns_cyl.cpp:81-172)  -> k_cyl_bound_r, k_cyl_bound_z, k_cyl_bound_p
//              FGH        (ns_cyl.cpp:175-277) -> k_cyl_fgh<false>
//              poisson    (ns_cyl.cpp:408-442) -> k_cyl_rhs + LaplCyl3FFT2 solve
//              update_uvwp(ns_cyl.cpp:445-484) -> k_cyl_update
//   L_step() = the same with L_FGH (ns_cyl.cpp:280-405), linearised about u0, v0, w0 -> k_cyl_fgh<true>
//
// phi is periodic (slowest index), z is Dirichlet (zflag none) or periodic, r is fastest.
// Index wrap follows the reference accessor (src/tensor.h:119-123).
#include <cmath>
#include <new>
#include <random>
#include <vector>

#include <cstring>

#include "common.h"
#include "lapl_cyl.h"

namespace fdmb {

// offset-indexed 3-D view [phi][z][r], row-major, last index fastest (src/tensor.h:207-219)
struct CFld {
    double* p;
    int li, lz, lr;       // lowest phi / z / r index (li != 0: a rank's window of a sharded run, indexed globally;
                          // phi = -1 and phi = nphi are the wrap-around halo planes there)
    long long sp, sz;     // strides (doubles) of the phi and z axes
    __host__ __device__ __forceinline__ double& at(int i, int k, int j) const
    {
        return p[(long long)(i - li) * sp + (long long)(k - lz) * sz + (j - lr)];
    }
};

struct CylGeom {
    int nr, nz, nphi;
    int wrap;                       // 1: phi neighbours wrap inside the array (one GPU); 0: halo planes hold them
    int zper;                       // z periodic?
    int z_, z0, z1, zn, znn;        // ns_cyl.h:71-75
    double U0, dt;
    double cRr, cRz, cRp;           // 1/Re/dr2, 1/Re/dz2, 1/Re/dphi2
    double idr, idz, idphi;         // 1/dr, 1/dz, 1/dphi
    double iRe;                     // 1/Re
    double iRdr, iRdz;              // 1/Re/dr, 1/Re/dz
    double iRdphi;                  // 1/dphi/Re
    double idr2, idz2, idt;
    double dtdr, dtdz, dtdphi;      // dt/dr, dt/dz, dt/dphi
    // r-dependent factors (host tables, same formulas as the reference):
    //   at u points r = r0 + dr*j (j = 0..nr):   fr2, fr1, firr (1/r^2), fir (1/r)
    //   at cell centres r = r0 + dr*j - dr/2 (j = 0..nr+1):   cr2, cr1, cirr, cir, crp = r+dr/2, crm = r-dr/2
    const double *fr2, *fr1, *firr, *fir;
    const double *cr2, *cr1, *cirr, *cir, *crp, *crm;
};

// ---- init_bound (ns_cyl.cpp:81-172) --------------------------------------------------------------
// r-direction ghosts of w, v, u (ns_cyl.cpp:83-104): one thread per (phi, z row)
__global__ void k_cyl_bound_r(CFld u, CFld v, CFld w, CylGeom g, int i0)
{
    pdl_wait();
    pdl_trigger();
    const int k = blockIdx.x * blockDim.x + threadIdx.x + g.z_;
    const int i = blockIdx.y + i0;
    if (k > g.znn) return;
    const int nr = g.nr;
    if (k >= g.z0) {
        w.at(i, k, 0) = 2 * g.U0 - w.at(i, k, 1);
        w.at(i, k, nr + 1) = -w.at(i, k, nr);
    }
    v.at(i, k, 0) = -v.at(i, k, 1);
    v.at(i, k, nr + 1) = -v.at(i, k, nr);
    if (k >= g.z0) {
        u.at(i, k, -1) = u.at(i, k, 1);
        u.at(i, k, nr + 1) = u.at(i, k, nr - 1);
    }
}

// z-direction ghosts, Dirichlet z only (ns_cyl.cpp:106-128).  The v mirror loops its r index over
// z0..znn (ns_cyl.cpp:108) -- reproduced, clamped to the allocated r range.
__global__ void k_cyl_bound_z(CFld u, CFld v, CFld w, CylGeom g, int jmax_v, int i0)
{
    pdl_wait();
    pdl_trigger();
    const int j = blockIdx.x * blockDim.x + threadIdx.x - 1;   // -1 .. nr+1
    const int i = blockIdx.y + i0;
    const int nr = g.nr, nz = g.nz;
    if (j > nr + 1) return;
    if (j >= g.z0 && j <= jmax_v) {
        v.at(i, -1, j) = v.at(i, 1, j);
        v.at(i, nz + 1, j) = v.at(i, nz - 1, j);
    }
    u.at(i, 0, j) = -u.at(i, 1, j);
    u.at(i, nz + 1, j) = -u.at(i, nz, j);
    if (j >= 0) {
        w.at(i, 0, j) = -w.at(i, 1, j);
        w.at(i, nz + 1, j) = -w.at(i, nz, j);
    }
}

// pressure ghosts (ns_cyl.cpp:132-171); blockIdx.z = 0: r faces, 1: z faces (Dirichlet z)
__global__ void k_cyl_bound_p(CFld u, CFld v, CFld p, CylGeom g, int i0)
{
    pdl_wait();
    pdl_trigger();
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y + i0;
    const int nr = g.nr, nz = g.nz;
    if (blockIdx.z == 0) {
        const int k = a + g.z1;
        if (k > g.zn) return;
        {
            const int j = 0;
            p.at(i, k, 0) = p.at(i, k, 1) -
                (g.crp[j] * u.at(i, k, 1) * g.cir[j] - 2 * u.at(i, k, 0) + g.crm[j] * u.at(i, k, -1) * g.cir[j]) * g.iRdr;
        }
        {
            const int j = nr;
            p.at(i, k, nr + 1) = p.at(i, k, nr) +
                (g.crp[j] * u.at(i, k, nr + 1) * g.cir[j] - 2 * u.at(i, k, nr) + g.crm[j] * u.at(i, k, nr - 1) * g.cir[j]) * g.iRdr;
        }
    } else {
        const int j = a + 1;
        if (j > nr) return;
        p.at(i, 0, j) = p.at(i, 1, j) - (v.at(i, 1, j) - 2 * v.at(i, 0, j) + v.at(i, -1, j)) * g.iRdz;
        p.at(i, nz + 1, j) = p.at(i, nz, j) + (v.at(i, nz + 1, j) - 2 * v.at(i, nz, j) + v.at(i, nz - 1, j)) * g.iRdz;
    }
}

__device__ __forceinline__ double sq(double x) { return x * x; }

// ---- FGH / L_FGH (ns_cyl.cpp:175-277, 280-405) ---------------------------------------------------
// One thread per (i, k, j) in [0..nphi-1] x [z0..zn] x [0..nr]:
//   F where k >= z1 (the reference loops i = 1..nphi and relies on the periodic wrap, :181),
//   G where j >= 1, H where k >= z1 and j >= 1.
template <bool LIN>
__global__ void __launch_bounds__(256, LIN ? 1 : 3)
k_cyl_fgh(CFld u, CFld v, CFld w, CFld u0, CFld v0, CFld w0, CFld F, CFld G, CFld H, CylGeom g, int i0)
{
    pdl_wait();
    pdl_trigger();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y * blockDim.y + threadIdx.y + g.z0;
    const int i = blockIdx.z + i0;
    if (j > g.nr || k > g.zn) return;
    const int ip = (g.wrap && i + 1 == g.nphi) ? 0 : i + 1, im = (g.wrap && i == 0) ? g.nphi - 1 : i - 1;
    int kp = k + 1, km = k - 1;
    if (g.zper) { if (kp == g.nz) kp = 0; if (km < 0) km = g.nz - 1; }
    if (!LIN && k >= g.z1 && j >= 1) {
        // Interior fast path of step() (all three of F, G, H are due here): the 27 distinct taps of the three stencils
        // are read once through one centre pointer per field plus the four neighbour strides (periodic wraps folded
        // into the strides), instead of a 64-bit index computation per tap and a separate set of loads per output
        // (SASS of the generic path: 54 loads and ~340 integer instructions around 137 fp64 instructions).
        // Taps are named by (dphi, dz, dr) with m = -1, p = +1; the expressions are those of the generic path below.
        // (32-bit neighbour offsets: a field holds fewer than 2^31 elements, checked at creation.)
        const double* __restrict__ uc_ = &u.at(i, k, j);
        const double* __restrict__ vc_ = &v.at(i, k, j);
        const double* __restrict__ wc_ = &w.at(i, k, j);
        const int ukp = (kp - k) * (int)u.sz, ukm = (km - k) * (int)u.sz;
        const int uip = (ip - i) * (int)u.sp, uim = (im - i) * (int)u.sp;
        const int vkp = (kp - k) * (int)v.sz, vkm = (km - k) * (int)v.sz;
        const int vip = (ip - i) * (int)v.sp, vim = (im - i) * (int)v.sp;
        const int wkp = (kp - k) * (int)w.sz, wkm = (km - k) * (int)w.sz;
        const int wip = (ip - i) * (int)w.sp, wim = (im - i) * (int)w.sp;
        const double u000 = uc_[0], u00p = uc_[1], u00m = uc_[-1], u0p0 = uc_[ukp], u0pm = uc_[ukp - 1], u0m0 = uc_[ukm];
        const double up00 = uc_[uip], up0m = uc_[uip - 1], um00 = uc_[uim];
        const double v000 = vc_[0], v00p = vc_[1], v00m = vc_[-1], v0p0 = vc_[vkp], v0m0 = vc_[vkm], v0mp = vc_[vkm + 1];
        const double vp00 = vc_[vip], vpm0 = vc_[vip + vkm], vm00 = vc_[vim];
        const double w000 = wc_[0], w00p = wc_[1], w00m = wc_[-1], w0p0 = wc_[wkp], w0m0 = wc_[wkm];
        const double wp00 = wc_[wip], wm00 = wc_[wim], wm0p = wc_[wim + 1], wmp0 = wc_[wim + wkp];
        {
            const double r2 = g.fr2[j], r1 = g.fr1[j], irr = g.firr[j], ir = g.fir[j];
            const double uc = u000;
            const double conv =
                (r2 * sq(0.5 * (uc + u00p)) - r1 * sq(0.5 * (u00m + uc))) * g.idr +
                0.25 * ((uc + u0p0) * (v00p + v000) -
                        (u0m0 + uc) * (v0mp + v0m0)) * g.idz +
                0.25 * ((uc + up00) * (w00p + w000) -
                        (um00 + uc) * (wm0p + wm00)) * g.idphi * ir -
                sq(0.5 * (w00p + w000)) * ir;
            F.at(i, k, j) = uc + g.dt * (
                (r2 * u00p - 2 * uc + r1 * u00m) * g.cRr +
                (u0p0 - 2 * uc + u0m0) * g.cRz +
                (up00 - 2 * uc + um00) * g.cRp * irr -
                conv -
                uc * irr * g.iRe -
                2 * (0.5 * (w00p + w000) - 0.5 * (wm0p + wm00)) * irr * g.iRdphi);
        }
        {
            const double r2 = g.cr2[j], r1 = g.cr1[j], irr = g.cirr[j], ir = g.cir[j];
            {
                const double vc = v000;
                const double conv =
                    (sq(0.5 * (vc + v0p0)) - sq(0.5 * (v0m0 + vc))) * g.idz +
                    0.25 * (r2 * (u000 + u0p0) * (v00p + vc) -
                            r1 * (u00m + u0pm) * (vc + v00m)) * g.idr +
                    0.25 * ((w000 + w0p0) * (vc + vp00) -
                            (wm00 + wmp0) * (vm00 + vc)) * g.idphi * ir;
                G.at(i, k, j) = vc + g.dt * (
                    (r2 * v00p - 2 * vc + r1 * v00m) * g.cRr +
                    (v0p0 - 2 * vc + v0m0) * g.cRz +
                    (vp00 - 2 * vc + vm00) * g.cRp * irr -
                    conv);
            }
            {
                const double wc = w000;
                const double conv =
                    (sq(0.5 * (wp00 + wc)) - sq(0.5 * (wm00 + wc))) * g.idphi * ir +
                    0.25 * (r2 * (up00 + u000) * (w00p + wc) -
                            r1 * (up0m + u00m) * (wc + w00m)) * g.idr +
                    0.25 * ((wc + w0p0) * (v000 + vp00) -
                            (w0m0 + wc) * (v0m0 + vpm0)) * g.idz +
                    wc * 0.5 * (up00 + u000) * ir;
                H.at(i, k, j) = wc + g.dt * (
                    (r2 * w00p - 2 * wc + r1 * w00m) * g.cRr +
                    (w0p0 - 2 * wc + w0m0) * g.cRz +
                    (wp00 - 2 * wc + wm00) * g.cRp * irr -
                    conv -
                    wc * irr * g.iRe +
                    2 * (0.5 * (up00 + u000) - 0.5 * (u000 + um00)) * irr * g.iRdphi);
            }
        }
        return;
    }
#define U(a, b, c) u.at(a, b, c)
#define V(a, b, c) v.at(a, b, c)
#define W(a, b, c) w.at(a, b, c)
#define U0(a, b, c) u0.at(a, b, c)
#define V0(a, b, c) v0.at(a, b, c)
#define W0(a, b, c) w0.at(a, b, c)
    if (k >= g.z1) {
        const double r2 = g.fr2[j], r1 = g.fr1[j], irr = g.firr[j], ir = g.fir[j];
        const double uc = U(i, k, j);
        double conv;
        if (!LIN) {
            conv =
                (r2 * sq(0.5 * (uc + U(i, k, j + 1))) - r1 * sq(0.5 * (U(i, k, j - 1) + uc))) * g.idr +
                0.25 * ((uc + U(i, kp, j)) * (V(i, k, j + 1) + V(i, k, j)) -
                        (U(i, km, j) + uc) * (V(i, km, j + 1) + V(i, km, j))) * g.idz +
                0.25 * ((uc + U(ip, k, j)) * (W(i, k, j + 1) + W(i, k, j)) -
                        (U(im, k, j) + uc) * (W(im, k, j + 1) + W(im, k, j))) * g.idphi * ir -
                sq(0.5 * (W(i, k, j + 1) + W(i, k, j))) * ir;
        } else {
            conv =
                (r2 * (0.5 * uc + U(i, k, j + 1)) * (U0(i, k, j) + U0(i, k, j + 1)) -
                 r1 * (0.5 * U(i, k, j - 1) + uc) * (U0(i, k, j - 1) + U0(i, k, j))) * g.idr +
                0.25 * ((uc + U(i, kp, j)) * (V0(i, k, j + 1) + V0(i, k, j)) -
                        (U(i, km, j) + uc) * (V0(i, km, j + 1) + V0(i, km, j))) * g.idz +
                0.25 * ((U0(i, k, j) + U0(i, kp, j)) * (V(i, k, j + 1) + V(i, k, j)) -
                        (U0(i, km, j) + U0(i, k, j)) * (V(i, km, j + 1) + V(i, km, j))) * g.idz +
                0.25 * ((uc + U(ip, k, j)) * (W0(i, k, j + 1) + W0(i, k, j)) -
                        (U(im, k, j) + uc) * (W0(im, k, j + 1) + W0(im, k, j))) * g.idphi * ir +
                0.25 * ((U0(i, k, j) + U0(ip, k, j)) * (W(i, k, j + 1) + W(i, k, j)) -
                        (U0(im, k, j) + U0(i, k, j)) * (W(im, k, j + 1) + W(im, k, j))) * g.idphi * ir -
                0.5 * (W(i, k, j + 1) + W(i, k, j)) * (W0(i, k, j + 1) + W0(i, k, j)) * ir;
        }
        F.at(i, k, j) = uc + g.dt * (
            (r2 * U(i, k, j + 1) - 2 * uc + r1 * U(i, k, j - 1)) * g.cRr +
            (U(i, kp, j) - 2 * uc + U(i, km, j)) * g.cRz +
            (U(ip, k, j) - 2 * uc + U(im, k, j)) * g.cRp * irr -
            conv -
            uc * irr * g.iRe -
            2 * (0.5 * (W(i, k, j + 1) + W(i, k, j)) - 0.5 * (W(im, k, j + 1) + W(im, k, j))) * irr * g.iRdphi);
    }
    if (j >= 1) {
        const double r2 = g.cr2[j], r1 = g.cr1[j], irr = g.cirr[j], ir = g.cir[j];
        {
            const double vc = V(i, k, j);
            double conv;
            if (!LIN) {
                conv =
                    (sq(0.5 * (vc + V(i, kp, j))) - sq(0.5 * (V(i, km, j) + vc))) * g.idz +
                    0.25 * (r2 * (U(i, k, j) + U(i, kp, j)) * (V(i, k, j + 1) + vc) -
                            r1 * (U(i, k, j - 1) + U(i, kp, j - 1)) * (vc + V(i, k, j - 1))) * g.idr +
                    0.25 * ((W(i, k, j) + W(i, kp, j)) * (vc + V(ip, k, j)) -
                            (W(im, k, j) + W(im, kp, j)) * (V(im, k, j) + vc)) * g.idphi * ir;
            } else {
                conv =
                    (0.5 * (vc + V(i, kp, j)) * (V0(i, k, j) + V0(i, kp, j)) -
                     0.5 * (V(i, km, j) + vc) * (V0(i, km, j) + V0(i, k, j))) * g.idz +
                    0.25 * (r2 * (U(i, k, j) + U(i, kp, j)) * (V0(i, k, j + 1) + V0(i, k, j)) -
                            r1 * (U(i, k, j - 1) + U(i, kp, j - 1)) * (V0(i, k, j) + V0(i, k, j - 1))) * g.idr +
                    0.25 * (r2 * (U0(i, k, j) + U0(i, kp, j)) * (V(i, k, j + 1) + vc) -
                            r1 * (U0(i, k, j - 1) + U0(i, kp, j - 1)) * (vc + V(i, k, j - 1))) * g.idr +
                    0.25 * ((W(i, k, j) + W(i, kp, j)) * (V0(i, k, j) + V0(ip, k, j)) -
                            (W(im, k, j) + W(im, kp, j)) * (V0(im, k, j) + V0(i, k, j))) * g.idphi * ir +
                    0.25 * ((W0(i, k, j) + W0(i, kp, j)) * (vc + V(ip, k, j)) -
                            (W0(im, k, j) + W0(im, kp, j)) * (V(im, k, j) + vc)) * g.idphi * ir;
            }
            G.at(i, k, j) = vc + g.dt * (
                (r2 * V(i, k, j + 1) - 2 * vc + r1 * V(i, k, j - 1)) * g.cRr +
                (V(i, kp, j) - 2 * vc + V(i, km, j)) * g.cRz +
                (V(ip, k, j) - 2 * vc + V(im, k, j)) * g.cRp * irr -
                conv);
        }
        if (k >= g.z1) {
            const double wc = W(i, k, j);
            double conv;
            if (!LIN) {
                conv =
                    (sq(0.5 * (W(ip, k, j) + wc)) - sq(0.5 * (W(im, k, j) + wc))) * g.idphi * ir +
                    0.25 * (r2 * (U(ip, k, j) + U(i, k, j)) * (W(i, k, j + 1) + wc) -
                            r1 * (U(ip, k, j - 1) + U(i, k, j - 1)) * (wc + W(i, k, j - 1))) * g.idr +
                    0.25 * ((wc + W(i, kp, j)) * (V(i, k, j) + V(ip, k, j)) -
                            (W(i, km, j) + wc) * (V(i, km, j) + V(ip, km, j))) * g.idz +
                    wc * 0.5 * (U(ip, k, j) + U(i, k, j)) * ir;
            } else {
                conv =
                    (0.5 * (W(ip, k, j) + wc) * (W0(ip, k, j) + W0(i, k, j)) -
                     0.5 * (W(im, k, j) + wc) * (W0(im, k, j) + W0(i, k, j))) * g.idphi * ir +
                    0.25 * (r2 * (U(ip, k, j) + U(i, k, j)) * (W0(i, k, j + 1) + W0(i, k, j)) -
                            r1 * (U(ip, k, j - 1) + U(i, k, j - 1)) * (W0(i, k, j) + W0(i, k, j - 1))) * g.idr +
                    0.25 * (r2 * (U0(ip, k, j) + U0(i, k, j)) * (W(i, k, j + 1) + wc) -
                            r1 * (U0(ip, k, j - 1) + U0(i, k, j - 1)) * (wc + W(i, k, j - 1))) * g.idr +
                    0.25 * ((wc + W(i, kp, j)) * (V0(i, k, j) + V0(ip, k, j)) -
                            (W(i, km, j) + wc) * (V0(i, km, j) + V0(ip, km, j))) * g.idz +
                    0.25 * ((W0(i, k, j) + W0(i, kp, j)) * (V(i, k, j) + V(ip, k, j)) -
                            (W0(i, km, j) + W0(i, k, j)) * (V(i, km, j) + V(ip, km, j))) * g.idz +
                    W0(i, k, j) * 0.5 * (U(ip, k, j) + U(i, k, j)) * ir +
                    wc * 0.5 * (U0(ip, k, j) + U0(i, k, j)) * ir;
            }
            H.at(i, k, j) = wc + g.dt * (
                (r2 * W(i, k, j + 1) - 2 * wc + r1 * W(i, k, j - 1)) * g.cRr +
                (W(i, kp, j) - 2 * wc + W(i, km, j)) * g.cRz +
                (W(ip, k, j) - 2 * wc + W(im, k, j)) * g.cRp * irr -
                conv -
                wc * irr * g.iRe +
                2 * (0.5 * (U(ip, k, j) + U(i, k, j)) - 0.5 * (U(i, k, j) + U(im, k, j))) * irr * g.iRdphi);
        }
    }
#undef U
#undef V
#undef W
#undef U0
#undef V0
#undef W0
}

// ---- poisson RHS (ns_cyl.cpp:408-439) --------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cyl_rhs(CFld F, CFld G, CFld H, CFld p, CFld R, CylGeom g, int i0)
{
    pdl_wait();
    pdl_trigger();
    const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int k = blockIdx.y * blockDim.y + threadIdx.y + g.z1;
    const int i = blockIdx.z + i0;
    if (j > g.nr || k > g.zn) return;
    const int im = (g.wrap && i == 0) ? g.nphi - 1 : i - 1;
    int km = k - 1;
    if (g.zper && km < 0) km = g.nz - 1;
    const double ir = g.cir[j];
    double r = ((g.crp[j] * F.at(i, k, j) - g.crm[j] * F.at(i, k, j - 1)) * ir * g.idr +
                (G.at(i, k, j) - G.at(i, km, j)) * g.idz +
                (H.at(i, k, j) - H.at(im, k, j)) * g.idphi * ir) * g.idt;
    if (!g.zper && k <= 1) r -= p.at(i, k - 1, j) * g.idz2;
    if (j <= 1) r -= g.crm[j] * ir * p.at(i, k, j - 1) * g.idr2;
    if (j >= g.nr) r -= g.crp[j] * ir * p.at(i, k, j + 1) * g.idr2;
    if (!g.zper && k >= g.nz) r -= p.at(i, k + 1, j) * g.idz2;
    R.at(i, k, j) = r;
}

// ---- update_uvwp (ns_cyl.cpp:445-484) --------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_cyl_update(CFld u, CFld v, CFld w, CFld p, CFld x, CFld F, CFld G, CFld H, CylGeom g, int i0)
{
    pdl_wait();
    pdl_trigger();
    const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int k = blockIdx.y * blockDim.y + threadIdx.y + g.z1;
    const int i = blockIdx.z + i0;
    if (j > g.nr || k > g.zn) return;
    const int ip = (g.wrap && i + 1 == g.nphi) ? 0 : i + 1;
    const double xc = x.at(i, k, j);
    if (j < g.nr) u.at(i, k, j) = F.at(i, k, j) - g.dtdr * (x.at(i, k, j + 1) - xc);
    if (k < g.nz) {                                  // k = z1 .. nz-1 (ns_cyl.cpp:462)
        int kp = k + 1;
        if (g.zper && kp == g.nz) kp = 0;
        v.at(i, k, j) = G.at(i, k, j) - g.dtdz * (x.at(i, kp, j) - xc);
    }
    w.at(i, k, j) = H.at(i, k, j) - g.dtdphi * g.cir[j] * (x.at(ip, k, j) - xc);
    p.at(i, k, j) = xc;                              // p = x over the index-range intersection (tensor.h:103-111)
}

// ---- halo pull: whole phi planes copied from the neighbours' windows (peer loads over NVLink) ---------
constexpr int CYL_MAX_PULLS = 8;
struct CylPullList {
    const double* src[CYL_MAX_PULLS];
    double* dst[CYL_MAX_PULLS];
    long long n[CYL_MAX_PULLS];
    int count;
};
__global__ void __launch_bounds__(256) k_cyl_pull(CylPullList pl)
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

// phi planes of field `fld` (u v w p x F G H RHS u0 v0 w0) that rank `rank` of `nranks` HOLDS: its own planes
// [rank*S, (rank+1)*S) plus the wrap-around halo planes its stencils read (phi = -1 and phi = nphi exist there)
static void cyl_window(int fld, int nphi, int rank, int nranks, int* wlo, int* whi)
{
    if (nranks == 1) { *wlo = 0; *whi = nphi - 1; return; }
    const int S = nphi / nranks, lo = rank * S, hi = lo + S - 1;
    switch (fld) {
    case 0: case 1: case 2: case 9: case 10: case 11: *wlo = lo - 1; *whi = hi + 1; break;   // u v w (u0 v0 w0): +-1
    case 4: *wlo = lo; *whi = hi + 1; break;                                                  // x: update reads x[i+1]
    case 7: *wlo = lo - 1; *whi = hi; break;                                                  // H: divergence reads H[i-1]
    default: *wlo = lo; *whi = hi; break;                                                     // p F G RHS
    }
}

struct CylLayout {
    int wlo[12], whi[12];
    int lz[12], lr[12];
    long long sz[12], sp[12], count[12];
    size_t off[12], bytes;
};
static void cyl_layout(int nr, int nz, int nphi, int zper, int rank, int nranks, CylLayout* L)
{
    const int z_ = zper ? 0 : -1, z0 = 0, z1 = zper ? 0 : 1, zn = zper ? nz - 1 : nz, znn = zper ? nz - 1 : nz + 1;
    // extents: ns_cyl.h:80-93
    const int Z0[12] = {z0, z_, z0, z0, z1, z1, z0, z1, z1, z0, z_, z0};
    const int Z1[12] = {znn, znn, znn, znn, zn, zn, zn, zn, zn, znn, znn, znn};
    const int R0[12] = {-1, 0, 0, 0, 1, 0, 1, 1, 1, -1, 0, 0};
    const int R1[12] = {nr + 1, nr + 1, nr + 1, nr + 1, nr, nr, nr, nr, nr, nr + 1, nr + 1, nr + 1};
    size_t o = 0;
    for (int f = 0; f < 12; f++) {
        cyl_window(f, nphi, rank, nranks, &L->wlo[f], &L->whi[f]);
        L->lz[f] = Z0[f]; L->lr[f] = R0[f];
        L->sz[f] = R1[f] - R0[f] + 1;
        L->sp[f] = (long long)(Z1[f] - Z0[f] + 1) * L->sz[f];
        L->count[f] = (long long)(L->whi[f] - L->wlo[f] + 1) * L->sp[f];
        L->off[f] = o;
        o += (sizeof(double) * (size_t)L->count[f] + 255) & ~(size_t)255;
    }
    L->bytes = o;
}

struct fdmb_ns_cyl {
    fdmb_ns_cyl_params prm{};
    int nr = 0, nz = 0, nphi = 0, zper = 0;
    double dr = 0, dz = 0, dphi = 0;
    CylGeom g{};
    CFld f[12]{};               // u v w p x F G H RHS u0 v0 w0 (windows over the global arrays when sharded)
    CylLayout lay{};
    double* d_tab = nullptr;    // r-dependent factor tables
    fdmb_lapl_cyl* lapl = nullptr;
    cudaStream_t stream = nullptr;
    long long time_index = 0;
    // phi-slab sharding (nranks == 1: the windows are the whole arrays and phi wraps inside them)
    int rank = 0, nranks = 1, plo = 0, phi = 0, device = 0;     // plo..phi: this rank's own phi planes
    void* block = nullptr;                                        // all twelve windows in one allocation
    void* peer_block[FDMB_MAX_RANKS] = {};
    bool peer_ipc[FDMB_MAX_RANKS] = {};
    bool attached = false;

    StepGraph graph[2];         // one step() / L_step() as a replayed CUDA graph (single-GPU handles)

    int init();
    int step(int nsteps, int linear, cudaStream_t st);
    int step_once(int linear, cudaStream_t st);
    int pull(const int* flds, const int* planes, const int* from, int n, cudaStream_t st);
    double* owned_ptr(int fld) const { return f[fld].p + (long long)(plo - lay.wlo[fld]) * lay.sp[fld]; }
    long long owned_count(int fld) const { return (long long)(phi - plo + 1) * lay.sp[fld]; }
    ~fdmb_ns_cyl();
};

int fdmb_ns_cyl::init()
{
    nr = prm.nr; nz = prm.nz; nphi = prm.nphi; zper = prm.zperiodic ? 1 : 0;
    if (nr < 3 || nz < 3 || nphi < 4) { set_error("NSCyl: nr, nz >= 3 and nphi >= 4 required"); return FDMB_ERR_INVALID; }
    if ((double)(nr + 3) * (nz + 3) * (nphi + 2) >= 2147483647.0) {
        set_error("NSCyl: a field of %d x %d x %d points exceeds 2^31 elements", nr, nz, nphi);
        return FDMB_ERR_INVALID;
    }
    if (nranks > 1 && nphi / nranks < 2) {
        set_error("NSCyl: the sharded step needs at least 2 phi planes per rank (nphi=%d, %d ranks)", nphi, nranks);
        return FDMB_ERR_INVALID;
    }
    const double R = prm.R, r0 = prm.r, h1 = prm.h1, h2 = prm.h2;
    dr = (R - r0) / nr; dz = (h2 - h1) / nz; dphi = 2 * M_PI / nphi;          // ns_cyl.h:77
    const double dr2 = dr * dr, dz2 = dz * dz, dphi2 = dphi * dphi;
    // ns_cyl.h:95-97
    int rc;
    if (nranks > 1)
        rc = fdmb_lapl_cyl_create_sharded(&lapl, dr, dz, r0 - dr / 2, R - r0 + dr, zper ? h2 - h1 : h2 - h1 + dz, nr, nz, nphi,
                                          zper, rank, nranks);
    else
        rc = fdmb_lapl_cyl_create(&lapl, dr, dz, r0 - dr / 2, R - r0 + dr, zper ? h2 - h1 : h2 - h1 + dz, nr, nz, nphi, zper);
    if (rc) return rc;
    FDMB_CUDA(cudaGetDevice(&device));
    FDMB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    {   // load this translation unit's kernels now (see the note in ns_cube.cu: lazy loading behind a spinning barrier)
        cudaFuncAttributes fa;
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_cyl_pull));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_cyl_bound_r));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_cyl_bound_z));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_cyl_bound_p));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_cyl_fgh<false>));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_cyl_fgh<true>));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_cyl_rhs));
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_cyl_update));
    }
    g.nr = nr; g.nz = nz; g.nphi = nphi; g.zper = zper; g.wrap = nranks == 1 ? 1 : 0;
    g.z_ = zper ? 0 : -1; g.z0 = 0; g.z1 = zper ? 0 : 1; g.zn = zper ? nz - 1 : nz; g.znn = zper ? nz - 1 : nz + 1;
    const int z1 = g.z1, zn = g.zn;
    plo = nranks > 1 ? rank * (nphi / nranks) : 0;
    phi = nranks > 1 ? plo + nphi / nranks - 1 : nphi - 1;
    cyl_layout(nr, nz, nphi, zper, rank, nranks, &lay);
    FDMB_CUDA(cudaMalloc(&block, lay.bytes));
    FDMB_CUDA(cudaMemset(block, 0, lay.bytes));
    // cudaMemset on device memory is asynchronous to the host and runs on the legacy default stream, which the
    // handle's non-blocking streams do not order against: finish it before the handle is handed out
    FDMB_CUDA(cudaDeviceSynchronize());
    for (int k = 0; k < 12; k++) {
        f[k].p = reinterpret_cast<double*>(static_cast<char*>(block) + lay.off[k]);
        f[k].li = lay.wlo[k]; f[k].lz = lay.lz[k]; f[k].lr = lay.lr[k];
        f[k].sz = lay.sz[k]; f[k].sp = lay.sp[k];
    }
    peer_block[rank] = block;
    g.U0 = prm.u0; g.dt = prm.dt;
    const double Re = prm.Re;
    g.cRr = 1.0 / Re / dr2; g.cRz = 1.0 / Re / dz2; g.cRp = 1.0 / Re / dphi2;
    g.idr = 1.0 / dr; g.idz = 1.0 / dz; g.idphi = 1.0 / dphi;
    g.iRe = 1.0 / Re; g.iRdr = 1.0 / Re / dr; g.iRdz = 1.0 / Re / dz; g.iRdphi = 1.0 / dphi / Re;
    g.idr2 = 1.0 / dr2; g.idz2 = 1.0 / dz2; g.idt = 1.0 / prm.dt;
    g.dtdr = prm.dt / dr; g.dtdz = prm.dt / dz; g.dtdphi = prm.dt / dphi;
    // r tables
    const int T = nr + 2;
    std::vector<double> tab(10 * (size_t)T, 0.0);
    double *fr2 = &tab[0], *fr1 = &tab[T], *firr = &tab[2 * T], *fir = &tab[3 * T];
    double *cr2 = &tab[4 * T], *cr1 = &tab[5 * T], *cirr = &tab[6 * T], *cir = &tab[7 * T], *crp = &tab[8 * T], *crm = &tab[9 * T];
    for (int j = 0; j <= nr; j++) {
        const double r = r0 + dr * j;                       // ns_cyl.cpp:184
        fr2[j] = (r + 0.5 * dr) / r; fr1[j] = (r - 0.5 * dr) / r; firr[j] = 1.0 / (r * r); fir[j] = 1.0 / r;
    }
    for (int j = 0; j <= nr + 1; j++) {
        const double r = r0 + dr * j - dr / 2;              // ns_cyl.cpp:217,246
        cr2[j] = (r + 0.5 * dr) / r; cr1[j] = (r - 0.5 * dr) / r; cirr[j] = 1.0 / (r * r); cir[j] = 1.0 / r;
        crp[j] = r + 0.5 * dr; crm[j] = r - 0.5 * dr;
    }
    FDMB_CUDA(cudaMalloc(&d_tab, sizeof(double) * tab.size()));
    FDMB_CUDA(cudaMemcpy(d_tab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
    g.fr2 = d_tab; g.fr1 = d_tab + T; g.firr = d_tab + 2 * T; g.fir = d_tab + 3 * T;
    g.cr2 = d_tab + 4 * T; g.cr1 = d_tab + 5 * T; g.cirr = d_tab + 6 * T; g.cir = d_tab + 7 * T;
    g.crp = d_tab + 8 * T; g.crm = d_tab + 9 * T;
    if (prm.vrandom == 1) {
        // ns_cyl.h:99-108: default-seeded std::default_random_engine, uniform(-1e-3, 1e-3), drawn in
        // (phi, z, r) order -- the same standard-library sequence as the reference build; a sharded rank draws the
        // whole sequence and keeps its own planes
        std::vector<double> hv((size_t)owned_count(1), 0.0);
        std::default_random_engine generator;
        std::uniform_real_distribution<double> distribution(-1e-3, 1e-3);
        for (int i = 0; i < nphi; i++)
            for (int k = z1; k <= zn; k++)
                for (int j = 1; j <= nr; j++) {
                    const double val = distribution(generator);
                    if (i >= plo && i <= phi)
                        hv[(size_t)(i - plo) * f[1].sp + (size_t)(k - f[1].lz) * f[1].sz + (j - f[1].lr)] = val;
                }
        FDMB_CUDA(cudaMemcpy(owned_ptr(1), hv.data(), sizeof(double) * hv.size(), cudaMemcpyHostToDevice));
    }
    return FDMB_OK;
}

fdmb_ns_cyl::~fdmb_ns_cyl()
{
    for (int q = 0; q < nranks; q++)
        if (q != rank && peer_ipc[q] && peer_block[q]) cudaIpcCloseMemHandle(peer_block[q]);
    cudaFree(block);
    cudaFree(d_tab);
    if (lapl) fdmb_lapl_cyl_destroy(lapl);
    if (stream) cudaStreamDestroy(stream);
}

// copy phi plane planes[s] (this rank's numbering: -1 and nphi are the wrap-around halos) of field flds[s] from the
// window of rank from[s], which owns it
int fdmb_ns_cyl::pull(const int* flds, const int* planes, const int* from, int n, cudaStream_t st)
{
    CylPullList pl{};
    long long nmax = 0;
    for (int s = 0; s < n; s++) {
        const int k = flds[s];
        CylLayout q;
        cyl_layout(nr, nz, nphi, zper, from[s], nranks, &q);
        const int gsrc = (planes[s] + nphi) % nphi;                   // the plane's index on its owner
        pl.src[s] = reinterpret_cast<const double*>(static_cast<const char*>(peer_block[from[s]]) + q.off[k]) +
                    (long long)(gsrc - q.wlo[k]) * q.sp[k];
        pl.dst[s] = f[k].p + (long long)(planes[s] - lay.wlo[k]) * lay.sp[k];
        pl.n[s] = lay.sp[k];
        if (pl.n[s] > nmax) nmax = pl.n[s];
    }
    pl.count = n;
    LaunchScope sc("nscyl_halo_pull", st);
    long long bx = (nmax + 256 * 4 - 1) / (256 * 4);
    if (bx > 512) bx = 512;
    if (bx < 1) bx = 1;
    k_cyl_pull<<<dim3((unsigned)bx, n), 256, 0, st>>>(pl);
    FDMB_CHECK_LAUNCH();
    return FDMB_OK;
}

int fdmb_ns_cyl::step(int nsteps, int linear, cudaStream_t st)
{
    if (nranks > 1 && !attached) { set_error("NSCyl: sharded handle used before attach_ipc/attach_local"); return FDMB_ERR_COMM; }
    for (int s = 0; s < nsteps; s++) {
        // (sharded steps too: the cross-GPU barriers keep their epoch in device memory, so the sequence replays)
        const int rc = graph[linear ? 1 : 0].run(st, this, nullptr, [&]() { return step_once(linear, st); });
        if (rc) return rc;
        time_index++;
    }
    return FDMB_OK;
}

int fdmb_ns_cyl::step_once(int linear, cudaStream_t st)
{
    const CFld &u = f[0], &v = f[1], &w = f[2], &p = f[3], &x = f[4], &F = f[5], &G = f[6], &H = f[7], &R = f[8];
    const CFld &u0 = f[9], &v0 = f[10], &w0 = f[11];
    const int nzrows = g.znn - g.z_ + 1;
    const int nown = phi - plo + 1;
    const int wlo = lay.wlo[0], nwin = lay.whi[0] - lay.wlo[0] + 1;     // planes of u, v, w held here (own + halos)
    const int prev = (rank + nranks - 1) % nranks, next = (rank + 1) % nranks;
    PdlScope pdl(nranks == 1 && pdl_small_grid((long long)g.nr * g.nz * g.nphi));   // launch-bound sizes (pdl.cuh)
    int rc;
    {
        if (nranks > 1) {
            // every rank has finished the previous update (or set_field): fetch the wrap-around halo planes
            if ((rc = lapl->barrier(st))) return rc;
            int flds[12], planes[12], from[12], n = 0;
            for (int k = 0; k < 3; k++) {
                flds[n] = k; planes[n] = plo - 1; from[n++] = prev;
                flds[n] = k; planes[n] = phi + 1; from[n++] = next;
            }
            if ((rc = pull(flds, planes, from, 6, st))) return rc;
            if (linear) {
                n = 0;
                for (int k = 9; k < 12; k++) {
                    flds[n] = k; planes[n] = plo - 1; from[n++] = prev;
                    flds[n] = k; planes[n] = phi + 1; from[n++] = next;
                }
                if ((rc = pull(flds, planes, from, 6, st))) return rc;
            }
        }
        {
            LaunchScope sc("nscyl_bound_r", st);
            dim3 grid((nzrows + 127) / 128, nwin);
            launch_pdl(k_cyl_bound_r, grid, dim3(128), 0, st, u, v, w, g, wlo);
        }
        if (!zper) {
            LaunchScope sc("nscyl_bound_z", st);
            const int jmax_v = g.znn < nr + 1 ? g.znn : nr + 1;
            dim3 grid((nr + 3 + 127) / 128, nwin);
            launch_pdl(k_cyl_bound_z, grid, dim3(128), 0, st, u, v, w, g, jmax_v, wlo);
        }
        {
            LaunchScope sc("nscyl_bound_p", st);
            const int na = (g.zn - g.z1 + 1) > nr ? (g.zn - g.z1 + 1) : nr;
            dim3 grid((na + 127) / 128, nown, zper ? 1 : 2);
            launch_pdl(k_cyl_bound_p, grid, dim3(128), 0, st, u, v, p, g, plo);
        }
        {
            LaunchScope sc(linear ? "nscyl_lfgh" : "nscyl_fgh", st);
            dim3 block(64, 4);
            dim3 grid((nr + 1 + 63) / 64, (g.zn - g.z0 + 1 + 3) / 4, nown);
            if (linear) launch_pdl(k_cyl_fgh<true>, grid, block, 0, st, u, v, w, u0, v0, w0, F, G, H, g, plo);
            else launch_pdl(k_cyl_fgh<false>, grid, block, 0, st, u, v, w, u0, v0, w0, F, G, H, g, plo);
        }
        if (nranks > 1) {
            // the divergence reads H on the plane below the slab: fetch it from its owner
            if ((rc = lapl->barrier(st))) return rc;
            int fld = 7, plane = plo - 1;
            if ((rc = pull(&fld, &plane, &prev, 1, st))) return rc;
        }
        {
            LaunchScope sc("nscyl_rhs", st);
            dim3 block(64, 4);
            dim3 grid((nr + 63) / 64, (g.zn - g.z1 + 1 + 3) / 4, nown);
            launch_pdl(k_cyl_rhs, grid, block, 0, st, F, G, H, p, R, g, plo);
        }
        FDMB_CHECK_LAUNCH();
        rc = lapl->solve_device(x.p, R.p, st);      // ns_cyl.cpp:441
        if (rc) return rc;
        if (nranks > 1) {
            // the neighbour above has written its x slab: fetch the plane the w update reads
            if ((rc = lapl->barrier(st))) return rc;
            int fld = 4, plane = phi + 1;
            if ((rc = pull(&fld, &plane, &next, 1, st))) return rc;
        }
        {
            LaunchScope sc("nscyl_update", st);
            dim3 block(64, 4);
            dim3 grid((nr + 63) / 64, (g.zn - g.z1 + 1 + 3) / 4, nown);
            launch_pdl(k_cyl_update, grid, block, 0, st, u, v, w, p, x, F, G, H, g, plo);
        }
        FDMB_CHECK_LAUNCH();
    }
    return FDMB_OK;
}

// runs `body` with the handle's device current (several ranks may share one process)
template <typename Fn> static int cyl_ns_on_device(fdmb_ns_cyl* h, Fn body)
{
    if (h->nranks < 2) return body();
    int cur = 0;
    FDMB_CUDA(cudaGetDevice(&cur));
    if (cur != h->device) FDMB_CUDA(cudaSetDevice(h->device));
    int rc = body();
    if (cur != h->device) cudaSetDevice(cur);
    return rc;
}

extern "C" {

int fdmb_ns_cyl_default_params(fdmb_ns_cyl_params* p)
{
    if (!p) { set_error("null argument"); return FDMB_ERR_INVALID; }
    // defaults of ns_cyl.h:57-68
    p->R = M_PI; p->r = M_PI / 2; p->h1 = 0; p->h2 = 10; p->u0 = 1.0; p->Re = 1.0; p->dt = 0.001;
    p->nr = 32; p->nz = 31; p->nphi = 32; p->verbose = 0; p->vrandom = 0; p->zperiodic = 0;
    return FDMB_OK;
}

static int ns_cyl_create(fdmb_ns_cyl** out, const fdmb_ns_cyl_params* p, int rank, int nranks)
{
    if (!out || !p) { set_error("null argument"); return FDMB_ERR_INVALID; }
    *out = nullptr;
    if (nranks != 1 && nranks != 2 && nranks != 4 && nranks != 8) {
        set_error("NSCyl: nranks must be 1, 2, 4 or 8 (got %d)", nranks);
        return FDMB_ERR_INVALID;
    }
    if (rank < 0 || rank >= nranks) { set_error("NSCyl: rank %d out of range", rank); return FDMB_ERR_INVALID; }
    auto* h = new (std::nothrow) fdmb_ns_cyl();
    if (!h) { set_error("out of host memory"); return FDMB_ERR_NOMEM; }
    h->prm = *p; h->rank = rank; h->nranks = nranks;
    int rc = h->init();
    if (rc) { delete h; return rc; }
    *out = h;
    return FDMB_OK;
}

int fdmb_ns_cyl_create(fdmb_ns_cyl** out, const fdmb_ns_cyl_params* p) { return ns_cyl_create(out, p, 0, 1); }

int fdmb_ns_cyl_create_sharded(fdmb_ns_cyl** out, const fdmb_ns_cyl_params* p, int rank, int nranks)
{
    return ns_cyl_create(out, p, rank, nranks);
}

int fdmb_ns_cyl_local_slab(fdmb_ns_cyl* h, int* phi_first, int* nphi_local)
{
    if (!h || !phi_first || !nphi_local) { set_error("null argument"); return FDMB_ERR_INVALID; }
    *phi_first = h->plo; *nphi_local = h->phi - h->plo + 1;
    return FDMB_OK;
}

int fdmb_ns_cyl_export_ipc(fdmb_ns_cyl* h, void* handles)
{
    if (!h || !handles || h->nranks < 2) { set_error("export_ipc needs a sharded handle"); return FDMB_ERR_INVALID; }
    int rc = fdmb_lapl_cyl_export_ipc(h->lapl, handles);
    if (rc) return rc;
    cudaIpcMemHandle_t ih;
    FDMB_CUDA(cudaIpcGetMemHandle(&ih, h->block));
    memcpy(static_cast<char*>(handles) + FDMB_IPC_HANDLE_BYTES, &ih, sizeof(ih));
    return FDMB_OK;
}

// handles: nranks records of 2 * FDMB_IPC_HANDLE_BYTES (solver block, field block), indexed by rank
int fdmb_ns_cyl_attach_ipc(fdmb_ns_cyl* h, const void* handles)
{
    if (!h || !handles || h->nranks < 2) { set_error("attach_ipc needs a sharded handle"); return FDMB_ERR_INVALID; }
    char solver[FDMB_MAX_RANKS * FDMB_IPC_HANDLE_BYTES];
    for (int q = 0; q < h->nranks; q++)
        memcpy(solver + (size_t)q * FDMB_IPC_HANDLE_BYTES, static_cast<const char*>(handles) + (size_t)q * 2 * FDMB_IPC_HANDLE_BYTES,
               FDMB_IPC_HANDLE_BYTES);
    int rc = fdmb_lapl_cyl_attach_ipc(h->lapl, solver);
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

int fdmb_ns_cyl_attach_local(fdmb_ns_cyl* h, fdmb_ns_cyl* const* all)
{
    if (!h || !all || h->nranks < 2) { set_error("attach_local needs a sharded handle"); return FDMB_ERR_INVALID; }
    fdmb_lapl_cyl* solvers[FDMB_MAX_RANKS] = {};
    for (int q = 0; q < h->nranks; q++) {
        if (!all[q] || all[q]->nranks != h->nranks || all[q]->rank != q || all[q]->nphi != h->nphi || all[q]->nr != h->nr) {
            set_error("attach_local: handle %d does not belong to this sharded run", q);
            return FDMB_ERR_INVALID;
        }
        solvers[q] = all[q]->lapl;
    }
    int rc = fdmb_lapl_cyl_attach_local(h->lapl, solvers);      // also enables peer access between the devices
    if (rc) return rc;
    for (int q = 0; q < h->nranks; q++) h->peer_block[q] = all[q]->block;
    h->attached = true;
    return FDMB_OK;
}

int fdmb_ns_cyl_synchronize(fdmb_ns_cyl* h)
{
    if (!h) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return cyl_ns_on_device(h, [&] {
        FDMB_CUDA(cudaStreamSynchronize(h->stream));
        return (int)FDMB_OK;
    });
}

int fdmb_ns_cyl_step(fdmb_ns_cyl* h, int nsteps)
{
    if (!h || nsteps < 0) { set_error("bad argument"); return FDMB_ERR_INVALID; }
    return cyl_ns_on_device(h, [&] {
        int rc = h->step(nsteps, 0, h->stream);
        if (rc) return rc;
        FDMB_CUDA(cudaStreamSynchronize(h->stream));
        return (int)FDMB_OK;
    });
}

int fdmb_ns_cyl_lstep(fdmb_ns_cyl* h, int nsteps)
{
    if (!h || nsteps < 0) { set_error("bad argument"); return FDMB_ERR_INVALID; }
    return cyl_ns_on_device(h, [&] {
        int rc = h->step(nsteps, 1, h->stream);
        if (rc) return rc;
        FDMB_CUDA(cudaStreamSynchronize(h->stream));
        return (int)FDMB_OK;
    });
}

int fdmb_ns_cyl_step_async(fdmb_ns_cyl* h, int nsteps, int linear, void* stream)
{
    if (!h || nsteps < 0) { set_error("bad argument"); return FDMB_ERR_INVALID; }
    return cyl_ns_on_device(h, [&] { return h->step(nsteps, linear ? 1 : 0, stream ? (cudaStream_t)stream : h->stream); });
}

int fdmb_ns_cyl_field_size(fdmb_ns_cyl* h, int field, long long* count)
{
    if (!h || field < 0 || field > 11 || !count) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    *count = h->owned_count(field);
    return FDMB_OK;
}

int fdmb_ns_cyl_get_field(fdmb_ns_cyl* h, int field, double* host)
{
    if (!h || field < 0 || field > 11 || !host) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    return cyl_ns_on_device(h, [&] {
        FDMB_CUDA(cudaMemcpyAsync(host, h->owned_ptr(field), sizeof(double) * h->owned_count(field), cudaMemcpyDeviceToHost,
                                  h->stream));
        FDMB_CUDA(cudaStreamSynchronize(h->stream));
        return (int)FDMB_OK;
    });
}

int fdmb_ns_cyl_set_field(fdmb_ns_cyl* h, int field, const double* host)
{
    if (!h || field < 0 || field > 11 || !host) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    return cyl_ns_on_device(h, [&] {
        FDMB_CUDA(cudaMemcpyAsync(h->owned_ptr(field), host, sizeof(double) * h->owned_count(field), cudaMemcpyHostToDevice,
                                  h->stream));
        FDMB_CUDA(cudaStreamSynchronize(h->stream));
        return (int)FDMB_OK;
    });
}

int fdmb_ns_cyl_field_device_ptr(fdmb_ns_cyl* h, int field, void** dptr)
{
    if (!h || field < 0 || field > 11 || !dptr) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    *dptr = h->owned_ptr(field);
    return FDMB_OK;
}

long long fdmb_ns_cyl_time_index(fdmb_ns_cyl* h) { return h ? h->time_index : -1; }

// NSCyl::U0 is a public, non-const member of the reference class (src/ns_cyl.h:23) that callers change between
// steps (test/test_ns_cyl_spectral.cpp sets ns.U0 = 0 before its L_step loop); takes effect at the next step
int fdmb_ns_cyl_set_u0(fdmb_ns_cyl* h, double u0)
{
    if (!h) { set_error("null handle"); return FDMB_ERR_INVALID; }
    h->prm.u0 = u0;
    h->g.U0 = u0;
    h->graph[0].reset(); h->graph[1].reset();      // the captured launches carry the old wall speed as an argument
    return FDMB_OK;
}

int fdmb_ns_cyl_destroy(fdmb_ns_cyl* h)
{
    delete h;
    return FDMB_OK;
}

}  // extern "C"
