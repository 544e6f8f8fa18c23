// LaplCyl3FFT2 on B200: Poisson equation in cylindrical coordinates (phi, z, r), r Dirichlet,
//   1/r d/dr(r du/dr) + d2u/dz2 + 1/r^2 d2u/dphi2 = f.
// Replaces fdm::LaplCyl3FFT2<double,check,zflag>::solve (reference src/lapl_cyl.cpp:11-128,
// init_solver :131-170, data src/lapl_cyl.h:12-37,172-249).
//
// Sweeps: phi forward (periodic, slowest axis) -> z forward (DST or periodic, middle axis) ->
// batched tridiagonal solves along the contiguous r axis -> z inverse -> phi inverse.
// The tridiagonal matrices are never stored: the r-dependent entries come from three small
// tables and the (phi,z)-mode shift from the eigenvalue tables, the LU recurrence (the no-pivot
// path of LAPACK gttrf/gttrs that the reference calls, src/lapl_cyl.cpp:83,166; the matrices are
// diagonally dominant so LAPACK never pivots) is recomputed on the fly.
#include <cmath>
#include <cstdint>
#include <new>
#include <vector>

#include "common.h"
#include "lapl_cyl.h"

namespace fdmb {

// ---- batched tridiagonal solve along the contiguous axis -------------------------------------
// A CTA stages TS systems (rows of nr doubles) in an odd-pitch shared-memory tile with coalesced
// loads; thread s then runs the Thomas recurrence of system s (conflict-free: consecutive lanes sit
// one odd pitch apart), keeping the reciprocal pivots in a second tile; coalesced store.
constexpr int TS = 32;

__global__ void __launch_bounds__(128) k_tridiag_rows(TridiagArgs a)
{
    extern __shared__ double smem[];
    const int P = a.nr | 1;                    // odd pitch
    double* tb = smem;                         // rhs / solution
    double* ti = smem + TS * P;                // reciprocal pivots
    double* cL = ti + TS * P;                  // coefficient tables, 1-based
    double* cU = cL + (a.nr + 1);
    double* cR = cU + (a.nr + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int j = tid + 1; j <= a.nr; j += blockDim.x) { cL[j] = a.L[j]; cU[j] = a.U[j]; cR[j] = a.ir2[j]; }
    const long long row0 = (long long)blockIdx.x * TS;
    for (int r = warp; r < TS; r += 4) {
        const long long row = row0 + r;
        if (row >= a.nsys) break;
        const double* src = a.data + row * a.pitch;
        for (int x = lane; x < a.nr; x += 32) tb[r * P + x] = src[x];
    }
    __syncthreads();
    if (tid < TS && row0 + tid < a.nsys) {
        const long long row = row0 + tid;
        const int outer = (int)(row / a.nmid), mid = (int)(row % a.nmid);
        const double lo = a.lm_outer[outer], lmz = a.lm_mid[mid + a.mid0];
        double* b = tb + tid * P - 1;          // 1-based
        double* iv = ti + tid * P - 1;
        // D_j = -2/dr2 - lm_phi/r/r - lm_z   (lapl_cyl.cpp:153)
        double d = a.c0 - lo * cR[1] - lmz;
        double inv = 1.0 / d;
        double bp = b[1];
        iv[1] = inv;
        for (int j = 2; j <= a.nr; j++) {
            const double fact = cL[j] * inv;                    // dl / d
            d = (a.c0 - lo * cR[j] - lmz) - fact * cU[j - 1];   // d_{j} -= fact * du_{j-1}
            bp = b[j] - fact * bp;                              // forward substitution
            inv = 1.0 / d;
            b[j] = bp;
            iv[j] = inv;
        }
        double x = bp * inv;
        b[a.nr] = x;
        for (int j = a.nr - 1; j >= 1; j--) {
            x = (b[j] - cU[j] * x) * iv[j];
            b[j] = x;
        }
    }
    __syncthreads();
    for (int r = warp; r < TS; r += 4) {
        const long long row = row0 + r;
        if (row >= a.nsys) break;
        double* dst = a.data + row * a.pitch;
        for (int x = lane; x < a.nr; x += 32) dst[x] = tb[r * P + x];
    }
}

size_t tridiag_rows_smem(int nr) { return sizeof(double) * (size_t)(2 * TS * (nr | 1) + 3 * (nr + 1)); }

cudaError_t launch_tridiag_rows(const TridiagArgs& a, cudaStream_t st, const char* tag)
{
    LaunchScope scope(tag, st);
    const size_t smem = tridiag_rows_smem(a.nr);
    static size_t set_smem_dev[64] = {0};     // function attributes are per device
    size_t& set_smem = set_smem_dev[current_device_slot()];
    if (smem > set_smem) {
        cudaError_t e = cudaFuncSetAttribute(k_tridiag_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        set_smem = smem;
    }
    const unsigned grid = (unsigned)((a.nsys + TS - 1) / TS);
    k_tridiag_rows<<<grid, 128, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace fdmb

using namespace fdmb;

static inline double sq(double x) { return x * x; }

int fdmb_lapl_cyl::init()
{
    Nz = zperiodic ? nz : nz + 1;
    if (nr < 2 || nz < 1 || nphi < 1 || !supported_N(nphi) || !supported_N(Nz)) {
        set_error("LaplCyl3FFT2: nphi=%d and the z transform length %d (nz%s) must be powers of two in [4,2048], nr >= 2 "
                  "(reference: verify((1<<n) == N), src/fft.cpp:67)", nphi, Nz, zperiodic ? "" : "+1");
        return FDMB_ERR_INVALID;
    }
    if (tridiag_rows_smem(nr) > 220 * 1024) {
        set_error("LaplCyl3FFT2: nr=%d exceeds the shared-memory tile of the tridiagonal kernel", nr);
        return FDMB_ERR_INVALID;
    }
    dphi = 2 * M_PI / nphi;
    slz = std::sqrt(2. / lz);
    pr = (nr + 15) / 16 * 16;
    int rc;
    if ((rc = get_tables(nphi, &tphi)) || (rc = get_tables(Nz, &tz))) return rc;
    const double dphi2 = dphi * dphi, dz2 = dz * dz, dr2 = dr * dr;
    std::vector<double> lm_phi(nphi), lm_z(Nz), L(nr + 1, 0.0), U(nr + 1, 0.0), ir2(nr + 1, 0.0);
    for (int i = 0; i < nphi; i++) lm_phi[i] = 4.0 / dphi2 * sq(sin(i * dphi * 0.5));
    for (int k = 0; k < Nz; k++)
        lm_z[k] = zperiodic ? 4. / dz2 * sq(sin(k * M_PI / Nz)) : 4. / dz2 * sq(sin(k * M_PI * 0.5 / Nz));
    for (int j = 1; j <= nr; j++) {
        const double r = r0 + j * dr;
        L[j] = (r - 0.5 * dr) / dr2 / r;
        U[j] = (r + 0.5 * dr) / dr2 / r;
        ir2[j] = 1.0 / r / r;
    }
    FDMB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    auto up = [&](double** d, const std::vector<double>& h) -> int {
        FDMB_CUDA(cudaMalloc(d, sizeof(double) * h.size()));
        FDMB_CUDA(cudaMemcpy(*d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
        return FDMB_OK;
    };
    if ((rc = up(&d_lmphi, lm_phi)) || (rc = up(&d_lmz, lm_z)) || (rc = up(&d_L, L)) || (rc = up(&d_U, U)) ||
        (rc = up(&d_ir2, ir2)))
        return rc;
    FDMB_CUDA(cudaMalloc(&d_work, sizeof(double) * (size_t)nphi * nz * pr));
    if (pipe_enabled()) {
        const unsigned long long s1 = 8ull * pr, s2 = 8ull * (unsigned long long)nz * pr;
        if (pipe_supported_N(Nz)) {
            if ((rc = make_cols_maps(&tm_z, d_work, Nz, 1, nr, nz, nphi, s1, s2, pipe_B(Nz)))) return rc;
            pipe_z = true;
        }
        if (pipe_supported_N(nphi)) {
            if ((rc = make_cols_maps(&tm_phi, d_work, nphi, 2, nr, nz, nphi, s1, s2, pipe_B(nphi)))) return rc;
            pipe_phi = true;
        }
    }
    return FDMB_OK;
}

fdmb_lapl_cyl::~fdmb_lapl_cyl()
{
    cudaFree(d_lmphi); cudaFree(d_lmz); cudaFree(d_L); cudaFree(d_U); cudaFree(d_ir2);
    cudaFree(d_work); cudaFree(d_rhs); cudaFree(d_ans);
    if (stream) cudaStreamDestroy(stream);
}

int fdmb_lapl_cyl::solve_device(double* d_out, const double* d_in, cudaStream_t st)
{
    const double SQRT_M_1_PI = 0.56418958354775629;   // lapl_cyl.h:175
    const long long wplane = (long long)nz * pr;      // work-array stride between phi planes
    const long long uplane = (long long)nz * nr;      // caller-array stride between phi planes
    const int kzf = zperiodic ? XF_PFWD : XF_DST, kzi = zperiodic ? XF_PINV : XF_DST;
    int rc;
    // phi forward: caller rhs -> work.  TMA-fed when the caller's rows are 16-byte multiples.
    {
        bool pipe_in = pipe_phi && (nr % 2 == 0) && ((reinterpret_cast<uintptr_t>(d_in) & 15) == 0);
        if (pipe_in && tm_in_ptr != d_in) {
            if ((rc = make_cols_maps(&tm_in, d_in, nphi, 2, nr, nz, nphi, 8ull * nr, 8ull * uplane, pipe_B(nphi)))) return rc;
            tm_in_ptr = d_in;
        }
        if (pipe_in) {
            ColsPipeArgs p{};
            p.out = d_work; p.out_sj = wplane; p.out_so = pr; p.nvalid = nphi; p.nb = nr; p.no = nz; p.taxis = 2;
            p.reverse = 0; p.scale = dphi * SQRT_M_1_PI;
            p.SN = tphi.SN; p.WM = tphi.WM;
            FDMB_CUDA(launch_cols_pipe(nphi, XF_PFWD, tm_in, p, st, "cyl_phi_fwd"));
        } else {
            ColsArgs c{};
            c.in = d_in; c.out = d_work; c.nvalid = nphi; c.in_sj = uplane; c.out_sj = wplane; c.nb = nr; c.no = nz;
            c.in_so = nr; c.out_so = pr; c.scale = dphi * SQRT_M_1_PI; c.SN = tphi.SN; c.WM = tphi.WM;
            FDMB_CUDA(launch_cols(nphi, XF_PFWD, c, st, "cyl_phi_fwd"));
        }
    }
    // z forward / inverse on the work array
    auto zsweep = [&](int kind, double scale, const char* tag, int reverse) -> cudaError_t {
        if (pipe_z) {
            ColsPipeArgs p{};
            p.out = d_work; p.out_sj = pr; p.out_so = wplane; p.nvalid = nz; p.nb = nr; p.no = nphi; p.taxis = 1;
            p.reverse = reverse; p.scale = scale; p.SN = tz.SN; p.WM = tz.WM;
            return launch_cols_pipe(Nz, kind, tm_z, p, st, tag);
        }
        ColsArgs c{};
        c.in = d_work; c.out = d_work; c.nvalid = nz; c.in_sj = c.out_sj = pr; c.nb = nr; c.no = nphi;
        c.in_so = c.out_so = wplane; c.scale = scale; c.SN = tz.SN; c.WM = tz.WM;
        return launch_cols(Nz, kind, c, st, tag);
    };
    FDMB_CUDA(zsweep(kzf, dz * slz, "cyl_z_fwd", 1));
    // tridiagonal solves along r for every (phi mode, z mode)
    {
        TridiagArgs t{};
        t.data = d_work; t.pitch = pr; t.nr = nr; t.nmid = nz; t.nsys = (long long)nphi * nz;
        t.lm_outer = d_lmphi; t.lm_mid = d_lmz; t.mid0 = zperiodic ? 0 : 1; t.c0 = -2 / (dr * dr);
        t.L = d_L; t.U = d_U; t.ir2 = d_ir2;
        FDMB_CUDA(launch_tridiag_rows(t, st, "cyl_r_tridiag"));
    }
    FDMB_CUDA(zsweep(kzi, slz, "cyl_z_inv", 0));
    // phi inverse: work -> caller ans
    if (pipe_phi) {
        ColsPipeArgs p{};
        p.out = d_out; p.out_sj = uplane; p.out_so = nr; p.nvalid = nphi; p.nb = nr; p.no = nz; p.taxis = 2;
            p.reverse = 0; p.scale = SQRT_M_1_PI; p.SN = tphi.SN; p.WM = tphi.WM;
        FDMB_CUDA(launch_cols_pipe(nphi, XF_PINV, tm_phi, p, st, "cyl_phi_inv"));
    } else {
        ColsArgs c{};
        c.in = d_work; c.out = d_out; c.nvalid = nphi; c.in_sj = wplane; c.out_sj = uplane; c.nb = nr; c.no = nz;
        c.in_so = pr; c.out_so = nr; c.scale = SQRT_M_1_PI; c.SN = tphi.SN; c.WM = tphi.WM;
        FDMB_CUDA(launch_cols(nphi, XF_PINV, c, st, "cyl_phi_inv"));
    }
    return FDMB_OK;
}

int fdmb_lapl_cyl::solve_host(double* ans, const double* rhs)
{
    const size_t bytes = sizeof(double) * (size_t)nr * nz * nphi;
    if (!d_rhs) FDMB_CUDA(cudaMalloc(&d_rhs, bytes));
    if (!d_ans) FDMB_CUDA(cudaMalloc(&d_ans, bytes));
    FDMB_CUDA(cudaMemcpyAsync(d_rhs, rhs, bytes, cudaMemcpyHostToDevice, stream));
    int rc = solve_device(d_ans, d_rhs, stream);
    if (rc) return rc;
    FDMB_CUDA(cudaMemcpyAsync(ans, d_ans, bytes, cudaMemcpyDeviceToHost, stream));
    FDMB_CUDA(cudaStreamSynchronize(stream));
    return FDMB_OK;
}

extern "C" {

int fdmb_lapl_cyl_create(fdmb_lapl_cyl** out, double dr, double dz, double r0, double lr, double lz, int nr, int nz,
                         int nphi, int zperiodic)
{
    if (!out) { set_error("null handle pointer"); return FDMB_ERR_INVALID; }
    *out = nullptr;
    auto* h = new (std::nothrow) fdmb_lapl_cyl();
    if (!h) { set_error("out of host memory"); return FDMB_ERR_NOMEM; }
    h->dr = dr; h->dz = dz; h->r0 = r0; h->lr = lr; h->lz = lz;
    h->nr = nr; h->nz = nz; h->nphi = nphi; h->zperiodic = zperiodic ? 1 : 0;
    int rc = h->init();
    if (rc) { delete h; return rc; }
    *out = h;
    return FDMB_OK;
}

int fdmb_lapl_cyl_solve(fdmb_lapl_cyl* h, double* ans, const double* rhs)
{
    if (!h || !ans || !rhs) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return h->solve_host(ans, rhs);
}

int fdmb_lapl_cyl_solve_device(fdmb_lapl_cyl* h, double* d_ans, const double* d_rhs, void* stream)
{
    if (!h || !d_ans || !d_rhs) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return h->solve_device(d_ans, d_rhs, stream ? (cudaStream_t)stream : h->stream);
}

int fdmb_lapl_cyl_destroy(fdmb_lapl_cyl* h)
{
    delete h;
    return FDMB_OK;
}

}  // extern "C"
