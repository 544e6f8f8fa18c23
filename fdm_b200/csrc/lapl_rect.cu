// LaplRect / LaplRectFFT2 on B200: 2-D Poisson-type solves on [y][x] arrays (x fastest).
// Replaces fdm::LaplRect<double,check,F>::solve (reference src/lapl_rect.cpp:63-110, init_Mat
// :44-60, constructor src/lapl_rect.h:40-68) and fdm::LaplRectFFT2<double,check,F>::solve
// (src/lapl_rect.cpp:113-207, constructor src/lapl_rect.h:89-105).
//
//   LaplRect     : y transform (strided axis) -> one tridiagonal system along x per y mode -> y inverse.
//                  x is always Dirichlet (:47).  The reference solves the systems with LAPACK gtsv
//                  (src/lapl_rect.cpp:90: Gaussian elimination with PARTIAL PIVOTING, info checked); here
//                  they are solved by the Thomas recurrence without pivoting, which gives the same
//                  answer (to round-off) exactly when gtsv never swaps rows, i.e. for row-wise diagonally
//                  dominant matrices.  That holds for the default scales and for the cylindrical column
//                  scales of src/velocity_plot.h:113-127 (|L| + |U| = 2 <= |D|); set_scales REJECTS scales
//                  for which it does not (FDMB_ERR_INVALID) instead of returning a silently different answer.
//   LaplRectFFT2 : y transform -> x transform -> divide by -(lm_y[k]*lm_y_scale[j] + lm_x[j])
//                  -> x inverse -> y inverse; the doubly periodic null mode is set to 1 (:169-172).
// The per-column scales lm_y_scale / L_scale / U_scale (src/lapl_rect.h:57-59) are public members
// of the reference that the cylindrical slice plotter overwrites (src/velocity_plot.h:113-127);
// fdmb_lapl_rect_set_scales mirrors that.
#include <cmath>
#include <cstdint>
#include <new>
#include <vector>

#include "common.h"
#include "lapl_cyl.h"

struct fdmb_lapl_rect {
    int kind, yperiodic, xperiodic;
    double dx, dy, lx, ly;
    int nx, ny;
    int Nx = 0, Ny = 0;                 // transform lengths
    double slx = 0, sly = 0;
    int px = 0;                         // x pitch of the work array
    fdmb::Tables tx{}, ty{};
    cudaStream_t stream = nullptr;
    double *d_lmy = nullptr, *d_lmx = nullptr;           // by array row / column (0-based)
    double *d_ysc = nullptr, *d_Lc = nullptr, *d_Uc = nullptr;   // lm_y_scale, L_scale/dx2, U_scale/dx2; nx+1 each
    double* d_zero = nullptr;
    std::vector<double> h_ysc, h_L, h_U;                 // host copies of the public scales (validation)
    double* d_work = nullptr;
    double *d_rhs = nullptr, *d_ans = nullptr;

    int init();
    int set_scales(const double* lm_y_scale, const double* L_scale, const double* U_scale);
    int solve_device(double* d_out, const double* d_in, cudaStream_t st);
    int solve_host(double* ans, const double* rhs);
    ~fdmb_lapl_rect();
};

namespace fdmb {

// RHSm[k][j] /= -lm_y[k]*lm_y_scale[j] - lm_x[j]   (lapl_rect.cpp:162-167), null-mode hack (:169-172)
__global__ void k_rect_divide(double* w, int ny, int nx, int px, const double* __restrict__ lmy,
                              const double* __restrict__ lmx, const double* __restrict__ ysc, int null_hack)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (j >= nx || k >= ny) return;
    double v = w[(long long)k * px + j] / (-lmy[k] * ysc[j] - lmx[j]);
    if (null_hack && j == 0 && k == 0) v = 1.0;
    w[(long long)k * px + j] = v;
}

}  // namespace fdmb

using namespace fdmb;

static inline double sq(double x) { return x * x; }

int fdmb_lapl_rect::init()
{
    if (kind == 0) xperiodic = 0;   // "TODO: matrix for periodic over x" -- the matrix is Dirichlet whatever F says
    Ny = yperiodic ? ny : ny + 1;
    Nx = xperiodic ? nx : nx + 1;
    if (nx < 2 || ny < 1 || !supported_N(Ny) || (kind == 1 && !supported_N(Nx))) {
        set_error("LaplRect: transform lengths (y %d%s) must be powers of two in [4,2048], nx >= 2 "
                  "(reference: verify((1<<n) == N), src/fft.cpp:67)", Ny, kind == 1 ? ", and x" : "");
        return FDMB_ERR_INVALID;
    }
    if (kind == 0 && tridiag_rows_smem(nx) > 220 * 1024) {
        set_error("LaplRect: nx=%d exceeds the shared-memory tile of the tridiagonal kernel", nx);
        return FDMB_ERR_INVALID;
    }
    slx = std::sqrt(2. / lx); sly = std::sqrt(2. / ly);
    px = (nx + 15) / 16 * 16;
    int rc;
    if ((rc = get_tables(Ny, &ty))) return rc;
    if (kind == 1 && (rc = get_tables(Nx, &tx))) return rc;
    const double dx2 = dx * dx, dy2 = dy * dy;
    const int y1 = yperiodic ? 0 : 1, x1 = xperiodic ? 0 : 1;
    std::vector<double> lm_y(ny), lm_x(nx);
    for (int r = 0; r < ny; r++) {
        const int k = r + y1;
        lm_y[r] = yperiodic ? 4. / dy2 * sq(sin(k * M_PI / ny)) : 4. / dy2 * sq(sin(k * M_PI * 0.5 / (ny + 1)));
    }
    for (int c = 0; c < nx; c++) {
        const int j = c + x1;
        lm_x[c] = xperiodic ? 4. / dx2 * sq(sin(j * M_PI / Nx)) : 4. / dx2 * sq(sin(j * M_PI * 0.5 / Nx));
    }
    // lm_x aliases lm_y only in the all-Dirichlet instantiation (lapl_rect.cpp:36-40)
    if (kind == 1 && !yperiodic && !xperiodic && Nx == Ny) lm_x = lm_y;
    FDMB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    FDMB_CUDA(cudaMalloc(&d_lmy, sizeof(double) * ny));
    FDMB_CUDA(cudaMalloc(&d_lmx, sizeof(double) * nx));
    FDMB_CUDA(cudaMemcpy(d_lmy, lm_y.data(), sizeof(double) * ny, cudaMemcpyHostToDevice));
    FDMB_CUDA(cudaMemcpy(d_lmx, lm_x.data(), sizeof(double) * nx, cudaMemcpyHostToDevice));
    FDMB_CUDA(cudaMalloc(&d_ysc, sizeof(double) * (nx + 1)));
    FDMB_CUDA(cudaMalloc(&d_Lc, sizeof(double) * (nx + 1)));
    FDMB_CUDA(cudaMalloc(&d_Uc, sizeof(double) * (nx + 1)));
    FDMB_CUDA(cudaMalloc(&d_zero, sizeof(double)));
    FDMB_CUDA(cudaMemset(d_zero, 0, sizeof(double)));
    // cudaMemset on device memory is asynchronous to the host and runs on the legacy default stream, which the
    // handle's non-blocking streams do not order against: finish it before the handle is handed out
    FDMB_CUDA(cudaDeviceSynchronize());
    FDMB_CUDA(cudaMalloc(&d_work, sizeof(double) * (size_t)ny * px));
    std::vector<double> ones(nx + 1, 1.0);   // lapl_rect.h:57-59
    return set_scales(ones.data(), ones.data(), ones.data());
}

int fdmb_lapl_rect::set_scales(const double* lm_y_scale, const double* L_scale, const double* U_scale)
{
    const double dx2 = dx * dx;
    std::vector<double> t(nx + 1);
    if (h_ysc.empty()) { h_ysc.assign(nx + 1, 1.0); h_L.assign(nx + 1, 1.0); h_U.assign(nx + 1, 1.0); }
    {   // validate before anything is changed: the no-pivot recurrence needs row-wise diagonal dominance,
        // |D_j| = 2/dx2 + lm_y[k] lm_y_scale[j] >= |L_j| + |U_j| for every mode k (lm_y >= 0, src/lapl_rect.cpp:44-60)
        const double* ys = lm_y_scale ? lm_y_scale : h_ysc.data();
        const double* Ls = L_scale ? L_scale : h_L.data();
        const double* Us = U_scale ? U_scale : h_U.data();
        for (int j = 1; j <= nx && kind == 0; j++) {
            const double off = (j > 1 ? std::fabs(Ls[j]) : 0.0) + (j < nx ? std::fabs(Us[j]) : 0.0);
            if (!(ys[j] >= 0.0) || !(off <= 2.0 * (1.0 + 1e-12))) {
                set_error("LaplRect: column scales at j=%d (lm_y_scale=%g, L_scale=%g, U_scale=%g) make the tridiagonal "
                          "systems non-dominant; the reference's gtsv would pivot (src/lapl_rect.cpp:90), the device "
                          "recurrence does not", j, ys[j], Ls[j], Us[j]);
                return FDMB_ERR_INVALID;
            }
        }
        if (lm_y_scale) h_ysc.assign(lm_y_scale, lm_y_scale + nx + 1);
        if (L_scale) h_L.assign(L_scale, L_scale + nx + 1);
        if (U_scale) h_U.assign(U_scale, U_scale + nx + 1);
    }
    if (lm_y_scale) FDMB_CUDA(cudaMemcpy(d_ysc, lm_y_scale, sizeof(double) * (nx + 1), cudaMemcpyHostToDevice));
    if (L_scale) {
        for (int j = 0; j <= nx; j++) t[j] = L_scale[j] / dx2;   // init_Mat: L_scale[j]/dx2
        FDMB_CUDA(cudaMemcpy(d_Lc, t.data(), sizeof(double) * (nx + 1), cudaMemcpyHostToDevice));
    }
    if (U_scale) {
        for (int j = 0; j <= nx; j++) t[j] = U_scale[j] / dx2;
        FDMB_CUDA(cudaMemcpy(d_Uc, t.data(), sizeof(double) * (nx + 1), cudaMemcpyHostToDevice));
    }
    return FDMB_OK;
}

fdmb_lapl_rect::~fdmb_lapl_rect()
{
    cudaFree(d_lmy); cudaFree(d_lmx); cudaFree(d_ysc); cudaFree(d_Lc); cudaFree(d_Uc); cudaFree(d_zero);
    cudaFree(d_work); cudaFree(d_rhs); cudaFree(d_ans);
    if (stream) cudaStreamDestroy(stream);
}

int fdmb_lapl_rect::solve_device(double* d_out, const double* d_in, cudaStream_t st)
{
    const int kyf = yperiodic ? XF_PFWD : XF_DST, kyi = yperiodic ? XF_PINV : XF_DST;
    const int kxf = xperiodic ? XF_PFWD : XF_DST, kxi = xperiodic ? XF_PINV : XF_DST;
    // y forward: caller rhs -> pitched work
    ColsArgs c{};
    c.in = d_in; c.out = d_work; c.nvalid = ny; c.in_sj = nx; c.out_sj = px; c.nb = nx; c.no = 1;
    c.in_so = 0; c.out_so = 0; c.scale = dy * sly; c.SN = ty.SN; c.WM = ty.WM;
    FDMB_CUDA(launch_cols(Ny, kyf, c, st, "rect_y_fwd"));
    if (kind == 0) {
        TridiagArgs t{};
        t.data = d_work; t.pitch = px; t.nr = nx; t.nmid = 1; t.nsys = ny;
        t.lm_outer = d_lmy; t.lm_mid = d_zero; t.mid0 = 0; t.c0 = -2 / (dx * dx);
        t.L = d_Lc; t.U = d_Uc; t.ir2 = d_ysc;
        FDMB_CUDA(launch_tridiag_rows(t, st, "rect_x_tridiag"));
    } else {
        RowsArgs r{};
        r.in = d_work; r.out = d_work; r.nrows = ny; r.nvalid = nx; r.in_pitch = px; r.out_pitch = px;
        r.scale = dx * slx; r.SN = tx.SN; r.WM = tx.WM;
        FDMB_CUDA(launch_rows(Nx, kxf, r, st, "rect_x_fwd"));
        {
            LaunchScope scope("rect_divide", st);
            dim3 grid((nx + 127) / 128, ny);
            // lm_y_scale is indexed by j = x1..xn: column c uses entry c + x1
            k_rect_divide<<<grid, 128, 0, st>>>(d_work, ny, nx, px, d_lmy, d_lmx, d_ysc + (xperiodic ? 0 : 1),
                                                (yperiodic && xperiodic) ? 1 : 0);
            FDMB_CHECK_LAUNCH();
        }
        r.scale = slx;
        FDMB_CUDA(launch_rows(Nx, kxi, r, st, "rect_x_inv"));
    }
    // y inverse: work -> caller ans
    c.in = d_work; c.out = d_out; c.in_sj = px; c.out_sj = nx; c.scale = sly;
    FDMB_CUDA(launch_cols(Ny, kyi, c, st, "rect_y_inv"));
    return FDMB_OK;
}

int fdmb_lapl_rect::solve_host(double* ans, const double* rhs)
{
    const size_t bytes = sizeof(double) * (size_t)nx * ny;
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

int fdmb_lapl_rect_create(fdmb_lapl_rect** out, int kind, int yperiodic, int xperiodic, double dx, double dy, double lx,
                          double ly, int nx, int ny)
{
    if (!out) { set_error("null handle pointer"); return FDMB_ERR_INVALID; }
    *out = nullptr;
    if (kind != 0 && kind != 1) { set_error("LaplRect: kind must be 0 (LaplRect) or 1 (LaplRectFFT2)"); return FDMB_ERR_INVALID; }
    auto* h = new (std::nothrow) fdmb_lapl_rect();
    if (!h) { set_error("out of host memory"); return FDMB_ERR_NOMEM; }
    h->kind = kind; h->yperiodic = yperiodic ? 1 : 0; h->xperiodic = xperiodic ? 1 : 0;
    h->dx = dx; h->dy = dy; h->lx = lx; h->ly = ly; h->nx = nx; h->ny = ny;
    int rc = h->init();
    if (rc) { delete h; return rc; }
    *out = h;
    return FDMB_OK;
}

int fdmb_lapl_rect_set_scales(fdmb_lapl_rect* h, const double* lm_y_scale, const double* L_scale, const double* U_scale)
{
    if (!h) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return h->set_scales(lm_y_scale, L_scale, U_scale);
}

int fdmb_lapl_rect_solve(fdmb_lapl_rect* h, double* ans, const double* rhs)
{
    if (!h || !ans || !rhs) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return h->solve_host(ans, rhs);
}

int fdmb_lapl_rect_solve_device(fdmb_lapl_rect* h, double* d_ans, const double* d_rhs, void* stream)
{
    if (!h || !d_ans || !d_rhs) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return h->solve_device(d_ans, d_rhs, stream ? (cudaStream_t)stream : h->stream);
}

int fdmb_lapl_rect_destroy(fdmb_lapl_rect* h)
{
    delete h;
    return FDMB_OK;
}

}  // extern "C"
