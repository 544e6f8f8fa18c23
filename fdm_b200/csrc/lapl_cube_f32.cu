// LaplCube in single precision: fdm::LaplCube<float,check,F> (reference instantiations src/lapl_cube.cpp:176-177,
// 181-182; solve src/lapl_cube.cpp:9-142 with T = float; constructor src/lapl_cube.h:58-100).
//
// Real fp32 kernels, not a cast around the fp64 path: the arrays on the device, the work array, the transform tables,
// the eigenvalue tables (the reference stores them as T, src/lapl_cube.h:53-54) and every butterfly are float, so a solve
// moves half the bytes of the fp64 one.  The sweeps are the plain tile kernels of xform_kernels.cuh instantiated for
// float (k_rows / k_cols: the shared-memory tile transforms of xform.cuh are templated on the element type); the TMA-fed
// and ring sweeps stay fp64-only: their tile shapes, bank-conflict swizzles and tensor-map boxes are derived for 8-byte
// elements.  Sweep structure as in lapl_cube.cu: x rows -> y columns -> [z forward, divide, z inverse] -> y -> x.
// Parity bar: the reference's own float instantiation (tests/test_f32_gpu.py).
#include <cmath>
#include <cstdint>
#include <new>
#include <vector>

#include "common.h"
#include "lapl_cube.h"

namespace fdmb {

struct TablesF {
    const float* SN;
    const cx<float>* WM;
};
int get_tables_f32(int N, TablesF* out);

// transform lengths instantiated in single precision
#define FDMB_FOR_EACH_N_F32(X) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024)
static inline bool supported_N_f32(int N) { return supported_N(N) && N <= 1024; }

template <int KIND> static cudaError_t rows_f(int N, const RowsArgsT<float>& a, cudaStream_t st)
{
#define X(NN) case NN: return launch_rows_t<NN, KIND, float>(a, st);
    switch (N) { FDMB_FOR_EACH_N_F32(X) }
#undef X
    return cudaErrorInvalidValue;
}
template <int KIND> static cudaError_t cols_f(int N, const ColsArgsT<float>& a, cudaStream_t st)
{
    MidNone mid;
#define X(NN) case NN: return launch_cols_t<NN, KIND, MidNone, XF_DST, float>(a, mid, st);
    switch (N) { FDMB_FOR_EACH_N_F32(X) }
#undef X
    return cudaErrorInvalidValue;
}
static cudaError_t cols_divide_f(int N, bool periodic, const ColsArgsT<float>& a, const MidCubeDivideF& mid, cudaStream_t st)
{
#define X(NN)                                                                                         \
    case NN:                                                                                          \
        if (periodic) return launch_cols_t<NN, XF_PFWD, MidCubeDivideF, XF_PINV, float>(a, mid, st);  \
        return launch_cols_t<NN, XF_DST, MidCubeDivideF, XF_DST, float>(a, mid, st);
    switch (N) { FDMB_FOR_EACH_N_F32(X) }
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t launch_rows_f32(int N, int kind, const RowsArgsT<float>& a, cudaStream_t st, const char* tag)
{
    LaunchScope scope(tag, st);
    if (kind == XF_DST) return rows_f<XF_DST>(N, a, st);
    if (kind == XF_PFWD) return rows_f<XF_PFWD>(N, a, st);
    return rows_f<XF_PINV>(N, a, st);
}
cudaError_t launch_cols_f32(int N, int kind, const ColsArgsT<float>& a, cudaStream_t st, const char* tag)
{
    LaunchScope scope(tag, st);
    if (kind == XF_DST) return cols_f<XF_DST>(N, a, st);
    if (kind == XF_PFWD) return cols_f<XF_PFWD>(N, a, st);
    return cols_f<XF_PINV>(N, a, st);
}

}  // namespace fdmb

using namespace fdmb;

struct fdmb_lapl_cube_f32 {
    double dx, dy, dz, lx, ly, lz;
    int nx, ny, nz, periodic;
    int Nx = 0, Ny = 0, Nz = 0;
    double slx = 0, sly = 0, slz = 0;
    int px = 0;                       // x pitch of the work array (floats): rows start 128-byte aligned
    TablesF tx{}, ty{}, tz{};
    cudaStream_t stream = nullptr;
    float *d_lmx = nullptr, *d_lmy = nullptr, *d_lmz = nullptr;
    float* d_work = nullptr;
    float *d_rhs = nullptr, *d_ans = nullptr;

    int init();
    int solve_device(float* d_out, const float* d_in, cudaStream_t st);
    int solve_host(float* ans, const float* rhs);
    ~fdmb_lapl_cube_f32()
    {
        cudaFree(d_lmx); cudaFree(d_lmy); cudaFree(d_lmz); cudaFree(d_work); cudaFree(d_rhs); cudaFree(d_ans);
        if (stream) cudaStreamDestroy(stream);
    }
};

static inline double sq(double x) { return x * x; }

int fdmb_lapl_cube_f32::init()
{
    Nx = periodic ? nx : nx + 1;
    Ny = periodic ? ny : ny + 1;
    Nz = periodic ? nz : nz + 1;
    if (nx < 1 || ny < 1 || nz < 1 || !supported_N_f32(Nx) || !supported_N_f32(Ny) || !supported_N_f32(Nz)) {
        set_error("LaplCube<float>: %s axis sizes (%d,%d,%d) need transform lengths that are powers of two in [4,1024] "
                  "(reference: verify((1<<n) == N), src/fft.cpp:67)", periodic ? "periodic" : "Dirichlet", nx, ny, nz);
        return FDMB_ERR_INVALID;
    }
    slx = std::sqrt(2. / lx); sly = std::sqrt(2. / ly); slz = std::sqrt(2. / lz);
    px = (nx + 31) / 32 * 32;
    int rc;
    if ((rc = get_tables_f32(Nx, &tx)) || (rc = get_tables_f32(Ny, &ty)) || (rc = get_tables_f32(Nz, &tz))) return rc;
    // eigenvalues: computed in double, stored as T = float like the reference (src/lapl_cube.cpp:145-172), with the
    // aliasing quirk of :162,171
    const int x1 = periodic ? 0 : 1, xn = periodic ? nx - 1 : nx;
    const int y1 = periodic ? 0 : 1, yn = periodic ? ny - 1 : ny;
    const int z1 = periodic ? 0 : 1, zn = periodic ? nz - 1 : nz;
    const double dx2 = dx * dx, dy2 = dy * dy, dz2 = dz * dz;
    std::vector<float> lm_y(ny + 1, 0.f), lm_x(nx + 1, 0.f), lm_z(nz + 1, 0.f);
    for (int k = y1; k <= yn; k++)
        lm_y[k] = (float)(periodic ? 4. / dy2 * sq(sin(k * M_PI / (ny))) : 4. / dy2 * sq(sin(k * M_PI * 0.5 / (ny + 1))));
    for (int j = x1; j <= xn; j++)
        lm_x[j] = (float)(periodic ? 4. / dx2 * sq(sin(j * M_PI / (nx))) : 4. / dx2 * sq(sin(j * M_PI * 0.5 / (nx + 1))));
    for (int i = z1; i <= zn; i++)
        lm_z[i] = (float)(periodic ? 4. / dz2 * sq(sin(i * M_PI / (nz))) : 4. / dz2 * sq(sin(i * M_PI * 0.5 / (nz + 1))));
    if (Nx == Ny) lm_x = lm_y;
    if (Nz == Ny) lm_z = lm_y;
    FDMB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    FDMB_CUDA(cudaMalloc(&d_lmx, sizeof(float) * (nx + 1)));
    FDMB_CUDA(cudaMalloc(&d_lmy, sizeof(float) * (ny + 1)));
    FDMB_CUDA(cudaMalloc(&d_lmz, sizeof(float) * (nz + 1)));
    FDMB_CUDA(cudaMemcpy(d_lmx, lm_x.data(), sizeof(float) * (nx + 1), cudaMemcpyHostToDevice));
    FDMB_CUDA(cudaMemcpy(d_lmy, lm_y.data(), sizeof(float) * (ny + 1), cudaMemcpyHostToDevice));
    FDMB_CUDA(cudaMemcpy(d_lmz, lm_z.data(), sizeof(float) * (nz + 1), cudaMemcpyHostToDevice));
    FDMB_CUDA(cudaMalloc(&d_work, sizeof(float) * (size_t)nz * ny * px));
    return FDMB_OK;
}

int fdmb_lapl_cube_f32::solve_device(float* d_out, const float* d_in, cudaStream_t st)
{
    const int kf = periodic ? XF_PFWD : XF_DST;
    const int ki = periodic ? XF_PINV : XF_DST;
    const long long plane = (long long)ny * px;
    // forward scale d * sqrt(2/l) per axis, inverse sqrt(2/l) (src/lapl_cube.h:64, src/lapl_cube.cpp:23,96), in float
    RowsArgsT<float> r{};
    r.in = d_in; r.out = d_work; r.nrows = (long long)nz * ny; r.nvalid = nx;
    r.in_pitch = nx; r.out_pitch = px; r.scale = (float)(dx * slx); r.SN = tx.SN; r.WM = tx.WM;
    FDMB_CUDA(launch_rows_f32(Nx, kf, r, st, "cube32_x_fwd"));
    ColsArgsT<float> c{};
    c.in = d_work; c.out = d_work; c.nvalid = ny; c.in_sj = c.out_sj = px; c.nb = nx; c.no = nz;
    c.in_so = c.out_so = plane; c.scale = (float)(dy * sly); c.SN = ty.SN; c.WM = ty.WM;
    FDMB_CUDA(launch_cols_f32(Ny, kf, c, st, "cube32_y_fwd"));
    ColsArgsT<float> z{};
    z.in = d_work; z.out = d_work; z.nvalid = nz; z.in_sj = z.out_sj = plane; z.nb = nx; z.no = ny;
    z.in_so = z.out_so = px; z.scale = (float)(dz * slz); z.scale2 = (float)slz; z.SN = tz.SN; z.WM = tz.WM;
    {
        LaunchScope scope("cube32_z_fwd_div_inv", st);
        MidCubeDivideF mid{d_lmz, d_lmx, d_lmy, periodic ? 1 : 0};
        FDMB_CUDA(cols_divide_f(Nz, periodic != 0, z, mid, st));
    }
    c.scale = (float)sly;
    FDMB_CUDA(launch_cols_f32(Ny, ki, c, st, "cube32_y_inv"));
    r.in = d_work; r.out = d_out; r.in_pitch = px; r.out_pitch = nx; r.scale = (float)slx;
    FDMB_CUDA(launch_rows_f32(Nx, ki, r, st, "cube32_x_inv"));
    return FDMB_OK;
}

int fdmb_lapl_cube_f32::solve_host(float* ans, const float* rhs)
{
    const size_t bytes = sizeof(float) * (size_t)nx * ny * nz;
    if (!d_rhs) FDMB_CUDA(cudaMalloc(&d_rhs, bytes));
    if (!d_ans) FDMB_CUDA(cudaMalloc(&d_ans, bytes));
    FDMB_CUDA(cudaMemcpyAsync(d_rhs, rhs, bytes, cudaMemcpyHostToDevice, stream));
    int rc = solve_device(d_ans, d_rhs, stream);
    if (rc) { cudaStreamSynchronize(stream); return rc; }
    FDMB_CUDA(cudaMemcpyAsync(ans, d_ans, bytes, cudaMemcpyDeviceToHost, stream));
    FDMB_CUDA(cudaStreamSynchronize(stream));
    return FDMB_OK;
}

extern "C" {

int fdmb_lapl_cube_f32_create(fdmb_lapl_cube_f32** out, double dx, double dy, double dz, double lx, double ly, double lz,
                              int nx, int ny, int nz, int periodic)
{
    if (!out) { set_error("null handle pointer"); return FDMB_ERR_INVALID; }
    *out = nullptr;
    auto* h = new (std::nothrow) fdmb_lapl_cube_f32();
    if (!h) { set_error("out of host memory"); return FDMB_ERR_NOMEM; }
    h->dx = dx; h->dy = dy; h->dz = dz; h->lx = lx; h->ly = ly; h->lz = lz;
    h->nx = nx; h->ny = ny; h->nz = nz; h->periodic = periodic ? 1 : 0;
    int rc = h->init();
    if (rc) { delete h; return rc; }
    *out = h;
    return FDMB_OK;
}

int fdmb_lapl_cube_f32_solve(fdmb_lapl_cube_f32* h, float* ans, const float* rhs)
{
    if (!h || !ans || !rhs) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return h->solve_host(ans, rhs);
}

int fdmb_lapl_cube_f32_solve_device(fdmb_lapl_cube_f32* h, float* d_ans, const float* d_rhs, void* stream)
{
    if (!h || !d_ans || !d_rhs) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return h->solve_device(d_ans, d_rhs, stream ? (cudaStream_t)stream : h->stream);
}

int fdmb_lapl_cube_f32_destroy(fdmb_lapl_cube_f32* h)
{
    delete h;
    return FDMB_OK;
}

}  // extern "C"
