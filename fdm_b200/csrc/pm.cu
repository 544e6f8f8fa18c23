// Particle-mesh N-body step on B200: the second in-tree consumer of the periodic LaplCube.
// Replaces the step of NBody<double,check,CIC3<double>> in the reference's test/nbody.cpp (:24-596, local = 0):
//   calc_a_pm()  :292-421  mean density + cloud-in-cell deposit -> 4 pi G rho -> periodic Poisson solve ->
//                          4-point field differencing -> cloud-in-cell gather
//   move()       :469-487  velocity Verlet with periodic wrap
// Bodies live on the device as structure-of-arrays (x[3][N], v[3][N], ...), the grids as [z][y][x] fp64 arrays; one
// step is six launches of this file (fill, deposit with fp64 atomics, rhs, field, gather + move) around the five
// sweeps of the periodic LaplCube on the same stream.  Nothing crosses PCIe between steps.
// The per-element arithmetic is in pm_math.h (shared with the host emulation test).
#include <cmath>
#include <new>
#include <vector>

#include "common.h"
#include "pm_math.h"

namespace fdmb {

__global__ void k_pm_fill(double* __restrict__ f, long long n3, double rho0)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n3; t += (long long)gridDim.x * blockDim.x)
        f[t] = rho0;
}

// one thread per body, lanes along the body index (coalesced SoA reads); eight fp64 atomics per deposited body
__global__ void k_pm_deposit(PMGeom g, const double* __restrict__ x, const double* __restrict__ mass, double* f)
{
    for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < g.N; b += (long long)gridDim.x * blockDim.x)
        pm_deposit_body(g, x[b], x[g.N + b], x[2 * g.N + b], mass[b], f, [](double* p, double v) { atomicAdd(p, v); });
}

__global__ void k_pm_rhs(PMGeom g, const double* __restrict__ f, double* __restrict__ rhs, long long n3)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n3; t += (long long)gridDim.x * blockDim.x)
        rhs[t] = pm_rhs(g, f[t]);
}

__global__ void k_pm_field(PMGeom g, const double* __restrict__ psi, double* __restrict__ E, long long n3)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n3; t += (long long)gridDim.x * blockDim.x)
        pm_field_elem(g, t, psi, E);
}

// gather (calc_accelerations) and, when do_move, the Verlet update of the same body
__global__ void k_pm_gather_move(PMGeom g, double* __restrict__ x, double* __restrict__ v, double* __restrict__ a,
                                 double* __restrict__ aprev, const double* __restrict__ E, int do_move)
{
    const long long N = g.N;
    for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < N; b += (long long)gridDim.x * blockDim.x) {
        double xb[3] = {x[b], x[N + b], x[2 * N + b]};
        double ab[3];
        pm_gather_body(g, xb[0], xb[1], xb[2], E, ab);
        for (int m = 0; m < 3; m++) a[m * N + b] = ab[m];
        if (do_move) {
            double vb[3] = {v[b], v[N + b], v[2 * N + b]};
            double pb[3] = {aprev[b], aprev[N + b], aprev[2 * N + b]};
            pm_move_body(g, xb, vb, ab, pb);
            for (int m = 0; m < 3; m++) { x[m * N + b] = xb[m]; v[m * N + b] = vb[m]; aprev[m * N + b] = pb[m]; }
        }
    }
}

}  // namespace fdmb

using namespace fdmb;

struct fdmb_pm {
    fdmb_pm_params p{};
    PMGeom g{};
    long long n3 = 0;
    fdmb_lapl_cube* solver = nullptr;
    double *d_f = nullptr, *d_rhs = nullptr, *d_psi = nullptr, *d_E = nullptr;
    double *d_x = nullptr, *d_v = nullptr, *d_a = nullptr, *d_aprev = nullptr, *d_mass = nullptr;   // [3][N] / [N]
    cudaStream_t stream = nullptr;

    int init();
    int set_bodies(long long N, const double* x, const double* v, const double* mass);
    int advance(int nsteps, int do_move);
    void free_bodies();
    ~fdmb_pm();
};

int fdmb_pm::init()
{
    if (p.n < 4 || !(p.l > 0)) { set_error("NBody: n >= 4 and l > 0 required"); return FDMB_ERR_INVALID; }
    g.n = p.n; g.N = 0; g.l = p.l; g.h = p.l / p.n; g.ox = p.x0; g.oy = p.y0; g.oz = p.z0; g.dt = p.dt; g.G = p.G;
    g.rho0 = 0; g.deposit_all = p.deposit_all ? 1 : 0; g.lgn = pm_log2_or_neg(p.n);
    n3 = (long long)p.n * p.n * p.n;
    // solver3(h,h,h,l,l,l,n,n,n), periodic on every axis (:128, :29-31)
    int rc = fdmb_lapl_cube_create(&solver, g.h, g.h, g.h, p.l, p.l, p.l, p.n, p.n, p.n, 1);
    if (rc) return rc;
    FDMB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    FDMB_CUDA(cudaMalloc(&d_f, sizeof(double) * n3));
    FDMB_CUDA(cudaMalloc(&d_rhs, sizeof(double) * n3));
    FDMB_CUDA(cudaMalloc(&d_psi, sizeof(double) * n3));
    FDMB_CUDA(cudaMalloc(&d_E, sizeof(double) * 3 * n3));
    FDMB_CUDA(cudaMemset(d_f, 0, sizeof(double) * n3));
    FDMB_CUDA(cudaMemset(d_rhs, 0, sizeof(double) * n3));
    FDMB_CUDA(cudaMemset(d_psi, 0, sizeof(double) * n3));
    FDMB_CUDA(cudaMemset(d_E, 0, sizeof(double) * 3 * n3));
    // cudaMemset on device memory is asynchronous to the host and runs on the legacy default stream, which the
    // handle's non-blocking streams do not order against: finish it before the handle is handed out
    FDMB_CUDA(cudaDeviceSynchronize());
    return FDMB_OK;
}

void fdmb_pm::free_bodies()
{
    cudaFree(d_x); cudaFree(d_v); cudaFree(d_a); cudaFree(d_aprev); cudaFree(d_mass);
    d_x = d_v = d_a = d_aprev = d_mass = nullptr;
}

fdmb_pm::~fdmb_pm()
{
    free_bodies();
    cudaFree(d_f); cudaFree(d_rhs); cudaFree(d_psi); cudaFree(d_E);
    if (solver) fdmb_lapl_cube_destroy(solver);
    if (stream) cudaStreamDestroy(stream);
}

int fdmb_pm::set_bodies(long long N, const double* x, const double* v, const double* mass)
{
    free_bodies();
    g.N = 0;
    if (N <= 0) { set_error("NBody: N must be positive"); return FDMB_ERR_INVALID; }
    // [N][3] (Body::x[3]) -> [3][N]; total mass summed in body order like init_points (:583)
    std::vector<double> t((size_t)(3 * N));
    double total = 0;
    for (long long b = 0; b < N; b++) {
        for (int m = 0; m < 3; m++) {
            const double c = x[3 * b + m], o = m == 0 ? p.x0 : m == 1 ? p.y0 : p.z0;
            if (!(c >= o && c < o + p.l)) {
                set_error("NBody: body %lld coordinate %d = %.17g outside [origin, origin + l)", b, m, c);
                return FDMB_ERR_INVALID;
            }
            t[(size_t)(m * N + b)] = c;
        }
        total += mass[b];
    }
    const size_t bytes = sizeof(double) * 3 * (size_t)N;
    FDMB_CUDA(cudaMalloc(&d_x, bytes));
    FDMB_CUDA(cudaMalloc(&d_v, bytes));
    FDMB_CUDA(cudaMalloc(&d_a, bytes));
    FDMB_CUDA(cudaMalloc(&d_aprev, bytes));
    FDMB_CUDA(cudaMalloc(&d_mass, sizeof(double) * (size_t)N));
    FDMB_CUDA(cudaMemcpy(d_x, t.data(), bytes, cudaMemcpyHostToDevice));
    for (long long b = 0; b < N; b++)
        for (int m = 0; m < 3; m++) t[(size_t)(m * N + b)] = v[3 * b + m];
    FDMB_CUDA(cudaMemcpy(d_v, t.data(), bytes, cudaMemcpyHostToDevice));
    FDMB_CUDA(cudaMemset(d_a, 0, bytes));
    FDMB_CUDA(cudaMemset(d_aprev, 0, bytes));
    // cudaMemset on device memory is asynchronous to the host and runs on the legacy default stream, which the
    // handle's non-blocking streams do not order against: finish it before the handle is handed out
    FDMB_CUDA(cudaDeviceSynchronize());
    FDMB_CUDA(cudaMemcpy(d_mass, mass, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice));
    g.N = N;
    g.rho0 = -total / p.l / p.l / p.l;
    return FDMB_OK;
}

int fdmb_pm::advance(int nsteps, int do_move)
{
    if (g.N <= 0) { set_error("NBody: step before set_bodies"); return FDMB_ERR_INVALID; }
    const int threads = 256;
    const long long cap = (long long)device_sm_count() * 8;
    auto blocks = [&](long long n) { long long b = (n + threads - 1) / threads; return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap); };
    for (int s = 0; s < nsteps; s++) {
        {
            LaunchScope scope("pm_fill", stream);
            k_pm_fill<<<blocks(n3), threads, 0, stream>>>(d_f, n3, g.rho0);
            FDMB_CHECK_LAUNCH();
        }
        {
            LaunchScope scope("pm_deposit", stream);
            k_pm_deposit<<<blocks(g.N), threads, 0, stream>>>(g, d_x, d_mass, d_f);
            FDMB_CHECK_LAUNCH();
        }
        {
            LaunchScope scope("pm_rhs", stream);
            k_pm_rhs<<<blocks(n3), threads, 0, stream>>>(g, d_f, d_rhs, n3);
            FDMB_CHECK_LAUNCH();
        }
        int rc = fdmb_lapl_cube_solve_device(solver, d_psi, d_rhs, stream);
        if (rc) return rc;
        {
            LaunchScope scope("pm_field", stream);
            k_pm_field<<<blocks(n3), threads, 0, stream>>>(g, d_psi, d_E, n3);
            FDMB_CHECK_LAUNCH();
        }
        {
            LaunchScope scope("pm_gather_move", stream);
            k_pm_gather_move<<<blocks(g.N), threads, 0, stream>>>(g, d_x, d_v, d_a, d_aprev, d_E, do_move);
            FDMB_CHECK_LAUNCH();
        }
    }
    FDMB_CUDA(cudaStreamSynchronize(stream));
    return FDMB_OK;
}

extern "C" {

int fdmb_pm_create(fdmb_pm** out, const fdmb_pm_params* p)
{
    if (!out || !p) { set_error("null argument"); return FDMB_ERR_INVALID; }
    *out = nullptr;
    auto* h = new (std::nothrow) fdmb_pm();
    if (!h) { set_error("out of host memory"); return FDMB_ERR_NOMEM; }
    h->p = *p;
    int rc = h->init();
    if (rc) { delete h; return rc; }
    *out = h;
    return FDMB_OK;
}

int fdmb_pm_set_bodies(fdmb_pm* h, long long N, const double* x, const double* v, const double* mass)
{
    if (!h || !x || !v || !mass) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return h->set_bodies(N, x, v, mass);
}

long long fdmb_pm_count(fdmb_pm* h) { return h ? h->g.N : -1; }

int fdmb_pm_calc_accel(fdmb_pm* h)
{
    if (!h) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return h->advance(1, 0);
}

int fdmb_pm_step(fdmb_pm* h, int nsteps)
{
    if (!h || nsteps < 0) { set_error("bad argument"); return FDMB_ERR_INVALID; }
    return h->advance(nsteps, 1);
}

int fdmb_pm_get_bodies(fdmb_pm* h, int field, double* host)
{
    if (!h || !host || field < FDMB_PM_X || field > FDMB_PM_MASS) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    const long long N = h->g.N;
    if (N <= 0) { set_error("NBody: no bodies"); return FDMB_ERR_INVALID; }
    if (field == FDMB_PM_MASS) {
        FDMB_CUDA(cudaMemcpy(host, h->d_mass, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost));
        return FDMB_OK;
    }
    const double* src = field == FDMB_PM_X ? h->d_x : field == FDMB_PM_V ? h->d_v : field == FDMB_PM_A ? h->d_a : h->d_aprev;
    std::vector<double> t((size_t)(3 * N));
    FDMB_CUDA(cudaMemcpy(t.data(), src, sizeof(double) * 3 * (size_t)N, cudaMemcpyDeviceToHost));
    for (long long b = 0; b < N; b++)
        for (int m = 0; m < 3; m++) host[3 * b + m] = t[(size_t)(m * N + b)];
    return FDMB_OK;
}

int fdmb_pm_get_grid(fdmb_pm* h, int grid, double* host)
{
    if (!h || !host || grid < FDMB_PM_F || grid > FDMB_PM_E) { set_error("bad grid id"); return FDMB_ERR_INVALID; }
    const double* src = grid == FDMB_PM_F ? h->d_f : grid == FDMB_PM_RHS ? h->d_rhs : grid == FDMB_PM_PSI ? h->d_psi : h->d_E;
    const size_t cnt = (size_t)h->n3 * (grid == FDMB_PM_E ? 3 : 1);
    FDMB_CUDA(cudaMemcpy(host, src, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
    return FDMB_OK;
}

int fdmb_pm_destroy(fdmb_pm* h)
{
    delete h;
    return FDMB_OK;
}

}  // extern "C"
