// velocity_plotter on B200: the step after NSCube::step / NSCyl::step in both reference drivers.
// Replaces fdm::velocity_plotter<double,check,F> (reference src/velocity_plot.h:11-143):
//   update()   src/velocity_plot.cpp:17-67   mid-plane face averages -> vorticity-like right-hand sides ->
//                                            three 2-D stream-function solves (LaplRectFFT2 in the x = const
//                                            plane, LaplRect in the other two)
//   vtk_out()  src/velocity_plot.cpp:117-220 ASCII VTK: structured points (box) or hexahedra (cylinder)
// The reference reads the 3-D host arrays of the NS object.  Here the NS state stays in HBM: one kernel gathers
// the six slices, one forms the three right-hand sides, the 2-D solvers run on the same stream, and only the
// 2-D results (or, for the VTK file, the cell-centred velocity triples) are copied to the host.
// The slice / right-hand-side arithmetic has no multiply-add pair, so it is bit-identical to the reference;
// the stream functions inherit the 1e-12 parity of the 2-D solvers.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "common.h"

namespace fdmb {

// Index ranges of src/velocity_plot.h:73-83 and the array extents of :85-99.
struct VGeom {
    int nx, ny, nz;
    int zper, yper;
    int y_, y1, yn, ynn;    // y0 = z0 = 0
    int z_, z1, zn, znn;
    int Yc, Zc;             // points in y0..ynn, z0..znn
    int Yv;                 // points in y_..ynn (v)
    int Yi, Zi;             // points in y1..yn, z1..zn
    double dx, dy, dz;
};

__device__ __forceinline__ int wrap_z(const VGeom& g, int i) { return g.zper ? (i + g.nz) % g.nz : i; }
__device__ __forceinline__ int wrap_y(const VGeom& g, int k) { return g.yper ? (k + g.ny) % g.ny : k; }
// u[z0..znn][y0..ynn][-1..nx+1], v[z0..znn][y_..ynn][0..nx+1], w[z_..znn][y0..ynn][0..nx+1]
__device__ __forceinline__ long long iu(const VGeom& g, int i, int k, int j)
{
    return ((long long)i * g.Yc + k) * (g.nx + 3) + (j + 1);
}
__device__ __forceinline__ long long iv(const VGeom& g, int i, int k, int j)
{
    return ((long long)i * g.Yv + (k - g.y_)) * (g.nx + 2) + j;
}
__device__ __forceinline__ long long iw(const VGeom& g, int i, int k, int j)
{
    return ((long long)(i - g.z_) * g.Yc + k) * (g.nx + 2) + j;
}

// src/velocity_plot.cpp:19-38: face averages on the planes x = nx/2, y = ny/2, z = nz/2
__global__ void k_vplot_slices(VGeom g, const double* __restrict__ u, const double* __restrict__ v,
                               const double* __restrict__ w, double* __restrict__ vx, double* __restrict__ wx,
                               double* __restrict__ uy, double* __restrict__ wy, double* __restrict__ uz,
                               double* __restrict__ vz)
{
    const int X2 = g.nx + 2;
    const long long n0 = (long long)g.Zc * g.Yc, n1 = (long long)g.Zc * X2, n2 = (long long)g.Yc * X2;
    const int jm = g.nx / 2, km = g.ny / 2, im = g.nz / 2;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n0 + n1 + n2;
         t += (long long)gridDim.x * blockDim.x) {
        if (t < n0) {
            const int i = (int)(t / g.Yc), k = (int)(t % g.Yc);
            vx[t] = 0.5 * (v[iv(g, i, wrap_y(g, k - 1), jm)] + v[iv(g, i, k, jm)]);
            wx[t] = 0.5 * (w[iw(g, wrap_z(g, i - 1), k, jm)] + w[iw(g, i, k, jm)]);
        } else if (t < n0 + n1) {
            const long long s = t - n0;
            const int i = (int)(s / X2), j = (int)(s % X2);
            uy[s] = 0.5 * (u[iu(g, i, km, j - 1)] + u[iu(g, i, km, j)]);
            wy[s] = 0.5 * (w[iw(g, wrap_z(g, i - 1), km, j)] + w[iw(g, i, km, j)]);
        } else {
            const long long s = t - n0 - n1;
            const int k = (int)(s / X2), j = (int)(s % X2);
            uz[s] = 0.5 * (u[iu(g, im, k, j - 1)] + u[iu(g, im, k, j)]);
            vz[s] = 0.5 * (v[iv(g, im, wrap_y(g, k - 1), j)] + v[iv(g, im, k, j)]);
        }
    }
}

// src/velocity_plot.cpp:40-64: centred differences of the slices (periodic axes wrap inside the slice)
__global__ void k_vplot_rhs(VGeom g, const double* __restrict__ vx, const double* __restrict__ wx,
                            const double* __restrict__ uy, const double* __restrict__ wy,
                            const double* __restrict__ uz, const double* __restrict__ vz, double* __restrict__ rx,
                            double* __restrict__ ry, double* __restrict__ rz)
{
    const int X2 = g.nx + 2;
    const long long n0 = (long long)g.Zi * g.Yi, n1 = (long long)g.Zi * g.nx, n2 = (long long)g.Yi * g.nx;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n0 + n1 + n2;
         t += (long long)gridDim.x * blockDim.x) {
        if (t < n0) {
            const int i = (int)(t / g.Yi) + g.z1, k = (int)(t % g.Yi) + g.y1;
            const int kp = wrap_y(g, k + 1), kmn = wrap_y(g, k - 1), ip = wrap_z(g, i + 1), imn = wrap_z(g, i - 1);
            const double a = wx[(long long)i * g.Yc + kp] - wx[(long long)i * g.Yc + kmn];
            const double b = vx[(long long)ip * g.Yc + k] - vx[(long long)imn * g.Yc + k];
            rx[t] = a / 2 / g.dy - b / 2 / g.dz;
        } else if (t < n0 + n1) {
            const long long s = t - n0;
            const int i = (int)(s / g.nx) + g.z1, j = (int)(s % g.nx) + 1;
            const int ip = wrap_z(g, i + 1), imn = wrap_z(g, i - 1);
            const double a = wy[(long long)i * X2 + j + 1] - wy[(long long)i * X2 + j - 1];
            const double b = uy[(long long)ip * X2 + j] - uy[(long long)imn * X2 + j];
            ry[s] = a / 2 / g.dx - b / 2 / g.dz;
        } else {
            const long long s = t - n0 - n1;
            const int k = (int)(s / g.nx) + g.y1, j = (int)(s % g.nx) + 1;
            const int kp = wrap_y(g, k + 1), kmn = wrap_y(g, k - 1);
            const double a = vz[(long long)k * X2 + j + 1] - vz[(long long)k * X2 + j - 1];
            const double b = uz[(long long)kp * X2 + j] - uz[(long long)kmn * X2 + j];
            rz[s] = a / 2 / g.dx - b / 2 / g.dy;
        }
    }
}

// src/velocity_plot.cpp:180-182,209-213: cell-centred velocity = average of the two faces, i=z1..zn, k=y1..yn, j=1..nx.
// One thread per cell, lanes along x; the three components are interleaved as the VTK file wants them.
__global__ void k_vplot_cells(VGeom g, const double* __restrict__ u, const double* __restrict__ v,
                              const double* __restrict__ w, double* __restrict__ out)
{
    const long long n = (long long)g.Zi * g.Yi * g.nx;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % g.nx) + 1;
        const long long r = t / g.nx;
        const int k = (int)(r % g.Yi) + g.y1, i = (int)(r / g.Yi) + g.z1;
        out[3 * t + 0] = 0.5 * (u[iu(g, i, k, j)] + u[iu(g, i, k, j - 1)]);
        out[3 * t + 1] = 0.5 * (v[iv(g, i, k, j)] + v[iv(g, i, wrap_y(g, k - 1), j)]);
        out[3 * t + 2] = 0.5 * (w[iw(g, i, k, j)] + w[iw(g, wrap_z(g, i - 1), k, j)]);
    }
}

}  // namespace fdmb

using namespace fdmb;

struct fdmb_vplot {
    fdmb_vplot_params p{};
    VGeom g{};
    double ly = 0, lz = 0;
    long long fsz[3] = {0, 0, 0};                           // elements of u, v, w
    const double* dfield[3] = {nullptr, nullptr, nullptr};  // device arrays update() reads
    const double* hfield[3] = {nullptr, nullptr, nullptr};  // use_host
    double* staging[3] = {nullptr, nullptr, nullptr};       // device copies of the host arrays
    fdmb_ns_cube* cube = nullptr;
    fdmb_ns_cyl* cylns = nullptr;
    fdmb_lapl_rect *lapl_x = nullptr, *lapl_y = nullptr, *lapl_z = nullptr;
    double* d_slices = nullptr;
    double* d_slice[FDMB_SLICE_COUNT] = {};
    int rows[FDMB_SLICE_COUNT] = {}, cols[FDMB_SLICE_COUNT] = {};
    double* d_cells = nullptr;
    cudaStream_t stream = nullptr;
    bool updated = false;

    int init();
    int refresh_inputs();
    int update();
    int cells_to_host(double* host);
    int vtk_out(const char* name, int time_index);
    ~fdmb_vplot();
};

int fdmb_vplot::init()
{
    if (p.nx < 2 || p.ny < 2 || p.nz < 2) { set_error("velocity_plotter: nx, ny, nz must be >= 2"); return FDMB_ERR_INVALID; }
    if (p.yperiodic && !p.zperiodic) {
        // the reference instantiates F = <>, <periodic>, <periodic,periodic> only (src/velocity_plot.cpp:222-235)
        set_error("velocity_plotter: periodic y needs periodic z (tensor_flags<periodic,periodic>)");
        return FDMB_ERR_INVALID;
    }
    g.nx = p.nx; g.ny = p.ny; g.nz = p.nz; g.zper = p.zperiodic ? 1 : 0; g.yper = p.yperiodic ? 1 : 0;
    g.dx = p.dx; g.dy = p.dy; g.dz = p.dz;
    g.y_ = g.yper ? 0 : -1; g.y1 = g.yper ? 0 : 1; g.yn = g.yper ? p.ny - 1 : p.ny; g.ynn = g.yper ? p.ny - 1 : p.ny + 1;
    g.z_ = g.zper ? 0 : -1; g.z1 = g.zper ? 0 : 1; g.zn = g.zper ? p.nz - 1 : p.nz; g.znn = g.zper ? p.nz - 1 : p.nz + 1;
    g.Yc = g.ynn + 1; g.Zc = g.znn + 1; g.Yv = g.ynn - g.y_ + 1;
    g.Yi = g.yn - g.y1 + 1; g.Zi = g.zn - g.z1 + 1;
    ly = g.yper ? p.yy2 - p.yy1 : p.yy2 - p.yy1 + p.dy;      // src/velocity_plot.h:67-68
    lz = g.zper ? p.zz2 - p.zz1 : p.zz2 - p.zz1 + p.dz;
    fsz[0] = (long long)g.Zc * g.Yc * (p.nx + 3);
    fsz[1] = (long long)g.Zc * g.Yv * (p.nx + 2);
    fsz[2] = (long long)(g.znn - g.z_ + 1) * g.Yc * (p.nx + 2);

    // src/velocity_plot.h:101-105: lapl_x(dy,dz,ly,lz,ny,nz) on [z][y]; lapl_y(dx,dz,..,nx,nz) on [z][x];
    // lapl_z(dx,dy,..,nx,ny) on [y][x]
    int rc;
    if ((rc = fdmb_lapl_rect_create(&lapl_x, 1, g.zper, g.yper, p.dy, p.dz, ly, lz, p.ny, p.nz))) return rc;
    if ((rc = fdmb_lapl_rect_create(&lapl_y, 0, g.zper, 0, p.dx, p.dz, p.xx2 - p.xx1 + p.dx, lz, p.nx, p.nz))) return rc;
    if ((rc = fdmb_lapl_rect_create(&lapl_z, 0, g.yper, 0, p.dx, p.dy, p.xx2 - p.xx1 + p.dx, ly, p.nx, p.ny))) return rc;
    if (p.cyl) {
        // src/velocity_plot.h:113-127
        std::vector<double> ysc(p.nx + 1, 1.0), Ls(p.nx + 1, 1.0), Us(p.nx + 1, 1.0);
        for (int j = 1; j <= p.nx; j++) {
            const double r = p.xx1 + j * p.dx - p.dx / 2;
            ysc[j] = 1. / r / r;
            Us[j] = (r + p.dx / 2) / r;
            Ls[j] = (r - p.dx / 2) / r;
        }
        if ((rc = fdmb_lapl_rect_set_scales(lapl_y, ysc.data(), Ls.data(), Us.data()))) return rc;
        if ((rc = fdmb_lapl_rect_set_scales(lapl_z, nullptr, Ls.data(), Us.data()))) return rc;
    }

    const int X2 = p.nx + 2;
    const int r_[FDMB_SLICE_COUNT] = {g.Zc, g.Zc, g.Zc, g.Zc, g.Yc, g.Yc, g.Zi, g.Zi, g.Yi, g.Zi, g.Zi, g.Yi};
    const int c_[FDMB_SLICE_COUNT] = {g.Yc, g.Yc, X2, X2, X2, X2, g.Yi, p.nx, p.nx, g.Yi, p.nx, p.nx};
    long long total = 0;
    for (int s = 0; s < FDMB_SLICE_COUNT; s++) { rows[s] = r_[s]; cols[s] = c_[s]; total += ((long long)r_[s] * c_[s] + 15) / 16 * 16; }
    FDMB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    FDMB_CUDA(cudaMalloc(&d_slices, sizeof(double) * total));
    FDMB_CUDA(cudaMemset(d_slices, 0, sizeof(double) * total));
    long long off = 0;
    for (int s = 0; s < FDMB_SLICE_COUNT; s++) { d_slice[s] = d_slices + off; off += ((long long)rows[s] * cols[s] + 15) / 16 * 16; }
    return FDMB_OK;
}

fdmb_vplot::~fdmb_vplot()
{
    if (lapl_x) fdmb_lapl_rect_destroy(lapl_x);
    if (lapl_y) fdmb_lapl_rect_destroy(lapl_y);
    if (lapl_z) fdmb_lapl_rect_destroy(lapl_z);
    for (double* s : staging) cudaFree(s);
    cudaFree(d_slices); cudaFree(d_cells);
    if (stream) cudaStreamDestroy(stream);
}

// Makes dfield[] current: uploads host arrays, or waits for the NS handle's stream.
int fdmb_vplot::refresh_inputs()
{
    if (hfield[0]) {
        for (int f = 0; f < 3; f++) {
            if (!staging[f]) FDMB_CUDA(cudaMalloc(&staging[f], sizeof(double) * fsz[f]));
            FDMB_CUDA(cudaMemcpyAsync(staging[f], hfield[f], sizeof(double) * fsz[f], cudaMemcpyHostToDevice, stream));
            dfield[f] = staging[f];
        }
    } else if (cube) {
        int rc = fdmb_ns_cube_synchronize(cube);
        if (rc) return rc;
    } else if (cylns) {
        int rc = fdmb_ns_cyl_synchronize(cylns);
        if (rc) return rc;
    }
    if (!dfield[0] || !dfield[1] || !dfield[2]) {
        set_error("velocity_plotter: update() before use() (the reference would dereference its 0xF placeholder, "
                  "src/velocity_plot.h:97-99)");
        return FDMB_ERR_INVALID;
    }
    return FDMB_OK;
}

int fdmb_vplot::update()
{
    int rc = refresh_inputs();
    if (rc) return rc;
    const int threads = 256;
    {
        const long long n = (long long)g.Zc * g.Yc + (long long)(g.Zc + g.Yc) * (p.nx + 2);
        LaunchScope scope("vplot_slices", stream);
        k_vplot_slices<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(
            g, dfield[0], dfield[1], dfield[2], d_slice[FDMB_SLICE_VX], d_slice[FDMB_SLICE_WX], d_slice[FDMB_SLICE_UY],
            d_slice[FDMB_SLICE_WY], d_slice[FDMB_SLICE_UZ], d_slice[FDMB_SLICE_VZ]);
        FDMB_CHECK_LAUNCH();
    }
    {
        const long long n = (long long)g.Zi * g.Yi + (long long)(g.Zi + g.Yi) * p.nx;
        LaunchScope scope("vplot_rhs", stream);
        k_vplot_rhs<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(
            g, d_slice[FDMB_SLICE_VX], d_slice[FDMB_SLICE_WX], d_slice[FDMB_SLICE_UY], d_slice[FDMB_SLICE_WY],
            d_slice[FDMB_SLICE_UZ], d_slice[FDMB_SLICE_VZ], d_slice[FDMB_SLICE_RHS_X], d_slice[FDMB_SLICE_RHS_Y],
            d_slice[FDMB_SLICE_RHS_Z]);
        FDMB_CHECK_LAUNCH();
    }
    if ((rc = fdmb_lapl_rect_solve_device(lapl_x, d_slice[FDMB_SLICE_PSI_X], d_slice[FDMB_SLICE_RHS_X], stream))) return rc;
    if ((rc = fdmb_lapl_rect_solve_device(lapl_y, d_slice[FDMB_SLICE_PSI_Y], d_slice[FDMB_SLICE_RHS_Y], stream))) return rc;
    if ((rc = fdmb_lapl_rect_solve_device(lapl_z, d_slice[FDMB_SLICE_PSI_Z], d_slice[FDMB_SLICE_RHS_Z], stream))) return rc;
    FDMB_CUDA(cudaStreamSynchronize(stream));
    updated = true;
    return FDMB_OK;
}

int fdmb_vplot::cells_to_host(double* host)
{
    int rc = refresh_inputs();
    if (rc) return rc;
    const long long n = (long long)g.Zi * g.Yi * p.nx;
    if (!d_cells) FDMB_CUDA(cudaMalloc(&d_cells, sizeof(double) * 3 * n));
    {
        const int threads = 256;
        long long blocks = (n + threads - 1) / threads;
        const long long cap = (long long)device_sm_count() * 16;
        if (blocks > cap) blocks = cap;
        LaunchScope scope("vplot_cells", stream);
        k_vplot_cells<<<(unsigned)blocks, threads, 0, stream>>>(g, dfield[0], dfield[1], dfield[2], d_cells);
        FDMB_CHECK_LAUNCH();
    }
    FDMB_CUDA(cudaMemcpyAsync(host, d_cells, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, stream));
    FDMB_CUDA(cudaStreamSynchronize(stream));
    return FDMB_OK;
}

namespace {

// Formats lines [0, n) with fmt(line, buf) -> length (each at most 256 bytes), several host threads per block of
// lines, and writes them in order: the file is byte-identical to a sequential fprintf loop.
template <typename Fmt>
void write_lines(FILE* f, long long n, Fmt fmt)
{
    const long long block = 1 << 20;
    unsigned hw = std::thread::hardware_concurrency();
    const int nt = (int)(hw == 0 ? 1 : (hw > 16 ? 16 : hw));
    std::vector<std::string> parts(nt);
    for (long long b0 = 0; b0 < n; b0 += block) {
        const long long b1 = b0 + block < n ? b0 + block : n;
        const long long per = (b1 - b0 + nt - 1) / nt;
        auto work = [&](int t) {
            std::string& s = parts[t];
            s.clear();
            char buf[256];
            const long long lo = b0 + t * per, hi = lo + per < b1 ? lo + per : b1;
            for (long long i = lo; i < hi; i++) {
                const int len = fmt(i, buf);
                s.append(buf, (size_t)len);
            }
        };
        if (nt == 1 || b1 - b0 < 4096) {
            for (int t = 0; t < nt; t++) work(t);
        } else {
            std::vector<std::thread> th;
            for (int t = 1; t < nt; t++) th.emplace_back(work, t);
            work(0);
            for (auto& x : th) x.join();
        }
        for (int t = 0; t < nt; t++) fwrite(parts[t].data(), 1, parts[t].size(), f);
    }
}

}  // namespace

int fdmb_vplot::vtk_out(const char* name, int time_index)
{
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    const double dx = p.dx, dy = p.dy, dz = p.dz, xx1 = p.xx1, yy1 = p.yy1, zz1 = p.zz1;
    if (p.cyl && !g.zper) {
        set_error("velocity_plotter::vtk_out: cyl needs periodic z (reference: verify(zflag == periodic), "
                  "src/velocity_plot.cpp:127)");
        return FDMB_ERR_INVALID;
    }
    const long long ncell = (long long)g.Zi * g.Yi * nx;
    std::vector<double> c((size_t)(3 * ncell));
    int rc = cells_to_host(c.data());
    if (rc) return rc;
    FILE* f = fopen(name, "wb");
    if (!f) { set_error("velocity_plotter::vtk_out: cannot open %s", name); return FDMB_ERR_INVALID; }
    fprintf(f, "# vtk DataFile Version 3.0\n");
    fprintf(f, "step %d\n", time_index);
    fprintf(f, "ASCII\n");
    if (p.cyl) {
        // hexahedra between the grid nodes (r, z, phi); the node count of the header uses yn like the reference
        fprintf(f, "DATASET UNSTRUCTURED_GRID\n");
        fprintf(f, "POINTS %d double\n", (nx + 1) * (g.yn + 1) * nz);
        const int X1 = nx + 1, Y1 = ny + 1;
        write_lines(f, (long long)nz * Y1 * X1, [&](long long t, char* buf) {
            const int j = (int)(t % X1), k = (int)((t / X1) % Y1), i = (int)(t / ((long long)X1 * Y1));
            const double r = xx1 + dx * j, z = yy1 + dy * k, phi = dz * i;
            return snprintf(buf, 256, "%f %f %f\n", r * cos(phi), r * sin(phi), z);
        });
        const int l = nz * ny * nx;
        fprintf(f, "CELLS %d %d\n", l, 9 * l);
        auto node = [&](int i, int k, int j) { return (i % nz) * X1 * Y1 + k * X1 + j; };
        write_lines(f, (long long)l, [&](long long t, char* buf) {
            const int j = (int)(t % nx), k = (int)((t / nx) % ny), i = (int)(t / ((long long)nx * ny));
            return snprintf(buf, 256, "8 %d %d %d %d %d %d %d %d\n", node(i, k, j), node(i + 1, k, j),
                            node(i + 1, k, j + 1), node(i, k, j + 1), node(i, k + 1, j), node(i + 1, k + 1, j),
                            node(i + 1, k + 1, j + 1), node(i, k + 1, j + 1));
        });
        fprintf(f, "CELL_TYPES %d\n", l);
        write_lines(f, (long long)l, [&](long long, char* buf) { memcpy(buf, "12\n", 3); return 3; });   // VTK_HEXAHEDRON
        fprintf(f, "CELL_DATA %d\n", nx * ny * nz);
        fprintf(f, "VECTORS u double\n");
        const int Yi = g.Yi;
        write_lines(f, ncell, [&](long long t, char* buf) {
            const int j = (int)(t % nx) + 1, i = (int)(t / ((long long)nx * Yi));
            const double phi = dz * i + dz / 2;
            const double r = xx1 + dx * j + dx / 2;
            double u0 = c[3 * t], v0 = c[3 * t + 1], w0 = c[3 * t + 2];
            double len = sqrt(u0 * u0 + v0 * v0 + r * r * w0 * w0);
            u0 /= len; v0 /= len; w0 /= len;
            double x = u0 * cos(phi) - w0 * sin(phi);
            double y = u0 * sin(phi) + w0 * cos(phi);
            double z = v0;
            if (std::abs(len) < 1e-7) len = 1e-4;   // the reference's "hack" for cells at rest
            x *= len; y *= len; z *= len;
            return snprintf(buf, 256, "%f %f %f\n", x, y, z);
        });
    } else {
        fprintf(f, "DATASET STRUCTURED_POINTS\n");
        fprintf(f, "DIMENSIONS %d %d %d\n", nx, ny, nz);
        fprintf(f, "ASPECT_RATIO 1 1 1\n");
        fprintf(f, "ORIGIN %f %f %f\n", xx1, yy1, zz1);
        fprintf(f, "SPACING %f %f %f\n", dx, dy, dz);
        fprintf(f, "POINT_DATA %d\n", nx * ny * nz);
        fprintf(f, "VECTORS u double\n");
        write_lines(f, ncell, [&](long long t, char* buf) {
            return snprintf(buf, 256, "%f %f %f\n", c[3 * t], c[3 * t + 1], c[3 * t + 2]);
        });
    }
    fclose(f);
    return FDMB_OK;
}

extern "C" {

int fdmb_vplot_create(fdmb_vplot** out, const fdmb_vplot_params* p)
{
    if (!out || !p) { set_error("null argument"); return FDMB_ERR_INVALID; }
    *out = nullptr;
    auto* h = new (std::nothrow) fdmb_vplot();
    if (!h) { set_error("out of host memory"); return FDMB_ERR_NOMEM; }
    h->p = *p;
    int rc = h->init();
    if (rc) { delete h; return rc; }
    *out = h;
    return FDMB_OK;
}

int fdmb_vplot_field_size(fdmb_vplot* h, int field, long long* count)
{
    if (!h || field < FDMB_FIELD_U || field > FDMB_FIELD_W || !count) { set_error("bad field id"); return FDMB_ERR_INVALID; }
    *count = h->fsz[field];
    return FDMB_OK;
}

static void vplot_clear_sources(fdmb_vplot* h)
{
    for (int f = 0; f < 3; f++) { h->hfield[f] = nullptr; h->dfield[f] = nullptr; }
    h->cube = nullptr; h->cylns = nullptr;
}

int fdmb_vplot_use_host(fdmb_vplot* h, const double* u, const double* v, const double* w)
{
    if (!h || !u || !v || !w) { set_error("null argument"); return FDMB_ERR_INVALID; }
    vplot_clear_sources(h);
    h->hfield[0] = u; h->hfield[1] = v; h->hfield[2] = w;
    return FDMB_OK;
}

int fdmb_vplot_use_device(fdmb_vplot* h, const double* d_u, const double* d_v, const double* d_w)
{
    if (!h || !d_u || !d_v || !d_w) { set_error("null argument"); return FDMB_ERR_INVALID; }
    vplot_clear_sources(h);
    h->dfield[0] = d_u; h->dfield[1] = d_v; h->dfield[2] = d_w;
    return FDMB_OK;
}

int fdmb_vplot_use_ns_cube(fdmb_vplot* h, fdmb_ns_cube* ns)
{
    if (!h || !ns) { set_error("null argument"); return FDMB_ERR_INVALID; }
    void* ptr[3];
    for (int f = 0; f < 3; f++) {
        long long cnt = 0;
        int rc = fdmb_ns_cube_field_size(ns, f, &cnt);
        if (rc) return rc;
        if (cnt != h->fsz[f]) {
            set_error("velocity_plotter: the NSCube field %d has %lld elements, the plotter expects %lld "
                      "(sharded handles and mismatched nx/ny/nz are not supported)", f, cnt, h->fsz[f]);
            return FDMB_ERR_INVALID;
        }
        if ((rc = fdmb_ns_cube_field_device_ptr(ns, f, &ptr[f]))) return rc;
    }
    vplot_clear_sources(h);
    for (int f = 0; f < 3; f++) h->dfield[f] = (const double*)ptr[f];
    h->cube = ns;
    return FDMB_OK;
}

int fdmb_vplot_use_ns_cyl(fdmb_vplot* h, fdmb_ns_cyl* ns)
{
    if (!h || !ns) { set_error("null argument"); return FDMB_ERR_INVALID; }
    void* ptr[3];
    for (int f = 0; f < 3; f++) {
        long long cnt = 0;
        int rc = fdmb_ns_cyl_field_size(ns, f, &cnt);
        if (rc) return rc;
        if (cnt != h->fsz[f]) {
            set_error("velocity_plotter: the NSCyl field %d has %lld elements, the plotter expects %lld "
                      "(sharded handles and mismatched nr/nz/nphi or z flag are not supported)", f, cnt, h->fsz[f]);
            return FDMB_ERR_INVALID;
        }
        if ((rc = fdmb_ns_cyl_field_device_ptr(ns, f, &ptr[f]))) return rc;
    }
    vplot_clear_sources(h);
    for (int f = 0; f < 3; f++) h->dfield[f] = (const double*)ptr[f];
    h->cylns = ns;
    return FDMB_OK;
}

int fdmb_vplot_update(fdmb_vplot* h)
{
    if (!h) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return h->update();
}

int fdmb_vplot_slice_dims(fdmb_vplot* h, int slice, int* rows, int* cols)
{
    if (!h || slice < 0 || slice >= FDMB_SLICE_COUNT || !rows || !cols) { set_error("bad slice id"); return FDMB_ERR_INVALID; }
    *rows = h->rows[slice]; *cols = h->cols[slice];
    return FDMB_OK;
}

int fdmb_vplot_get_slice(fdmb_vplot* h, int slice, double* host)
{
    if (!h || slice < 0 || slice >= FDMB_SLICE_COUNT || !host) { set_error("bad slice id"); return FDMB_ERR_INVALID; }
    FDMB_CUDA(cudaMemcpyAsync(host, h->d_slice[slice], sizeof(double) * (size_t)h->rows[slice] * h->cols[slice],
                              cudaMemcpyDeviceToHost, h->stream));
    FDMB_CUDA(cudaStreamSynchronize(h->stream));
    return FDMB_OK;
}

int fdmb_vplot_cell_velocity(fdmb_vplot* h, double* host)
{
    if (!h || !host) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return h->cells_to_host(host);
}

int fdmb_vplot_vtk_out(fdmb_vplot* h, const char* name, int time_index)
{
    if (!h || !name) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return h->vtk_out(name, time_index);
}

int fdmb_vplot_destroy(fdmb_vplot* h)
{
    delete h;
    return FDMB_OK;
}

}  // extern "C"
