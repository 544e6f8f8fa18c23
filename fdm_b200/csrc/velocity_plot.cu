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
#include "vplot_math.h"

namespace fdmb {

// The per-element arithmetic lives in vplot_math.h (shared with the host emulation test).

__global__ void k_vplot_slices(VGeom g, const double* __restrict__ u, const double* __restrict__ v,
                               const double* __restrict__ w, double* __restrict__ vx, double* __restrict__ wx,
                               double* __restrict__ uy, double* __restrict__ wy, double* __restrict__ uz,
                               double* __restrict__ vz)
{
    const long long n = vplot_slice_elems(g);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        vplot_slice_elem(g, t, u, v, w, vx, wx, uy, wy, uz, vz);
}

__global__ void k_vplot_rhs(VGeom g, const double* __restrict__ vx, const double* __restrict__ wx,
                            const double* __restrict__ uy, const double* __restrict__ wy,
                            const double* __restrict__ uz, const double* __restrict__ vz, double* __restrict__ rx,
                            double* __restrict__ ry, double* __restrict__ rz)
{
    const long long n = vplot_rhs_elems(g);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        vplot_rhs_elem(g, t, vx, wx, uy, wy, uz, vz, rx, ry, rz);
}

// One thread per cell, lanes along x; the three components are interleaved as the VTK file wants them.
__global__ void k_vplot_cells(VGeom g, const double* __restrict__ u, const double* __restrict__ v,
                              const double* __restrict__ w, double* __restrict__ out)
{
    const long long n = vplot_cell_elems(g);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        vplot_cell_elem(g, t, u, v, w, out);
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
    g = vplot_make_geom(p.nx, p.ny, p.nz, p.zperiodic, p.yperiodic, p.dx, p.dy, p.dz);
    ly = g.yper ? p.yy2 - p.yy1 : p.yy2 - p.yy1 + p.dy;      // src/velocity_plot.h:67-68
    lz = g.zper ? p.zz2 - p.zz1 : p.zz2 - p.zz1 + p.dz;
    for (int f = 0; f < 3; f++) fsz[f] = vplot_field_elems(g, f);

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
    // cudaMemset on device memory is asynchronous to the host and runs on the legacy default stream, which the
    // handle's non-blocking streams do not order against: finish it before the handle is handed out
    FDMB_CUDA(cudaDeviceSynchronize());
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
        const long long n = vplot_slice_elems(g);
        LaunchScope scope("vplot_slices", stream);
        k_vplot_slices<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(
            g, dfield[0], dfield[1], dfield[2], d_slice[FDMB_SLICE_VX], d_slice[FDMB_SLICE_WX], d_slice[FDMB_SLICE_UY],
            d_slice[FDMB_SLICE_WY], d_slice[FDMB_SLICE_UZ], d_slice[FDMB_SLICE_VZ]);
        FDMB_CHECK_LAUNCH();
    }
    {
        const long long n = vplot_rhs_elems(g);
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
