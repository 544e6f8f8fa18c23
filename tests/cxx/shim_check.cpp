// Exercises the drop-in C++ class headers (fdm_b200/cxx) the way the reference's own callers do
// (ut/ut_lapl_cube.cpp:39-125, test/test_ns_cube.cpp:24-53) and dumps results for the Python tests.
//   shim_check cube <n> <rhs.bin> <ans.bin>       LaplCube<double,true> Dirichlet n^3
//   shim_check cubep <n> <rhs.bin> <ans.bin>      LaplCube<double,true,periodic^3>
//   shim_check cubef <n> <rhs.bin> <ans.bin>      LaplCube<float,false> (fp32 at the boundary)
//   shim_check ns <n> <steps> <prefix> [--ns:...] NSCube<double,true>, dumps <prefix>_{u,v,w,p}.bin
//   shim_check rect|rectfft2 <nx> <ny> <rhs.bin> <ans.bin>   LaplRect / LaplRectFFT2 <double,true>, cylindrical
//                                                 column scales written through the public vectors
//   shim_check cyl <nr> <nz> <nphi> <rhs.bin> <ans.bin>      LaplCyl3FFT2<double,true>
//   shim_check nscyl <steps> <lsteps> <prefix> [--ns:...]    NSCyl<double,true>, dumps <prefix>_{u,v,w,p}.bin
//   shim_check nscylspec <lsteps> <prefix> <x.bin> [--ns:...]  the host-write pattern of test/test_ns_cyl_spectral.cpp
//   shim_check vplot <n> <steps> <prefix> [--ns:...]         NSCube + velocity_plotter as in test/test_ns_cube.cpp:24-50:
//                                                 host use(u,v,w), device use(ns), float instantiation; dumps
//                                                 <prefix>_{host,dev,flt}_psi_{x,y,z}.bin, <prefix>_{host,dev}.vtk, .ppm
//   shim_check vplotcyl <steps> <zero> <prefix> [--ns:...]   NSCyl + velocity_plotter as in test/test_ns_cyl.cpp:53-92
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ns_cube.h"
#include "ns_cyl.h"
#include "lapl_rect.h"
#include "velocity_plot.h"

using namespace fdm;

// every member of the veneers compiles, also the ones no mode below calls (solve_device, solve_batch, ...)
template class fdm::LaplCube<double, true>;
template class fdm::LaplCube<float, false>;

static std::vector<double> slurp(const char* fn, size_t n)
{
    std::vector<double> v(n);
    FILE* f = fopen(fn, "rb");
    if (!f || fread(v.data(), 8, n, f) != n) { fprintf(stderr, "cannot read %s\n", fn); exit(2); }
    fclose(f);
    return v;
}
static void spit(const std::string& fn, const double* p, size_t n)
{
    FILE* f = fopen(fn.c_str(), "wb");
    fwrite(p, 8, n, f);
    fclose(f);
}

int main(int argc, char** argv)
{
    if (argc < 5) return 1;
    std::string mode = argv[1];
    int n = atoi(argv[2]);
    if (mode == "cube" || mode == "cubep" || mode == "cubef") {
        size_t cnt = (size_t)n * n * n;
        auto rhs = slurp(argv[3], cnt);
        std::vector<double> ans(cnt);
        if (mode == "cube") {
            double dx = 1.0 / n, l = 1 + dx;
            LaplCube<double, true> lapl(dx, dx, dx, l, l, l, n, n, n);
            lapl.solve(ans.data(), rhs.data());
        } else if (mode == "cubep") {
            double dx = 2 * M_PI / n, l = 2 * M_PI;
            using P3 = tensor_flags<tensor_flag::periodic, tensor_flag::periodic, tensor_flag::periodic>;
            LaplCube<double, true, P3> lapl(dx, dx, dx, l, l, l, n, n, n);
            lapl.solve(ans.data(), rhs.data());
        } else {
            double dx = 1.0 / n, l = 1 + dx;
            LaplCube<float, false> lapl(dx, dx, dx, l, l, l, n, n, n);
            std::vector<float> r(rhs.begin(), rhs.end()), a(cnt);
            lapl.solve(a.data(), r.data());
            for (size_t i = 0; i < cnt; i++) ans[i] = a[i];
        }
        spit(argv[4], ans.data(), cnt);
        return 0;
    }
    if (mode == "rect" || mode == "rectfft2") {
        // the slice solver of the cylindrical plotter: src/velocity_plot.h:113-127
        int nx = atoi(argv[2]), ny = atoi(argv[3]);
        double x1 = M_PI / 2, dx = (M_PI / 2) / nx, dy = 10.0 / ny;
        auto rhs = slurp(argv[4], (size_t)nx * ny);
        std::vector<double> ans((size_t)nx * ny);
        auto scales = [&](auto& lapl) {
            for (int j = 1; j <= nx; j++) {
                double r = x1 + j * dx - dx / 2;
                lapl.lm_y_scale[j] = 1. / r / r; lapl.U_scale[j] = (r + dx / 2) / r; lapl.L_scale[j] = (r - dx / 2) / r;
            }
        };
        if (mode == "rect") {
            LaplRect<double, true> lapl(dx, dy, M_PI / 2 + dx, 10.0 + dy, nx, ny);
            scales(lapl);
            lapl.solve(ans.data(), rhs.data());
        } else {
            LaplRectFFT2<double, true> lapl(dx, dy, M_PI / 2 + dx, 10.0 + dy, nx, ny);
            scales(lapl);
            lapl.solve(ans.data(), rhs.data());
        }
        spit(argv[5], ans.data(), ans.size());
        return 0;
    }
    if (mode == "cyl") {
        if (argc < 7) return 1;
        int nr = atoi(argv[2]), nz = atoi(argv[3]), nphi = atoi(argv[4]);
        double R = M_PI, r0 = M_PI / 2, h = 10.0, dr = (R - r0) / nr, dz = h / nz;
        size_t cnt = (size_t)nr * nz * nphi;
        auto rhs = slurp(argv[5], cnt);
        std::vector<double> ans(cnt);
        LaplCyl3FFT2<double, true> lapl(dr, dz, r0 - dr / 2, R - r0 + dr, h + dz, nr, nz, nphi);   // src/ns_cyl.h:94-97
        lapl.solve(ans.data(), rhs.data());
        spit(argv[6], ans.data(), cnt);
        return 0;
    }
    if (mode == "nscyl") {
        int steps = atoi(argv[2]), lsteps = atoi(argv[3]);
        std::string prefix = argv[4];
        std::vector<char*> args{argv[0]};
        for (int i = 5; i < argc; i++) args.push_back(argv[i]);
        Config c;
        c.rewrite((int)args.size(), args.data());
        NSCyl<double, true> ns(c);
        for (int i = 0; i < steps; i++) ns.step();
        if (lsteps) {
            // linearise about the current state, like test/test_ns_cyl_spectral.cpp:47-58
            for (long long i = 0; i < (long long)ns.u.size; i++) ns.u0.vec[i] = ns.u.vec[i];
            for (long long i = 0; i < (long long)ns.v.size; i++) ns.v0.vec[i] = ns.v.vec[i];
            for (long long i = 0; i < (long long)ns.w.size; i++) ns.w0.vec[i] = ns.w.vec[i];
            ns.sync_to_device();
            for (int i = 0; i < lsteps; i++) ns.L_step();
        }
        spit(prefix + "_u.bin", ns.u.vec, ns.u.size); spit(prefix + "_v.bin", ns.v.vec, ns.v.size);
        spit(prefix + "_w.bin", ns.w.vec, ns.w.size); spit(prefix + "_p.bin", ns.p.vec, ns.p.size);
        printf("time_index %d size %d\n", ns.time_index, ns.size());
        return 0;
    }
    if (mode == "nscylspec") {
        // The access pattern of test/test_ns_cyl_spectral.cpp:16,72-104 with NO B200-specific call: write the Couette
        // profile into ns.w0 through the public tensor, set the public member U0 to 0, assign the state tensors from
        // caller-owned storage (range intersection, tensor.h:103-111), run L_step(), read the state back -- twice,
        // like two ARPACK iterations.  Arguments: <lsteps> <prefix> <x.bin> [--ns:...]; x.bin holds u|v|w|p.
        int lsteps = atoi(argv[2]);
        std::string prefix = argv[3];
        std::vector<char*> args{argv[0]};
        for (int i = 5; i < argc; i++) args.push_back(argv[i]);
        Config c;
        c.rewrite((int)args.size(), args.data());
        using Task = NSCyl<double, true, tensor_flag::periodic>;
        using tensor = typename Task::tensor;
        Task ns(c);
        const int nphi = ns.nphi, nz = ns.nz, nr = ns.nr;
        tensor u{{0, nphi - 1, 0, nz - 1, 1, nr - 1}};
        tensor v{{0, nphi - 1, 0, nz - 1, 1, nr}};
        tensor w{{0, nphi - 1, 0, nz - 1, 1, nr}};
        tensor p({0, nphi - 1, 0, nz - 1, 1, nr});
        const size_t n = u.size + v.size + w.size + p.size;
        auto xin = slurp(argv[4], n);
        std::vector<double> y(n);
        for (int i = 0; i < nphi; i++)
            for (int k = 0; k < nz; k++)
                for (int j = 0; j <= nr; j++) {
                    double r = ns.r0 + ns.dr * j + ns.dr / 2;
                    ns.w0[i][k][j] = -ns.U0 * ns.r0 * ns.r0 / (ns.R * ns.R - ns.r0 * ns.r0)
                                     + ns.U0 * ns.r0 * ns.r0 * ns.R * ns.R / (ns.R * ns.R - ns.r0 * ns.r0) / r / r;
                }
        ns.U0 = 0;
        for (int iter = 0; iter < 2; iter++) {
            std::memcpy(y.data(), xin.data(), n * sizeof(double));
            size_t off = 0;
            u.use(y.data() + off); off += u.size;
            v.use(y.data() + off); off += v.size;
            w.use(y.data() + off); off += w.size;
            p.use(y.data() + off); off += p.size;
            ns.u = u; ns.v = v; ns.w = w; ns.p = p;
            for (int i = 0; i < lsteps; i++) ns.L_step();
            u = ns.u; v = ns.v; w = ns.w; p = ns.p;
            spit(prefix + "_y" + std::to_string(iter) + ".bin", y.data(), n);
            for (size_t i = 0; i < n; i++) xin[i] = y[i];          // next "ARPACK vector"
        }
        printf("time_index %d n %zu\n", ns.time_index, n);
        return 0;
    }
    if (mode == "vplot") {
        int steps = atoi(argv[3]);
        std::string prefix = argv[4];
        std::string a1 = "--ns:nx=" + std::to_string(n), a2 = "--ns:nz=" + std::to_string(n);
        std::vector<char*> args{argv[0], a1.data(), a2.data()};
        for (int i = 5; i < argc; i++) args.push_back(argv[i]);
        Config c;
        c.rewrite((int)args.size(), args.data());
        NSCube<double, true> ns(c);
        velocity_plotter<double, true> plot(ns.dx, ns.dy, ns.dz, ns.nx, ns.ny, ns.nz, ns.x1, ns.x2, ns.y1, ns.y2, ns.z1, ns.z2);
        plot.use(ns.u.vec, ns.v.vec, ns.w.vec);            // the reference's way: host mirrors
        for (int i = 0; i < steps; i++) ns.step();
        plot.update();
        spit(prefix + "_host_psi_x.bin", plot.psi_x.vec, plot.psi_x.size);
        spit(prefix + "_host_psi_y.bin", plot.psi_y.vec, plot.psi_y.size);
        spit(prefix + "_host_psi_z.bin", plot.psi_z.vec, plot.psi_z.size);
        plot.vtk_out(prefix + "_host.vtk", ns.time_index);
        velocity_plotter<double, true> dplot(ns.dx, ns.dy, ns.dz, ns.nx, ns.ny, ns.nz, ns.x1, ns.x2, ns.y1, ns.y2, ns.z1, ns.z2);
        dplot.use(ns);                                     // device state in place
        dplot.update();
        spit(prefix + "_dev_psi_x.bin", dplot.psi_x.vec, dplot.psi_x.size);
        spit(prefix + "_dev_psi_y.bin", dplot.psi_y.vec, dplot.psi_y.size);
        spit(prefix + "_dev_psi_z.bin", dplot.psi_z.vec, dplot.psi_z.size);
        dplot.vtk_out(prefix + "_dev.vtk", ns.time_index);
        dplot.plot(prefix + "_dev.png", ns.time_index * ns.dt);
        // index like a reference caller: psi_y[i][j], i = z1..zn, j = 1..nx
        if (!(std::abs(dplot.psi_y[ns.nz / 2][ns.nx / 2]) > 0)) { fprintf(stderr, "psi_y mirror empty\n"); return 3; }
        std::vector<float> fu(ns.u.vec, ns.u.vec + ns.u.size), fv(ns.v.vec, ns.v.vec + ns.v.size),
            fw(ns.w.vec, ns.w.vec + ns.w.size);
        velocity_plotter<float, false> fplot(ns.dx, ns.dy, ns.dz, ns.nx, ns.ny, ns.nz, ns.x1, ns.x2, ns.y1, ns.y2, ns.z1, ns.z2);
        fplot.use(fu.data(), fv.data(), fw.data());
        fplot.update();
        std::vector<double> t(fplot.psi_y.vec, fplot.psi_y.vec + fplot.psi_y.size);
        spit(prefix + "_flt_psi_y.bin", t.data(), t.size());
        spit(prefix + "_u.bin", ns.u.vec, ns.u.size); spit(prefix + "_v.bin", ns.v.vec, ns.v.size);
        spit(prefix + "_w.bin", ns.w.vec, ns.w.size);
        printf("time_index %d\n", ns.time_index);
        return 0;
    }
    if (mode == "vplotcyl") {
        int steps = atoi(argv[2]);
        std::string prefix = argv[4];
        std::vector<char*> args{argv[0]};
        for (int i = 5; i < argc; i++) args.push_back(argv[i]);
        Config c;
        c.rewrite((int)args.size(), args.data());
        using Task = NSCyl<double, true, tensor_flag::periodic>;
        Task ns(c);
        velocity_plotter<double, true, typename Task::tensor_flags> plot(ns.dr, ns.dz, ns.dphi, ns.nr, ns.nz, ns.nphi,
                                                                         ns.r0, ns.R, ns.h1, ns.h2, 0, 2 * M_PI, true);
        plot.set_labels("R", "Z", "PHI");
        plot.use(ns);
        for (int i = 0; i < steps; i++) ns.step();
        plot.update();
        spit(prefix + "_dev_psi_x.bin", plot.psi_x.vec, plot.psi_x.size);
        spit(prefix + "_dev_psi_y.bin", plot.psi_y.vec, plot.psi_y.size);
        spit(prefix + "_dev_psi_z.bin", plot.psi_z.vec, plot.psi_z.size);
        plot.vtk_out(prefix + "_dev.vtk", ns.time_index);
        spit(prefix + "_u.bin", ns.u.vec, ns.u.size); spit(prefix + "_v.bin", ns.v.vec, ns.v.size);
        spit(prefix + "_w.bin", ns.w.vec, ns.w.size);
        printf("time_index %d\n", ns.time_index);
        return 0;
    }
    if (mode == "ns") {
        int steps = atoi(argv[3]);
        std::string prefix = argv[4];
        std::string a1 = "--ns:nx=" + std::to_string(n), a2 = "--ns:nz=" + std::to_string(n);
        std::vector<char*> args{argv[0], a1.data(), a2.data()};
        for (int i = 5; i < argc; i++) args.push_back(argv[i]);
        Config c;
        c.rewrite((int)args.size(), args.data());
        NSCube<double, true> ns(c);
        for (int i = 0; i < steps; i++) ns.step();
        // index through the host mirrors like a reference caller would
        double lid = ns.u[ns.nz + 1][1][1];
        if (!(std::abs(lid) > 0)) { fprintf(stderr, "lid ghost row not mirrored\n"); return 3; }
        spit(prefix + "_u.bin", ns.u.vec, ns.u.size); spit(prefix + "_v.bin", ns.v.vec, ns.v.size);
        spit(prefix + "_w.bin", ns.w.vec, ns.w.size); spit(prefix + "_p.bin", ns.p.vec, ns.p.size);
        printf("time_index %d\n", ns.time_index);
        return 0;
    }
    return 1;
}
