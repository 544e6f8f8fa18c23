// Taylor-Couette driver with the command line of the reference's test/test_ns_cyl.cpp (README.md:8-10):
//   fdm_ns_cyl --ns:nr=32 --ns:nz=31 --ns:nphi=32 --ns:Re=200 --ns:dt=0.01 --ns:steps=10000 [--ns:zperiod=1]
//              [--ns:vrandom=1] [--plot:interval=100 --plot:png=1 --plot:vtk=0] [--out:prefix=run]
// (Dirichlet z needs nz = 2^k-1, periodic z nz = 2^k; the README's nz=32 with Dirichlet z aborts in the reference.)
// The state stays on the device between plot intervals; the (r, z, phi) slice plotter with the cylindrical column
// scales reads it in place.  VTK output needs periodic z, like the reference (src/velocity_plot.cpp:127).
// The reference driver's eigenvector stabilisation branch ([st] enable=1, test/test_ns_cyl.cpp:48-52,94-98) reads
// NetCDF eigenvector files produced by its ARPACK tooling and is not part of this path: it is refused, not ignored.
#include <chrono>
#include <cstdio>
#include <string>

#include "ns_cyl.h"
#include "velocity_plot.h"

using namespace fdm;

static std::string step_name(int time_index, const char* ext)
{
    char buf[64];
    snprintf(buf, sizeof(buf), "step_%07d.%s", time_index, ext);
    return buf;
}

template <typename V>
static void dump(const std::string& fn, V& t)
{
    FILE* f = fopen(fn.c_str(), "wb");
    if (!f) { perror(fn.c_str()); return; }
    fwrite(t.vec, sizeof(*t.vec), (size_t)t.size, f);
    fclose(f);
}

template <tensor_flag zflag>
static int calc(const Config& c)
{
    using Task = NSCyl<double, false, zflag>;
    Task ns(c);
    const int steps = c.get("ns", "steps", 1);
    const int interval = c.get("plot", "interval", 100);
    const int png = c.get("plot", "png", 1);
    int vtk = c.get("plot", "vtk", 0);
    const std::string prefix = c.get("out", "prefix", "");
    if (vtk && zflag != tensor_flag::periodic) {
        fprintf(stderr, "plot:vtk needs ns:zperiod=1 (reference: verify(zflag == periodic), src/velocity_plot.cpp:127)\n");
        vtk = 0;
    }
    ns.auto_sync = false;
    velocity_plotter<double, false, typename Task::tensor_flags> plot(ns.dr, ns.dz, ns.dphi, ns.nr, ns.nz, ns.nphi, ns.r0,
                                                                      ns.R, ns.h1, ns.h2, 0, 2 * M_PI, true);
    plot.set_labels("R", "Z", "PHI");
    plot.use(ns);
    auto output = [&]() {
        if (!png && !vtk) return;
        plot.update();
        if (png) plot.plot(step_name(ns.time_index, "png"), ns.time_index * ns.dt);
        if (vtk) plot.vtk_out(step_name(ns.time_index, "vtk"), ns.time_index);
    };
    output();
    auto t1 = std::chrono::steady_clock::now();
    for (int done = 0; done < steps;) {
        int n = std::min(interval, steps - done);
        ns.steps(n);
        done += n;
        if (n == interval) output();
    }
    ns.sync_to_host(false);
    auto t2 = std::chrono::steady_clock::now();
    printf("It took me '%f' seconds\n", std::chrono::duration<double>(t2 - t1).count());
    if (!prefix.empty()) {
        dump(prefix + "_u.bin", ns.u); dump(prefix + "_v.bin", ns.v);
        dump(prefix + "_w.bin", ns.w); dump(prefix + "_p.bin", ns.p);
    }
    return 0;
}

int main(int argc, char** argv)
{
    Config c;
    c.open("ns_rect.ini");      // the reference driver's file name (test/test_ns_cyl.cpp:126)
    c.rewrite(argc, argv);
    if (c.get("st", "enable", 0)) {
        fprintf(stderr, "[st] enable=1 (eigenvector stabilisation from NetCDF input) is not supported by this driver\n");
        return 2;
    }
    if (c.get("ns", "zperiod", 0) == 1) return calc<tensor_flag::periodic>(c);
    return calc<tensor_flag::none>(c);
}
