// Lid-driven cavity driver with the command line of the reference's test/test_ns_cube.cpp (README.md:16-17):
//   fdm_ns_cube --ns:nx=31 --ns:nz=31 --ns:Re=250 --ns:dt=0.01 --ns:steps=10000
//               [--plot:interval=100 --plot:png=1 --plot:vtk=0] [--out:prefix=run]
// (the README's nx=32 aborts in the reference itself: Dirichlet axes need 2^k-1 interior points.)
// Unlike the unmodified reference driver (which also builds against the drop-in headers) this one
// keeps the state on the device between plot intervals: the plotter reads it in place (plot.use(ns)), so a run moves
// only 2-D slices per interval, plus 24 B per cell when a VTK file is written.  --plot:png=1 writes the six panels of
// velocity_plotter::plot as step_NNNNNNN.ppm (no plplot needed); --out:prefix dumps u,v,w,p as raw fp64.
#include <chrono>
#include <cstdio>
#include <string>

#include "ns_cube.h"
#include "velocity_plot.h"

using namespace fdm;

template <typename T>
static void dump(const std::string& fn, tensor<T, 3, false>& t)
{
    FILE* f = fopen(fn.c_str(), "wb");
    if (!f) { perror(fn.c_str()); return; }
    fwrite(t.vec, sizeof(T), (size_t)t.size, f);
    fclose(f);
}

static std::string step_name(int time_index, const char* ext)
{
    char buf[64];
    snprintf(buf, sizeof(buf), "step_%07d.%s", time_index, ext);
    return buf;
}

int main(int argc, char** argv)
{
    Config c;
    c.open("ns_cube.ini");
    c.rewrite(argc, argv);
    NSCube<double, false> ns(c);
    const int steps = c.get("ns", "steps", 1);
    const int interval = c.get("plot", "interval", 100);
    const int png = c.get("plot", "png", 1);
    const int vtk = c.get("plot", "vtk", 0);
    const std::string prefix = c.get("out", "prefix", "");
    ns.auto_sync = false;
    velocity_plotter<double, false> plot(ns.dx, ns.dy, ns.dz, ns.nx, ns.ny, ns.nz, ns.x1, ns.x2, ns.y1, ns.y2, ns.z1, ns.z2);
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
        if (ns.verbose) {
            ns.sync_to_host(false);
            printf("%.1e: %.1e %.1e %.1e %.1e\n", ns.time_index * ns.dt, (double)ns.p.maxabs(), (double)ns.u.maxabs(),
                   (double)ns.v.maxabs(), (double)ns.w.maxabs());
        }
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
