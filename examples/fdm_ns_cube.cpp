// Lid-driven cavity driver in the shape of the reference's test/test_ns_cube.cpp (README.md:16-17):
//   fdm_ns_cube --ns:nx=31 --ns:nz=31 --ns:Re=250 --ns:dt=0.01 --ns:steps=10000 [--out:prefix=run]
// (the README's nx=32 aborts in the reference itself: Dirichlet axes need 2^k-1 interior points.)
// Plotting (plplot PNG / VTK) is out of scope; --out:prefix dumps u,v,w,p as raw fp64 instead.
#include <chrono>
#include <cstdio>
#include <string>

#include "ns_cube.h"

using namespace fdm;

template <typename T>
static void dump(const std::string& fn, tensor<T, 3, false>& t)
{
    FILE* f = fopen(fn.c_str(), "wb");
    if (!f) { perror(fn.c_str()); return; }
    fwrite(t.vec, sizeof(T), (size_t)t.size, f);
    fclose(f);
}

int main(int argc, char** argv)
{
    Config c;
    c.open("ns_cube.ini");
    c.rewrite(argc, argv);
    NSCube<double, false> ns(c);
    const int steps = c.get("ns", "steps", 1);
    const int interval = c.get("plot", "interval", 100);
    const std::string prefix = c.get("out", "prefix", "");
    ns.auto_sync = false;
    auto t1 = std::chrono::steady_clock::now();
    for (int done = 0; done < steps;) {
        int n = std::min(interval, steps - done);
        ns.steps(n);
        done += n;
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
