// Particle-mesh N-body driver with the command line of the reference's test/nbody.cpp (:598-690):
//   fdm_nbody --nbody:n=32 --nbody:N=100000 --nbody:steps=100 [--nbody:dt=0.001 --nbody:G=1 --nbody:vel=4
//             --nbody:x0=-10 --nbody:y0=-10 --nbody:z0=-10 --nbody:l=20] [--nbody:deposit_all=0] [--out:prefix=run]
// Bodies are seeded exactly like init_points (:541-587: std::default_random_engine, uniform positions, masses
// 0.2 + 1.5 u, solid-rotation velocities scaled by vel / sqrt(R)), so with the same libstdc++ the run starts from
// the reference's state; the steps run on the device through fdmb_pm_*.  The short-range pair correction
// (--nbody:local=1), the solar-system preset and the O(N^2) error report are not built: they are refused.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "lapl_cube.h"      // FDMB_VERIFY, the C ABI
#if __has_include("config.h")
#include "config.h"
#else
#include "fdm_compat_config.h"
#endif

using namespace fdm;

int main(int argc, char** argv)
{
    Config c;
    c.open("ns_rect.ini");      // the reference program's file name (:682)
    c.rewrite(argc, argv);
    const int n = c.get("nbody", "n", 32);
    const int N = c.get("nbody", "N", 100000);
    const int steps = c.get("nbody", "steps", 50000);
    const double x0 = c.get("nbody", "x0", -10.0), y0 = c.get("nbody", "y0", -10.0), z0 = c.get("nbody", "z0", -10.0);
    const double l = c.get("nbody", "l", 20.0);
    const double dt = c.get("nbody", "dt", 0.001);
    const double G = c.get("nbody", "G", 1.0);
    const double vel = c.get("nbody", "vel", 4.0);
    const int interval = c.get("plot", "interval", 100);
    const std::string prefix = c.get("out", "prefix", "");
    if (c.get("nbody", "local", 0) || c.get("nbody", "solar", 0) || c.get("nbody", "error", 0)) {
        fprintf(stderr, "nbody:local / nbody:solar / nbody:error are not supported by this driver\n");
        return 2;
    }

    // init_points (:541-587)
    const double origin[3] = {x0, y0, z0};
    std::vector<double> x(3 * (size_t)N), v(3 * (size_t)N, 0.0), mass((size_t)N);
    std::default_random_engine generator;
    std::uniform_real_distribution<double> distribution(0.0, 1.0);
    for (int b = 0; b < N; b++) {
        double* xb = &x[3 * (size_t)b];
        for (int i = 0; i < 3; i++) xb[i] = l * distribution(generator) + origin[i];
        mass[b] = 0.2 + 1.5 * distribution(generator);
        const double R = std::sqrt(xb[0] * xb[0] + xb[1] * xb[1] + xb[2] * xb[2]);
        const double V = vel / std::sqrt(R);
        v[3 * (size_t)b + 0] = V * xb[1];
        v[3 * (size_t)b + 1] = -V * xb[0];
    }

    auto dump = [&](const char* name, const std::vector<double>& arr) {
        FILE* fp = fopen((prefix + name).c_str(), "wb");
        if (!fp) { perror(prefix.c_str()); exit(1); }
        fwrite(arr.data(), sizeof(double), arr.size(), fp);
        fclose(fp);
    };
    if (!prefix.empty()) { dump("_x0.bin", x); dump("_v0.bin", v); dump("_mass.bin", mass); }

    fdmb_pm_params prm;
    prm.x0 = x0; prm.y0 = y0; prm.z0 = z0; prm.l = l; prm.dt = dt; prm.G = G; prm.n = n;
    prm.deposit_all = c.get("nbody", "deposit_all", 0);
    fdmb_pm* pm = nullptr;
    FDMB_VERIFY(fdmb_pm_create(&pm, &prm));
    FDMB_VERIFY(fdmb_pm_set_bodies(pm, N, x.data(), v.data(), mass.data()));

    std::vector<double> a(3 * (size_t)N);
    auto t1 = std::chrono::steady_clock::now();
    for (int done = 0; done < steps;) {
        const int k = std::min(interval, steps - done);
        FDMB_VERIFY(fdmb_pm_step(pm, k));
        done += k;
        FDMB_VERIFY(fdmb_pm_get_bodies(pm, FDMB_PM_X, x.data()));
        FDMB_VERIFY(fdmb_pm_get_bodies(pm, FDMB_PM_A, a.data()));
        // the line move() prints every step for body 2 (:505-507), here once per interval
        if (N > 2) printf("step=%d %e %e %e %e \n", done, a[6], a[7], x[6], x[7]);
    }
    auto t2 = std::chrono::steady_clock::now();
    const double sec = std::chrono::duration<double>(t2 - t1).count();
    printf("total: %.2fms\n", steps > 0 ? 1000.0 * sec / steps : 0.0);
    if (!prefix.empty()) {
        FDMB_VERIFY(fdmb_pm_get_bodies(pm, FDMB_PM_V, v.data()));
        dump("_x.bin", x); dump("_v.bin", v); dump("_a.bin", a);
    }
    fdmb_pm_destroy(pm);
    return 0;
}
