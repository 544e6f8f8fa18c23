// Exercises the drop-in C++ class headers (fdm_b200/cxx) the way the reference's own callers do
// (ut/ut_lapl_cube.cpp:39-125, test/test_ns_cube.cpp:24-53) and dumps results for the Python tests.
//   shim_check cube <n> <rhs.bin> <ans.bin>       LaplCube<double,true> Dirichlet n^3
//   shim_check cubep <n> <rhs.bin> <ans.bin>      LaplCube<double,true,periodic^3>
//   shim_check cubef <n> <rhs.bin> <ans.bin>      LaplCube<float,false> (fp32 at the boundary)
//   shim_check ns <n> <steps> <prefix> [--ns:...] NSCube<double,true>, dumps <prefix>_{u,v,w,p}.bin
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ns_cube.h"

using namespace fdm;

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
