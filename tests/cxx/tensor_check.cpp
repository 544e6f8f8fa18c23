// The constructions of the reference's ut/ut_tensor.cpp:46-118 (read/write with offsets, range-intersection
// assignment, periodic wrap) plus use(), maxabs(), norm2(), index(); prints every observable.  Built twice by
// tests/test_tensor_cpu.py: against the reference's own src/tensor.h and against fdm_b200/cxx/fdm_compat_tensor.h.
#include <cmath>
#include <cstdio>
#include <vector>
#ifdef USE_REFERENCE_TENSOR
#include "tensor.h"
extern "C" double cblas_dnrm2(int n, const double* x, int incx)
{
    long double s = 0;
    for (int i = 0; i < n; i++) s += (long double)x[(long)i * incx] * (long double)x[(long)i * incx];
    return std::sqrt((double)s);
}
#else
#include "fdm_compat_tensor.h"
#endif

using namespace fdm;

template <typename TT> static void dump(const char* name, TT& t)
{
    printf("%s size %d:", name, (int)t.size);
    for (long long i = 0; i < (long long)t.size; i++) printf(" %g", (double)t.vec[i]);
    printf("\n");
}

int main()
{
    using P1 = tensor_flags<tensor_flag::periodic>;
    using P2 = tensor_flags<tensor_flag::periodic, tensor_flag::periodic>;
    using NP = tensor_flags<tensor_flag::none, tensor_flag::periodic>;
    // 3-D with negative lower bounds, like NSCube::u (src/ns_cube.h:66)
    tensor<double, 3, true> u({0, 3, 0, 2, -1, 3});
    int c = 0;
    for (int i = 0; i <= 3; i++) for (int k = 0; k <= 2; k++) for (int j = -1; j <= 3; j++) u[i][k][j] = ++c;
    dump("u", u);
    printf("u[2][1][-1]=%g u[3][2][3]=%g maxabs=%g norm2=%.12g index=%d\n", u[2][1][-1], u[3][2][3], (double)u.maxabs(),
           (double)u.norm2(), u.index({2, 1, 0}));
    // range-intersection assignment: interior-only x into the larger p, and back (src/ns_cube.cpp:275)
    tensor<double, 3, true> p({0, 4, 0, 4, 0, 4}), x({1, 3, 1, 3, 1, 3}), q({2, 6, -2, 2, 3, 8});
    c = 100;
    for (int i = 1; i <= 3; i++) for (int k = 1; k <= 3; k++) for (int j = 1; j <= 3; j++) x[i][k][j] = ++c;
    p = x;
    dump("p=x", p);
    q = p;
    dump("q=p", q);
    x = q;
    dump("x=q", x);
    // periodic wrap on the first axis only (NSCyl fields, src/ns_cyl.h:21-22), reads and writes
    tensor<double, 3, false, P1> w({0, 3, -1, 2, 0, 1});
    for (int i = 0; i < (int)w.size; i++) w.vec[i] = i;
    // (indices more than one period BELOW the range are undefined in the reference: (y + len) % len stays negative,
    //  src/tensor.h:119-123; the stand-in wraps them, the probe stays within what the reference defines)
    printf("w[-1][0][1]=%g w[4][2][0]=%g w[9][-1][1]=%g w[-4][1][0]=%g\n", w[-1][0][1], w[4][2][0], w[9][-1][1], w[-4][1][0]);
    w[5][0][0] = -1; w[-2][2][1] = -2;
    dump("w", w);
    printf("w.index({-1,0,1})=%d\n", w.index({-1, 0, 1}));
    // 2-D: both axes periodic with a non-zero lower bound; second axis only
    tensor<double, 2, true, P2> m({1, 3, -2, 1});
    for (int i = 0; i < (int)m.size; i++) m.vec[i] = 10 + i;
    printf("m[0][-3]=%g m[4][2]=%g m[7][-6]=%g\n", m[0][-3], m[4][2], m[7][-6]);
    tensor<double, 2, false, NP> n2({0, 1, 0, 3});
    for (int i = 0; i < (int)n2.size; i++) n2.vec[i] = 20 + i;
    printf("n2[1][-1]=%g n2[0][5]=%g\n", n2[1][-1], n2[0][5]);
    // use(): rebinding to caller storage (test/test_ns_cube.cpp:39, src/velocity_plot.cpp:12-15)
    std::vector<double> ext(u.size, 0.5);
    tensor<double, 3, false> view({0, 3, 0, 2, -1, 3}, reinterpret_cast<double*>(0xF));
    view.use(ext.data());
    view[1][1][1] = 7;
    printf("ext[%d]=%g maxabs=%g\n", view.index({1, 1, 1}), ext[view.index({1, 1, 1})], (double)view.maxabs());
    // float instantiation
    tensor<float, 2, true> f({1, 2, 1, 3});
    for (int i = 1; i <= 2; i++) for (int j = 1; j <= 3; j++) f[i][j] = (float)(i * 10 + j) / 4;
    dump("f", f);
    printf("f.maxabs=%g\n", (double)f.maxabs());
    return 0;
}
