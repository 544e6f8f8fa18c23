// Host emulation of the particle-mesh N-body kernels: the very functions the CUDA kernels call
// (fdm_b200/csrc/pm_math.h) run on the CPU in the kernels' order.  TEST INFRASTRUCTURE; no GPU, no product path.
#include <vector>

#include "../../fdm_b200/csrc/pm_math.h"

using namespace fdmb;

extern "C" {

// One calc_a_pm() without the Poisson solve, in two halves so that the caller can put any solver in between:
//   part 1: x[3][N], mass -> f, rhs          part 2: psi -> E, a; optional move of x, v, aprev
void emul_pm_deposit(int n, long long N, double l, double ox, double oy, double oz, double G, double total_mass,
                     int deposit_all, const double* x, const double* mass, double* f, double* rhs)
{
    PMGeom g{};
    g.n = n; g.N = N; g.l = l; g.h = l / n; g.ox = ox; g.oy = oy; g.oz = oz; g.G = G; g.deposit_all = deposit_all; g.lgn = pm_log2_or_neg(n);
    g.rho0 = -total_mass / l / l / l;
    const long long n3 = (long long)n * n * n;
    for (long long t = 0; t < n3; t++) f[t] = g.rho0;
    for (long long b = 0; b < N; b++)
        pm_deposit_body(g, x[b], x[N + b], x[2 * N + b], mass[b], f, [](double* p, double v) { *p += v; });
    for (long long t = 0; t < n3; t++) rhs[t] = pm_rhs(g, f[t]);
}

void emul_pm_gather_move(int n, long long N, double l, double ox, double oy, double oz, double dt, const double* psi,
                         double* E, double* x, double* v, double* a, double* aprev, int do_move)
{
    PMGeom g{};
    g.n = n; g.N = N; g.l = l; g.h = l / n; g.ox = ox; g.oy = oy; g.oz = oz; g.dt = dt; g.lgn = pm_log2_or_neg(n);
    const long long n3 = (long long)n * n * n;
    for (long long t = 0; t < n3; t++) pm_field_elem(g, t, psi, E);
    for (long long b = 0; b < N; b++) {
        double xb[3] = {x[b], x[N + b], x[2 * N + b]}, ab[3];
        pm_gather_body(g, xb[0], xb[1], xb[2], E, ab);
        for (int m = 0; m < 3; m++) a[m * N + b] = ab[m];
        if (do_move) {
            double vb[3] = {v[b], v[N + b], v[2 * N + b]}, pb[3] = {aprev[b], aprev[N + b], aprev[2 * N + b]};
            pm_move_body(g, xb, vb, ab, pb);
            for (int m = 0; m < 3; m++) { x[m * N + b] = xb[m]; v[m * N + b] = vb[m]; aprev[m * N + b] = pb[m]; }
        }
    }
}

}  // extern "C"
