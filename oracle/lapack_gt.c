/*
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Restatement of the three LAPACK tridiagonal routines the reference's
 * cylindrical / rectangular solvers call through src/blas.h:12-22,65-91:
 *   dgtsv_  (lapl_rect.cpp:90)   dgttrf_ (lapl_cyl.cpp:166)   dgttrs_ (lapl_cyl.cpp:83)
 * and their float twins.  The reference resolves these from a system
 * BLAS/LAPACK found by pkg-config with NO pinned version (CMakeLists.txt:87-109);
 * the source is not under /root/reference.  What follows restates the published
 * reference-LAPACK 3.x algorithms (Gaussian elimination with partial pivoting
 * on a tridiagonal matrix, row interchanges recorded in ipiv, second
 * super-diagonal du2 holding the fill-in) so that oracle/_ref links without an
 * external LAPACK.  tests/test_oracle_cpu.py checks these against SciPy's
 * bundled LAPACK (scipy.linalg.lapack.dgtsv/dgttrf/dgttrs) to 1e-15.
 *
 * Fortran calling convention (all arguments by pointer, trailing underscore).
 */
#include <math.h>

#define GT_IMPL(T, P)                                                            \
void P##gttrf_(int* pn, T* dl, T* d, T* du, T* du2, int* ipiv, int* info)        \
{                                                                                \
    int n = *pn;                                                                 \
    *info = 0;                                                                   \
    if (n < 0) { *info = -1; return; }                                           \
    if (n == 0) return;                                                          \
    for (int i = 0; i < n; i++) ipiv[i] = i + 1;                                 \
    for (int i = 0; i < n - 2; i++) du2[i] = 0;                                  \
    for (int i = 0; i < n - 2; i++) {                                            \
        if (fabs((double)d[i]) >= fabs((double)dl[i])) {                         \
            /* no interchange */                                                 \
            if (d[i] != 0) {                                                     \
                T fact = dl[i] / d[i];                                           \
                dl[i] = fact;                                                    \
                d[i + 1] = d[i + 1] - fact * du[i];                              \
            }                                                                    \
        } else {                                                                 \
            /* interchange rows i and i+1 */                                     \
            T fact = d[i] / dl[i];                                               \
            d[i] = dl[i];                                                        \
            dl[i] = fact;                                                        \
            T temp = du[i];                                                      \
            du[i] = d[i + 1];                                                    \
            d[i + 1] = temp - fact * d[i + 1];                                   \
            du2[i] = du[i + 1];                                                  \
            du[i + 1] = -fact * du[i + 1];                                       \
            ipiv[i] = i + 2;                                                     \
        }                                                                        \
    }                                                                            \
    if (n > 1) {                                                                 \
        int i = n - 2;                                                           \
        if (fabs((double)d[i]) >= fabs((double)dl[i])) {                         \
            if (d[i] != 0) {                                                     \
                T fact = dl[i] / d[i];                                           \
                dl[i] = fact;                                                    \
                d[i + 1] = d[i + 1] - fact * du[i];                              \
            }                                                                    \
        } else {                                                                 \
            T fact = d[i] / dl[i];                                               \
            d[i] = dl[i];                                                        \
            dl[i] = fact;                                                        \
            T temp = du[i];                                                      \
            du[i] = d[i + 1];                                                    \
            d[i + 1] = temp - fact * d[i + 1];                                   \
            ipiv[i] = i + 2;                                                     \
        }                                                                        \
    }                                                                            \
    for (int i = 0; i < n; i++) {                                                \
        if (d[i] == 0) { *info = i + 1; return; }                                \
    }                                                                            \
}                                                                                \
                                                                                 \
void P##gttrs_(const char* trans, int* pn, int* pnrhs, T* dl, T* d, T* du,       \
               T* du2, int* ipiv, T* b, int* pldb, int* info)                    \
{                                                                                \
    int n = *pn, nrhs = *pnrhs, ldb = *pldb;                                     \
    *info = 0;                                                                   \
    if (n == 0 || nrhs == 0) return;                                             \
    int notran = (trans[0] == 'N' || trans[0] == 'n');                           \
    for (int j = 0; j < nrhs; j++) {                                             \
        T* x = b + (long)j * ldb;                                                \
        if (notran) {                                                            \
            /* solve L*x = b */                                                  \
            for (int i = 0; i < n - 1; i++) {                                    \
                if (ipiv[i] == i + 1) {                                          \
                    x[i + 1] = x[i + 1] - dl[i] * x[i];                          \
                } else {                                                         \
                    T temp = x[i];                                               \
                    x[i] = x[i + 1];                                             \
                    x[i + 1] = temp - dl[i] * x[i];                              \
                }                                                                \
            }                                                                    \
            /* solve U*x = b */                                                  \
            x[n - 1] = x[n - 1] / d[n - 1];                                      \
            if (n > 1)                                                           \
                x[n - 2] = (x[n - 2] - du[n - 2] * x[n - 1]) / d[n - 2];         \
            for (int i = n - 3; i >= 0; i--)                                     \
                x[i] = (x[i] - du[i] * x[i + 1] - du2[i] * x[i + 2]) / d[i];     \
        } else {                                                                 \
            /* solve U**T * x = b */                                             \
            x[0] = x[0] / d[0];                                                  \
            if (n > 1) x[1] = (x[1] - du[0] * x[0]) / d[1];                      \
            for (int i = 2; i < n; i++)                                          \
                x[i] = (x[i] - du[i - 1] * x[i - 1] - du2[i - 2] * x[i - 2]) / d[i]; \
            /* solve L**T * x = b */                                             \
            for (int i = n - 2; i >= 0; i--) {                                   \
                if (ipiv[i] == i + 1) {                                          \
                    x[i] = x[i] - dl[i] * x[i + 1];                              \
                } else {                                                         \
                    T temp = x[i + 1];                                           \
                    x[i + 1] = x[i] - dl[i] * temp;                              \
                    x[i] = temp;                                                 \
                }                                                                \
            }                                                                    \
        }                                                                        \
    }                                                                            \
}                                                                                \
                                                                                 \
void P##gtsv_(int* pn, int* pnrhs, T* dl, T* d, T* du, T* b, int* pldb, int* info) \
{                                                                                \
    int n = *pn, nrhs = *pnrhs, ldb = *pldb;                                     \
    *info = 0;                                                                   \
    if (n == 0) return;                                                          \
    for (int i = 0; i < n - 2; i++) {                                            \
        if (fabs((double)d[i]) >= fabs((double)dl[i])) {                         \
            if (d[i] != 0) {                                                     \
                T fact = dl[i] / d[i];                                           \
                d[i + 1] = d[i + 1] - fact * du[i];                              \
                for (int j = 0; j < nrhs; j++)                                   \
                    b[i + 1 + (long)j * ldb] -= fact * b[i + (long)j * ldb];     \
            } else { *info = i + 1; return; }                                    \
            dl[i] = 0;                                                           \
        } else {                                                                 \
            T fact = d[i] / dl[i];                                               \
            d[i] = dl[i];                                                        \
            T temp = d[i + 1];                                                   \
            d[i + 1] = du[i] - fact * temp;                                      \
            dl[i] = du[i + 1];                                                   \
            du[i + 1] = -fact * dl[i];                                           \
            du[i] = temp;                                                        \
            for (int j = 0; j < nrhs; j++) {                                     \
                T t2 = b[i + (long)j * ldb];                                     \
                b[i + (long)j * ldb] = b[i + 1 + (long)j * ldb];                 \
                b[i + 1 + (long)j * ldb] = t2 - fact * b[i + 1 + (long)j * ldb]; \
            }                                                                    \
        }                                                                        \
    }                                                                            \
    if (n > 1) {                                                                 \
        int i = n - 2;                                                           \
        if (fabs((double)d[i]) >= fabs((double)dl[i])) {                         \
            if (d[i] != 0) {                                                     \
                T fact = dl[i] / d[i];                                           \
                d[i + 1] = d[i + 1] - fact * du[i];                              \
                for (int j = 0; j < nrhs; j++)                                   \
                    b[i + 1 + (long)j * ldb] -= fact * b[i + (long)j * ldb];     \
            } else { *info = i + 1; return; }                                    \
        } else {                                                                 \
            T fact = d[i] / dl[i];                                               \
            d[i] = dl[i];                                                        \
            T temp = d[i + 1];                                                   \
            d[i + 1] = du[i] - fact * temp;                                      \
            du[i] = temp;                                                        \
            for (int j = 0; j < nrhs; j++) {                                     \
                T t2 = b[i + (long)j * ldb];                                     \
                b[i + (long)j * ldb] = b[i + 1 + (long)j * ldb];                 \
                b[i + 1 + (long)j * ldb] = t2 - fact * b[i + 1 + (long)j * ldb]; \
            }                                                                    \
        }                                                                        \
    }                                                                            \
    if (d[n - 1] == 0) { *info = n; return; }                                    \
    /* back substitution with U (dl now holds the second super-diagonal) */      \
    for (int j = 0; j < nrhs; j++) {                                             \
        T* x = b + (long)j * ldb;                                                \
        x[n - 1] = x[n - 1] / d[n - 1];                                          \
        if (n > 1) x[n - 2] = (x[n - 2] - du[n - 2] * x[n - 1]) / d[n - 2];      \
        for (int i = n - 3; i >= 0; i--)                                         \
            x[i] = (x[i] - du[i] * x[i + 1] - dl[i] * x[i + 2]) / d[i];          \
    }                                                                            \
}

GT_IMPL(double, d)
GT_IMPL(float, s)

/* cblas symbols referenced by src/tensor.h (norm2) -- not on the path, but
 * the reference headers need them at link time. */
double cblas_dnrm2(int n, const double* x, int incx)
{
    double s = 0;
    for (int i = 0; i < n; i++) s += x[(long)i * incx] * x[(long)i * incx];
    return sqrt(s);
}
float cblas_snrm2(int n, const float* x, int incx)
{
    double s = 0;
    for (int i = 0; i < n; i++) s += (double)x[(long)i * incx] * x[(long)i * incx];
    return (float)sqrt(s);
}
