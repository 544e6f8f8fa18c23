// Internal definition of the LaplCube handle (shared with ns_cube.cu).
#pragma once
#include "common.h"
#include "xform_kernels.cuh"

namespace fdmb {
cudaError_t launch_rows(int N, int kind, const RowsArgs& a, cudaStream_t st, const char* tag);
cudaError_t launch_cols(int N, int kind, const ColsArgs& a, cudaStream_t st, const char* tag);
cudaError_t launch_cols_cube_divide(int N, bool periodic, const ColsArgs& a, const MidCubeDivide& mid,
                                    cudaStream_t st, const char* tag);
}  // namespace fdmb

struct fdmb_lapl_cube {
    double dx, dy, dz, lx, ly, lz;
    int nx, ny, nz, periodic;
    int Nx = 0, Ny = 0, Nz = 0;       // transform lengths
    double slx = 0, sly = 0, slz = 0; // sqrt(2/l), lapl_cube.h:64
    int px = 0;                       // x pitch of the work array (doubles)
    fdmb::Tables tx{}, ty{}, tz{};
    cudaStream_t stream = nullptr;
    double *d_lmx = nullptr, *d_lmy = nullptr, *d_lmz = nullptr;
    double* d_work = nullptr;
    double *d_rhs = nullptr, *d_ans = nullptr;   // staging for the host-pointer entry point

    int init();
    int solve_device(double* d_out, const double* d_in, cudaStream_t st);
    int solve_host(double* ans, const double* rhs);
    ~fdmb_lapl_cube();
};
