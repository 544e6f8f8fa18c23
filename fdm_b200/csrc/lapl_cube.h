// Internal definition of the LaplCube handle (shared with ns_cube.cu).
#pragma once
#include "common.h"
#include "xform_kernels.cuh"
#include "xform_pipe.cuh"

namespace fdmb {
cudaError_t launch_rows(int N, int kind, const RowsArgs& a, cudaStream_t st, const char* tag);
cudaError_t launch_cols(int N, int kind, const ColsArgs& a, cudaStream_t st, const char* tag);
cudaError_t launch_cols_cube_divide(int N, bool periodic, const ColsArgs& a, const MidCubeDivide& mid,
                                    cudaStream_t st, const char* tag);
// persistent TMA-fed variants (transform lengths >= 32)
inline bool pipe_supported_N(int N) { return supported_N(N) && N >= 32; }
inline int pipe_B(int N) { return N <= 512 ? 16 : 8; }
cudaError_t launch_rows_pipe(int N, int kind, const RowsPipeArgs& a, cudaStream_t st, const char* tag);
cudaError_t launch_cols_pipe(int N, int kind, const CUtensorMap& tm, const ColsPipeArgs& a, cudaStream_t st, const char* tag);
cudaError_t launch_cols_pipe_cube_divide(int N, bool periodic, const CUtensorMap& tm, const ColsPipeArgs& a,
                                         const MidCubeDivide& mid, cudaStream_t st, const char* tag);
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
    bool pipe_y = false, pipe_z = false;         // tensor maps over d_work are valid
    CUtensorMap tm_y{}, tm_z{};
    int boxrows_y = 0, nchunk_y = 0, boxrows_z = 0, nchunk_z = 0;

    int init();
    int solve_device(double* d_out, const double* d_in, cudaStream_t st);
    int solve_host(double* ans, const double* rhs);
    ~fdmb_lapl_cube();
};
