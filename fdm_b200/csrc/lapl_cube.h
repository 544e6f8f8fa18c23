// Internal definition of the LaplCube handle (shared with ns_cube.cu).
#pragma once
#include <map>
#include <utility>
#include "common.h"
#include "xform_kernels.cuh"
#include "xform_pipe.cuh"
#include "xform_ring.cuh"

namespace fdmb {
cudaError_t launch_rows(int N, int kind, const RowsArgs& a, cudaStream_t st, const char* tag);
cudaError_t launch_cols(int N, int kind, const ColsArgs& a, cudaStream_t st, const char* tag);
cudaError_t launch_cols_cube_divide(int N, bool periodic, const ColsArgs& a, const MidCubeDivide& mid,
                                    cudaStream_t st, const char* tag);
// persistent TMA-fed variants (transform lengths >= 32)
inline bool pipe_supported_N(int N) { return supported_N(N) && N >= 32; }
// the contiguous-axis sweep keeps a staging buffer AND a compute tile of 8 rows: N = 2048 (260 KB) does not fit one SM
inline bool rows_pipe_supported_N(int N) { return pipe_supported_N(N) && N <= 1024; }
inline int pipe_B(int N) { return N <= 512 ? 16 : 8; }
bool ring_enabled();     // FDMB_RING=0 selects the one-tile-per-CTA sweeps at N = 1024 (A/B measurements)
// cross-GPU barrier on a stream over the flag arrays at `off_flags` inside every rank's peer-mapped block
// (FDMB_MAX_RANKS u64 arrival epochs + this rank's own barrier count, zero-initialised; every rank must run the same
// sequence of barriers)
int launch_mg_barrier(void* const* peer_blocks, size_t off_flags, int rank, int nranks, cudaStream_t st);
int preload_mg_barrier();
inline int pipe_B_sharded(int N) { return N <= 1024 ? 16 : 8; }     // PipeCfg<N, true>::B
cudaError_t launch_rows_pipe(int N, int kind, const RowsPipeArgs& a, cudaStream_t st, const char* tag);
cudaError_t launch_cols_pipe(int N, int kind, const ColsMaps& tm, ColsPipeArgs a, cudaStream_t st, const char* tag);
cudaError_t launch_cols_pipe_cube_divide(int N, bool periodic, const ColsMaps& tm, ColsPipeArgs a,
                                         const MidCubeDivide& mid, cudaStream_t st, const char* tag);
cudaError_t launch_cols_pipe_blocked(int N, const ColsMaps& tm, ColsPipeArgs a, const OutBlocked& ob, cudaStream_t st,
                                     const char* tag);
// multi-GPU variants: the sweep's stores scatter over the peers' buffers (slab <-> pencil transpose)
cudaError_t launch_cols_pipe_shard(int N, int kind, const ColsMaps& tm, ColsPipeArgs a, const OutShard& om,
                                   cudaStream_t st, const char* tag);
cudaError_t launch_cols_pipe_cube_divide_shard(int N, bool periodic, const ColsMaps& tm, ColsPipeArgs a,
                                               const MidCubeDivide& mid, const OutShard& om, cudaStream_t st,
                                               const char* tag);
// DST sweeps whose finished tiles leave as TMA tile stores (OutShardTma)
cudaError_t launch_cols_pipe_shard_tma(int N, const ColsMaps& tm, ColsPipeArgs a, const OutShardTma& om, cudaStream_t st,
                                       const char* tag);
cudaError_t launch_cols_pipe_cube_divide_shard_tma(int N, const ColsMaps& tm, ColsPipeArgs a, const MidCubeDivide& mid,
                                                   const OutShardTma& om, cudaStream_t st, const char* tag);

// Even split of the N index slots of one axis over P ranks (N, P powers of two).  Slot 0 of a
// Dirichlet axis is the implicit zero boundary, so rank 0 owns one interior entry fewer.
// first/count are in 0-based interior entries (what the caller's arrays hold).
inline void slab_range(int n, int periodic, int nranks, int rank, int* first, int* count)
{
    const int N = periodic ? n : n + 1, S = N / nranks, J0 = periodic ? 0 : 1;
    const int lo = rank * S < J0 ? J0 : rank * S;   // first slot owned
    *first = lo - J0;
    *count = (rank + 1) * S - lo;
}
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
    // second staging pair, copy streams and events of the pipelined multi-solve entry point (solve_batch)
    double *d_rhs1 = nullptr, *d_ans1 = nullptr;
    cudaStream_t s_up = nullptr, s_dn = nullptr;
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr}, ev_dn[2] = {nullptr, nullptr};
    int solve_batch(int count, double* const* ans, const double* const* rhs);
    bool pipe_y = false, pipe_z = false;         // tensor maps over d_work are valid
    fdmb::ColsMaps tm_y{}, tm_z{}, tm_yw{};      // tm_yw: wide-tile maps of the sharded y forward sweep
    // y-sweep maps over a chunk [z0, z0 + nzc) of the planes (streamed host solve, x || y overlap of the sharded solve):
    // the same maps with a shifted base, encoded on first use (a sweep kernel's tiles then start at outer index 0)
    std::map<std::pair<int, int>, fdmb::ColsMaps> tm_y_chunk, tm_yw_chunk;
    const fdmb::ColsMaps* y_chunk_maps(bool wide, int z0, int nzc);
    int blog = 0, nyb = 0;                       // blocked work array [yb][z][yi][x], 1 << blog rows per block (0: natural)

    // ---- z-slab sharding over `nranks` GPUs (nranks == 1: everything above is the whole problem) ----
    // Rank r owns the z slots [r*Sz, (r+1)*Sz) of the caller's arrays and, between the two transposes,
    // the y slots [r*Sy, (r+1)*Sy) of the pencil buffer.  d_A ([Sz][ny][px]) and d_T ([nz][Sy][px]) are
    // carved out of one allocation that the peers map (IPC or same-process peer access).
    int rank = 0, nranks = 1;
    int Sy = 0, Sz = 0;                 // slots per rank along y / z
    int z_first = 0, nzl = 0;           // local interior planes: global z' in [z_first, z_first + nzl)
    int y_first = 0, nyl = 0;           // local pencil rows
    void* mg_block = nullptr;           // A | T | flags
    size_t mg_bytes = 0, off_T = 0, off_flags = 0;
    double *d_A = nullptr, *d_T = nullptr;                 // uniform allocations (slot 0 row/plane included)
    unsigned long long* d_flags = nullptr;                 // [FDMB_MAX_RANKS] arrival epochs, written by the peers
    void* peer_block[fdmb::FDMB_MAX_RANKS] = {};           // mapped bases of every rank's mg_block (own included)
    bool peer_ipc[fdmb::FDMB_MAX_RANKS] = {};
    bool attached = false;
    int device = 0;
    // TMA tile-store views of every rank's pencil buffer (y forward sweep) and slab (z sweep), built at attach()
    fdmb::OutShardTma st_y{}, st_z{};
    bool st_ready = false;
    int build_store_maps();
    // overlap of the local x sweep with the NVLink-bound transposing y sweep (solve_device_sharded)
    cudaStream_t s_side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_chunk[16] = {}, ev_done[16] = {};

    int init();
    int init_sharded();
    int attach(void* const* bases);
    int barrier(cudaStream_t st);
    int solve_device(double* d_out, const double* d_in, cudaStream_t st);
    int sweeps(double* d_out, const double* d_in, cudaStream_t st, int phases, int z0, int nzc);
    int solve_device_sharded(double* d_out, const double* d_in, cudaStream_t st);
    int solve_host(double* ans, const double* rhs);
    ~fdmb_lapl_cube();
};
