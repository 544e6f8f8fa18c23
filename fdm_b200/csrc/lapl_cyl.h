// Internal definition of the LaplCyl3FFT2 handle (shared with ns_cyl.cu).
#pragma once
#include "lapl_cube.h"

namespace fdmb {
struct TridiagArgs {
    double* data;             // [nsys rows][pitch], solved in place
    long long pitch;
    int nr;
    int nmid;                 // systems are (outer, mid) pairs: row = outer * nmid + mid
    long long nsys;
    const double* lm_outer;   // eigenvalue by outer index (phi mode), scaled by ir2[j]
    const double* lm_mid;     // eigenvalue by mid index (z mode), offset by mid0
    int mid0;
    int swap;                 // 1: rows are (z slot, phi mode) pairs instead: lm_outer is read with `mid`, lm_mid with
                              //    `outer + mid0` (the pencil layout of the multi-GPU solve)
    double c0;                // -2/dr^2
    const double* L;          // L[j], U[j], ir2[j], j = 1..nr
    const double* U;
    const double* ir2;
    double* piv;              // optional reciprocal pivots, filled once by launch_tridiag_pivots; takes the divisions out of
                              // the per-solve recurrence.  Layout (tridiag_piv_rowmajor(nr)): row-major [nsys][nr] like the
                              // data when a tile of data AND pivots fits in shared memory (the recurrence then runs out of
                              // shared memory only), else j-major [nr][nsys] (read coalesced over systems on the fly)
};
cudaError_t launch_tridiag_rows(const TridiagArgs& a, cudaStream_t st, const char* tag);
cudaError_t launch_tridiag_pivots(const TridiagArgs& a, cudaStream_t st);
cudaError_t prepare_tridiag_rows(int nr);
size_t tridiag_rows_smem(int nr);
bool tridiag_piv_rowmajor(int nr);
}  // namespace fdmb

struct fdmb_lapl_cyl {
    double dr, dz, r0, lr, lz;
    int nr, nz, nphi, zperiodic;
    double dphi = 0, slz = 0;
    int Nz = 0;                        // z transform length (nz+1 Dirichlet, nz periodic)
    int pr = 0;                        // r pitch of the work array (doubles)
    fdmb::Tables tphi{}, tz{};
    cudaStream_t stream = nullptr;
    double *d_lmphi = nullptr, *d_lmz = nullptr;   // eigenvalues (lapl_cyl.cpp:132-141)
    double *d_L = nullptr, *d_U = nullptr, *d_ir2 = nullptr;   // r-dependent matrix entries (lapl_cyl.cpp:151-159)
    double* d_work = nullptr;
    double* d_piv = nullptr;           // reciprocal pivots of every (phi mode, z mode) system, 8 B per grid point
    double *d_rhs = nullptr, *d_ans = nullptr;
    bool pipe_z = false, pipe_phi = false;
    fdmb::ColsMaps tm_z{}, tm_phi{}, tm_in{};
    const void* tm_in_ptr = nullptr;

    // ---- phi-slab sharding over `nranks` GPUs.  Rank r owns phi in [r*Sphi, (r+1)*Sphi) of the caller's
    // arrays and, between the two transposes, the z slots [r*Sz, (r+1)*Sz) for ALL phi (pencil buffer T).  The sweep
    // order is z forward -> (transpose) -> phi forward, r tridiagonals, phi inverse -> (transpose) -> z inverse; the
    // axis transforms commute, so the result differs from the reference's phi,z,r,z,phi order by round-off only.
    int rank = 0, nranks = 1, device = 0;
    int Sphi = 0, Sz = 0, phi_first = 0, z_first = 0, nzl = 0;
    void* mg_block = nullptr;                 // [slab A | pencils T | flags]
    size_t off_T = 0, off_flags = 0, mg_bytes = 0;
    double *d_A = nullptr, *d_T = nullptr;
    void* peer_block[fdmb::FDMB_MAX_RANKS] = {};
    bool peer_ipc[fdmb::FDMB_MAX_RANKS] = {};
    bool attached = false;
    fdmb::ColsMaps tm_tphi{}, tm_tphi_l{}, tm_za{};   // phi sweeps over the local pencils (wide tiles for the transposing
                                                      // one, normal tiles for the local one), z inverse over the slab
    int init_sharded();
    int solve_device_sharded(double* d_out, const double* d_in, cudaStream_t st);
    int barrier(cudaStream_t st);

    int init();
    int solve_device(double* d_out, const double* d_in, cudaStream_t st);
    int solve_host(double* ans, const double* rhs);
    ~fdmb_lapl_cyl();
};
