// LaplCyl3FFT2 on B200: Poisson equation in cylindrical coordinates (phi, z, r), r Dirichlet,
//   1/r d/dr(r du/dr) + d2u/dz2 + 1/r^2 d2u/dphi2 = f.
// Replaces fdm::LaplCyl3FFT2<double,check,zflag>::solve (reference src/lapl_cyl.cpp:11-128,
// init_solver :131-170, data src/lapl_cyl.h:12-37,172-249).
//
// Sweeps: phi forward (periodic, slowest axis) -> z forward (DST or periodic, middle axis) ->
// batched tridiagonal solves along the contiguous r axis -> z inverse -> phi inverse.
// The tridiagonal matrices are never stored: the r-dependent entries come from three small
// tables and the (phi,z)-mode shift from the eigenvalue tables, the LU recurrence (the no-pivot
// path of LAPACK gttrf/gttrs that the reference calls, src/lapl_cyl.cpp:83,166; the matrices are
// diagonally dominant so LAPACK never pivots) is recomputed on the fly.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.h"
#include "lapl_cyl.h"

namespace fdmb {

// ---- batched tridiagonal solve along the contiguous axis -------------------------------------
// A CTA stages ts systems (rows of nr doubles) in an odd-pitch shared-memory tile with coalesced
// loads; thread s then runs the Thomas recurrence of system s (conflict-free: consecutive lanes sit
// one odd pitch apart), keeping the reciprocal pivots in a second tile; coalesced store.
// ts = TS for systems that fit (nr <= ~430); longer systems get a lower tile (tridiag_tile_rows) so that
// e.g. the reference's 511 x 511 LaplRect case (ut/ut_lapl_rect.cpp:384-455) still runs.
constexpr int TS = 32;

__global__ void __launch_bounds__(128) k_tridiag_rows(TridiagArgs a, int ts)
{
    pdl_wait();
    pdl_trigger();
    extern __shared__ double smem[];
    const int P = a.nr | 1;                    // odd pitch
    double* tb = smem;                         // rhs / solution
    double* ti = smem + ts * P;                // reciprocal pivots
    double* cL = ti + ts * P;                  // coefficient tables, 1-based
    double* cU = cL + (a.nr + 1);
    double* cR = cU + (a.nr + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int j = tid + 1; j <= a.nr; j += blockDim.x) { cL[j] = a.L[j]; cU[j] = a.U[j]; cR[j] = a.ir2[j]; }
    const long long row0 = (long long)blockIdx.x * ts;
    for (int r = warp; r < ts; r += 4) {
        const long long row = row0 + r;
        if (row >= a.nsys) break;
        const double* src = a.data + row * a.pitch;
        for (int x = lane; x < a.nr; x += 32) tb[r * P + x] = src[x];
    }
    __syncthreads();
    if (tid < ts && row0 + tid < a.nsys) {
        const long long row = row0 + tid;
        const int outer = (int)(row / a.nmid), mid = (int)(row % a.nmid);
        const double lo = a.swap ? a.lm_outer[mid] : a.lm_outer[outer];
        const double lmz = a.swap ? a.lm_mid[outer + a.mid0] : a.lm_mid[mid + a.mid0];
        double* b = tb + tid * P - 1;          // 1-based
        double* iv = ti + tid * P - 1;
        // D_j = -2/dr2 - lm_phi/r/r - lm_z   (lapl_cyl.cpp:153)
        double d = a.c0 - lo * cR[1] - lmz;
        double inv = 1.0 / d;
        double bp = b[1];
        iv[1] = inv;
        for (int j = 2; j <= a.nr; j++) {
            const double fact = cL[j] * inv;                    // dl / d
            d = (a.c0 - lo * cR[j] - lmz) - fact * cU[j - 1];   // d_{j} -= fact * du_{j-1}
            bp = b[j] - fact * bp;                              // forward substitution
            inv = 1.0 / d;
            b[j] = bp;
            iv[j] = inv;
        }
        double x = bp * inv;
        b[a.nr] = x;
        for (int j = a.nr - 1; j >= 1; j--) {
            x = (b[j] - cU[j] * x) * iv[j];
            b[j] = x;
        }
    }
    __syncthreads();
    for (int r = warp; r < ts; r += 4) {
        const long long row = row0 + r;
        if (row >= a.nsys) break;
        double* dst = a.data + row * a.pitch;
        for (int x = lane; x < a.nr; x += 32) dst[x] = tb[r * P + x];
    }
}

// ---- the same with the reciprocal pivots precomputed ------------------------------------------------
// The pivots depend on the (phi mode, z mode) pair only, not on the right-hand side.  k_tridiag_pivots runs the
// pivot recurrence once per handle and stores 1/d_j j-major ([nr][nsys]: consecutive systems are contiguous, so the
// thread-per-system recurrence reads them coalesced, no staging).  The per-solve recurrence is then one dependent
// FMA per element forward and an FMA + multiply backward instead of a division chain: 43 -> ~12 us at 128 x 127 x 128.
// The reference keeps its LU factors too (`matrices`, `ipivs`, src/lapl_cyl.h:235-236, 36 B per point; here 8 B).
__global__ void __launch_bounds__(128) k_tridiag_pivots(TridiagArgs a, int rowmajor)
{
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.nsys) return;
    const int outer = (int)(row / a.nmid), mid = (int)(row % a.nmid);
    const double lo = a.swap ? a.lm_outer[mid] : a.lm_outer[outer];
    const double lmz = a.swap ? a.lm_mid[outer + a.mid0] : a.lm_mid[mid + a.mid0];
    double d = a.c0 - lo * a.ir2[1] - lmz;
    double inv = 1.0 / d;
    a.piv[rowmajor ? row * a.nr : row] = inv;
    for (int j = 2; j <= a.nr; j++) {
        const double fact = a.L[j] * inv;
        d = (a.c0 - lo * a.ir2[j] - lmz) - fact * a.U[j - 1];
        inv = 1.0 / d;
        a.piv[rowmajor ? row * a.nr + (j - 1) : (long long)(j - 1) * a.nsys + row] = inv;
    }
}

// ---- pivot table staged in shared memory ---------------------------------------------------------------------
// ncu (profiles/r02m_full_nscyl128.md): k_tridiag_rows_piv spends its time waiting -- 16 round trips to L2 for the
// pivots of each system, issued by a quarter of the CTA while the rest sits on the barrier (54 % of the stall samples).
// Here a tile of TSS systems AND its pivots (row-major table, same shape as the data) are fetched with coalesced loads
// by the whole CTA, the recurrence of a system then touches shared memory only (one dependent FMA per element forward,
// FMA + multiply backward), and the solution leaves with coalesced stores.
constexpr int TSS = 32;

__global__ void __launch_bounds__(128) k_tridiag_rows_pivs(TridiagArgs a)
{
    pdl_wait();
    pdl_trigger();
    extern __shared__ double smem[];
    const int P = a.nr | 1;                    // odd pitch: lane s walks row s without bank conflicts
    double* tb = smem;                         // rhs / solution
    double* tp = tb + TSS * P;                 // reciprocal pivots
    double* cL = tp + TSS * P;                 // coefficient tables, 1-based
    double* cU = cL + (a.nr + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    for (int j = tid + 1; j <= a.nr; j += blockDim.x) { cL[j] = a.L[j]; cU[j] = a.U[j]; }
    const long long row0 = (long long)blockIdx.x * TSS;
    for (int r = warp; r < TSS; r += nwarp) {
        const long long row = row0 + r;
        if (row >= a.nsys) break;
        const double* src = a.data + row * a.pitch;
        const double* pv = a.piv + row * a.nr;
#pragma unroll 4
        for (int x = lane; x < a.nr; x += 32) { tb[r * P + x] = src[x]; tp[r * P + x] = pv[x]; }
    }
    __syncthreads();
    if (tid < TSS && row0 + tid < a.nsys) {
        double* b = tb + tid * P - 1;          // 1-based
        const double* iv = tp + tid * P - 1;   // iv[j] = 1 / d_j
        double bp = b[1];
#pragma unroll 4
        for (int j = 2; j <= a.nr; j++) {
            const double fact = cL[j] * iv[j - 1];              // dl_j / d_{j-1}: off the dependent chain
            bp = b[j] - fact * bp;
            b[j] = bp;
        }
        double x = bp * iv[a.nr];
        b[a.nr] = x;
#pragma unroll 4
        for (int j = a.nr - 1; j >= 1; j--) {
            x = (b[j] - cU[j] * x) * iv[j];
            b[j] = x;
        }
    }
    __syncthreads();
    for (int r = warp; r < TSS; r += nwarp) {
        const long long row = row0 + r;
        if (row >= a.nsys) break;
        double* dst = a.data + row * a.pitch;
#pragma unroll 4
        for (int x = lane; x < a.nr; x += 32) dst[x] = tb[r * P + x];
    }
}

constexpr int TSP = 64;       // systems per CTA of the pivot-table variant

__global__ void __launch_bounds__(256) k_tridiag_rows_piv(TridiagArgs a, int ts)
{
    pdl_wait();
    pdl_trigger();
    extern __shared__ double smem[];
    const int P = a.nr | 1;                    // odd pitch
    double* tb = smem;                         // rhs / solution
    double* cL = tb + ts * P;                 // coefficient tables, 1-based
    double* cU = cL + (a.nr + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    for (int j = tid + 1; j <= a.nr; j += blockDim.x) { cL[j] = a.L[j]; cU[j] = a.U[j]; }
    const long long row0 = (long long)blockIdx.x * ts;
    for (int r = warp; r < ts; r += nwarp) {
        const long long row = row0 + r;
        if (row >= a.nsys) break;
        const double* src = a.data + row * a.pitch;
        for (int x = lane; x < a.nr; x += 32) tb[r * P + x] = src[x];
    }
    __syncthreads();
    if (tid < ts && row0 + tid < a.nsys) {
        const double* __restrict__ iv = a.piv + row0 + tid;      // iv[(j-1) * nsys] = 1 / d_j
        double* b = tb + tid * P - 1;                            // 1-based
        constexpr int CH = 16;                                   // pivots fetched CH at a time: CH global loads in flight
        double bp = b[1];
        double ip = iv[0];
        int j = 2;
        for (; j + CH - 1 <= a.nr; j += CH) {
            double pv[CH], bb[CH];
#pragma unroll
            for (int c = 0; c < CH; c++) pv[c] = iv[(long long)(j - 1 + c) * a.nsys];
#pragma unroll
            for (int c = 0; c < CH; c++) bb[c] = b[j + c];
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const double fact = cL[j + c] * ip;              // dl_j / d_{j-1}: off the dependent chain
                ip = pv[c];
                bp = bb[c] - fact * bp;                          // forward substitution
                b[j + c] = bp;
            }
        }
        for (; j <= a.nr; j++) {
            const double fact = cL[j] * ip;
            ip = iv[(long long)(j - 1) * a.nsys];
            bp = b[j] - fact * bp;
            b[j] = bp;
        }
        double x = bp * ip;
        b[a.nr] = x;
        j = a.nr - 1;
        for (; j - CH + 1 >= 1; j -= CH) {
            double pv[CH], bb[CH];
#pragma unroll
            for (int c = 0; c < CH; c++) pv[c] = iv[(long long)(j - 1 - c) * a.nsys];
#pragma unroll
            for (int c = 0; c < CH; c++) bb[c] = b[j - c];
#pragma unroll
            for (int c = 0; c < CH; c++) {
                x = (bb[c] - cU[j - c] * x) * pv[c];
                b[j - c] = x;
            }
        }
        for (; j >= 1; j--) {
            x = (b[j] - cU[j] * x) * iv[(long long)(j - 1) * a.nsys];
            b[j] = x;
        }
    }
    __syncthreads();
    for (int r = warp; r < ts; r += nwarp) {
        const long long row = row0 + r;
        if (row >= a.nsys) break;
        double* dst = a.data + row * a.pitch;
        for (int x = lane; x < a.nr; x += 32) dst[x] = tb[r * P + x];
    }
}

constexpr size_t TRIDIAG_SMEM_MAX = 220 * 1024;
static size_t rows_smem(int nr, int ts) { return sizeof(double) * (size_t)(2 * ts * (nr | 1) + 3 * (nr + 1)); }
static size_t rows_piv_smem(int nr, int ts) { return sizeof(double) * (size_t)(ts * (nr | 1) + 2 * (nr + 1)); }
static size_t rows_pivs_smem(int nr) { return sizeof(double) * (size_t)(2 * TSS * (nr | 1) + 2 * (nr + 1)); }
// row-major pivots + the shared-memory staged recurrence while three CTAs still fit on an SM (nr <= ~140)
static int tridiag_variant()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("FDMB_TRIDIAG"); v = e ? atoi(e) : 0; }
    return v;
}
bool tridiag_piv_rowmajor(int nr) { return tridiag_variant() >= 2 && rows_pivs_smem(nr) <= 74 * 1024; }

// systems per CTA: the full tile (TS / TSP) when it fits in shared memory, else halved down to 4
static int tridiag_tile_rows(int nr, bool piv)
{
    int ts = piv ? TSP : TS;
    while (ts > 4 && (piv ? rows_piv_smem(nr, ts) : rows_smem(nr, ts)) > TRIDIAG_SMEM_MAX) ts >>= 1;
    return ts;
}

// shared memory of the on-the-fly variant for systems of length nr; callers reject nr when it exceeds 220 KB
// (nr > ~2500: not even four systems fit)
size_t tridiag_rows_smem(int nr) { return rows_smem(nr, tridiag_tile_rows(nr, false)); }

// raises both kernels' dynamic shared-memory limit on the current device (this also loads the kernels)
cudaError_t prepare_tridiag_rows(int nr)
{
    (void)nr;
    static bool done_dev[64] = {false};       // function attributes are per device
    bool& done = done_dev[current_device_slot()];
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(k_tridiag_rows, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)TRIDIAG_SMEM_MAX);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_tridiag_rows_piv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIDIAG_SMEM_MAX);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_tridiag_rows_pivs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIDIAG_SMEM_MAX);
        if (e != cudaSuccess) return e;
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, k_tridiag_pivots);
        if (e != cudaSuccess) return e;
        done = true;
    }
    return cudaSuccess;
}

cudaError_t launch_tridiag_pivots(const TridiagArgs& a, cudaStream_t st)
{
    LaunchScope scope("tridiag_pivots", st);
    k_tridiag_pivots<<<(unsigned)((a.nsys + 127) / 128), 128, 0, st>>>(a, tridiag_piv_rowmajor(a.nr) ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_tridiag_rows(const TridiagArgs& a, cudaStream_t st, const char* tag)
{
    LaunchScope scope(tag, st);
    cudaError_t e = prepare_tridiag_rows(a.nr);
    if (e != cudaSuccess) return e;
    if (a.piv && tridiag_piv_rowmajor(a.nr)) {
        return launch_pdl(k_tridiag_rows_pivs, dim3((unsigned)((a.nsys + TSS - 1) / TSS)), dim3(tridiag_variant() == 3 ? 32 : 128), rows_pivs_smem(a.nr), st, a);
    }
    if (a.piv) {
        const int ts = tridiag_tile_rows(a.nr, true);
        const int nt = (tridiag_variant() == 1 && ts >= 32) ? ts : 256;      // 1: every thread of the CTA owns a system
        return launch_pdl(k_tridiag_rows_piv, dim3((unsigned)((a.nsys + ts - 1) / ts)), dim3(nt), rows_piv_smem(a.nr, ts), st, a, ts);
    }
    const int ts = tridiag_tile_rows(a.nr, false);
    return launch_pdl(k_tridiag_rows, dim3((unsigned)((a.nsys + ts - 1) / ts)), dim3(128), rows_smem(a.nr, ts), st, a, ts);
}

}  // namespace fdmb

using namespace fdmb;

static inline double sq(double x) { return x * x; }

int fdmb_lapl_cyl::init()
{
    Nz = zperiodic ? nz : nz + 1;
    if (nr < 2 || nz < 1 || nphi < 1 || !supported_N(nphi) || !supported_N(Nz)) {
        set_error("LaplCyl3FFT2: nphi=%d and the z transform length %d (nz%s) must be powers of two in [4,2048], nr >= 2 "
                  "(reference: verify((1<<n) == N), src/fft.cpp:67)", nphi, Nz, zperiodic ? "" : "+1");
        return FDMB_ERR_INVALID;
    }
    if (tridiag_rows_smem(nr) > 220 * 1024) {
        set_error("LaplCyl3FFT2: nr=%d exceeds the shared-memory tile of the tridiagonal kernel", nr);
        return FDMB_ERR_INVALID;
    }
    dphi = 2 * M_PI / nphi;
    slz = std::sqrt(2. / lz);
    pr = (nr + 15) / 16 * 16;
    int rc;
    if ((rc = get_tables(nphi, &tphi)) || (rc = get_tables(Nz, &tz))) return rc;
    const double dphi2 = dphi * dphi, dz2 = dz * dz, dr2 = dr * dr;
    std::vector<double> lm_phi(nphi), lm_z(Nz), L(nr + 1, 0.0), U(nr + 1, 0.0), ir2(nr + 1, 0.0);
    for (int i = 0; i < nphi; i++) lm_phi[i] = 4.0 / dphi2 * sq(sin(i * dphi * 0.5));
    for (int k = 0; k < Nz; k++)
        lm_z[k] = zperiodic ? 4. / dz2 * sq(sin(k * M_PI / Nz)) : 4. / dz2 * sq(sin(k * M_PI * 0.5 / Nz));
    for (int j = 1; j <= nr; j++) {
        const double r = r0 + j * dr;
        L[j] = (r - 0.5 * dr) / dr2 / r;
        U[j] = (r + 0.5 * dr) / dr2 / r;
        ir2[j] = 1.0 / r / r;
    }
    FDMB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    auto up = [&](double** d, const std::vector<double>& h) -> int {
        FDMB_CUDA(cudaMalloc(d, sizeof(double) * h.size()));
        FDMB_CUDA(cudaMemcpy(*d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
        return FDMB_OK;
    };
    if ((rc = up(&d_lmphi, lm_phi)) || (rc = up(&d_lmz, lm_z)) || (rc = up(&d_L, L)) || (rc = up(&d_U, U)) ||
        (rc = up(&d_ir2, ir2)))
        return rc;
    FDMB_CUDA(cudaGetDevice(&device));
    if (nranks > 1) return init_sharded();
    FDMB_CUDA(cudaMalloc(&d_work, sizeof(double) * (size_t)nphi * nz * pr));
    {
        FDMB_CUDA(prepare_tridiag_rows(nr));
        FDMB_CUDA(cudaMalloc(&d_piv, sizeof(double) * (size_t)nphi * nz * nr));
        TridiagArgs t{};
        t.nr = nr; t.nmid = nz; t.nsys = (long long)nphi * nz; t.lm_outer = d_lmphi; t.lm_mid = d_lmz;
        t.mid0 = zperiodic ? 0 : 1; t.c0 = -2 / (dr * dr); t.L = d_L; t.U = d_U; t.ir2 = d_ir2; t.piv = d_piv;
        FDMB_CUDA(launch_tridiag_pivots(t, stream));
        FDMB_CUDA(cudaStreamSynchronize(stream));
    }
    if (pipe_enabled()) {
        const unsigned long long s1 = 8ull * pr, s2 = 8ull * (unsigned long long)nz * pr;
        if (pipe_supported_N(Nz)) {
            if ((rc = make_cols_maps(&tm_z, d_work, Nz, 1, nr, nz, nphi, s1, s2, pipe_B(Nz)))) return rc;
            pipe_z = true;
        }
        if (pipe_supported_N(nphi)) {
            if ((rc = make_cols_maps(&tm_phi, d_work, nphi, 2, nr, nz, nphi, s1, s2, pipe_B(nphi)))) return rc;
            pipe_phi = true;
        }
    }
    return FDMB_OK;
}

static int ilog2i(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

int fdmb_lapl_cyl::init_sharded()
{
    const int J0 = zperiodic ? 0 : 1;
    if (nranks != 2 && nranks != 4 && nranks != 8) {
        set_error("LaplCyl3FFT2: nranks must be 1, 2, 4 or 8 (got %d)", nranks);
        return FDMB_ERR_INVALID;
    }
    if (rank < 0 || rank >= nranks) { set_error("LaplCyl3FFT2: rank %d out of range", rank); return FDMB_ERR_INVALID; }
    if (!pipe_supported_N(nphi) || !pipe_supported_N(Nz) || nphi / nranks < 2 || Nz / nranks < 2 || (nr & 1)) {
        set_error("LaplCyl3FFT2: the sharded solve needs nphi and the z transform length >= 32 and >= 2*nranks, and an "
                  "even nr (got nphi=%d, Nz=%d, nr=%d)", nphi, Nz, nr);
        return FDMB_ERR_INVALID;
    }
    int rc;
    if ((rc = preload_mg_barrier())) return rc;
    FDMB_CUDA(prepare_tridiag_rows(nr));
    Sphi = nphi / nranks; Sz = Nz / nranks;
    phi_first = rank * Sphi;
    slab_range(nz, zperiodic, nranks, rank, &z_first, &nzl);
    const size_t a_bytes = (sizeof(double) * (size_t)Sphi * nz * pr + 255) & ~(size_t)255;
    const size_t t_bytes = (sizeof(double) * (size_t)Sz * nphi * pr + 255) & ~(size_t)255;
    off_T = a_bytes; off_flags = a_bytes + t_bytes;
    mg_bytes = off_flags + 256;
    FDMB_CUDA(cudaMalloc(&mg_block, mg_bytes));
    FDMB_CUDA(cudaMemset(mg_block, 0, mg_bytes));
    // cudaMemset on device memory is asynchronous to the host and runs on the legacy default stream, which the
    // handle's non-blocking streams do not order against: finish it before the handle is handed out
    FDMB_CUDA(cudaDeviceSynchronize());
    d_A = reinterpret_cast<double*>(mg_block);
    d_T = reinterpret_cast<double*>(static_cast<char*>(mg_block) + off_T);
    peer_block[rank] = mg_block;
    d_work = d_A;
    // local pencils: the uniform allocation keeps the slot-0 row block on rank 0 of a Dirichlet z axis
    double* t_loc = d_T + (size_t)(z_first + J0 - rank * Sz) * nphi * pr;
    if ((rc = make_cols_maps(&tm_tphi, t_loc, nphi, 1, nr, nphi, nzl, 8ull * pr, 8ull * (unsigned long long)nphi * pr,
                             pipe_B_sharded(nphi))))
        return rc;
    if ((rc = make_cols_maps(&tm_tphi_l, t_loc, nphi, 1, nr, nphi, nzl, 8ull * pr, 8ull * (unsigned long long)nphi * pr,
                             pipe_B(nphi))))
        return rc;
    if ((rc = make_cols_maps(&tm_za, d_A, Nz, 1, nr, nz, Sphi, 8ull * pr, 8ull * (unsigned long long)nz * pr, pipe_B(Nz))))
        return rc;
    pipe_z = pipe_phi = true;
    {   // reciprocal pivots of this rank's (z slot, phi mode) systems
        FDMB_CUDA(cudaMalloc(&d_piv, sizeof(double) * (size_t)nzl * nphi * nr));
        TridiagArgs t{};
        t.nr = nr; t.nmid = nphi; t.nsys = (long long)nzl * nphi; t.swap = 1; t.lm_outer = d_lmphi; t.lm_mid = d_lmz;
        t.mid0 = z_first + J0; t.c0 = -2 / (dr * dr); t.L = d_L; t.U = d_U; t.ir2 = d_ir2; t.piv = d_piv;
        FDMB_CUDA(launch_tridiag_pivots(t, stream));
        FDMB_CUDA(cudaStreamSynchronize(stream));
    }
    // load every kernel of the sharded solve on this device now (preload_only(), xform_pipe.cuh)
    preload_only() = true;
    attached = true;
    rc = solve_device_sharded(d_A, d_A, stream);
    attached = false;
    preload_only() = false;
    tm_in_ptr = nullptr;
    return rc;
}

int fdmb_lapl_cyl::barrier(cudaStream_t st)
{
    if (preload_only()) return FDMB_OK;
    return launch_mg_barrier(peer_block, off_flags, rank, nranks, st);
}

// d_in / d_out: this rank's phi-slab [Sphi][nz][nr] of the caller's arrays
int fdmb_lapl_cyl::solve_device_sharded(double* d_out, const double* d_in, cudaStream_t st)
{
    if (!attached) { set_error("LaplCyl3FFT2: sharded handle used before attach_ipc/attach_local"); return FDMB_ERR_COMM; }
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) != 0) {
        set_error("LaplCyl3FFT2: the sharded solve needs a 16-byte aligned rhs slab");
        return FDMB_ERR_INVALID;
    }
    const double SQRT_M_1_PI = 0.56418958354775629;   // lapl_cyl.h:175
    const int J0 = zperiodic ? 0 : 1;
    const int kzf = zperiodic ? XF_PFWD : XF_DST, kzi = zperiodic ? XF_PINV : XF_DST;
    const long long wplane = (long long)nz * pr, uplane = (long long)nz * nr, tplane = (long long)nphi * pr;
    int rc;
    if (tm_in_ptr != d_in) {
        if ((rc = make_cols_maps(&tm_in, d_in, Nz, 1, nr, nz, Sphi, 8ull * nr, 8ull * uplane, pipe_B_sharded(Nz)))) return rc;
        tm_in_ptr = d_in;
    }
    {   // z forward straight from the caller's slab, transposing into the pencil buffers T_q[z slot & (Sz-1)][phi][r]
        ColsPipeArgs p{};
        p.out = nullptr; p.nvalid = nz; p.nb = nr; p.no = Sphi; p.taxis = 1; p.reverse = 0; p.scale = dz * slz;
        p.SN = tz.SN; p.WM = tz.WM;
        OutShard om{};
        for (int q = 0; q < nranks; q++) om.base[q] = reinterpret_cast<double*>(static_cast<char*>(peer_block[q]) + off_T);
        om.logS = ilog2i(Sz); om.maskS = Sz - 1; om.sj = tplane; om.so = pr; om.o_off = phi_first;
        FDMB_CUDA(launch_cols_pipe_shard(Nz, kzf, tm_in, p, om, st, "cyl_z_fwd_xpose"));
    }
    if ((rc = barrier(st))) return rc;
    double* t_loc = d_T + (size_t)(z_first + J0 - rank * Sz) * nphi * pr;
    {   // phi forward on the local pencils (in place)
        ColsPipeArgs p{};
        p.out = t_loc; p.out_sj = pr; p.out_so = tplane; p.nvalid = nphi; p.nb = nr; p.no = nzl; p.taxis = 1;
        p.reverse = 0; p.scale = dphi * SQRT_M_1_PI; p.SN = tphi.SN; p.WM = tphi.WM;
        FDMB_CUDA(launch_cols_pipe(nphi, XF_PFWD, tm_tphi_l, p, st, "cyl_phi_fwd"));
    }
    if (!preload_only()) {   // tridiagonal solves along r; rows are (z slot, phi mode) pairs here
        TridiagArgs t{};
        t.data = t_loc; t.pitch = pr; t.nr = nr; t.nmid = nphi; t.nsys = (long long)nzl * nphi; t.swap = 1;
        t.lm_outer = d_lmphi; t.lm_mid = d_lmz; t.mid0 = z_first + J0; t.c0 = -2 / (dr * dr);
        t.L = d_L; t.U = d_U; t.ir2 = d_ir2; t.piv = d_piv;
        FDMB_CUDA(launch_tridiag_rows(t, st, "cyl_r_tridiag"));
    }
    {   // phi inverse, stores scattered back into the peers' slabs A_q[phi & (Sphi-1)][z][r]
        ColsPipeArgs p{};
        p.out = nullptr; p.nvalid = nphi; p.nb = nr; p.no = nzl; p.taxis = 1; p.reverse = 0; p.scale = SQRT_M_1_PI;
        p.SN = tphi.SN; p.WM = tphi.WM;
        OutShard om{};
        for (int q = 0; q < nranks; q++) om.base[q] = reinterpret_cast<double*>(peer_block[q]);
        om.logS = ilog2i(Sphi); om.maskS = Sphi - 1; om.sj = wplane; om.so = pr; om.o_off = z_first;
        FDMB_CUDA(launch_cols_pipe_shard(nphi, XF_PINV, tm_tphi, p, om, st, "cyl_phi_inv_xpose"));
    }
    if ((rc = barrier(st))) return rc;
    {   // z inverse on the slab, written to the caller's array
        ColsPipeArgs p{};
        p.out = d_out; p.out_sj = nr; p.out_so = uplane; p.nvalid = nz; p.nb = nr; p.no = Sphi; p.taxis = 1;
        p.reverse = 0; p.scale = slz; p.SN = tz.SN; p.WM = tz.WM;
        FDMB_CUDA(launch_cols_pipe(Nz, kzi, tm_za, p, st, "cyl_z_inv"));
    }
    return FDMB_OK;
}

fdmb_lapl_cyl::~fdmb_lapl_cyl()
{
    cudaFree(d_lmphi); cudaFree(d_lmz); cudaFree(d_L); cudaFree(d_U); cudaFree(d_ir2); cudaFree(d_piv);
    if (nranks > 1) {
        for (int q = 0; q < nranks; q++)
            if (q != rank && peer_ipc[q] && peer_block[q]) cudaIpcCloseMemHandle(peer_block[q]);
        cudaFree(mg_block);
    } else {
        cudaFree(d_work);
    }
    cudaFree(d_rhs); cudaFree(d_ans);
    if (stream) cudaStreamDestroy(stream);
}

int fdmb_lapl_cyl::solve_device(double* d_out, const double* d_in, cudaStream_t st)
{
    if (nranks > 1) return solve_device_sharded(d_out, d_in, st);
    const double SQRT_M_1_PI = 0.56418958354775629;   // lapl_cyl.h:175
    const long long wplane = (long long)nz * pr;      // work-array stride between phi planes
    const long long uplane = (long long)nz * nr;      // caller-array stride between phi planes
    const int kzf = zperiodic ? XF_PFWD : XF_DST, kzi = zperiodic ? XF_PINV : XF_DST;
    PdlScope pdl(pdl_small_grid((long long)nr * nz * nphi));    // launch-bound sizes (pdl.cuh)
    int rc;
    // phi forward: caller rhs -> work.  TMA-fed when the caller's rows are 16-byte multiples.
    {
        bool pipe_in = pipe_phi && (nr % 2 == 0) && ((reinterpret_cast<uintptr_t>(d_in) & 15) == 0);
        if (pipe_in && tm_in_ptr != d_in) {
            if ((rc = make_cols_maps(&tm_in, d_in, nphi, 2, nr, nz, nphi, 8ull * nr, 8ull * uplane, pipe_B(nphi)))) return rc;
            tm_in_ptr = d_in;
        }
        if (pipe_in) {
            ColsPipeArgs p{};
            p.out = d_work; p.out_sj = wplane; p.out_so = pr; p.nvalid = nphi; p.nb = nr; p.no = nz; p.taxis = 2;
            p.reverse = 0; p.scale = dphi * SQRT_M_1_PI;
            p.SN = tphi.SN; p.WM = tphi.WM;
            FDMB_CUDA(launch_cols_pipe(nphi, XF_PFWD, tm_in, p, st, "cyl_phi_fwd"));
        } else {
            ColsArgs c{};
            c.in = d_in; c.out = d_work; c.nvalid = nphi; c.in_sj = uplane; c.out_sj = wplane; c.nb = nr; c.no = nz;
            c.in_so = nr; c.out_so = pr; c.scale = dphi * SQRT_M_1_PI; c.SN = tphi.SN; c.WM = tphi.WM;
            FDMB_CUDA(launch_cols(nphi, XF_PFWD, c, st, "cyl_phi_fwd"));
        }
    }
    // z forward / inverse on the work array
    auto zsweep = [&](int kind, double scale, const char* tag, int reverse) -> cudaError_t {
        if (pipe_z) {
            ColsPipeArgs p{};
            p.out = d_work; p.out_sj = pr; p.out_so = wplane; p.nvalid = nz; p.nb = nr; p.no = nphi; p.taxis = 1;
            p.reverse = reverse; p.scale = scale; p.SN = tz.SN; p.WM = tz.WM;
            return launch_cols_pipe(Nz, kind, tm_z, p, st, tag);
        }
        ColsArgs c{};
        c.in = d_work; c.out = d_work; c.nvalid = nz; c.in_sj = c.out_sj = pr; c.nb = nr; c.no = nphi;
        c.in_so = c.out_so = wplane; c.scale = scale; c.SN = tz.SN; c.WM = tz.WM;
        return launch_cols(Nz, kind, c, st, tag);
    };
    FDMB_CUDA(zsweep(kzf, dz * slz, "cyl_z_fwd", 1));
    // tridiagonal solves along r for every (phi mode, z mode)
    {
        TridiagArgs t{};
        t.data = d_work; t.pitch = pr; t.nr = nr; t.nmid = nz; t.nsys = (long long)nphi * nz;
        t.lm_outer = d_lmphi; t.lm_mid = d_lmz; t.mid0 = zperiodic ? 0 : 1; t.c0 = -2 / (dr * dr);
        t.L = d_L; t.U = d_U; t.ir2 = d_ir2; t.piv = d_piv;
        FDMB_CUDA(launch_tridiag_rows(t, st, "cyl_r_tridiag"));
    }
    FDMB_CUDA(zsweep(kzi, slz, "cyl_z_inv", 0));
    // phi inverse: work -> caller ans
    if (pipe_phi) {
        ColsPipeArgs p{};
        p.out = d_out; p.out_sj = uplane; p.out_so = nr; p.nvalid = nphi; p.nb = nr; p.no = nz; p.taxis = 2;
            p.reverse = 0; p.scale = SQRT_M_1_PI; p.SN = tphi.SN; p.WM = tphi.WM;
        FDMB_CUDA(launch_cols_pipe(nphi, XF_PINV, tm_phi, p, st, "cyl_phi_inv"));
    } else {
        ColsArgs c{};
        c.in = d_work; c.out = d_out; c.nvalid = nphi; c.in_sj = wplane; c.out_sj = uplane; c.nb = nr; c.no = nz;
        c.in_so = pr; c.out_so = nr; c.scale = SQRT_M_1_PI; c.SN = tphi.SN; c.WM = tphi.WM;
        FDMB_CUDA(launch_cols(nphi, XF_PINV, c, st, "cyl_phi_inv"));
    }
    return FDMB_OK;
}

int fdmb_lapl_cyl::solve_host(double* ans, const double* rhs)
{
    const size_t bytes = sizeof(double) * (size_t)nr * nz * (nranks > 1 ? Sphi : nphi);   // this rank's slab
    if (!d_rhs) FDMB_CUDA(cudaMalloc(&d_rhs, bytes));
    if (!d_ans) FDMB_CUDA(cudaMalloc(&d_ans, bytes));
    FDMB_CUDA(cudaMemcpyAsync(d_rhs, rhs, bytes, cudaMemcpyHostToDevice, stream));
    int rc = solve_device(d_ans, d_rhs, stream);
    if (rc) return rc;
    FDMB_CUDA(cudaMemcpyAsync(ans, d_ans, bytes, cudaMemcpyDeviceToHost, stream));
    FDMB_CUDA(cudaStreamSynchronize(stream));
    return FDMB_OK;
}

// runs `body` with the handle's device current (several ranks may share one process)
template <typename Fn> static int cyl_on_device(fdmb_lapl_cyl* h, Fn body)
{
    if (h->nranks < 2) return body();
    int cur = 0;
    FDMB_CUDA(cudaGetDevice(&cur));
    if (cur != h->device) FDMB_CUDA(cudaSetDevice(h->device));
    int rc = body();
    if (cur != h->device) cudaSetDevice(cur);
    return rc;
}

extern "C" {

int fdmb_lapl_cyl_create(fdmb_lapl_cyl** out, double dr, double dz, double r0, double lr, double lz, int nr, int nz,
                         int nphi, int zperiodic)
{
    if (!out) { set_error("null handle pointer"); return FDMB_ERR_INVALID; }
    *out = nullptr;
    auto* h = new (std::nothrow) fdmb_lapl_cyl();
    if (!h) { set_error("out of host memory"); return FDMB_ERR_NOMEM; }
    h->dr = dr; h->dz = dz; h->r0 = r0; h->lr = lr; h->lz = lz;
    h->nr = nr; h->nz = nz; h->nphi = nphi; h->zperiodic = zperiodic ? 1 : 0;
    int rc = h->init();
    if (rc) { delete h; return rc; }
    *out = h;
    return FDMB_OK;
}

int fdmb_lapl_cyl_create_sharded(fdmb_lapl_cyl** out, double dr, double dz, double r0, double lr, double lz, int nr, int nz,
                                 int nphi, int zperiodic, int rank, int nranks)
{
    if (!out) { set_error("null handle pointer"); return FDMB_ERR_INVALID; }
    *out = nullptr;
    auto* h = new (std::nothrow) fdmb_lapl_cyl();
    if (!h) { set_error("out of host memory"); return FDMB_ERR_NOMEM; }
    h->dr = dr; h->dz = dz; h->r0 = r0; h->lr = lr; h->lz = lz;
    h->nr = nr; h->nz = nz; h->nphi = nphi; h->zperiodic = zperiodic ? 1 : 0;
    h->rank = rank; h->nranks = nranks;
    int rc = h->init();
    if (rc) { delete h; return rc; }
    *out = h;
    return FDMB_OK;
}

int fdmb_lapl_cyl_local_slab(fdmb_lapl_cyl* h, int* phi_first, int* nphi_local)
{
    if (!h || !phi_first || !nphi_local) { set_error("null argument"); return FDMB_ERR_INVALID; }
    *phi_first = h->nranks > 1 ? h->phi_first : 0;
    *nphi_local = h->nranks > 1 ? h->Sphi : h->nphi;
    return FDMB_OK;
}

int fdmb_lapl_cyl_export_ipc(fdmb_lapl_cyl* h, void* handle)
{
    if (!h || !handle || h->nranks < 2) { set_error("export_ipc needs a sharded handle"); return FDMB_ERR_INVALID; }
    cudaIpcMemHandle_t ih;
    FDMB_CUDA(cudaIpcGetMemHandle(&ih, h->mg_block));
    memcpy(handle, &ih, sizeof(ih));
    return FDMB_OK;
}

int fdmb_lapl_cyl_attach_ipc(fdmb_lapl_cyl* h, const void* handles)
{
    if (!h || !handles || h->nranks < 2) { set_error("attach_ipc needs a sharded handle"); return FDMB_ERR_INVALID; }
    for (int q = 0; q < h->nranks; q++) {
        if (q == h->rank) continue;
        cudaIpcMemHandle_t ih;
        memcpy(&ih, (const char*)handles + (size_t)q * FDMB_IPC_HANDLE_BYTES, sizeof(ih));
        cudaError_t e = cudaIpcOpenMemHandle(&h->peer_block[q], ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("cudaIpcOpenMemHandle for rank %d failed: %s", q, cudaGetErrorString(e));
            return FDMB_ERR_COMM;
        }
        h->peer_ipc[q] = true;
    }
    h->attached = true;
    return FDMB_OK;
}

int fdmb_lapl_cyl_attach_local(fdmb_lapl_cyl* h, fdmb_lapl_cyl* const* all)
{
    if (!h || !all || h->nranks < 2) { set_error("attach_local needs a sharded handle"); return FDMB_ERR_INVALID; }
    int cur = 0;
    FDMB_CUDA(cudaGetDevice(&cur));
    FDMB_CUDA(cudaSetDevice(h->device));
    for (int q = 0; q < h->nranks; q++) {
        if (q == h->rank) continue;
        if (!all[q] || all[q]->nranks != h->nranks || all[q]->rank != q || all[q]->mg_bytes != h->mg_bytes) {
            set_error("attach_local: handle %d does not belong to this sharded solve", q);
            cudaSetDevice(cur);
            return FDMB_ERR_INVALID;
        }
        if (all[q]->device != h->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(all[q]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) {
                set_error("cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", h->device, all[q]->device, cudaGetErrorString(e));
                cudaSetDevice(cur);
                return FDMB_ERR_COMM;
            }
        }
        h->peer_block[q] = all[q]->mg_block;
    }
    FDMB_CUDA(cudaSetDevice(cur));
    h->attached = true;
    return FDMB_OK;
}

int fdmb_lapl_cyl_solve(fdmb_lapl_cyl* h, double* ans, const double* rhs)
{
    if (!h || !ans || !rhs) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return cyl_on_device(h, [&] { return h->solve_host(ans, rhs); });
}

int fdmb_lapl_cyl_solve_device(fdmb_lapl_cyl* h, double* d_ans, const double* d_rhs, void* stream)
{
    if (!h || !d_ans || !d_rhs) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return cyl_on_device(h, [&] { return h->solve_device(d_ans, d_rhs, stream ? (cudaStream_t)stream : h->stream); });
}

int fdmb_lapl_cyl_destroy(fdmb_lapl_cyl* h)
{
    delete h;
    return FDMB_OK;
}

}  // extern "C"
