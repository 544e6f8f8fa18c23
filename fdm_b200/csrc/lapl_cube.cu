// LaplCube on B200: 3-D Poisson solve, all-Dirichlet or all-periodic.
// Replaces fdm::LaplCube<double,check,F>::solve (reference src/lapl_cube.cpp:9-142,
// constructor src/lapl_cube.h:58-100, eigenvalues src/lapl_cube.cpp:145-172).
//
// Sweep structure (v1): x rows -> y columns -> [z forward, divide, z inverse] -> y -> x.
// The work array is pitched to a multiple of 16 doubles in x so that every strided
// tile segment is 128-byte aligned; the caller's arrays keep the reference layout.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <new>

#include "common.h"
#include "xform_kernels.cuh"
#include "lapl_cube.h"

namespace fdmb {

template <int N, int KIND>
static cudaError_t rows_n(const RowsArgs& a, cudaStream_t st) { return launch_rows_t<N, KIND>(a, st); }

cudaError_t launch_rows(int N, int kind, const RowsArgs& a, cudaStream_t st, const char* tag)
{
    LaunchScope scope(tag, st);
#define X(NN)                                                                  \
    case NN:                                                                   \
        if (kind == XF_DST) return launch_rows_t<NN, XF_DST>(a, st);           \
        if (kind == XF_PFWD) return launch_rows_t<NN, XF_PFWD>(a, st);         \
        if (kind == XF_DCT) return launch_rows_t<NN, XF_DCT>(a, st);           \
        return launch_rows_t<NN, XF_PINV>(a, st);
    switch (N) { FDMB_FOR_EACH_N(X) }
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t launch_cols(int N, int kind, const ColsArgs& a, cudaStream_t st, const char* tag)
{
    LaunchScope scope(tag, st);
    MidNone mid;
#define X(NN)                                                                            \
    case NN:                                                                             \
        if (kind == XF_DST) return launch_cols_t<NN, XF_DST, MidNone, XF_DST>(a, mid, st);   \
        if (kind == XF_PFWD) return launch_cols_t<NN, XF_PFWD, MidNone, XF_DST>(a, mid, st); \
        return launch_cols_t<NN, XF_PINV, MidNone, XF_DST>(a, mid, st);
    switch (N) { FDMB_FOR_EACH_N(X) }
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t launch_cols_cube_divide(int N, bool periodic, const ColsArgs& a, const MidCubeDivide& mid,
                                    cudaStream_t st, const char* tag)
{
    LaunchScope scope(tag, st);
#define X(NN)                                                                                   \
    case NN:                                                                                    \
        if (periodic) return launch_cols_t<NN, XF_PFWD, MidCubeDivide, XF_PINV>(a, mid, st);    \
        return launch_cols_t<NN, XF_DST, MidCubeDivide, XF_DST>(a, mid, st);
    switch (N) { FDMB_FOR_EACH_N(X) }
#undef X
    return cudaErrorInvalidValue;
}


// picks the tensor maps and box shape a sweep kernel expects: planar pair for the fused DST kernels
#define FDMB_PICK_MAPS(fused)                                                                     \
    const bool use_p = (fused);                                                                   \
    if (use_p && !tm.planar) return cudaErrorInvalidValue;                                        \
    const CUtensorMap& m1 = use_p ? tm.podd : tm.nat;                                             \
    const CUtensorMap& m2 = use_p ? tm.peven : tm.nat;                                            \
    a.boxrows = use_p ? tm.boxrows_p : tm.boxrows;                                                \
    a.nchunk = use_p ? tm.nchunk_p : tm.nchunk;                                                   \
    if (tm.taxis) { a.taxis = tm.taxis; a.boxhi = use_p ? tm.boxhi_p : tm.boxhi; a.blog = tm.blog; }

#define FDMB_FOR_EACH_PIPE_N(X) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)
#define FDMB_FOR_EACH_ROWS_PIPE_N(X) X(32) X(64) X(128) X(256) X(512) X(1024)

cudaError_t launch_rows_pipe(int N, int kind, const RowsPipeArgs& a, cudaStream_t st, const char* tag)
{
    LaunchScope scope(tag, st);
    if (N == 1024 && kind == XF_DST && ring_enabled() && rows_ring_fits<1024>(a)) return launch_rows_ring_t<1024>(a, st);
#define X(NN)                                                                       \
    case NN:                                                                        \
        if (kind == XF_DST) return launch_rows_pipe_t<NN, XF_DST>(a, st);           \
        if (kind == XF_PFWD) return launch_rows_pipe_t<NN, XF_PFWD>(a, st);         \
        return launch_rows_pipe_t<NN, XF_PINV>(a, st);
    switch (N) { FDMB_FOR_EACH_ROWS_PIPE_N(X) }
#undef X
    return cudaErrorInvalidValue;
}

// FDMB_RING_PAIR (bit 0: y sweeps, bit 1: z sweep): the two groups of a ring CTA take ADJACENT tiles (the two 64-byte
// halves of the same lines and pages) instead of tiles a grid stride apart.  Measured (r02x): 22.25-22.31 ms against
// 22.42-22.43 ms -- within the run-to-run noise; off.
static int ring_pair()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("FDMB_RING_PAIR"); v = e ? atoi(e) : 0; }
    return v;
}

cudaError_t launch_cols_pipe(int N, int kind, const ColsMaps& tm, ColsPipeArgs a, cudaStream_t st, const char* tag)
{
    LaunchScope scope(tag, st);
    MidNone mid;
    a.pair_tiles = ring_pair() & 1;
    FDMB_PICK_MAPS(kind == XF_DST)
    if (N == 1024 && kind == XF_DST && ring_enabled()) return launch_cols_ring_t<1024, MidNone>(m1, m2, a, mid, st);
#define X(NN)                                                                                          \
    case NN:                                                                                           \
        if (kind == XF_DST) return launch_cols_pipe_t<NN, XF_DST, MidNone, XF_DST>(m1, m2, a, mid, st);    \
        if (kind == XF_PFWD) return launch_cols_pipe_t<NN, XF_PFWD, MidNone, XF_DST>(m1, m2, a, mid, st);  \
        return launch_cols_pipe_t<NN, XF_PINV, MidNone, XF_DST>(m1, m2, a, mid, st);
    switch (N) { FDMB_FOR_EACH_PIPE_N(X) }
#undef X
    return cudaErrorInvalidValue;
}

// DST sweep along the blocked axis of a blocked work array (ColsMaps::taxis == 3), written back in place
cudaError_t launch_cols_pipe_blocked(int N, const ColsMaps& tm, ColsPipeArgs a, const OutBlocked& ob, cudaStream_t st,
                                     const char* tag)
{
    LaunchScope scope(tag, st);
    MidNone mid;
    FDMB_PICK_MAPS(true)
    if (N == 1024 && ring_enabled()) return launch_cols_ring_t<1024, MidNone, OutBlocked>(m1, m2, a, mid, st, ob);
#define X(NN) case NN: return launch_cols_pipe_t<NN, XF_DST, MidNone, XF_DST, OutBlocked>(m1, m2, a, mid, st, ob);
    switch (N) { FDMB_FOR_EACH_PIPE_N(X) }
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t launch_cols_pipe_cube_divide(int N, bool periodic, const ColsMaps& tm, ColsPipeArgs a,
                                         const MidCubeDivide& mid, cudaStream_t st, const char* tag)
{
    LaunchScope scope(tag, st);
    a.pair_tiles = (ring_pair() >> 1) & 1;
    FDMB_PICK_MAPS(!periodic)
    if (N == 1024 && !periodic && ring_enabled()) return launch_cols_ring_t<1024, MidCubeDivide>(m1, m2, a, mid, st);
#define X(NN)                                                                                              \
    case NN:                                                                                               \
        if (periodic) return launch_cols_pipe_t<NN, XF_PFWD, MidCubeDivide, XF_PINV>(m1, m2, a, mid, st);  \
        return launch_cols_pipe_t<NN, XF_DST, MidCubeDivide, XF_DST>(m1, m2, a, mid, st);
    switch (N) { FDMB_FOR_EACH_PIPE_N(X) }
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t launch_cols_pipe_shard(int N, int kind, const ColsMaps& tm, ColsPipeArgs a, const OutShard& om,
                                   cudaStream_t st, const char* tag)
{
    LaunchScope scope(tag, st);
    MidNone mid;
    FDMB_PICK_MAPS(kind == XF_DST)
#define X(NN)                                                                                                         \
    case NN:                                                                                                          \
        if (kind == XF_DST) return launch_cols_pipe_t<NN, XF_DST, MidNone, XF_DST, OutShard>(m1, m2, a, mid, st, om);   \
        if (kind == XF_PFWD) return launch_cols_pipe_t<NN, XF_PFWD, MidNone, XF_DST, OutShard>(m1, m2, a, mid, st, om); \
        return launch_cols_pipe_t<NN, XF_PINV, MidNone, XF_DST, OutShard>(m1, m2, a, mid, st, om);
    switch (N) { FDMB_FOR_EACH_PIPE_N(X) }
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t launch_cols_pipe_cube_divide_shard(int N, bool periodic, const ColsMaps& tm, ColsPipeArgs a,
                                               const MidCubeDivide& mid, const OutShard& om, cudaStream_t st,
                                               const char* tag)
{
    LaunchScope scope(tag, st);
    FDMB_PICK_MAPS(!periodic)
#define X(NN)                                                                                                            \
    case NN:                                                                                                             \
        if (periodic) return launch_cols_pipe_t<NN, XF_PFWD, MidCubeDivide, XF_PINV, OutShard>(m1, m2, a, mid, st, om);  \
        return launch_cols_pipe_t<NN, XF_DST, MidCubeDivide, XF_DST, OutShard>(m1, m2, a, mid, st, om);
    switch (N) { FDMB_FOR_EACH_PIPE_N(X) }
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t launch_cols_pipe_shard_tma(int N, const ColsMaps& tm, ColsPipeArgs a, const OutShardTma& om, cudaStream_t st,
                                       const char* tag)
{
    LaunchScope scope(tag, st);
    MidNone mid;
    FDMB_PICK_MAPS(true)
#define X(NN) case NN: return launch_cols_pipe_t<NN, XF_DST, MidNone, XF_DST, OutShardTma>(m1, m2, a, mid, st, om);
    switch (N) { FDMB_FOR_EACH_ROWS_PIPE_N(X) }        // lengths up to 1024: 16-column, unswizzled tiles
#undef X
    return cudaErrorInvalidValue;
}

cudaError_t launch_cols_pipe_cube_divide_shard_tma(int N, const ColsMaps& tm, ColsPipeArgs a, const MidCubeDivide& mid,
                                                   const OutShardTma& om, cudaStream_t st, const char* tag)
{
    LaunchScope scope(tag, st);
    FDMB_PICK_MAPS(true)
#define X(NN) case NN: return launch_cols_pipe_t<NN, XF_DST, MidCubeDivide, XF_DST, OutShardTma>(m1, m2, a, mid, st, om);
    switch (N) { FDMB_FOR_EACH_ROWS_PIPE_N(X) }
#undef X
    return cudaErrorInvalidValue;
}

// ---- cross-GPU barrier on the handle's stream -----------------------------------------------------
// Thread t publishes this rank's arrival epoch into rank t's flag array (release, system scope: all
// peer stores of the kernels that ran before on this stream are complete at the kernel boundary),
// then waits until rank t's epoch has arrived here.  A lost peer traps after ~30 s instead of hanging.
struct PeerFlags { unsigned long long* f[FDMB_MAX_RANKS]; };

// The epoch is NOT a kernel argument: every rank counts its own barriers in device memory (slot FDMB_MAX_RANKS of its
// flag array; all ranks execute the same sequence of barriers, so the counters agree).  A launch sequence that contains
// barriers can therefore be captured once and replayed as a CUDA graph (the sharded NS steps).
__global__ void k_mg_barrier(PeerFlags pf, int rank, int nranks)
{
    __shared__ unsigned long long s_epoch;
    const int t = threadIdx.x;
    if (t == 0) {
        unsigned long long* cnt = pf.f[rank] + FDMB_MAX_RANKS;
        s_epoch = *cnt + 1;
        *cnt = s_epoch;
    }
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    if (t < nranks) {
        __threadfence_system();
        unsigned long long* dst = pf.f[t] + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(epoch) : "memory");
        const unsigned long long* src = pf.f[rank] + t;
        unsigned long long t0, t1, seen;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(src) : "memory");
            if (seen >= epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 30000000000ull) __trap();
            __nanosleep(200);
        }
    }
    __syncthreads();
    __threadfence_system();
}

int launch_mg_barrier(void* const* peer_blocks, size_t off_flags, int rank, int nranks, cudaStream_t st)
{
    PeerFlags pf{};
    for (int q = 0; q < nranks; q++)
        pf.f[q] = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(peer_blocks[q]) + off_flags);
    LaunchScope scope("mg_barrier", st);
    k_mg_barrier<<<1, 32, 0, st>>>(pf, rank, nranks);
    FDMB_CHECK_LAUNCH();
    return FDMB_OK;
}

int preload_mg_barrier()
{
    cudaFuncAttributes fa;
    FDMB_CUDA(cudaFuncGetAttributes(&fa, k_mg_barrier));
    return FDMB_OK;
}

}  // namespace fdmb

using namespace fdmb;

static inline double sq(double x) { return x * x; }

int fdmb_lapl_cube::init()
{
    // transform lengths: Dirichlet n+1, periodic n (lapl_cube.h:74-76)
    Nx = periodic ? nx : nx + 1;
    Ny = periodic ? ny : ny + 1;
    Nz = periodic ? nz : nz + 1;
    if (nx < 1 || ny < 1 || nz < 1 || !supported_N(Nx) || !supported_N(Ny) || !supported_N(Nz)) {
        set_error("LaplCube: %s axis sizes (%d,%d,%d) need transform lengths that are powers of two in [4,2048] "
                  "(reference: verify((1<<n) == N), src/fft.cpp:67)",
                  periodic ? "periodic" : "Dirichlet", nx, ny, nz);
        return FDMB_ERR_INVALID;
    }
    slx = std::sqrt(2. / lx); sly = std::sqrt(2. / ly); slz = std::sqrt(2. / lz);
    px = (nx + 15) / 16 * 16;
    int rc;
    if ((rc = get_tables(Nx, &tx)) || (rc = get_tables(Ny, &ty)) || (rc = get_tables(Nz, &tz))) return rc;

    // eigenvalues, including the aliasing quirk of lapl_cube.cpp:162,171
    const int x1 = periodic ? 0 : 1, xn = periodic ? nx - 1 : nx;
    const int y1 = periodic ? 0 : 1, yn = periodic ? ny - 1 : ny;
    const int z1 = periodic ? 0 : 1, zn = periodic ? nz - 1 : nz;
    const double dx2 = dx * dx, dy2 = dy * dy, dz2 = dz * dz;
    std::vector<double> lm_y(ny + 1, 0.0), lm_x(nx + 1, 0.0), lm_z(nz + 1, 0.0);
    for (int k = y1; k <= yn; k++)
        lm_y[k] = periodic ? 4. / dy2 * sq(sin(k * M_PI / (ny))) : 4. / dy2 * sq(sin(k * M_PI * 0.5 / (ny + 1)));
    for (int j = x1; j <= xn; j++)
        lm_x[j] = periodic ? 4. / dx2 * sq(sin(j * M_PI / (nx))) : 4. / dx2 * sq(sin(j * M_PI * 0.5 / (nx + 1)));
    for (int i = z1; i <= zn; i++)
        lm_z[i] = periodic ? 4. / dz2 * sq(sin(i * M_PI / (nz))) : 4. / dz2 * sq(sin(i * M_PI * 0.5 / (nz + 1)));
    if (Nx == Ny) lm_x = lm_y;   // lm_x = xpoints == ypoints ? &lm_y[0] : &lm_x_[0]
    if (Nz == Ny) lm_z = lm_y;

    FDMB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    FDMB_CUDA(cudaMalloc(&d_lmx, sizeof(double) * (nx + 1)));
    FDMB_CUDA(cudaMalloc(&d_lmy, sizeof(double) * (ny + 1)));
    FDMB_CUDA(cudaMalloc(&d_lmz, sizeof(double) * (nz + 1)));
    FDMB_CUDA(cudaMemcpy(d_lmx, lm_x.data(), sizeof(double) * (nx + 1), cudaMemcpyHostToDevice));
    FDMB_CUDA(cudaMemcpy(d_lmy, lm_y.data(), sizeof(double) * (ny + 1), cudaMemcpyHostToDevice));
    FDMB_CUDA(cudaMemcpy(d_lmz, lm_z.data(), sizeof(double) * (nz + 1), cudaMemcpyHostToDevice));
    FDMB_CUDA(cudaGetDevice(&device));
    if (nranks > 1) return init_sharded();
    // Optional BLOCKED work array ([yb][z][yi][x], 64 rows per block; FDMB_BLOCKED=1).  In the natural layout the
    // 1023 rows of a z tile at 1023^3 sit on 1023 different 2 MB pages and a plain copy with that access pattern
    // runs at 2.4 TB/s against 5.0 TB/s with a short stride (profiles/r01g_access_pattern.md).  Blocking fixes the
    // copy rate, but the z sweep is bound by the SMs' shared-memory pipe today, so the solve gains only ~1 %: off
    // by default until the sweeps' compute side is faster.
    {
        const char* e = getenv("FDMB_BLOCKED");
        const bool can = pipe_enabled() && !periodic && rows_pipe_supported_N(Nx) && pipe_supported_N(Ny) && pipe_supported_N(Nz);
        blog = (can && e && e[0] == '1') ? 6 : 0;
        if (blog && (1 << blog) > Ny / 2) blog = 4;      // small grids (tests): keep at least two blocks
        if (const char* b = getenv("FDMB_BLOG")) if (blog) blog = atoi(b);
    }
    if (blog) {
        const int YB = 1 << blog;
        nyb = (ny + YB - 1) / YB;
        const size_t welems = (size_t)nyb * nz * YB * px;
        FDMB_CUDA(cudaMalloc(&d_work, sizeof(double) * welems));
        FDMB_CUDA(cudaMemset(d_work, 0, sizeof(double) * welems));      // the padding rows of the last block stay zero
        // cudaMemset on device memory is asynchronous to the host and runs on the legacy default stream, which the
        // handle's non-blocking streams do not order against: finish it before the handle is handed out
        FDMB_CUDA(cudaDeviceSynchronize());
        const unsigned long long s_lo = 8ull * px, s_mid = s_lo * YB, s_hi = s_mid * (unsigned long long)nz;
        if ((rc = make_cols_maps_blocked(&tm_y, d_work, Ny, true, blog, nx, ny, nz, s_lo, s_mid, s_hi, pipe_B(Ny)))) return rc;
        if ((rc = make_cols_maps_blocked(&tm_z, d_work, Nz, false, blog, nx, ny, nz, s_lo, s_mid, s_hi, pipe_B(Nz)))) return rc;
        pipe_y = pipe_z = true;
        return FDMB_OK;
    }
    FDMB_CUDA(cudaMalloc(&d_work, sizeof(double) * (size_t)nz * ny * px));
    if (pipe_enabled()) {
        // tensor maps over the pitched work array: dims (x, y, z), tiles [N][B] along y or z
        const unsigned long long s1 = 8ull * px, s2 = 8ull * (unsigned long long)ny * px;
        if (pipe_supported_N(Ny)) {
            if ((rc = make_cols_maps(&tm_y, d_work, Ny, 1, nx, ny, nz, s1, s2, pipe_B(Ny)))) return rc;
            pipe_y = true;
        }
        if (pipe_supported_N(Nz)) {
            if ((rc = make_cols_maps(&tm_z, d_work, Nz, 2, nx, ny, nz, s1, s2, pipe_B(Nz)))) return rc;
            pipe_z = true;
        }
    }
    return FDMB_OK;
}

static int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

const ColsMaps* fdmb_lapl_cube::y_chunk_maps(bool wide, int z0, int nzc)
{
    const int nz_here = nranks > 1 ? nzl : nz;
    if (z0 == 0 && nzc == nz_here) return wide ? &tm_yw : &tm_y;
    auto& cache = wide ? tm_yw_chunk : tm_y_chunk;
    auto it = cache.find({z0, nzc});
    if (it != cache.end()) return &it->second;
    ColsMaps m{};
    const unsigned long long plane_b = 8ull * (unsigned long long)ny * px;
    const int B = wide ? pipe_B_sharded(Ny) : pipe_B(Ny);
    if (make_cols_maps(&m, d_work + (size_t)z0 * ny * px, Ny, 1, nx, ny, nzc, 8ull * px, plane_b, B)) return nullptr;
    return &cache.emplace(std::make_pair(z0, nzc), m).first->second;
}

static bool xinv_inplace()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("FDMB_XINV_INPLACE"); v = (e && e[0] == '1') ? 1 : 0; }
    return v != 0;
}

// FDMB_MG_OVERLAP: z chunks of the x-sweep / transposing-y-sweep overlap of the sharded solve (0 or 1: off)
static int mg_overlap_chunks(int nranks)
{
    const char* e = getenv("FDMB_MG_OVERLAP");
    return e ? atoi(e) : (nranks <= 2 ? 8 : 4);      // measured (r02p, r02q): 2 GPUs 14.19 -> 13.73 ms, 8 GPUs 4.64 -> 4.37 ms
}
// FDMB_MG_SPLIT: SMs the x sweep gets while it runs beside the transposing y sweep.  The y sweep needs SM time in
// proportion to 1 / nranks of a full sweep but NVLink time that barely shrinks with nranks: the more ranks, the
// fewer SMs it needs.
static int mg_overlap_split(int nranks, int sms)
{
    if (const char* e = getenv("FDMB_MG_SPLIT")) { int v = atoi(e); if (v > 0 && v < sms) return v; }
    return nranks <= 2 ? sms / 2 : (nranks == 4 ? (sms * 3) / 8 : sms / 3);
}

int fdmb_lapl_cube::init_sharded()
{
    const int J0 = periodic ? 0 : 1;
    if (nranks != 2 && nranks != 4 && nranks != 8) {
        set_error("LaplCube: nranks must be 1, 2, 4 or 8 (got %d)", nranks);
        return FDMB_ERR_INVALID;
    }
    if (rank < 0 || rank >= nranks) { set_error("LaplCube: rank %d out of range", rank); return FDMB_ERR_INVALID; }
    if (!pipe_supported_N(Ny) || !pipe_supported_N(Nz) || Ny / nranks < 2 || Nz / nranks < 2) {
        set_error("LaplCube: the sharded solve needs y/z transform lengths >= 32 and >= 2*nranks (got %d, %d)", Ny, Nz);
        return FDMB_ERR_INVALID;
    }
    Sy = Ny / nranks; Sz = Nz / nranks;
    slab_range(nz, periodic, nranks, rank, &z_first, &nzl);
    slab_range(ny, periodic, nranks, rank, &y_first, &nyl);
    const size_t plane = (size_t)ny * px;
    const size_t a_bytes = (sizeof(double) * (size_t)Sz * plane + 255) & ~(size_t)255;
    const size_t t_bytes = (sizeof(double) * (size_t)nz * Sy * px + 255) & ~(size_t)255;
    off_T = a_bytes; off_flags = a_bytes + t_bytes;
    mg_bytes = off_flags + 256;
    FDMB_CUDA(cudaMalloc(&mg_block, mg_bytes));
    FDMB_CUDA(cudaMemset(mg_block, 0, mg_bytes));
    // cudaMemset on device memory is asynchronous to the host and runs on the legacy default stream, which the
    // handle's non-blocking streams do not order against: finish it before the handle is handed out
    FDMB_CUDA(cudaDeviceSynchronize());
    d_A = reinterpret_cast<double*>(mg_block);
    d_T = reinterpret_cast<double*>(reinterpret_cast<char*>(mg_block) + off_T);
    d_flags = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(mg_block) + off_flags);
    peer_block[rank] = mg_block;
    // local views: the uniform allocations keep a slot-0 plane / row on rank 0 of a Dirichlet axis
    d_work = d_A + (size_t)(z_first + J0 - rank * Sz) * plane;
    double* t_loc = d_T + (size_t)(y_first + J0 - rank * Sy) * px;
    int rc;
    // the transposing sweeps (y forward, z) use the wide tiles of PipeCfg<N, true>; the local y inverse the normal ones
    if ((rc = make_cols_maps(&tm_y, d_work, Ny, 1, nx, ny, nzl, 8ull * px, 8ull * plane, pipe_B(Ny)))) return rc;
    if ((rc = make_cols_maps(&tm_yw, d_work, Ny, 1, nx, ny, nzl, 8ull * px, 8ull * plane, pipe_B_sharded(Ny)))) return rc;
    if ((rc = make_cols_maps(&tm_z, t_loc, Nz, 2, nx, nyl, nz, 8ull * px, 8ull * (unsigned long long)Sy * px,
                             pipe_B_sharded(Nz))))
        return rc;
    pipe_y = pipe_z = true;
    {   // load every kernel of the sharded solve on this device now (preload_only(), xform_pipe.cuh)
        cudaFuncAttributes fa;
        FDMB_CUDA(cudaFuncGetAttributes(&fa, k_mg_barrier));
        preload_only() = true;
        attached = true;
        rc = solve_device_sharded(d_work, d_work, stream);
        attached = false;
        preload_only() = false;
        if (rc) return rc;
    }
    return FDMB_OK;
}

int fdmb_lapl_cube::attach(void* const* bases)
{
    for (int q = 0; q < nranks; q++)
        if (q != rank) peer_block[q] = bases[q];
    attached = true;
    return build_store_maps();
}

// FDMB_MG_TMA_STORE=1: the transposing sweeps hand their finished tiles to the TMA unit (OutShardTma) instead of storing
// every value from the transform's registers (EmitShard).  Parity-green, but measured SLOWER on 2 GPUs (r02p: 14.93 vs
// 13.71 ms per solve; y forward 733 vs 584 us per chunk): with one 128 KB tile per CTA the tile has to be written to
// shared memory, read by the TMA unit and released before the next load can start, while the register stores stream out
// during the last two transform stages; and the NVLink packets are the same 128-byte rows either way.  Off by default.
static bool mg_tma_store()
{
    const char* e = getenv("FDMB_MG_TMA_STORE");
    return e && e[0] == '1';
}

// Tensor maps over every rank's pencil buffer T_q[z'][y slot][x] and slab A_q[z slot][y'][x], split into the views of
// the odd and the even local slots (a planar tile holds a rank's odd slots and its even slots as two contiguous row
// blocks).  Dirichlet sweeps only (the planar layout); boxes hold half a rank's slots (<= 256 rows).
int fdmb_lapl_cube::build_store_maps()
{
    st_ready = false;
    if (periodic || !mg_tma_store() || Sy / 2 > 256 || Sz / 2 > 256 || Sy < 2 || Sz < 2 || Ny > 1024 || Nz > 1024) return FDMB_OK;
    const unsigned long long row = 8ull * px, plane_b = 8ull * (unsigned long long)ny * px;
    const unsigned By = pipe_B_sharded(Ny), Bz = pipe_B_sharded(Nz);
    int rc;
    for (int q = 0; q < nranks; q++) {
        char* T = reinterpret_cast<char*>(peer_block[q]) + off_T;
        char* A = reinterpret_cast<char*>(peer_block[q]);
        // y forward: destination rows are y slots (stride px), planes are z' (stride Sy * px)
        if ((rc = make_tensor_map_3d(&st_y.even[q], T, nx, Sy / 2, nz, 2 * row, (unsigned long long)Sy * row, By, Sy / 2, 1))) return rc;
        if ((rc = make_tensor_map_3d(&st_y.odd[q], T + row, nx, Sy / 2, nz, 2 * row, (unsigned long long)Sy * row, By, Sy / 2, 1))) return rc;
        // z sweep: destination "rows" are z slots (stride one plane), the middle index is y'
        if ((rc = make_tensor_map_3d(&st_z.even[q], A, nx, ny, Sz / 2, row, 2 * plane_b, Bz, 1, Sz / 2))) return rc;
        if ((rc = make_tensor_map_3d(&st_z.odd[q], A + plane_b, nx, ny, Sz / 2, row, 2 * plane_b, Bz, 1, Sz / 2))) return rc;
    }
    st_y.nranks = st_z.nranks = nranks;
    st_y.half = Sy / 2; st_y.taxis = 1; st_y.o_off = z_first;
    st_z.half = Sz / 2; st_z.taxis = 2; st_z.o_off = y_first;
    st_ready = true;
    return FDMB_OK;
}

int fdmb_lapl_cube::barrier(cudaStream_t st)
{
    if (preload_only()) return FDMB_OK;
    PeerFlags pf{};
    for (int q = 0; q < nranks; q++)
        pf.f[q] = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(peer_block[q]) + off_flags);
    LaunchScope scope("cube_mg_barrier", st);
    k_mg_barrier<<<1, 32, 0, st>>>(pf, rank, nranks);
    FDMB_CHECK_LAUNCH();
    return FDMB_OK;
}

fdmb_lapl_cube::~fdmb_lapl_cube()
{
    cudaFree(d_lmx); cudaFree(d_lmy); cudaFree(d_lmz);
    if (nranks > 1) {
        for (int q = 0; q < nranks; q++)
            if (q != rank && peer_ipc[q] && peer_block[q]) cudaIpcCloseMemHandle(peer_block[q]);
        cudaFree(mg_block);
    } else {
        cudaFree(d_work);
    }
    cudaFree(d_rhs); cudaFree(d_ans); cudaFree(d_rhs1); cudaFree(d_ans1);
    for (int b = 0; b < 2; b++) {
        if (ev_up[b]) cudaEventDestroy(ev_up[b]);
        if (ev_cmp[b]) cudaEventDestroy(ev_cmp[b]);
        if (ev_dn[b]) cudaEventDestroy(ev_dn[b]);
    }
    if (s_up) cudaStreamDestroy(s_up);
    if (s_dn) cudaStreamDestroy(s_dn);
    if (s_side) cudaStreamDestroy(s_side);
    if (ev_fork) cudaEventDestroy(ev_fork);
    for (auto& e : ev_chunk) if (e) cudaEventDestroy(e);
    for (auto& e : ev_done) if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
}

// Sharded solve: x rows (local) -> y columns, stores scattered into the peers' pencil buffers ->
// barrier -> z forward, divide, z inverse on the local pencils, stores scattered back into the
// peers' slabs -> barrier -> y columns, x rows (local).  d_in / d_out are this rank's z-slab.
int fdmb_lapl_cube::solve_device_sharded(double* d_out, const double* d_in, cudaStream_t st)
{
    if (!attached) { set_error("LaplCube: sharded handle used before attach_ipc/attach_local"); return FDMB_ERR_COMM; }
    const int kf = periodic ? XF_PFWD : XF_DST;
    const int ki = periodic ? XF_PINV : XF_DST;
    const long long plane = (long long)ny * px;
    int rc;
    auto rows = [&](const double* in, double* out, int in_pitch, int out_pitch, double scale, int kind, const char* tag,
                    int reverse, long long nrows, cudaStream_t s, int max_ctas) -> cudaError_t {
        if (rows_pipe_supported_N(Nx) && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
            RowsPipeArgs p{};
            p.in = in; p.out = out; p.nrows = nrows; p.nvalid = nx; p.in_pitch = in_pitch;
            p.out_pitch = out_pitch; p.reverse = reverse; p.scale = scale; p.SN = tx.SN; p.WM = tx.WM; p.max_ctas = max_ctas;
            return launch_rows_pipe(Nx, kind, p, s, tag);
        }
        RowsArgs r{};
        r.in = in; r.out = out; r.nrows = nrows; r.nvalid = nx; r.in_pitch = in_pitch;
        r.out_pitch = out_pitch; r.scale = scale; r.SN = tx.SN; r.WM = tx.WM;
        return launch_rows(Nx, kind, r, s, tag);
    };
    auto y_fwd_xpose = [&](int o0, int no, int max_ctas) -> cudaError_t {
        // y forward, transposing into the pencil buffers T_q[z'][y slot & (Sy-1)][x]
        const ColsMaps* maps = y_chunk_maps(true, o0, no);
        if (!maps) return cudaErrorInvalidValue;
        ColsPipeArgs p{};
        p.out = nullptr; p.nvalid = ny; p.nb = nx; p.no = no; p.taxis = 1; p.max_ctas = max_ctas;
        p.reverse = 1; p.scale = dy * sly; p.SN = ty.SN; p.WM = ty.WM;
        if (kf == XF_DST && (st_ready || (preload_only() && Ny <= 1024 && Nz <= 1024)) && mg_tma_store()) {
            OutShardTma ot = st_y;
            ot.o_off = z_first + o0;
            return launch_cols_pipe_shard_tma(Ny, *maps, p, ot, st, "cube_y_fwd_xpose");
        }
        OutShard om{};
        for (int q = 0; q < nranks; q++)
            om.base[q] = reinterpret_cast<double*>(reinterpret_cast<char*>(peer_block[q]) + off_T);
        om.logS = ilog2(Sy); om.maskS = Sy - 1; om.sj = px; om.so = (long long)Sy * px; om.o_off = z_first + o0;
        return launch_cols_pipe_shard(Ny, kf, *maps, p, om, st, "cube_y_fwd_xpose");
    };
    // The transposing y sweep is bound by its NVLink stores (r01g: ~640 GB/s per direction), not by the SMs: cut the slab
    // into z chunks and let the x sweep of chunk c+1 run beside the y sweep of chunk c, each on its share of the SMs
    // (side stream + events; the x sweep gets `split` SMs, the y sweep the rest).  FDMB_MG_OVERLAP = chunks (0: off).
    int nch = mg_overlap_chunks(nranks);
    {   // a chunk has to be worth two launches: at least 16 M points (1023^3: 4-8 chunks; 255^3: none)
        const long long by_size = (long long)nzl * ny * nx / (16ll << 20);
        if (nch > by_size) nch = (int)by_size;
    }
    if (preload_only() || nch < 2 || nzl < 2 * nch || nch > 16 || !pipe_supported_N(Nx)) nch = 1;
    if (nch == 1) {
        FDMB_CUDA(rows(d_in, d_work, nx, px, dx * slx, kf, "cube_x_fwd", 0, (long long)nzl * ny, st, 0));
        FDMB_CUDA(y_fwd_xpose(0, nzl, 0));
    } else {
        if (!s_side) FDMB_CUDA(cudaStreamCreateWithFlags(&s_side, cudaStreamNonBlocking));
        if (!ev_fork) FDMB_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        for (int c = 0; c < nch; c++)
            if (!ev_chunk[c]) FDMB_CUDA(cudaEventCreateWithFlags(&ev_chunk[c], cudaEventDisableTiming));
        const int sms = device_sm_count();
        int split = mg_overlap_split(nranks, sms);
        FDMB_CUDA(cudaEventRecord(ev_fork, st));
        FDMB_CUDA(cudaStreamWaitEvent(s_side, ev_fork, 0));
        for (int c = 0; c < nch; c++) {
            // even chunk boundaries: the caller's unpitched planes start 16-byte aligned only every other plane
            const int z0 = (int)((long long)c * nzl / nch) & ~1;
            const int z1 = c == nch - 1 ? nzl : (int)((long long)(c + 1) * nzl / nch) & ~1;
            // the first x chunk has the machine to itself
            FDMB_CUDA(rows(d_in + (long long)z0 * ny * nx, d_work + (long long)z0 * plane, nx, px, dx * slx, kf, "cube_x_fwd", 0,
                           (long long)(z1 - z0) * ny, s_side, c == 0 ? 0 : split));
            FDMB_CUDA(cudaEventRecord(ev_chunk[c], s_side));
            FDMB_CUDA(cudaStreamWaitEvent(st, ev_chunk[c], 0));
            FDMB_CUDA(y_fwd_xpose(z0, z1 - z0, c == nch - 1 ? 0 : sms - split));
        }
    }
    if ((rc = barrier(st))) return rc;
    {   // z forward, divide, z inverse; stores go back to the slabs A_r[z slot & (Sz-1)][y'][x]
        ColsPipeArgs p{};
        p.out = nullptr; p.nvalid = nz; p.nb = nx; p.no = nyl; p.taxis = 2;
        p.reverse = 0; p.mid_o_off = y_first; p.scale = dz * slz; p.scale2 = slz; p.SN = tz.SN; p.WM = tz.WM;
        MidCubeDivide mid{d_lmz, d_lmx, d_lmy, periodic ? 1 : 0};
        if (!periodic && (st_ready || (preload_only() && Ny <= 1024 && Nz <= 1024)) && mg_tma_store()) {
            FDMB_CUDA(launch_cols_pipe_cube_divide_shard_tma(Nz, tm_z, p, mid, st_z, st, "cube_z_fwd_div_inv_xpose"));
        } else {
        OutShard om{};
        for (int q = 0; q < nranks; q++) om.base[q] = reinterpret_cast<double*>(peer_block[q]);
        om.logS = ilog2(Sz); om.maskS = Sz - 1; om.sj = plane; om.so = px; om.o_off = y_first;
        FDMB_CUDA(launch_cols_pipe_cube_divide_shard(Nz, periodic != 0, tm_z, p, mid, om, st, "cube_z_fwd_div_inv_xpose"));
        }
    }
    if ((rc = barrier(st))) return rc;
    {   // y inverse (local)
        ColsPipeArgs p{};
        p.out = d_work; p.out_sj = px; p.out_so = plane; p.nvalid = ny; p.nb = nx; p.no = nzl; p.taxis = 1;
        p.reverse = 0; p.scale = sly; p.SN = ty.SN; p.WM = ty.WM;
        FDMB_CUDA(launch_cols_pipe(Ny, ki, tm_y, p, st, "cube_y_inv"));
    }
    FDMB_CUDA(rows(d_work, d_out, px, nx, slx, ki, "cube_x_inv", 1, (long long)nzl * ny, st, 0));
    return FDMB_OK;
}

int fdmb_lapl_cube::solve_device(double* d_out, const double* d_in, cudaStream_t st)
{
    if (nranks > 1) return solve_device_sharded(d_out, d_in, st);
    return sweeps(d_out, d_in, st, 7, 0, nz);
}

// The single-GPU solve in three phases, so that the host-pointer entry point can stream z chunks through it:
//   phase 1  x and y forward sweeps of the planes [z0, z0 + nzc) of d_in into the work array
//   phase 2  z forward / divide / inverse over the whole work array
//   phase 4  y and x inverse sweeps of the planes [z0, z0 + nzc) into d_out
int fdmb_lapl_cube::sweeps(double* d_out, const double* d_in, cudaStream_t st, int phases, int z0, int nzc)
{
    PdlScope pdl(pdl_small_grid((long long)nx * ny * nz));    // launch-bound sizes: programmatic dependent launch (pdl.cuh)
    const int kf = periodic ? XF_PFWD : XF_DST;
    const int ki = periodic ? XF_PINV : XF_DST;
    const long long plane = (long long)ny * px;
    // x forward: rhs rows -> pitched work
    RowsArgs r{};
    r.in = d_in ? d_in + (long long)z0 * ny * nx : nullptr; r.out = d_work + (long long)z0 * plane; r.nrows = (long long)nzc * ny;
    r.nvalid = nx; r.in_pitch = nx; r.out_pitch = px; r.scale = dx * slx; r.SN = tx.SN; r.WM = tx.WM;
    const bool pipe_x = pipe_enabled() && rows_pipe_supported_N(Nx);
    if (blog) {
        if (phases != 7 || z0 != 0 || nzc != nz) { set_error("LaplCube: the blocked work layout does not stream z chunks"); return FDMB_ERR_INVALID; }
        // blocked work array W[yb][z][yi][x]
        if ((reinterpret_cast<uintptr_t>(d_in) & 15) != 0) {
            set_error("LaplCube: rhs must be 16-byte aligned for grids this large");
            return FDMB_ERR_INVALID;
        }
        const int YB = 1 << blog;
        const long long s_lo = px, s_mid = (long long)YB * px, s_hi = s_mid * nz;
        RowsPipeArgs rp{};
        rp.in = d_in; rp.out = d_work; rp.nrows = (long long)nz * ny; rp.nvalid = nx; rp.in_pitch = nx; rp.out_pitch = px;
        rp.reverse = 0; rp.scale = dx * slx; rp.SN = tx.SN; rp.WM = tx.WM; rp.blk = 1; rp.blog = blog; rp.ny = ny; rp.nz = nz;
        FDMB_CUDA(launch_rows_pipe(Nx, XF_DST, rp, st, "cube_x_fwd"));
        ColsPipeArgs py{};
        py.out = d_work; py.out_sj = s_lo; py.out_so = s_mid; py.nvalid = ny; py.nb = nx; py.no = nz;
        py.reverse = 1; py.scale = dy * sly; py.SN = ty.SN; py.WM = ty.WM;
        const OutBlocked ob{s_hi, blog};
        FDMB_CUDA(launch_cols_pipe_blocked(Ny, tm_y, py, ob, st, "cube_y_fwd"));
        ColsPipeArgs pz{};
        pz.out = d_work; pz.out_sj = s_mid; pz.out_so = s_lo; pz.out_so_hi = s_hi; pz.nvalid = nz; pz.nb = nx; pz.no = ny;
        pz.reverse = 0; pz.scale = dz * slz; pz.scale2 = slz; pz.SN = tz.SN; pz.WM = tz.WM;
        MidCubeDivide midb{d_lmz, d_lmx, d_lmy, 0};
        FDMB_CUDA(launch_cols_pipe_cube_divide(Nz, false, tm_z, pz, midb, st, "cube_z_fwd_div_inv"));
        py.reverse = 0; py.scale = sly;
        FDMB_CUDA(launch_cols_pipe_blocked(Ny, tm_y, py, ob, st, "cube_y_inv"));
        rp.in = d_work; rp.out = d_out; rp.in_pitch = px; rp.out_pitch = nx; rp.reverse = 1; rp.scale = slx; rp.blk = 2;
        FDMB_CUDA(launch_rows_pipe(Nx, XF_DST, rp, st, "cube_x_inv"));
        return FDMB_OK;
    }
    auto rows = [&](const RowsArgs& q, int kind, const char* tag, int reverse) -> cudaError_t {
        if (pipe_x && (reinterpret_cast<uintptr_t>(q.in) & 15) == 0) {
            RowsPipeArgs p{};
            p.in = q.in; p.out = q.out; p.nrows = q.nrows; p.nvalid = q.nvalid; p.in_pitch = (int)q.in_pitch;
            p.out_pitch = (int)q.out_pitch; p.reverse = reverse; p.scale = q.scale; p.SN = q.SN; p.WM = q.WM;
            return launch_rows_pipe(Nx, kind, p, st, tag);
        }
        return launch_rows(Nx, kind, q, st, tag);
    };
    auto cols_y = [&](const ColsArgs& q, int kind, const char* tag, int reverse) -> cudaError_t {
        if (pipe_y) {
            const ColsMaps* maps = y_chunk_maps(false, z0, nzc);     // loads AND stores start at plane z0 (q.out does)
            if (!maps) return cudaErrorInvalidValue;
            ColsPipeArgs p{};
            p.out = q.out; p.out_sj = q.out_sj; p.out_so = q.out_so; p.nvalid = q.nvalid; p.nb = q.nb; p.no = q.no;
            p.taxis = 1; p.reverse = reverse; p.scale = q.scale;
            p.scale2 = q.scale2; p.SN = q.SN; p.WM = q.WM;
            return launch_cols_pipe(Ny, kind, *maps, p, st, tag);
        }
        return launch_cols(Ny, kind, q, st, tag);
    };
    if (phases & 1) FDMB_CUDA(rows(r, kf, "cube_x_fwd", 0));
    // y forward (the plain kernel takes the chunk through its pointers, the tensor-map kernel through o0)
    ColsArgs c{};
    c.out = d_work + (long long)z0 * plane; c.in = c.out; c.nvalid = ny; c.in_sj = c.out_sj = px; c.nb = nx; c.no = nzc;
    c.in_so = c.out_so = plane; c.scale = dy * sly; c.SN = ty.SN; c.WM = ty.WM;
    if (phases & 1) FDMB_CUDA(cols_y(c, kf, "cube_y_fwd", 1));
    // z forward, divide by -(lm_z+lm_y+lm_x), z inverse
    ColsArgs z{};
    z.in = d_work; z.out = d_work; z.nvalid = nz; z.in_sj = z.out_sj = plane; z.nb = nx; z.no = ny;
    z.in_so = z.out_so = px; z.scale = dz * slz; z.scale2 = slz; z.SN = tz.SN; z.WM = tz.WM;
    MidCubeDivide mid{d_lmz, d_lmx, d_lmy, periodic ? 1 : 0};
    if (!(phases & 2)) {
    } else if (pipe_z) {
        ColsPipeArgs p{};
        p.out = z.out; p.out_sj = z.out_sj; p.out_so = z.out_so; p.nvalid = z.nvalid; p.nb = z.nb; p.no = z.no;
        p.taxis = 2; p.reverse = 0; p.scale = z.scale; p.scale2 = z.scale2;
        p.SN = z.SN; p.WM = z.WM;
        FDMB_CUDA(launch_cols_pipe_cube_divide(Nz, periodic != 0, tm_z, p, mid, st, "cube_z_fwd_div_inv"));
    } else {
        FDMB_CUDA(launch_cols_cube_divide(Nz, periodic != 0, z, mid, st, "cube_z_fwd_div_inv"));
    }
    if (phases & 4) {
        double* out_z0 = d_out + (long long)z0 * ny * nx;
        // Where the ring row sweep runs (Nx = 1024), the y inverse sweep stores straight into the caller's UNPITCHED ans
        // rows and the x inverse sweep transforms them in place: rows of odd pitch land at offsets of both parities,
        // which is what makes the row sweep's dense reads conflict-free (xform_ring.cuh); the pitched work array's rows
        // cannot (all 16-byte aligned).  Measured (r02x): the row sweep gains 0.5 ms (4.30 -> 3.80 ms), but the column
        // sweep's 64-byte segments no longer start on sector boundaries in rows of 8184 bytes and its stores turn into
        // read-modify-write traffic: 3.54 -> 9.64 ms.  OFF unless FDMB_XINV_INPLACE=1.
        const bool inplace = pipe_x && pipe_y && Nx == 1024 && !periodic && ring_enabled() && xinv_inplace() &&
                             (reinterpret_cast<uintptr_t>(out_z0) & 15) == 0;
        // y inverse
        c.scale = sly;
        if (inplace) { c.out = out_z0; c.out_sj = nx; c.out_so = (long long)ny * nx; }
        FDMB_CUDA(cols_y(c, ki, "cube_y_inv", 0));
        // x inverse: pitched work (or ans itself) -> ans rows
        r.in = inplace ? out_z0 : d_work + (long long)z0 * plane; r.out = out_z0;
        r.in_pitch = inplace ? nx : px; r.out_pitch = nx; r.scale = slx;
        FDMB_CUDA(rows(r, ki, "cube_x_inv", 1));
    }
    return FDMB_OK;
}

// Host-pointer solve.  Large single-GPU grids stream through in z chunks: the x / y forward sweeps of chunk c run while
// chunk c+1 is still crossing PCIe, and the downloads of the finished planes start while the y / x inverse sweeps of the
// later chunks run; only the z sweep (which needs every plane) and one chunk's worth of sweeps stay exposed.  Chunk
// boundaries are even (the caller's unpitched planes start 16-byte aligned every other plane).
int fdmb_lapl_cube::solve_host(double* ans, const double* rhs)
{
    const size_t bytes = sizeof(double) * (size_t)nx * ny * (nranks > 1 ? nzl : nz);   // this rank's slab
    if (!d_rhs) FDMB_CUDA(cudaMalloc(&d_rhs, bytes));
    if (!d_ans) FDMB_CUDA(cudaMalloc(&d_ans, bytes));
    int nch = 1;
    if (nranks == 1 && !blog && bytes >= (size_t(64) << 20)) {
        const char* e = getenv("FDMB_HOST_CHUNKS");
        nch = e ? atoi(e) : 8;
        if (nch > 16) nch = 16;
        if (nch < 1 || nz < 4 * nch) nch = 1;
    }
    if (nch == 1) {
        FDMB_CUDA(cudaMemcpyAsync(d_rhs, rhs, bytes, cudaMemcpyHostToDevice, stream));
        int rc = solve_device(d_ans, d_rhs, stream);
        if (rc) { cudaStreamSynchronize(stream); return rc; }
        FDMB_CUDA(cudaMemcpyAsync(ans, d_ans, bytes, cudaMemcpyDeviceToHost, stream));
        FDMB_CUDA(cudaStreamSynchronize(stream));
        return FDMB_OK;
    }
    if (!s_up) FDMB_CUDA(cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking));
    if (!s_dn) FDMB_CUDA(cudaStreamCreateWithFlags(&s_dn, cudaStreamNonBlocking));
    if (!ev_fork) FDMB_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    for (int c = 0; c < nch; c++) {
        if (!ev_chunk[c]) FDMB_CUDA(cudaEventCreateWithFlags(&ev_chunk[c], cudaEventDisableTiming));
        if (!ev_done[c]) FDMB_CUDA(cudaEventCreateWithFlags(&ev_done[c], cudaEventDisableTiming));
    }
    auto drain = [&]() { cudaStreamSynchronize(s_up); cudaStreamSynchronize(stream); cudaStreamSynchronize(s_dn); };
    auto bound = [&](int c) { return c >= nch ? nz : (int)((long long)c * nz / nch) & ~1; };
    const size_t pl = sizeof(double) * (size_t)nx * ny;
    int rc = FDMB_OK;
    cudaError_t e = cudaEventRecord(ev_fork, stream);                 // earlier work on the handle's stream owns d_rhs / d_ans
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s_up, ev_fork, 0);
    for (int c = 0; c < nch && e == cudaSuccess && !rc; c++) {
        const int z0 = bound(c), z1 = bound(c + 1);
        e = cudaMemcpyAsync(d_rhs + (size_t)z0 * nx * ny, rhs + (size_t)z0 * nx * ny, pl * (z1 - z0), cudaMemcpyHostToDevice, s_up);
        if (e == cudaSuccess) e = cudaEventRecord(ev_chunk[c], s_up);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, ev_chunk[c], 0);
        if (e == cudaSuccess) rc = sweeps(d_ans, d_rhs, stream, 1, z0, z1 - z0);
    }
    if (e == cudaSuccess && !rc) rc = sweeps(d_ans, d_rhs, stream, 2, 0, nz);
    for (int c = 0; c < nch && e == cudaSuccess && !rc; c++) {
        const int z0 = bound(c), z1 = bound(c + 1);
        rc = sweeps(d_ans, d_rhs, stream, 4, z0, z1 - z0);
        if (!rc) e = cudaEventRecord(ev_done[c], stream);
        if (e == cudaSuccess && !rc) e = cudaStreamWaitEvent(s_dn, ev_done[c], 0);
        if (e == cudaSuccess && !rc)
            e = cudaMemcpyAsync(ans + (size_t)z0 * nx * ny, d_ans + (size_t)z0 * nx * ny, pl * (z1 - z0), cudaMemcpyDeviceToHost, s_dn);
    }
    drain();      // every exit waits for the copies: they read and write the caller's arrays
    if (e != cudaSuccess) { set_error("LaplCube solve (streamed): %s", cudaGetErrorString(e)); return FDMB_ERR_CUDA; }
    return rc;
}

// `count` independent solves with host arrays, software-pipelined over two staging pairs: the upload of solve i+1
// and the download of solve i-1 run on their own streams while solve i computes (PCIe is full duplex, so a long
// batch costs max(upload, download, compute) per solve instead of their sum).  Same kernels, same results as
// `count` calls of solve_host.
int fdmb_lapl_cube::solve_batch(int count, double* const* ans, const double* const* rhs)
{
    if (nranks > 1) { set_error("LaplCube: solve_batch is for single-GPU handles"); return FDMB_ERR_INVALID; }
    const size_t bytes = sizeof(double) * (size_t)nx * ny * nz;
    if (!d_rhs) FDMB_CUDA(cudaMalloc(&d_rhs, bytes));
    if (!d_ans) FDMB_CUDA(cudaMalloc(&d_ans, bytes));
    if (count > 1) {
        if (!d_rhs1) FDMB_CUDA(cudaMalloc(&d_rhs1, bytes));
        if (!d_ans1) FDMB_CUDA(cudaMalloc(&d_ans1, bytes));
    }
    if (!s_up) FDMB_CUDA(cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking));
    if (!s_dn) FDMB_CUDA(cudaStreamCreateWithFlags(&s_dn, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
        if (!ev_up[b]) FDMB_CUDA(cudaEventCreateWithFlags(&ev_up[b], cudaEventDisableTiming));
        if (!ev_cmp[b]) FDMB_CUDA(cudaEventCreateWithFlags(&ev_cmp[b], cudaEventDisableTiming));
        if (!ev_dn[b]) FDMB_CUDA(cudaEventCreateWithFlags(&ev_dn[b], cudaEventDisableTiming));
    }
    double* dr[2] = {d_rhs, d_rhs1};
    double* da[2] = {d_ans, d_ans1};
    // Every exit, the failing ones included, first waits for the three streams: the copies read and write the
    // CALLER's host arrays, which must not be in flight once this function has returned.
    auto drain = [&]() { cudaStreamSynchronize(s_dn); cudaStreamSynchronize(stream); cudaStreamSynchronize(s_up); };
#define FDMB_BATCH(call)                                                                                       \
    do {                                                                                                       \
        cudaError_t e__ = (call);                                                                              \
        if (e__ != cudaSuccess) {                                                                              \
            drain();                                                                                           \
            set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__));           \
            return FDMB_ERR_CUDA;                                                                              \
        }                                                                                                      \
    } while (0)
    // earlier work on the handle's stream (a previous solve_host / solve_device) may still use the staging buffers;
    // a previous fdmb_lapl_cube_solve_device on a CALLER's stream uses d_work and is the caller's to order
    FDMB_BATCH(cudaStreamSynchronize(stream));
    for (int i = 0; i < count; i++) {
        const int b = i & 1;
        if (i >= 2) FDMB_BATCH(cudaStreamWaitEvent(s_up, ev_cmp[b], 0));      // solve i-2 has consumed dr[b]
        FDMB_BATCH(cudaMemcpyAsync(dr[b], rhs[i], bytes, cudaMemcpyHostToDevice, s_up));
        FDMB_BATCH(cudaEventRecord(ev_up[b], s_up));
        FDMB_BATCH(cudaStreamWaitEvent(stream, ev_up[b], 0));
        if (i >= 2) FDMB_BATCH(cudaStreamWaitEvent(stream, ev_dn[b], 0));     // download i-2 has drained da[b]
        int rc = solve_device(da[b], dr[b], stream);
        if (rc) { drain(); return rc; }
        FDMB_BATCH(cudaEventRecord(ev_cmp[b], stream));
        FDMB_BATCH(cudaStreamWaitEvent(s_dn, ev_cmp[b], 0));
        FDMB_BATCH(cudaMemcpyAsync(ans[i], da[b], bytes, cudaMemcpyDeviceToHost, s_dn));
        FDMB_BATCH(cudaEventRecord(ev_dn[b], s_dn));
    }
    FDMB_BATCH(cudaStreamSynchronize(s_dn));
    FDMB_BATCH(cudaStreamSynchronize(stream));
    FDMB_BATCH(cudaStreamSynchronize(s_up));
#undef FDMB_BATCH
    return FDMB_OK;
}

extern "C" {

int fdmb_lapl_cube_create(fdmb_lapl_cube** out, double dx, double dy, double dz, double lx, double ly, double lz,
                          int nx, int ny, int nz, int periodic)
{
    if (!out) { set_error("null handle pointer"); return FDMB_ERR_INVALID; }
    *out = nullptr;
    auto* h = new (std::nothrow) fdmb_lapl_cube();
    if (!h) { set_error("out of host memory"); return FDMB_ERR_NOMEM; }
    h->dx = dx; h->dy = dy; h->dz = dz; h->lx = lx; h->ly = ly; h->lz = lz;
    h->nx = nx; h->ny = ny; h->nz = nz; h->periodic = periodic ? 1 : 0;
    int rc = h->init();
    if (rc) { delete h; return rc; }
    *out = h;
    return FDMB_OK;
}

int fdmb_slab_range(int n, int periodic, int nranks, int rank, int* first, int* count)
{
    const int N = periodic ? n : n + 1;
    if (n < 1 || nranks < 1 || rank < 0 || rank >= nranks || !first || !count || !is_pow2(N) || !is_pow2(nranks) ||
        N / nranks < 2) {
        set_error("fdmb_slab_range: n=%d (transform length %d) cannot be split over %d ranks", n, N, nranks);
        return FDMB_ERR_INVALID;
    }
    slab_range(n, periodic ? 1 : 0, nranks, rank, first, count);
    return FDMB_OK;
}

int fdmb_lapl_cube_create_sharded(fdmb_lapl_cube** out, double dx, double dy, double dz, double lx, double ly, double lz,
                                  int nx, int ny, int nz, int periodic, int rank, int nranks)
{
    if (!out) { set_error("null handle pointer"); return FDMB_ERR_INVALID; }
    *out = nullptr;
    auto* h = new (std::nothrow) fdmb_lapl_cube();
    if (!h) { set_error("out of host memory"); return FDMB_ERR_NOMEM; }
    h->dx = dx; h->dy = dy; h->dz = dz; h->lx = lx; h->ly = ly; h->lz = lz;
    h->nx = nx; h->ny = ny; h->nz = nz; h->periodic = periodic ? 1 : 0;
    h->rank = rank; h->nranks = nranks;
    int rc = h->init();
    if (rc) { delete h; return rc; }
    *out = h;
    return FDMB_OK;
}

int fdmb_lapl_cube_local_slab(fdmb_lapl_cube* h, int* z_first, int* nz_local)
{
    if (!h || !z_first || !nz_local) { set_error("null argument"); return FDMB_ERR_INVALID; }
    *z_first = h->nranks > 1 ? h->z_first : 0;
    *nz_local = h->nranks > 1 ? h->nzl : h->nz;
    return FDMB_OK;
}

int fdmb_lapl_cube_export_ipc(fdmb_lapl_cube* h, void* handle)
{
    if (!h || !handle || h->nranks < 2) { set_error("export_ipc needs a sharded handle"); return FDMB_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == FDMB_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t ih;
    FDMB_CUDA(cudaIpcGetMemHandle(&ih, h->mg_block));
    memcpy(handle, &ih, sizeof(ih));
    return FDMB_OK;
}

int fdmb_lapl_cube_attach_ipc(fdmb_lapl_cube* h, const void* handles)
{
    if (!h || !handles || h->nranks < 2) { set_error("attach_ipc needs a sharded handle"); return FDMB_ERR_INVALID; }
    void* bases[FDMB_MAX_RANKS] = {};
    for (int q = 0; q < h->nranks; q++) {
        if (q == h->rank) continue;
        cudaIpcMemHandle_t ih;
        memcpy(&ih, (const char*)handles + (size_t)q * FDMB_IPC_HANDLE_BYTES, sizeof(ih));
        cudaError_t e = cudaIpcOpenMemHandle(&bases[q], ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("cudaIpcOpenMemHandle for rank %d failed: %s", q, cudaGetErrorString(e));
            return FDMB_ERR_COMM;
        }
        h->peer_ipc[q] = true;
    }
    return h->attach(bases);
}

int fdmb_lapl_cube_attach_local(fdmb_lapl_cube* h, fdmb_lapl_cube* const* all)
{
    if (!h || !all || h->nranks < 2) { set_error("attach_local needs a sharded handle"); return FDMB_ERR_INVALID; }
    void* bases[FDMB_MAX_RANKS] = {};
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
        bases[q] = all[q]->mg_block;
    }
    FDMB_CUDA(cudaSetDevice(cur));
    return h->attach(bases);
}

int fdmb_lapl_cube_solve(fdmb_lapl_cube* h, double* ans, const double* rhs)
{
    if (!h || !ans || !rhs) { set_error("null argument"); return FDMB_ERR_INVALID; }
    return h->solve_host(ans, rhs);
}

int fdmb_lapl_cube_solve_batch(fdmb_lapl_cube* h, int count, double* const* ans, const double* const* rhs)
{
    if (!h || count < 0 || (count > 0 && (!ans || !rhs))) { set_error("null argument"); return FDMB_ERR_INVALID; }
    for (int i = 0; i < count; i++)
        if (!ans[i] || !rhs[i]) { set_error("solve_batch: null array %d", i); return FDMB_ERR_INVALID; }
    return h->solve_batch(count, ans, rhs);
}

int fdmb_lapl_cube_solve_device(fdmb_lapl_cube* h, double* d_ans, const double* d_rhs, void* stream)
{
    if (!h || !d_ans || !d_rhs) { set_error("null argument"); return FDMB_ERR_INVALID; }
    if (h->nranks > 1) {   // several ranks may share one process: launch on the handle's device
        int cur = 0;
        FDMB_CUDA(cudaGetDevice(&cur));
        if (cur != h->device) FDMB_CUDA(cudaSetDevice(h->device));
        int rc = h->solve_device(d_ans, d_rhs, stream ? (cudaStream_t)stream : h->stream);
        if (cur != h->device) cudaSetDevice(cur);
        return rc;
    }
    return h->solve_device(d_ans, d_rhs, stream ? (cudaStream_t)stream : h->stream);
}

int fdmb_lapl_cube_destroy(fdmb_lapl_cube* h)
{
    delete h;
    return FDMB_OK;
}

}  // extern "C"
