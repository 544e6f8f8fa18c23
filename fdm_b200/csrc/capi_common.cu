// Library-wide pieces of the C ABI: error string, device tables, raw memory helpers,
// batched 1-D transforms (fdm::FFT<double>::{sFFT,pFFT_1,pFFT}).
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "common.h"
#include "lapl_cube.h"
#include "pipe.cuh"

namespace fdmb {

static thread_local std::string g_error;
std::atomic<unsigned long long> g_launch_count{0};

void set_error(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}
const char* get_error() { return g_error.c_str(); }

// Per-launch profile records.  Several host threads may drive handles (one per device) at the same time: the
// record list is guarded by a mutex, and every record remembers the device its events were created on.
struct ProfRec { const char* tag; cudaEvent_t e0, e1; int dev; };
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mutex;
static std::vector<ProfRec> g_prof;

LaunchScope::LaunchScope(const char* tag_, cudaStream_t st_) : tag(tag_), st(st_), slot(-1)
{
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    if (g_prof_on.load(std::memory_order_relaxed)) {
        ProfRec r{tag, nullptr, nullptr, 0};
        cudaGetDevice(&r.dev);
        if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) { cudaGetLastError(); return; }
        cudaEventRecord(r.e0, st);
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        slot = (int)g_prof.size();
        g_prof.push_back(r);
    }
}
LaunchScope::~LaunchScope()
{
    if (slot < 0) return;
    cudaEvent_t e1;
    {
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        if (slot >= (int)g_prof.size()) return;      // profile_end() ran in between
        e1 = g_prof[slot].e1;
    }
    cudaEventRecord(e1, st);
}

static std::mutex g_tab_mutex;
static std::map<std::pair<int, int>, Tables> g_tables;

int get_tables(int N, Tables* out)
{
    int dev = 0;
    FDMB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    auto it = g_tables.find({dev, N});
    if (it != g_tables.end()) { *out = it->second; return FDMB_OK; }
    const int M = N / 2;
    std::vector<double> sn(N / 2 + 1);
    std::vector<cd> wm(M);
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int j = 0; j <= N / 2; j++) sn[j] = (double)sinl(pi * j / N);
    sn[N / 2] = 1.0;
    for (int t = 0; t < M; t++) {
        // exact octant symmetries keep the table accurate to the last bit
        long double ang = 2 * pi * t / M;
        wm[t].x = (double)cosl(ang);
        wm[t].y = (double)(-sinl(ang));
    }
    if (M >= 4) { wm[M / 4].x = 0.0; wm[M / 4].y = -1.0; wm[3 * M / 4].x = 0.0; wm[3 * M / 4].y = 1.0; }
    if (M >= 2) { wm[M / 2].x = -1.0; wm[M / 2].y = 0.0; }
    double* d_sn = nullptr;
    cd* d_wm = nullptr;
    FDMB_CUDA(cudaMalloc(&d_sn, sizeof(double) * sn.size()));
    FDMB_CUDA(cudaMalloc(&d_wm, sizeof(cd) * wm.size()));
    FDMB_CUDA(cudaMemcpy(d_sn, sn.data(), sizeof(double) * sn.size(), cudaMemcpyHostToDevice));
    FDMB_CUDA(cudaMemcpy(d_wm, wm.data(), sizeof(cd) * wm.size(), cudaMemcpyHostToDevice));
    Tables t{d_sn, d_wm};
    g_tables[{dev, N}] = t;
    *out = t;
    return FDMB_OK;
}

// single-precision tables of the fp32 instantiations: the same values rounded once from long double
struct TablesF { const float* SN; const cx<float>* WM; };
static std::map<std::pair<int, int>, TablesF> g_tables_f32;

int get_tables_f32(int N, TablesF* out)
{
    int dev = 0;
    FDMB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    auto it = g_tables_f32.find({dev, N});
    if (it != g_tables_f32.end()) { *out = it->second; return FDMB_OK; }
    const int M = N / 2;
    std::vector<float> sn(N / 2 + 1);
    std::vector<cx<float>> wm(M);
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int j = 0; j <= N / 2; j++) sn[j] = (float)sinl(pi * j / N);
    sn[N / 2] = 1.0f;
    for (int t = 0; t < M; t++) { wm[t].x = (float)cosl(2 * pi * t / M); wm[t].y = (float)(-sinl(2 * pi * t / M)); }
    if (M >= 4) { wm[M / 4] = {0.0f, -1.0f}; wm[3 * M / 4] = {0.0f, 1.0f}; }
    if (M >= 2) wm[M / 2] = {-1.0f, 0.0f};
    float* d_sn = nullptr;
    cx<float>* d_wm = nullptr;
    FDMB_CUDA(cudaMalloc(&d_sn, sizeof(float) * sn.size()));
    FDMB_CUDA(cudaMalloc(&d_wm, sizeof(cx<float>) * wm.size()));
    FDMB_CUDA(cudaMemcpy(d_sn, sn.data(), sizeof(float) * sn.size(), cudaMemcpyHostToDevice));
    FDMB_CUDA(cudaMemcpy(d_wm, wm.data(), sizeof(cx<float>) * wm.size(), cudaMemcpyHostToDevice));
    TablesF t{d_sn, d_wm};
    g_tables_f32[{dev, N}] = t;
    *out = t;
    return FDMB_OK;
}

int device_sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int current_device_slot()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    return dev;
}

bool pipe_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FDMB_PIPE");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

bool profiling_on() { return g_prof_on.load(std::memory_order_relaxed); }

int pdl_mode()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("FDMB_PDL"); v = e ? atoi(e) : 1; if (v < 0 || v > 2) v = 1; }
    return v;
}
bool graphs_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FDMB_GRAPH");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

bool ring_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FDMB_RING");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_tiled(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                        const cuuint32_t* box)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        FDMB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return FDMB_ERR_CUDA;
        }
        fn = (EncodeTiledFn)p;
    }
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with code %d (rank %d dims %llu,%llu,%llu,%llu strides %llu,%llu,%llu "
                  "box %u,%u,%u,%u)", (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                  (unsigned long long)dims[2], rank > 3 ? (unsigned long long)dims[3] : 0ull,
                  (unsigned long long)strides[0], (unsigned long long)strides[1], rank > 3 ? (unsigned long long)strides[2] : 0ull,
                  box[0], box[1], box[2], rank > 3 ? box[3] : 0u);
        return FDMB_ERR_CUDA;
    }
    return FDMB_OK;
}

int make_tensor_map_3d(CUtensorMap* tm, const void* base, unsigned long long e0, unsigned long long e1,
                       unsigned long long e2, unsigned long long s1, unsigned long long s2, unsigned b0, unsigned b1,
                       unsigned b2)
{
    cuuint64_t dims[3] = {e0, e1, e2};
    cuuint64_t strides[2] = {s1, s2};
    cuuint32_t box[3] = {b0, b1, b2};
    return encode_tiled(tm, base, 3, dims, strides, box);
}

int make_cols_maps_blocked(ColsMaps* m, const void* base, int N, bool pair_axis, int blog, unsigned long long e_x,
                           unsigned long long n_lohi, unsigned long long n_mid, unsigned long long s_lo,
                           unsigned long long s_mid, unsigned long long s_hi, unsigned B)
{
    const unsigned LO = 1u << blog;
    const unsigned long long n_hi = (n_lohi + LO - 1) / LO;
    const unsigned long long n = pair_axis ? n_lohi : n_mid;     // entries along the transform axis
    const char* b = static_cast<const char*>(base);
    int rc;
    m->blog = blog;
    m->planar = ((int)n == N - 1 && N >= 32 && LO >= 2);
    if (pair_axis) {
        m->taxis = 3;
        m->boxhi = n_hi < 256 ? (int)n_hi : 256;
        m->boxrows = m->boxhi * (int)LO;
        m->nchunk = (int)((n_hi + m->boxhi - 1) / m->boxhi);
        {
            cuuint64_t dims[4] = {e_x, LO, n_mid, n_hi};
            cuuint64_t strides[3] = {s_lo, s_mid, s_hi};
            cuuint32_t box[4] = {B, LO, 1, (cuuint32_t)m->boxhi};
            if ((rc = encode_tiled(&m->nat, base, 4, dims, strides, box))) return rc;
        }
        if (m->planar) {
            m->boxhi_p = m->boxhi;
            m->boxrows_p = m->boxhi * (int)(LO / 2);
            m->nchunk_p = m->nchunk;
            cuuint64_t dims[4] = {e_x, LO / 2, n_mid, n_hi};
            cuuint64_t strides[3] = {2 * s_lo, s_mid, s_hi};
            cuuint32_t box[4] = {B, LO / 2, 1, (cuuint32_t)m->boxhi};
            if ((rc = encode_tiled(&m->podd, b, 4, dims, strides, box))) return rc;
            if ((rc = encode_tiled(&m->peven, b + s_lo, 4, dims, strides, box))) return rc;
        }
    } else {
        m->taxis = 4;
        m->boxrows = n < 256 ? (int)n : 256;
        m->nchunk = (int)((n + m->boxrows - 1) / m->boxrows);
        {
            cuuint64_t dims[4] = {e_x, LO, n_mid, n_hi};
            cuuint64_t strides[3] = {s_lo, s_mid, s_hi};
            cuuint32_t box[4] = {B, 1, (cuuint32_t)m->boxrows, 1};
            if ((rc = encode_tiled(&m->nat, base, 4, dims, strides, box))) return rc;
        }
        m->planar = ((int)n == N - 1 && N >= 32);
        if (m->planar) {
            const int M = N / 2;
            m->boxrows_p = M < 256 ? M : 256;
            m->nchunk_p = M / m->boxrows_p;
            cuuint64_t dO[4] = {e_x, LO, (n + 1) / 2, n_hi}, dE[4] = {e_x, LO, n / 2, n_hi};
            cuuint64_t strides[3] = {s_lo, 2 * s_mid, s_hi};
            cuuint32_t box[4] = {B, 1, (cuuint32_t)m->boxrows_p, 1};
            if ((rc = encode_tiled(&m->podd, b, 4, dO, strides, box))) return rc;
            if ((rc = encode_tiled(&m->peven, b + s_mid, 4, dE, strides, box))) return rc;
        }
    }
    return FDMB_OK;
}

int make_cols_maps(ColsMaps* m, const void* base, int N, int taxis, unsigned long long e0, unsigned long long e1,
                   unsigned long long e2, unsigned long long s1, unsigned long long s2, unsigned B)
{
    const unsigned long long n = taxis == 1 ? e1 : e2;       // entries along the transform axis
    int rc;
    m->boxrows = n < 256 ? (int)n : 256;
    m->nchunk = (int)((n + m->boxrows - 1) / m->boxrows);
    if (taxis == 1) rc = make_tensor_map_3d(&m->nat, base, e0, e1, e2, s1, s2, B, m->boxrows, 1);
    else rc = make_tensor_map_3d(&m->nat, base, e0, e1, e2, s1, s2, B, 1, m->boxrows);
    if (rc) return rc;
    m->planar = false;
    if ((int)n == N - 1 && N >= 32) {
        const int M = N / 2;
        m->boxrows_p = M < 256 ? M : 256;
        m->nchunk_p = M / m->boxrows_p;
        const unsigned long long no = (n + 1) / 2, ne = n / 2;     // even rows 0,2,.. / odd rows 1,3,..
        const char* b1 = static_cast<const char*>(base) + (taxis == 1 ? s1 : s2);
        if (taxis == 1) {
            if ((rc = make_tensor_map_3d(&m->podd, base, e0, no, e2, 2 * s1, s2, B, m->boxrows_p, 1))) return rc;
            if ((rc = make_tensor_map_3d(&m->peven, b1, e0, ne, e2, 2 * s1, s2, B, m->boxrows_p, 1))) return rc;
        } else {
            if ((rc = make_tensor_map_3d(&m->podd, base, e0, e1, no, s1, 2 * s2, B, 1, m->boxrows_p))) return rc;
            if ((rc = make_tensor_map_3d(&m->peven, b1, e0, e1, ne, s1, 2 * s2, B, 1, m->boxrows_p))) return rc;
        }
        m->planar = true;
    }
    return FDMB_OK;
}

}  // namespace fdmb

using namespace fdmb;

extern "C" {

const char* fdmb_last_error(void) { return get_error(); }
int fdmb_version(void) { return 100; }

int fdmb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int fdmb_set_device(int device)
{
    FDMB_CUDA(cudaSetDevice(device));
    return FDMB_OK;
}

unsigned long long fdmb_launch_count(void) { return g_launch_count.load(); }

int fdmb_malloc(void** dptr, unsigned long long bytes)
{
    FDMB_CUDA(cudaMalloc(dptr, bytes));
    return FDMB_OK;
}
int fdmb_free(void* dptr)
{
    FDMB_CUDA(cudaFree(dptr));
    return FDMB_OK;
}
int fdmb_memcpy_h2d(void* dst, const void* src, unsigned long long bytes)
{
    FDMB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return FDMB_OK;
}
int fdmb_memcpy_d2h(void* dst, const void* src, unsigned long long bytes)
{
    FDMB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return FDMB_OK;
}
int fdmb_device_synchronize(void)
{
    FDMB_CUDA(cudaDeviceSynchronize());
    return FDMB_OK;
}

int fdmb_profile_begin(void)
{
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof.clear();
    g_prof_on = true;
    return FDMB_OK;
}

int fdmb_profile_end(char* buf, int buflen)
{
    g_prof_on = false;
    int cur = 0;
    FDMB_CUDA(cudaGetDevice(&cur));
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    std::map<std::string, std::pair<int, double>> agg;
    std::vector<std::string> order;
    int failed = 0;
    for (auto& r : g_prof) {
        float ms = 0;
        // events belong to the device they were created on: measure there, skip (and count) what cannot be read
        cudaError_t e = cudaSetDevice(r.dev);
        if (e == cudaSuccess) e = cudaEventSynchronize(r.e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.e0, r.e1);
        if (e == cudaSuccess) {
            if (!agg.count(r.tag)) order.push_back(r.tag);
            agg[r.tag].first++;
            agg[r.tag].second += ms;
        } else {
            cudaGetLastError();
            failed++;
        }
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    cudaSetDevice(cur);
    g_prof.clear();
    if (failed) set_error("fdmb_profile_end: %d launch records could not be timed and were skipped", failed);
    std::string out;
    for (auto& t : order) {
        char line[256];
        snprintf(line, sizeof(line), "%s %d %.6f\n", t.c_str(), agg[t].first, agg[t].second);
        out += line;
    }
    if (buf && buflen > 0) {
        snprintf(buf, buflen, "%s", out.c_str());
    }
    return FDMB_OK;
}

// impl: 0 = what the solvers use for this length (the persistent bulk-copy-fed sweep for N >= 32, the plain kernel below
// that), 1 = plain kernel, 2 = persistent sweep (error if N < 32)
int fdmb_fft_batch_impl(int kind, int N, long long batch, double dx, const double* in, double* out, int impl)
{
    if (kind < 0 || kind > 3 || !supported_N(N) || batch < 0 || !in || !out || impl < 0 || impl > 2) {
        set_error("fdmb_fft_batch: kind must be 0..3 and N a power of two in [4,2048] (got kind=%d N=%d impl=%d)", kind, N, impl);
        return FDMB_ERR_INVALID;
    }
    const bool can_pipe = kind != XF_DCT && rows_pipe_supported_N(N);
    if (impl == 2 && !can_pipe) {
        set_error("fdmb_fft_batch: the persistent sweep needs 32 <= N <= 1024 and kind 0..2 (got kind=%d N=%d)", kind, N);
        return FDMB_ERR_INVALID;
    }
    if (batch == 0) return FDMB_OK;
    const int nvalid = kind == XF_DST ? N - 1 : (kind == XF_DCT ? N + 1 : N);
    const size_t bytes = sizeof(double) * (size_t)batch * nvalid;
    Tables t;
    int rc = get_tables(N, &t);
    if (rc) return rc;
    double *d_in = nullptr, *d_out = nullptr;
    FDMB_CUDA(cudaMalloc(&d_in, bytes));
    if (cudaMalloc(&d_out, bytes) != cudaSuccess) { cudaFree(d_in); set_error("fdmb_fft_batch: out of device memory"); return FDMB_ERR_NOMEM; }
    cudaError_t e = cudaMemcpy(d_in, in, bytes, cudaMemcpyHostToDevice);
    const bool use_pipe = can_pipe && (impl == 2 || (impl == 0 && pipe_enabled()));
    if (e == cudaSuccess) {
        if (use_pipe) {
            RowsPipeArgs p{};
            p.in = d_in; p.out = d_out; p.nrows = batch; p.nvalid = nvalid; p.in_pitch = p.out_pitch = nvalid;
            p.scale = dx; p.SN = t.SN; p.WM = t.WM;
            e = launch_rows_pipe(N, kind, p, 0, "fft_batch");
        } else {
            RowsArgs r{};
            r.in = d_in; r.out = d_out; r.nrows = batch; r.nvalid = nvalid;
            r.in_pitch = r.out_pitch = nvalid; r.scale = dx; r.SN = t.SN; r.WM = t.WM;
            e = launch_rows(N, kind, r, 0, "fft_batch");
        }
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, bytes, cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) { set_error("fdmb_fft_batch: %s", cudaGetErrorString(e)); return FDMB_ERR_CUDA; }
    return FDMB_OK;
}

int fdmb_fft_batch(int kind, int N, long long batch, double dx, const double* in, double* out)
{
    return fdmb_fft_batch_impl(kind, N, batch, dx, in, out, 0);
}

}  // extern "C"
