// Host emulation of fdm_b200/csrc/xform.cuh: the device tile transforms are compiled
// for the CPU with __syncthreads() mapped onto a std::barrier over G host threads
// (one emulated column at a time).  Lets the CPU test-suite check the transform
// algebra for every instantiated length without a GPU.  Test infrastructure only.
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <functional>

#define FDMB_HOST_EMUL 1
#define __device__
#define __forceinline__ inline
static thread_local std::barrier<>* tl_barrier = nullptr;
static inline void __syncthreads() { tl_barrier->arrive_and_wait(); }

#include "../../fdm_b200/csrc/xform.cuh"

using namespace fdmb;

static void make_tables(int N, std::vector<double>& sn, std::vector<cd>& wm)
{
    const int M = N / 2;
    sn.resize(N / 2 + 1); wm.resize(M);
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int j = 0; j <= N / 2; j++) sn[j] = (double)sinl(pi * j / N);
    sn[N / 2] = 1.0;
    for (int t = 0; t < M; t++) { wm[t].x = (double)cosl(2 * pi * t / M); wm[t].y = (double)(-sinl(2 * pi * t / M)); }
    if (M >= 4) { wm[M / 4] = {0.0, -1.0}; wm[3 * M / 4] = {0.0, 1.0}; }
    if (M >= 2) wm[M / 2] = {-1.0, 0.0};
}

template <int N, int KIND>
static void run_one(double* data, int sj, double scale)
{
    constexpr int G = Plan<N>::G;
    std::vector<double> sn; std::vector<cd> wm;
    make_tables(N, sn, wm);
    std::vector<double> scr((G + G / 8 + 2));
    std::barrier<> bar(G);
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++)
        th.emplace_back([&, g] {
            tl_barrier = &bar;
            xform_tile<N, G, KIND>(data, sj, g, scale, sn.data(), wm.data(), scr.data(), 1);
        });
    for (auto& t : th) t.join();
}

// the same tile transforms in single precision (the fp32 instantiations: lapl_cube_f32.cu)
template <int N, int KIND>
static void run_one_f32(float* data, int sj, float scale)
{
    constexpr int G = Plan<N>::G;
    std::vector<double> snd; std::vector<cd> wmd;
    make_tables(N, snd, wmd);
    std::vector<float> sn(snd.begin(), snd.end());
    std::vector<cx<float>> wm(wmd.size());
    for (size_t i = 0; i < wmd.size(); i++) wm[i] = {(float)wmd[i].x, (float)wmd[i].y};
    std::vector<float> scr((G + G / 8 + 2));
    std::barrier<> bar(G);
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++)
        th.emplace_back([&, g] {
            tl_barrier = &bar;
            xform_tile<N, G, KIND>(data, sj, g, scale, (const float*)sn.data(), (const cx<float>*)wm.data(), scr.data(), 1);
        });
    for (auto& t : th) t.join();
}

extern "C" int emul_xform_f32(int kind, int N, float* data, int sj, float scale)
{
#define X(NN)                                                          \
    case NN:                                                           \
        if (kind == 0) run_one_f32<NN, XF_DST>(data, sj, scale);       \
        else if (kind == 1) run_one_f32<NN, XF_PFWD>(data, sj, scale); \
        else if (kind == 3) run_one_f32<NN, XF_DCT>(data, sj, scale);  \
        else run_one_f32<NN, XF_PINV>(data, sj, scale);                \
        return 0;
    switch (N) { X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) }
#undef X
    return -1;
}

// fused DST (dst_tile_fused) on the planar tile layout: mode 0 = OutTile, 1 = OutGlobal-like policy
// writing to a separate array, 2 = PREFOLD (fold applied by the caller, like k_rows_pipe's first touch).
// swz selects the 8-column-tile swizzles (three-pass plans only).  data: natural slots 0..N-1, stride sj.
struct OutSep {
    double* dst;
    void emit(int j, double v) const { dst[j - 1] = v; }
};

template <int N, int GAP, bool SWZ, int GDIV = 1>
static void run_fused_t(double* data, int sj, double scale, int mode, double* sep)
{
    constexpr int G = Plan<N>::G / GDIV;   // GDIV = 2: several first/middle-pass butterflies per thread
    constexpr int M = N / 2;
    using PL = Planar<N, GAP>;
    std::vector<double> sn; std::vector<cd> wm;
    make_tables(N, sn, wm);
    const double hs = 0.5 * scale;
    std::vector<double> sf(M + 1);
    for (int j = 0; j <= M; j++) sf[j] = hs * sn[j];
    std::vector<double> tile((size_t)PL::ROWS * sj, 777.0);     // garbage in the unused rows
    if (mode == 2) {
        for (int j = 1; j < M; j++) {
            double a = data[j * sj], c = data[(N - j) * sj];
            double y1 = sf[j] * (a + c), y2 = 0.5 * hs * (a - c);
            tile[prefold_row<N, GAP, SWZ>(j) * sj] = y1 + y2; tile[prefold_row<N, GAP, SWZ>(N - j) * sj] = y1 - y2;
        }
        tile[prefold_row<N, GAP, SWZ>(0) * sj] = 0.0; tile[prefold_row<N, GAP, SWZ>(M) * sj] = scale * data[M * sj];
    } else {
        for (int j = 1; j < N; j++) tile[PL::row(j) * sj] = data[j * sj];
    }
    std::vector<double> scr((G + G / 8 + 2));
    std::barrier<> bar(G);
    std::vector<std::thread> th;
    double* t = tile.data();
    for (int g = 0; g < G; g++)
        th.emplace_back([&, g] {
            tl_barrier = &bar;
            if (mode == 0) dst_tile_fused<N, G, GAP, false, SWZ>(t, sj, g, hs, sn.data(), sf.data(), wm.data(), scr.data(), 1, OutTile<N, GAP>{t, sj});
            else if (mode == 1) dst_tile_fused<N, G, GAP, false, SWZ>(t, sj, g, hs, sn.data(), sf.data(), wm.data(), scr.data(), 1, OutSep{sep});
            else dst_tile_fused<N, G, GAP, true, SWZ>(t, sj, g, hs, sn.data(), sf.data(), wm.data(), scr.data(), 1, OutTile<N, GAP>{t, sj});
        });
    for (auto& x : th) x.join();
    if (mode != 1)
        for (int j = 1; j < N; j++) data[j * sj] = tile[PL::row(j) * sj];
}

template <int N>
static void run_fused(double* data, int sj, double scale, int mode, double* sep, int swz)
{
    if constexpr (PlanInfo<N>::NP == 3) {
        if (swz) { run_fused_t<N, 1, true>(data, sj, scale, mode, sep); return; }
    }
    if (mode == 2) run_fused_t<N, 0, false>(data, sj, scale, mode, sep);
    else run_fused_t<N, 1, false>(data, sj, scale, mode, sep);
}

// N = 1024 with 32 threads per sequence (the strided-axis sweeps' configuration, PipeCfg<1024>::GC)
extern "C" int emul_dst_fused_half(double* data, int sj, double scale, int mode, double* sep, int swz)
{
    if (swz) run_fused_t<1024, 1, true, 2>(data, sj, scale, mode, sep);
    else run_fused_t<1024, 1, false, 2>(data, sj, scale, mode, sep);
    return 0;
}

extern "C" int emul_dst_fused(int N, double* data, int sj, double scale, int mode, double* sep, int swz)
{
#define X(NN) case NN: run_fused<NN>(data, sj, scale, mode, sep, swz); return 0;
    switch (N) { X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) }
#undef X
    return -1;
}

// data: N slots with stride sj (slot 0 unused for kind 0; N + 1 slots for kind 3 = cFFT)
extern "C" int emul_xform(int kind, int N, double* data, int sj, double scale)
{
#define X(NN)                                                      \
    case NN:                                                       \
        if (kind == 0) run_one<NN, XF_DST>(data, sj, scale);       \
        else if (kind == 1) run_one<NN, XF_PFWD>(data, sj, scale); \
        else if (kind == 3) run_one<NN, XF_DCT>(data, sj, scale);  \
        else run_one<NN, XF_PINV>(data, sj, scale);                \
        return 0;
    switch (N) { X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) }
#undef X
    return -1;
}
