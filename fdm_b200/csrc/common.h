// Host-side helpers shared by every translation unit of libfdm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <atomic>
#include <cstdarg>
#include <string>
#include <utility>

#include "../../include/fdm_b200.h"
#include "xform.cuh"
#include "pdl.cuh"

namespace fdmb {

void set_error(const char* fmt, ...);
const char* get_error();

#define FDMB_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t err__ = (call);                                                       \
        if (err__ != cudaSuccess) {                                                       \
            ::fdmb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,        \
                              cudaGetErrorString(err__));                                 \
            return FDMB_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define FDMB_CHECK_LAUNCH()                                                               \
    do {                                                                                  \
        cudaError_t err__ = cudaGetLastError();                                           \
        if (err__ != cudaSuccess) {                                                       \
            ::fdmb::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,    \
                              cudaGetErrorString(err__));                                 \
            return FDMB_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

// Device-resident twiddle tables for transform length N (per device, cached).
//   SN[j] = sin(pi j / N),                 j = 0..N/2
//   WM[t] = (cos(2 pi t/M), -sin(2 pi t/M)), t = 0..M-1,  M = N/2
struct Tables {
    const double* SN;
    const cd* WM;
};
int get_tables(int N, Tables* out);

inline bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
// transform lengths for which kernels are instantiated
inline bool supported_N(int N) { return is_pow2(N) && N >= 4 && N <= 2048; }

bool pipe_enabled();     // FDMB_PIPE=0 selects the synchronous sweep kernels (A/B measurements)
int device_sm_count();
int current_device_slot();   // cudaGetDevice() clamped to [0, 63]

// counts kernel launches issued by this library (bench.py reports it as gpu_launches)
extern std::atomic<unsigned long long> g_launch_count;

// Optional per-launch CUDA-event timing (fdmb_profile_begin/end).  Every launch site
// wraps its kernel in a LaunchScope; outside profiling it only bumps the counter.
struct LaunchScope {
    const char* tag;
    cudaStream_t st;
    int slot;
    LaunchScope(const char* tag, cudaStream_t st);
    ~LaunchScope();
};

inline bool stream_capturing(cudaStream_t st)
{
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
    return cs != cudaStreamCaptureStatusNone;
}
bool profiling_on();      // fdmb_profile_begin() is active: launches are timed one by one, graphs are bypassed
bool graphs_enabled();    // FDMB_GRAPH=0 launches every kernel individually (A/B measurements)

// One captured-and-replayed launch sequence (a solve, a time step) of a single-GPU handle.  The small grids of the
// reference's own runs (31^3, 127^3, 128x127x128) are launch-bound: a step is 11 kernels of a few microseconds each,
// and the host-side launch cost shows between them.  The first call on a stream runs the body directly (it also
// warms every lazily initialised attribute); the second captures it (thread-local capture mode) and from then on the
// instantiated graph is replayed.  The pointers a body bakes in must therefore be fixed per key (the handles'
// own field and work arrays are; a solve keys on its caller's ans / rhs pair).
struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    unsigned long long launches = 0;     // kernels per replay (fdmb_launch_count keeps counting them)
    const void* key0 = nullptr;
    const void* key1 = nullptr;
    int calls = 0;
    bool failed = false;
    ~StepGraph() { if (exec) cudaGraphExecDestroy(exec); }
    void reset() { if (exec) cudaGraphExecDestroy(exec); exec = nullptr; calls = 0; }

    template <typename Body> int run(cudaStream_t st, const void* k0, const void* k1, Body body)
    {
        if (!graphs_enabled() || profiling_on() || failed || st == nullptr || st == cudaStreamLegacy ||
            st == cudaStreamPerThread)
            return body();
        if (k0 != key0 || k1 != key1) { reset(); key0 = k0; key1 = k1; }
        if (!exec) {
            if (calls++ == 0) return body();                      // plain first call: lazy initialisation happens here
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return body(); }
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); failed = true; return body(); }
            const unsigned long long c0 = g_launch_count.load();
            const int rc = body();
            cudaGraph_t graph = nullptr;
            const cudaError_t e = cudaStreamEndCapture(st, &graph);
            launches = g_launch_count.load() - c0;
            g_launch_count.fetch_sub(launches);                   // nothing has run yet
            if (rc != 0 || e != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                failed = true;
                return rc != 0 ? rc : body();
            }
            const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ei != cudaSuccess) { cudaGetLastError(); exec = nullptr; failed = true; return body(); }
        }
        FDMB_CUDA(cudaGraphLaunch(exec, st));
        g_launch_count.fetch_add(launches);
        return 0;
    }
};

}  // namespace fdmb
