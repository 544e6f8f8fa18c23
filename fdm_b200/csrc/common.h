// Host-side helpers shared by every translation unit of libfdm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <atomic>
#include <cstdarg>
#include <string>

#include "../../include/fdm_b200.h"
#include "xform.cuh"

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

}  // namespace fdmb
