// Programmatic dependent launch helpers (device + launch side); see the comment below.
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace fdmb {

// Programmatic dependent launch (sm_90+): a kernel launched with the attribute may become resident while its predecessor
// in the stream is still running, and runs its prologue (tables into shared memory, mbarrier initialisation, descriptor
// prefetch) there; pdl_wait() -- griddepcontrol.wait, placed before the first access to data another kernel may have
// produced -- returns when the predecessor has completed and its writes are visible.  Every kernel calls pdl_trigger()
// right AFTER its own wait, so its successor can be scheduled once all of its own blocks are running.  (Triggering first
// thing lets the blocks of the next several kernels pile up on the SMs while the first one is still running; they hold
// shared memory a persistent kernel's own late blocks then wait for: 1023^3 solve + 4 %, r02pdl.)
// In a kernel launched without the attribute both instructions do nothing.
// Measured (r02pdl, on / off / on): NSCube 31^3 step 35.2 -> 29.3 us, LaplCube 127^3 88.4 -> 74.4 us, but 255^3 336 -> 356 us
// and 1023^3 23.5 -> 24.6 ms: the launch-bound sizes gain the launch latency and the prologue, the large ones lose a few
// per cent.  So the attribute is only set inside a PdlScope that a handle opens when its grid is small
// (pdl_small_grid); FDMB_PDL=0 never sets it, FDMB_PDL=2 sets it for every size.
int pdl_mode();                                   // 0 off, 1 small grids (default), 2 always
inline bool& pdl_scope_on() { static thread_local bool v = false; return v; }
inline bool pdl_small_grid(long long points) { return pdl_mode() == 2 || (pdl_mode() == 1 && points <= (1ll << 22)); }
struct PdlScope {
    bool prev;
    explicit PdlScope(bool on) : prev(pdl_scope_on()) { pdl_scope_on() = on; }
    ~PdlScope() { pdl_scope_on() = prev; }
};
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_scope_on() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

}  // namespace fdmb
