// sm_100a asynchronous-copy plumbing: mbarrier, 1-D bulk copies (UBLKCP) and tensor-map tile
// copies (UTMALDG) issued by one elected thread, consumed by the whole CTA.  Inline PTX only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace fdmb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence()
{
    // make the initialised barriers visible to the async proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// plain arrival (release, CTA scope): counts towards the barrier's pending arrivals, no transaction bytes
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// generic-proxy accesses to shared memory before this fence are ordered before later async-proxy
// (TMA) accesses to the same locations
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// global -> shared, contiguous; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// with an L2 eviction-priority hint (createpolicy value)
__device__ __forceinline__ void bulk_load_1d_hint(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}

// global (tensor map, 3-D tile) -> shared; OOB elements are zero-filled and still count as bytes
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar,
                                                 uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], "
        "[%5], %6;" ::"r"(smem_u32(smem_dst)),
        "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
// shared -> global (tensor map, 3-D tile), bulk async-group completion; elements outside the tensor are not written
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(c0), "r"(c1),
                 "r"(c2), "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk groups of this thread have finished READING their shared-memory sources
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... have completed altogether (their global writes are performed)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_hint(double* p, double v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// ---- host: tensor-map encoding through the driver entry point (no link-time libcuda dependency)
// 3-D fp64 tensor (x fastest): extents e0,e1,e2 (elements), strides s1,s2 (BYTES, multiples of 16),
// box b0,b1,b2 (elements, each <= 256, b0*8 a multiple of 16).
int make_tensor_map_3d(CUtensorMap* tm, const void* base, unsigned long long e0, unsigned long long e1,
                       unsigned long long e2, unsigned long long s1, unsigned long long s2, unsigned b0, unsigned b1,
                       unsigned b2);


// Tensor maps of one strided-axis sweep over a 3-D fp64 array (x fastest; taxis = 1: tiles run along
// dimension 1, taxis = 2: along dimension 2; B = tile width in x):
//   nat          natural landing: row r of the transform axis -> tile row r (+ slot offset)
//   podd / peven planar landing of the fused DST (built when the axis holds N-1 entries): the even
//                global rows (odd slots) and the odd global rows (even slots) as two strided views
struct ColsMaps {
    CUtensorMap nat, podd, peven;
    int boxrows = 0, nchunk = 0;          // natural
    int boxrows_p = 0, nchunk_p = 0;      // planar, per parity
    bool planar = false;
    int taxis = 0;                        // set by make_cols_maps_blocked: 3 (pair axis) or 4 (mid axis)
    int boxhi = 0, boxhi_p = 0;           // taxis 3: blocks per box along the hi dimension
    int blog = 0;                         // log2(block)
};
int make_cols_maps(ColsMaps* m, const void* base, int N, int taxis, unsigned long long e0, unsigned long long e1,
                   unsigned long long e2, unsigned long long s1, unsigned long long s2, unsigned B);


// The same for a BLOCKED 4-D array (x, lo, mid, hi): an axis of n entries is split as index = hi * LO + lo with
// LO = 1 << blog consecutive entries kept next to each other (stride s_lo) inside every `mid` slice:
//   pair_axis = true  : tiles run along the blocked axis (n entries), outer index = mid
//   pair_axis = false : tiles run along mid (n entries), outer index = hi * LO + lo
// Keeps the 2 MB pages one tile touches to a few dozen for BOTH strided axes of a large 3-D array (a natural
// layout puts the 1023 rows of a third-axis tile on 1023 different pages, which the TLBs do not cover).
// Strides in BYTES.  n_lohi / n_mid are the real entry counts along the blocked axis / mid.
int make_cols_maps_blocked(ColsMaps* m, const void* base, int N, bool pair_axis, int blog, unsigned long long e_x,
                           unsigned long long n_lohi, unsigned long long n_mid, unsigned long long s_lo,
                           unsigned long long s_mid, unsigned long long s_hi, unsigned B);

}  // namespace fdmb
