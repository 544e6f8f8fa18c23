// Group-ring versions of the DST sweeps at the lengths where ONE tile per CTA leaves the SM waiting (sm_100a).
//
// k_cols_pipe / k_rows_pipe at N = 1024 hold one 64 KB tile per CTA: a CTA asks for its next tile only after it has
// finished the current one, so it sits on the mbarrier for the whole load latency (ncu r01h: 37 % of the y sweep's
// stall samples), and the rows sweep additionally serialises first touch / transform / copy-out behind CTA barriers.
// Here one 512-thread CTA per SM runs TWO consumer groups of 256 threads over a ring of THREE tile buffers:
//
//     tile i of this CTA lives in buffer i % 3 and is transformed by group i % 2;
//     when a group has finished tile i it re-arms buffer i % 3 with tile i + 3 (which the OTHER group will consume),
//     so each group always finds its next tile already in flight or landed: the load latency hides behind one
//     whole tile transform, and the two groups' shared-memory and fp64 phases overlap like two resident CTAs.
//
// Groups synchronise with named barriers (bar.sync id, 256); the three `full` mbarriers are shared.  The transform
// itself is dst_tile_fused_x (xform.cuh), unchanged arithmetic: parity is that of the pipe kernels.
//
// Hand-off protocol.  A parity wait on an mbarrier is only meaningful while the waiter is at most ONE phase ahead of
// the barrier.  The consumer of use j of a buffer can, however, arrive before the OTHER group has finished use j-1 and
// re-armed the buffer -- and if use j-1 has not even landed yet (a slow load, e.g. a concurrent kernel hogging DRAM),
// full[s] is still two phases back and the parity test passes on the stale phase.  Therefore every arming also
// arrives on a second barrier armed[s]: the consumer of use j first waits for armed[s] (the previous arming of this
// buffer, use j-1, was done by the consumer's own group, so it is never more than one phase ahead there) and only then
// for full[s], which by then is in phase j or j+1.
//
//   k_cols_ring : strided-axis DST sweep, optionally forward -> spectral multiply -> inverse (the z sweep).
//                 Tiles land in the planar layout through two tensor maps exactly as in k_cols_pipe.
//   k_rows_ring : contiguous-axis DST sweep.  The 8 rows of a tile land DENSE (one bulk copy for the caller's
//                 unpitched arrays, one bulk copy per row for the pitched work array, then with a skewed shared-memory
//                 pitch); stage A folds straight out of the dense rows into registers and, after its barrier, writes the
//                 planar tile over the same buffer -- the separate first-touch pass of k_rows_pipe (128 KB of
//                 shared-memory traffic per tile) and its staging buffer are gone.
#pragma once
#include "xform_pipe.cuh"

namespace fdmb {

// named barrier over the NT threads of one consumer group (ids 1.. ; 0 is __syncthreads)
template <int NT> struct GroupSync {
    int id;
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NT) : "memory"); }
};

template <int N> struct RingCfg {
    static constexpr int B = 8;                         // sequences per tile
    static constexpr int G = 32;                        // threads per sequence
    static constexpr int GT = B * G;                    // threads per consumer group
    static constexpr int NG = 2;                        // consumer groups
    static constexpr int THREADS = GT * NG;
    static constexpr int NBUF = 3;
    static constexpr int SCR = (G + G / 8 + 1) * B;     // per group
    static constexpr int M = N / 2;
    // ---- strided-axis sweep: planar tile [N + GAP + 1][B], as PipeCfg<N>
    static constexpr int COLS_GAP = 1;
    static constexpr int COLS_BUF = ((N + 2) * B + 15) / 16 * 16;
    static constexpr int COLS_TAB = (N / 2 + 2) + 2 * (N / 2) + 2 * (N / 2 + 2);       // SN, WM, SF1, SF2
    static constexpr size_t cols_smem() { return 8 * (size_t)(NBUF * COLS_BUF + NG * SCR + COLS_TAB + 2) + 8 * 8 + 128; }
    // ---- contiguous-axis sweep: planar rows [B][P]; GAP = 7 puts the 8 odd + 8 even slots a half-warp touches when it
    //      walks along a row (copy-out) into 16 different banks; P = 2 (mod 16) keeps lanes (row b, thread g) apart
    static constexpr int ROWS_GAP = 7;
    static constexpr int ROWS_P = N + 18;
    static constexpr int ROWS_BUF = B * ROWS_P;
    static constexpr int ROWS_TAB = (N / 2 + 2) + 2 * (N / 2) + (N / 2 + 2);            // SN, WM, SF1
    static constexpr size_t rows_smem() { return 8 * (size_t)(NBUF * ROWS_BUF + NG * SCR + ROWS_TAB + 2) + 8 * 8 + 128; }
    static_assert(ROWS_P % 16 == 2 && ROWS_P >= N + ROWS_GAP + 1, "rows tile pitch");
};

// ------------------------------------------------------------------------------------------------------------------
template <int N, typename MID, typename OMAP>
__global__ void __launch_bounds__(RingCfg<N>::THREADS, 1)
k_cols_ring(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm2, ColsPipeArgs a, MID mid,
            const __grid_constant__ OMAP omap)
{
    using C = RingCfg<N>;
    constexpr int B = C::B, G = C::G, M = N / 2, GAP = C::COLS_GAP, BUF = C::COLS_BUF, NBUF = C::NBUF, GT = C::GT;
    constexpr int J0 = 1;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* bufs = reinterpret_cast<double*>(smem_raw);
    double* scr_all = bufs + NBUF * BUF;
    double* SNs = scr_all + C::NG * C::SCR;
    cd* WMs = reinterpret_cast<cd*>(SNs + (N / 2 + 2));
    double* SF1 = reinterpret_cast<double*>(WMs + N / 2);
    double* SF2 = SF1 + (N / 2 + 2);
    uint64_t* full = reinterpret_cast<uint64_t*>(SF2 + (N / 2 + 2));
    uint64_t* armed = full + NBUF;

    const int tid = threadIdx.x;
    const int grp = tid / GT, t = tid % GT;
    const int b = t % B, g = t / B;
    const int nbt = (a.nb + B - 1) / B;
    const int ntiles = nbt * a.no;
    const unsigned tx_bytes = (unsigned)a.nchunk * a.boxrows * B * 8 * 2;
    const GroupSync<GT> gs{1 + grp};
    double* scr = scr_all + grp * C::SCR;

    auto issue = [&](int tile, int s) {
        if (a.reverse) tile = ntiles - 1 - tile;
        const int o = tile / nbt, b0 = (tile % nbt) * B;
        mbar_expect_tx(&full[s], tx_bytes);
        double* buf = bufs + s * BUF;
        for (int c = 0; c < a.nchunk; c++) {
            const int r0 = c * a.boxrows;
            double* dO = buf + r0 * B;                              // O[h]   <- global row 2h
            double* dE = buf + (M + GAP + 1 + r0) * B;              // E[h+1] <- global row 2h + 1
            switch (a.taxis) {
            case 1:
                tma_load_3d(dO, &tm, b0, r0, o, &full[s]);
                tma_load_3d(dE, &tm2, b0, r0, o, &full[s]);
                break;
            case 2:
                tma_load_3d(dO, &tm, b0, o, r0, &full[s]);
                tma_load_3d(dE, &tm2, b0, o, r0, &full[s]);
                break;
            case 3:
                tma_load_4d(dO, &tm, b0, 0, o, c * a.boxhi, &full[s]);
                tma_load_4d(dE, &tm2, b0, 0, o, c * a.boxhi, &full[s]);
                break;
            default:
                tma_load_4d(dO, &tm, b0, o & ((1 << a.blog) - 1), r0, o >> a.blog, &full[s]);
                tma_load_4d(dE, &tm2, b0, o & ((1 << a.blog) - 1), r0, o >> a.blog, &full[s]);
                break;
            }
        }
        mbar_arrive(&armed[s]);
    };

    if (tid == 0) {
        tma_prefetch_desc(&tm);
        tma_prefetch_desc(&tm2);
        for (int s = 0; s < NBUF; s++) { mbar_init(&full[s], 1); mbar_init(&armed[s], 1); }
        mbar_init_fence();
    }
    load_tables<N>(SNs, WMs, a.SN, a.WM);
    load_fold_table<N>(SF1, a.SN, 0.5 * a.scale);
    if constexpr (MID::active) load_fold_table<N>(SF2, a.SN, 0.5 * a.scale2);
    __syncthreads();
    pdl_wait();      // everything above is independent of the previous kernel's output (pdl.cuh)
    pdl_trigger();
    // tile of this CTA's i-th slot: grid-strided, or in adjacent pairs (i even / odd = the two groups)
    auto tile_of = [&](int i) -> long long {
        return a.pair_tiles ? 2ll * blockIdx.x + (i & 1) + (long long)(i >> 1) * 2 * gridDim.x
                            : blockIdx.x + (long long)i * gridDim.x;
    };
    if (tid == 0) {
        for (int i = 0; i < NBUF; i++) {
            const long long tile = tile_of(i);
            if (tile < ntiles) issue((int)tile, i);
        }
    }

    for (int i = grp;; i += C::NG) {
        const long long tile_l = tile_of(i);
        if (tile_l >= ntiles) break;
        const int s = i % NBUF;
        const unsigned parity = (i / NBUF) & 1;
        const int tt = a.reverse ? ntiles - 1 - (int)tile_l : (int)tile_l;
        const int o = tt / nbt, b0 = (tt % nbt) * B;
        const bool bok = b0 + b < a.nb;
        const long long ooff = (a.taxis == 4)
                                   ? (long long)(o >> a.blog) * a.out_so_hi + (long long)(o & ((1 << a.blog) - 1)) * a.out_so
                                   : (long long)o * a.out_so;
        double* tile = bufs + s * BUF;
        mbar_wait(&armed[s], parity);
        mbar_wait(&full[s], parity);

        const auto og = omap.emitter(a.out, a.out_sj, ooff, J0, o, b0 + b, bok);
        const InPlanar<N, GAP> in{tile + b, B};
        if constexpr (MID::active) {
            const OutMidTile<N, GAP, MID> om{tile + b, B, mid, mid.ctx(bok ? b0 + b + 1 : 0, o + a.mid_o_off + 1), bok};
            dst_tile_fused_x<N, G, GAP, false, true, false>(tile + b, B, g, 0.5 * a.scale, SNs, SF1, WMs, scr + b, B, om,
                                                            (double*)nullptr, gs, in);
            dst_tile_fused_x<N, G, GAP, false, true, false>(tile + b, B, g, 0.5 * a.scale2, SNs, SF2, WMs, scr + b, B, og,
                                                            (double*)nullptr, gs, in);
        } else {
            dst_tile_fused_x<N, G, GAP, false, true, false>(tile + b, B, g, 0.5 * a.scale, SNs, SF1, WMs, scr + b, B, og,
                                                            (double*)nullptr, gs, in);
        }
        // the buffer is free once every thread of the group has read it; order those generic accesses before the
        // async-proxy writes of the tile that lands here next (consumed by the other group)
        fence_proxy_async();
        gs.sync();
        const long long tn = tile_of(i + NBUF);
        if (t == 0 && tn < ntiles) issue((int)tn, s);
    }
}

// ------------------------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(RingCfg<N>::THREADS, 1) k_rows_ring(RowsPipeArgs a)
{
    using C = RingCfg<N>;
    constexpr int B = C::B, G = C::G, M = N / 2, GAP = C::ROWS_GAP, P = C::ROWS_P, BUF = C::ROWS_BUF, NBUF = C::NBUF,
                  GT = C::GT;
    using PL = Planar<N, GAP>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* bufs = reinterpret_cast<double*>(smem_raw);
    double* scr_all = bufs + NBUF * BUF;
    double* SNs = scr_all + C::NG * C::SCR;
    cd* WMs = reinterpret_cast<cd*>(SNs + (N / 2 + 2));
    double* SF1 = reinterpret_cast<double*>(WMs + N / 2);
    uint64_t* full = reinterpret_cast<uint64_t*>(SF1 + (N / 2 + 2));
    uint64_t* armed = full + NBUF;

    const int tid = threadIdx.x;
    const int grp = tid / GT, t = tid % GT;
    const int b = t % B, g = t / B;
    const int warp = t >> 5, lane = t & 31;            // within the group: 8 warps, one per row in the copy-out
    const GroupSync<GT> gs{1 + grp};
    double* scr = scr_all + grp * C::SCR;
    const double hs = 0.5 * a.scale;

    const int YB = 1 << a.blog, SUB = YB / B > 0 ? YB / B : 1;   // blk = 2: B divides YB
    const int nyb = (a.ny + YB - 1) >> a.blog;
    const long long ntiles = (a.blk == 2) ? (long long)a.nz * nyb * SUB : (a.nrows + B - 1) / B;
    // Landing of the 8 dense rows of a tile (stage A reads them with lanes = (row b, thread g): a half-warp's 16 accesses
    // sit at off[b] + 2 g' + const, so it needs row offsets that cover BOTH parities to reach all 16 double-banks):
    //   parity  the caller's unpitched arrays (pitch = N - 1, odd): row r of a tile starts 16-byte aligned for even r
    //           and 8 bytes off for odd r.  One bulk copy per row of N doubles -- odd rows from one element earlier --
    //           lands element j of row r at off[r] + j with off[r] = 1040 r + 4 (r >> 1) + (r & 1), i.e.
    //           off[r] mod 16 = 0, 1, 4, 5, 8, 9, 12, 13: conflict-free (r02d: the chunk landing below was 2-way
    //           conflicted, 32 % of the kernel's wavefronts).  The last row of the whole array is copied two doubles
    //           short and patched (the copy would run past the allocation).
    //   per_row pitched inputs (even pitch: every row 16-byte aligned): one copy per row, pitch skewed by 2 doubles;
    //           all offsets even, so a half-warp reaches 8 banks: 2-way conflicts remain (work array -> ans sweep)
    //   chunk   any other odd pitch: ONE contiguous copy with the pitch it has
    const bool land_par = (a.in_pitch & 1) && a.in_pitch == N - 1 && a.nvalid == N - 1 && a.blk == 0;
    const bool per_row = (a.in_pitch & 1) == 0;
    const int LP = land_par ? 1040 : ((per_row && (a.in_pitch & 15) == 0) ? a.in_pitch + 2 : a.in_pitch);
    const unsigned row_bytes = (unsigned)a.in_pitch * 8u;
    auto row_off = [&](int r) { return land_par ? r * 1040 + 4 * (r >> 1) + (r & 1) : r * LP; };

    auto locate = [&](long long tl, long long& in_row0, long long& nat_row0, int& rows) {
        if (a.reverse) tl = ntiles - 1 - tl;
        if (a.blk == 2) {
            const int sub = (int)(tl % SUB);
            const long long t2 = tl / SUB;
            const int yb = (int)(t2 % nyb);
            const long long z = t2 / nyb;
            const int y0 = yb * YB + sub * B;
            in_row0 = (((long long)yb * a.nz + z) << a.blog) + sub * B;
            nat_row0 = z * a.ny + y0;
            rows = a.ny - y0 < B ? a.ny - y0 : B;
            if (rows < 0) rows = 0;
        } else {
            in_row0 = nat_row0 = tl * B;
            rows = (int)((a.nrows - nat_row0) < B ? (a.nrows - nat_row0) : B);
        }
    };
    auto out_row = [&](long long row) -> long long {
        if (a.blk != 1) return row;
        const long long z = row / a.ny;
        const int y = (int)(row - z * a.ny);
        return ((((long long)(y >> a.blog)) * a.nz + z) << a.blog) + (y & (YB - 1));
    };
    auto issue = [&](long long tl, int s) {
        long long in_row0, nat_row0; int rows;
        locate(tl, in_row0, nat_row0, rows);
        double* buf = bufs + s * BUF;
        const double* src = a.in + in_row0 * a.in_pitch;
        if (land_par) {
            unsigned total = 0;
            for (int r = 0; r < rows; r++) total += (in_row0 + r == a.nrows - 1 && !(r & 1)) ? (N - 2) * 8u : N * 8u;
            mbar_expect_tx(&full[s], total);
            for (int r = 0; r < rows; r++) {
                const unsigned bytes = (in_row0 + r == a.nrows - 1 && !(r & 1)) ? (N - 2) * 8u : N * 8u;
                bulk_load_1d(buf + row_off(r) - (r & 1), src + (long long)r * a.in_pitch - (r & 1), bytes, &full[s]);
            }
        } else if (per_row) {
            mbar_expect_tx(&full[s], (unsigned)rows * row_bytes);
            for (int r = 0; r < rows; r++) bulk_load_1d(buf + r * LP, src + (long long)r * a.in_pitch, row_bytes, &full[s]);
        } else {
            const unsigned bytes = ((unsigned)rows * row_bytes) & ~15u;
            mbar_expect_tx(&full[s], bytes);
            if (bytes) bulk_load_1d(buf, src, bytes, &full[s]);
        }
        mbar_arrive(&armed[s]);
    };

    if (tid == 0) {
        for (int s = 0; s < NBUF; s++) { mbar_init(&full[s], 1); mbar_init(&armed[s], 1); }
        mbar_init_fence();
    }
    load_tables<N>(SNs, WMs, a.SN, a.WM);
    load_fold_table<N>(SF1, a.SN, hs);
    __syncthreads();
    pdl_wait();      // everything above is independent of the previous kernel's output (pdl.cuh)
    pdl_trigger();
    if (tid == 0) {
        for (int i = 0; i < NBUF; i++) {
            const long long tl = blockIdx.x + (long long)i * gridDim.x;
            if (tl < ntiles) issue(tl, i);
        }
    }

    for (int i = grp;; i += C::NG) {
        const long long tl = blockIdx.x + (long long)i * gridDim.x;
        if (tl >= ntiles) break;
        const int s = i % NBUF;
        const unsigned parity = (i / NBUF) & 1;
        long long in_row0, row0; int rows;
        locate(tl, in_row0, row0, rows);
        double* buf = bufs + s * BUF;
        mbar_wait(&armed[s], parity);
        mbar_wait(&full[s], parity);
        if (land_par) {   // the array's last row (even position in its tile) was copied two doubles short: element N - 2
            const int r = (int)(a.nrows - 1 - in_row0);
            if (r >= 0 && r < B && !(r & 1)) {
                if (t == 0) buf[row_off(r) + N - 2] = a.in[(a.nrows - 1) * a.in_pitch + N - 2];
                gs.sync();
            }
        } else if (!per_row) {   // an odd tail (rows * pitch odd) leaves one double outside the 16-byte granularity of the bulk copy
            const long long cnt = (long long)rows * a.in_pitch;
            if (cnt & 1) {
                if (t == 0) buf[cnt - 1] = a.in[in_row0 * a.in_pitch + cnt - 1];
                gs.sync();
            }
        }
        // fold + three radix passes + untangle + running sum, in place: dense rows in, planar rows [B][P] out
        dst_tile_fused_x<N, G, GAP, false, true, false>(buf + b * P, 1, g, hs, SNs, SF1, WMs, scr + b, B,
                                                        OutTile<N, GAP>{buf + b * P, 1}, (double*)nullptr, gs,
                                                        InDense{buf + row_off(b), b < rows});
        // copy-out: one warp per row, lanes along the row
        if (warp < rows) {
            double* dst = a.out + out_row(row0 + warp) * a.out_pitch;
            const double* src = buf + warp * P;
#pragma unroll 4
            for (int x = lane; x < a.nvalid; x += 32) dst[x] = src[PL::row(x + 1)];
        }
        fence_proxy_async();
        gs.sync();
        const long long tn = tl + (long long)NBUF * gridDim.x;
        if (t == 0 && tn < ntiles) issue(tn, s);
    }
}

// ---- host-side launchers ----------------------------------------------------------------------------------------
template <int N, typename MID, typename OMAP = OutLinear>
inline cudaError_t launch_cols_ring_t(const CUtensorMap& tm, const CUtensorMap& tm2, const ColsPipeArgs& a, const MID& mid,
                                      cudaStream_t st, const OMAP& omap = OMAP{})
{
    using C = RingCfg<N>;
    auto kern = k_cols_ring<N, MID, OMAP>;
    constexpr size_t smem = C::cols_smem();
    static bool done_dev[64] = {false};
    bool& done = done_dev[current_device_slot()];
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        done = true;
    }
    if (preload_only()) return cudaSuccess;
    const long long ntiles = (long long)((a.nb + C::B - 1) / C::B) * a.no;
    long long grid = device_sm_count();
    if (a.max_ctas > 0 && grid > a.max_ctas) grid = a.max_ctas;
    if (grid > (ntiles + 1) / 2) grid = (ntiles + 1) / 2;      // both groups of a CTA get a tile
    if (grid < 1) return cudaSuccess;
    return launch_pdl(kern, dim3((unsigned)grid), dim3(C::THREADS), smem, st, tm, tm2, a, mid, omap);
}

template <int N>
inline cudaError_t launch_rows_ring_t(const RowsPipeArgs& a, cudaStream_t st)
{
    using C = RingCfg<N>;
    auto kern = k_rows_ring<N>;
    constexpr size_t smem = C::rows_smem();
    static bool done_dev[64] = {false};
    bool& done = done_dev[current_device_slot()];
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        done = true;
    }
    if (preload_only()) return cudaSuccess;
    const long long ntiles = (a.nrows + C::B - 1) / C::B;       // (blocked input: an upper bound is enough for the grid)
    long long grid = device_sm_count();
    if (a.max_ctas > 0 && grid > a.max_ctas) grid = a.max_ctas;
    if (grid > (ntiles + 1) / 2) grid = (ntiles + 1) / 2;
    if (grid < 1) return cudaSuccess;
    return launch_pdl(kern, dim3((unsigned)grid), dim3(C::THREADS), smem, st, a);
}

// can the rows ring kernel take this sweep?  (dense rows + planar rows must fit the tile buffer)
template <int N> inline bool rows_ring_fits(const RowsPipeArgs& a)
{
    using C = RingCfg<N>;
    const bool per_row = (a.in_pitch & 1) == 0;
    const bool parity = (a.in_pitch & 1) && a.in_pitch == N - 1 && a.blk == 0;
    const int LP = parity ? 1040 : ((per_row && (a.in_pitch & 15) == 0) ? a.in_pitch + 2 : a.in_pitch);
    static_assert(7 * 1040 + 12 + 1 + N <= C::ROWS_BUF || N != 1024, "parity landing fits the tile buffer");
    return a.nvalid == N - 1 && a.in_pitch >= a.nvalid && (long long)C::B * LP <= C::ROWS_BUF &&
           (reinterpret_cast<uintptr_t>(a.in) & 15) == 0;
}

}  // namespace fdmb
