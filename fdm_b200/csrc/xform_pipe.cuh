// Persistent, asynchronously fed versions of the sweep kernels (sm_100a).
//
//   k_cols_pipe : strided-axis sweep.  Each CTA walks a static stride of [N][B] tiles.  One elected
//                 thread arms an mbarrier and issues tensor-map tile copies (UTMALDG) into a ring of
//                 NSTAGE shared-memory buffers; the CTA transforms a buffer in place, writes it back
//                 with coalesced stores, and re-arms the buffer for the tile NSTAGE strides ahead.
//                 Loads of tile i+1.. are in flight while tile i is transformed.
//   k_rows_pipe : contiguous-axis sweep.  BR consecutive rows are ONE contiguous global chunk (also
//                 for the caller's unpitched arrays), fetched with a 1-D bulk copy (UBLKCP) into a
//                 dense staging ring; the first pass of the transform (the DST fold) reads the
//                 staging buffer with lanes along the row and writes the odd-pitch compute tile, so
//                 the repack is free.
// Twiddle tables live in shared memory (warp-uniform broadcast reads).
#pragma once
#include "pipe.cuh"
#include "xform_kernels.cuh"

namespace fdmb {

// WIDE: the multi-GPU transposing sweeps.  Their stores cross NVLink, where 128-byte segments move at ~1.5x the
// rate of 64-byte ones (measured: 620 vs 420 GB/s per direction), so N = 1024 keeps 16-column tiles there: one
// 128 KB tile and one 512-thread CTA per SM, no second stage (those sweeps are NVLink-bound, not SM-bound).
template <int N, bool WIDE = false> struct PipeCfg {
    static constexpr int B = (N <= 512 || (WIDE && N == 1024)) ? 16 : 8;      // sequences per tile
    static constexpr int G = Plan<N>::G;               // threads per sequence, contiguous-axis sweep
    static constexpr int THREADS = B * G;
    // strided-axis sweeps at N = 1024: the last radix pass has only 32 units (block pairs) per sequence, so 64
    // threads per sequence leave half the CTA idle in the heaviest stage.  32 threads per sequence (two
    // first/middle-pass butterflies each) keep every warp busy, and two 256-thread CTAs share an SM, so one CTA's
    // shared-memory phases overlap the other's fp64 phases.
    static constexpr int GC = (N == 1024) ? 32 : G;
    static constexpr int THREADS_COLS = B * GC;
    // resident CTAs the register allocation must leave room for (plain sweep / fused forward-multiply-inverse sweep)
    static constexpr int COLS_MIN_CTAS = (N == 1024 && !WIDE) ? 2 : (N == 256 ? 4 : 1);
    static constexpr int COLS_MIN_CTAS_MID = (N == 1024 && !WIDE) ? 2 : (N == 256 ? 4 : 1);
    static constexpr int SCR = (G + G / 8 + 1) * B;
    static constexpr bool SWZ = (B == 8);              // 64-byte tile rows: keep half-warps on rows of different parity
    // SN + WM + two fold tables SF (doubles)
    static constexpr int TAB = (N / 2 + 2) + 2 * (N / 2) + 2 * (N / 2 + 2);
    static constexpr int COLS_GAP = 1;                 // planar tile: E region starts one row late (128-B aligned landing)
    static constexpr int COLS_BUF = ((N + 2) * B + 15) / 16 * 16;      // doubles per stage buffer (128-B multiple)
    static constexpr int COLS_STAGES = (N <= 512) ? 2 : 1;
    // fused DST sweeps write their results into a second tile instead of back into the working one
    // (dst_tile_fused SEP); only where one CTA per SM is resident anyway and the tile still fits
    static constexpr bool COLS_SEP = false;
    static constexpr size_t cols_smem(int nstage)
    {
        return 8 * (size_t)(16 + (nstage + (COLS_SEP ? 1 : 0)) * COLS_BUF + SCR + TAB + 2) + 8 * 8 + 128;
    }
    static constexpr int BR = B;
    // planar rows tile: the E region starts 8 doubles after a multiple of 16 so that the 8 even + 8 odd slots a
    // half-warp touches when it walks along a row fall into different banks (Planar<N, ROWS_GAP>)
    static constexpr int ROWS_GAP = 8;
    static constexpr int ROWS_P = (BR == 8) ? N + 18 : N + 9;   // tile pitch: = 2 (mod 16) for 8 rows, odd for 16 rows
    static constexpr int ROWS_STAGES = 1;
    static constexpr size_t rows_smem(int nstage)
    {
        return 8 * (size_t)(nstage * BR * N + BR * ROWS_P + SCR + TAB + 2) + 8 * 8 + 128;
    }
};

template <int N>
__device__ __forceinline__ void load_tables(double* SNs, cd* WMs, const double* __restrict__ SN, const cd* __restrict__ WM)
{
    for (int i = threadIdx.x; i <= N / 2; i += blockDim.x) SNs[i] = SN[i];
    for (int i = threadIdx.x; i < N / 2; i += blockDim.x) WMs[i] = WM[i];
}
// fold table of the fused DST: SF[j] = hs * sin(pi j / N)
template <int N>
__device__ __forceinline__ void load_fold_table(double* SFs, const double* __restrict__ SN, double hs)
{
    for (int i = threadIdx.x; i <= N / 2; i += blockDim.x) SFs[i] = hs * SN[i];
}

// ---- output maps of the strided-axis sweep ----------------------------------------------------
// A map turns (transform slot, outer index, contiguous index) into a destination address.
//   OutLinear : one pitched array on this GPU (slot j -> row j - J0)
//   OutShard  : the slab <-> pencil transpose of the multi-GPU solve fused into the sweep's stores:
//               slot j belongs to rank j >> logS and lands in that rank's buffer through its
//               peer-mapped base pointer (NVLink stores, 128-byte segments along the contiguous axis)
constexpr int FDMB_MAX_RANKS = 8;

struct EmitLinear {
    double* dst; long long stride; bool ok;
    __device__ __forceinline__ void emit(int j, double v) const { if (ok) dst[(long long)j * stride] = v; }
};
struct OutLinear {
    static constexpr bool sharded = false;
    // ooff: offset (doubles) of the tile's outer index, computed by the kernel (linear or blocked)
    __device__ __forceinline__ EmitLinear emitter(double* out, long long sj, long long ooff, int j0, int, int x, bool ok) const
    {
        return EmitLinear{out + ooff + x - (long long)j0 * sj, sj, ok};
    }
};
// the transform axis itself is blocked in the destination (taxis 3): entry r = hi * LO + lo
struct EmitBlocked {
    double* dst; long long s_lo, s_hi; int j0, blog, bmask; bool ok;
    __device__ __forceinline__ void emit(int j, double v) const
    {
        const int r = j - j0;
        if (ok) dst[(long long)(r >> blog) * s_hi + (long long)(r & bmask) * s_lo] = v;
    }
};
struct OutBlocked {
    static constexpr bool sharded = false;
    long long s_hi; int blog;
    __device__ __forceinline__ EmitBlocked emitter(double* out, long long sj, long long ooff, int j0, int, int x, bool ok) const
    {
        return EmitBlocked{out + ooff + x, sj, s_hi, j0, blog, (1 << blog) - 1, ok};
    }
};
struct EmitShard {
    double* const* base; int logS, maskS; long long sj, off; bool ok;
    __device__ __forceinline__ void emit(int j, double v) const
    {
        if (ok) base[j >> logS][(long long)(j & maskS) * sj + off] = v;
    }
};
struct OutShard {
    static constexpr bool sharded = true;
    double* base[FDMB_MAX_RANKS];   // peer-mapped destination buffers, indexed by owning rank
    int logS, maskS;                // slots per rank = 1 << logS
    long long sj, so;               // destination strides (doubles) along the transform / outer axis
    int o_off;                      // destination outer index of this rank's first outer entry
    __device__ __forceinline__ EmitShard emitter(double*, long long, long long, int, int o, int x, bool ok) const
    {   // (the kernel's outer offset is ignored: the destination buffers have their own strides)
        return EmitShard{base, logS, maskS, sj, (long long)(o + o_off) * so + x, ok};
    }
};

// The transposing sweeps' stores as TMA tile stores out of the finished planar tile (DST sweeps): per destination rank
// one box of its odd slots and one of its even slots (the same strided views the planar LOADS use).  The SMs only
// issue 2 * nranks descriptors per tile; the NVLink writes drain in the background while the CTA fetches and
// transforms its next tile -- with per-value st.global the store instructions themselves stalled on the link.
struct OutShardTma {
    static constexpr bool sharded = true;
    static constexpr bool tma_store = true;
    CUtensorMap odd[FDMB_MAX_RANKS], even[FDMB_MAX_RANKS];   // destination views of rank q: local slots 1,3,.. / 0,2,..
    int nranks, half;               // half = slots per rank / 2 = rows per box
    int taxis;                      // 1: box (cols, rows, 1), coordinates (b0, 0, o + o_off); 2: box (cols, 1, rows), (b0, o + o_off, 0)
    int o_off;
};
template <typename OMAP, typename = void> struct UsesTmaStore { static constexpr bool value = false; };
template <typename OMAP> struct UsesTmaStore<OMAP, decltype((void)OMAP::tma_store)> { static constexpr bool value = OMAP::tma_store; };

struct ColsPipeArgs {
    double* out;
    long long out_sj, out_so;   // output strides (doubles) along the transform / outer axis
    int nvalid;                 // entries along the transform axis
    int nb, no;                 // extents of the contiguous / outer axis
    int taxis;                  // 1 / 2: tensor-map dimension of the transform axis (3-D maps);
                                // 3 / 4: blocked 4-D maps, tiles along the blocked axis / along mid (ColsMaps)
    int boxrows, nchunk;        // rows per tensor-map box, boxes per tile
    int boxhi;                  // taxis 3: blocks per box
    int blog;                   // blocked layouts: log2(block); taxis 4 splits the outer index with it
    long long out_so_hi;        // taxis 4: output stride of the outer index's block number (out_so: within a block)
    int reverse;                // walk the tiles back to front (L2 reuse against the previous sweep)
    int mid_o_off;              // added to the outer index handed to the mid functor (sharded sweeps)
    int max_ctas;               // > 0: upper bound of the grid (leaves SMs to a kernel running beside this one)
    int pair_tiles;             // ring sweeps: the two groups of a CTA take ADJACENT tiles (the two 64-byte halves of the
                                // same 128-byte lines, the same pages) instead of tiles a whole grid stride apart
    double scale, scale2;
    const double* SN;
    const cd* WM;
};

// FUSED (DST, and DST -> multiply -> DST): the tile lands in the planar layout of dst_tile_fused through two
// tensor maps (tm: even global rows = odd slots, tm2: odd global rows = even slots); otherwise tm lands the
// tile in natural order and tm2 is unused.
template <int KIND, typename MID, int KIND2> struct ColsFused {
    static constexpr bool value = (KIND == XF_DST) && (!MID::active || KIND2 == XF_DST);
};

template <int N, int KIND, typename MID, int KIND2, int NSTAGE, typename OMAP>
__global__ void __launch_bounds__(PipeCfg<N, OMAP::sharded>::THREADS_COLS,
                                  MID::active ? PipeCfg<N, OMAP::sharded>::COLS_MIN_CTAS_MID
                                              : PipeCfg<N, OMAP::sharded>::COLS_MIN_CTAS)
k_cols_pipe(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm2, ColsPipeArgs a, MID mid,
            const __grid_constant__ OMAP omap)
{
    using C = PipeCfg<N, OMAP::sharded>;
    constexpr int B = C::B, G = C::GC, M = N / 2, GAP = C::COLS_GAP;
    constexpr bool FUSED = ColsFused<KIND, MID, KIND2>::value;
    constexpr int J0 = (KIND == XF_DST) ? 1 : 0;
    constexpr int BUF = C::COLS_BUF;
    // natural landing: keep the first loaded slot 128-B aligned
    constexpr int PRE = FUSED ? 0 : ((J0 * B) % 16 ? 16 - (J0 * B) % 16 : 0);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    double* bufs = sm + PRE;
    constexpr bool SEP = FUSED && C::COLS_SEP;
    double* alt = bufs + NSTAGE * BUF;                         // second tile of the SEP variant
    double* scr = sm + 16 + (NSTAGE + (C::COLS_SEP ? 1 : 0)) * BUF;
    double* SNs = scr + C::SCR;
    cd* WMs = reinterpret_cast<cd*>(SNs + (N / 2 + 2));
    double* SF1 = reinterpret_cast<double*>(WMs + N / 2);
    double* SF2 = SF1 + (N / 2 + 2);
    uint64_t* full = reinterpret_cast<uint64_t*>(SF2 + (N / 2 + 2));

    const int tid = threadIdx.x;
    const int b = tid % B, g = tid / B;
    const int nbt = (a.nb + B - 1) / B;
    const int ntiles = nbt * a.no;
    const unsigned tx_bytes = (unsigned)a.nchunk * a.boxrows * B * 8 * (FUSED ? 2 : 1);

    auto issue = [&](int t, int s) {
        if (a.reverse) t = ntiles - 1 - t;
        const int o = t / nbt, b0 = (t % nbt) * B;
        mbar_expect_tx(&full[s], tx_bytes);
        double* buf = bufs + s * BUF;
        for (int c = 0; c < a.nchunk; c++) {
            const int r0 = c * a.boxrows;
            double* dO = FUSED ? buf + r0 * B : buf + J0 * B + r0 * B;      // FUSED: O[h] <- global row 2h
            double* dE = buf + (M + GAP + 1 + r0) * B;                       // FUSED: E[h+1] <- global row 2h + 1
            switch (a.taxis) {
            case 1:
                tma_load_3d(dO, &tm, b0, r0, o, &full[s]);
                if constexpr (FUSED) tma_load_3d(dE, &tm2, b0, r0, o, &full[s]);
                break;
            case 2:
                tma_load_3d(dO, &tm, b0, o, r0, &full[s]);
                if constexpr (FUSED) tma_load_3d(dE, &tm2, b0, o, r0, &full[s]);
                break;
            case 3:
                tma_load_4d(dO, &tm, b0, 0, o, c * a.boxhi, &full[s]);
                if constexpr (FUSED) tma_load_4d(dE, &tm2, b0, 0, o, c * a.boxhi, &full[s]);
                break;
            default:
                tma_load_4d(dO, &tm, b0, o & ((1 << a.blog) - 1), r0, o >> a.blog, &full[s]);
                if constexpr (FUSED) tma_load_4d(dE, &tm2, b0, o & ((1 << a.blog) - 1), r0, o >> a.blog, &full[s]);
                break;
            }
        }
    };

    if (tid == 0) {
        tma_prefetch_desc(&tm);
        if constexpr (FUSED) tma_prefetch_desc(&tm2);
        for (int s = 0; s < NSTAGE; s++) mbar_init(&full[s], 1);
        mbar_init_fence();
    }
    load_tables<N>(SNs, WMs, a.SN, a.WM);
    if constexpr (FUSED) {
        load_fold_table<N>(SF1, a.SN, 0.5 * a.scale);
        if constexpr (MID::active) load_fold_table<N>(SF2, a.SN, 0.5 * a.scale2);
    }
    __syncthreads();
    pdl_wait();      // everything above is independent of the previous kernel's output (pdl.cuh)
    pdl_trigger();
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            int t = blockIdx.x + s * gridDim.x;
            if (t < ntiles) issue(t, s);
        }
    }

    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
        const int s = it % NSTAGE;
        const unsigned parity = (it / NSTAGE) & 1;
        const int tt = a.reverse ? ntiles - 1 - t : t;
        const int o = tt / nbt, b0 = (tt % nbt) * B;
        const bool bok = b0 + b < a.nb;
        // offset of the outer index in the output: linear, or block number / position in block (taxis 4)
        const long long ooff = (a.taxis == 4)
                                   ? (long long)(o >> a.blog) * a.out_so_hi + (long long)(o & ((1 << a.blog) - 1)) * a.out_so
                                   : (long long)o * a.out_so;
        double* tile = bufs + s * BUF;
        mbar_wait(&full[s], parity);

        if constexpr (FUSED && UsesTmaStore<OMAP>::value) {
            // transposing sweep: the last transform leaves its spectrum in the planar tile, one thread hands the tile to
            // the TMA unit as 2 * nranks boxes (odd / even slots of every destination rank)
            static_assert(!C::SWZ && !SEP, "TMA tile stores need the unswizzled single-tile layout");
            const OutTile<N, GAP> ot{tile + b, B};
            if constexpr (MID::active) {
                const OutMidTile<N, GAP, MID> om{tile + b, B, mid, mid.ctx(bok ? b0 + b + 1 : 0, o + a.mid_o_off + 1), bok};
                dst_tile_fused<N, G, GAP, false, false, false>(tile + b, B, g, 0.5 * a.scale, SNs, SF1, WMs, scr + b, B, om);
                dst_tile_fused<N, G, GAP, false, false, false>(tile + b, B, g, 0.5 * a.scale2, SNs, SF2, WMs, scr + b, B, ot);
            } else {
                dst_tile_fused<N, G, GAP, false, false, false>(tile + b, B, g, 0.5 * a.scale, SNs, SF1, WMs, scr + b, B, ot);
            }
            fence_proxy_async();          // the transform's generic-proxy writes, before the async proxy reads the tile
            __syncthreads();
            if (tid == 0) {
                const int oc = o + omap.o_off;
                for (int q = 0; q < omap.nranks; q++) {
                    const double* so = tile + (size_t)(q * omap.half) * B;                    // O[k]: slot 2k + 1
                    const double* se = tile + (size_t)(M + GAP + q * omap.half) * B;          // E[k]: slot 2k
                    if (omap.taxis == 1) {
                        tma_store_3d(&omap.odd[q], so, b0, 0, oc);
                        tma_store_3d(&omap.even[q], se, b0, 0, oc);
                    } else {
                        tma_store_3d(&omap.odd[q], so, b0, oc, 0);
                        tma_store_3d(&omap.even[q], se, b0, oc, 0);
                    }
                }
                bulk_commit();
                bulk_wait_read();         // the tile may be overwritten once the TMA unit has read it
            }
        } else if constexpr (FUSED) {
            // smem-lean path: finished spectral values leave the registers straight to global memory
            const auto og = omap.emitter(a.out, a.out_sj, ooff, J0, o, b0 + b, bok);
            if constexpr (MID::active) {
                // forward -> multiply -> inverse; with SEP the two transforms ping-pong between the tiles
                double* t2 = SEP ? alt : tile;
                const OutMidTile<N, GAP, MID> om{t2 + b, B, mid, mid.ctx(bok ? b0 + b + 1 : 0, o + a.mid_o_off + 1), bok};
                dst_tile_fused<N, G, GAP, false, C::SWZ, SEP>(tile + b, B, g, 0.5 * a.scale, SNs, SF1, WMs, scr + b, B, om,
                                                              t2 + b);
                dst_tile_fused<N, G, GAP, false, C::SWZ, SEP>(t2 + b, B, g, 0.5 * a.scale2, SNs, SF2, WMs, scr + b, B, og,
                                                              tile + b);
            } else {
                dst_tile_fused<N, G, GAP, false, C::SWZ, SEP>(tile + b, B, g, 0.5 * a.scale, SNs, SF1, WMs, scr + b, B, og,
                                                              alt + b);
            }
        } else {
            xform_tile<N, G, KIND>(tile + b, B, g, a.scale, SNs, WMs, scr + b, B);
            if constexpr (MID::active) {
                for (int j = g; j < a.nvalid; j += G) {
                    double v = tile[(j + J0) * B + b];
                    tile[(j + J0) * B + b] = bok ? mid(v, j + J0, b0 + b + J0, o + a.mid_o_off + J0) : 0.0;
                }
                __syncthreads();
                xform_tile<N, G, KIND2>(tile + b, B, g, a.scale2, SNs, WMs, scr + b, B);
            }
            {
                const auto og = omap.emitter(a.out, a.out_sj, ooff, J0, o, b0 + b, bok);
#pragma unroll 4
                for (int j = g; j < a.nvalid; j += G) og.emit(j + J0, tile[(j + J0) * B + b]);
            }
        }
        // the buffer is free once every thread has read it; order those generic reads/writes before
        // the async-proxy writes of the next tile
        fence_proxy_async();
        __syncthreads();
        const int tn = t + NSTAGE * gridDim.x;
        if (tid == 0 && tn < ntiles) issue(tn, s);
    }
    if constexpr (UsesTmaStore<OMAP>::value) {
        if (tid == 0) bulk_wait_all();    // the tile stores are performed before the CTA (and the kernel) ends
    }
}

struct RowsPipeArgs {
    const double* in;
    double* out;
    long long nrows;
    int nvalid;
    int in_pitch, out_pitch;    // doubles; BR*in_pitch*8 is a multiple of 16, in is 16-byte aligned
    int reverse;
    double scale;
    const double* SN;
    const cd* WM;
    // Blocked work-array layout [yb][z][yi][x] (make_cols_maps_blocked): rows are (z, y) pairs, y = yb * YB + yi.
    //   blk = 0 : both sides natural (row = z * ny + y)
    //   blk = 1 : natural in, blocked out (forward sweep: the caller's array -> work array)
    //   blk = 2 : blocked in, natural out (inverse sweep); a tile is BR consecutive yi of one (yb, z) block
    int blk, blog, ny, nz;
    int max_ctas;               // > 0: upper bound of the grid (leaves SMs to a kernel running beside this one)
};

template <int N, int KIND, int NSTAGE>
__global__ void __launch_bounds__(PipeCfg<N>::THREADS) k_rows_pipe(RowsPipeArgs a)
{
    using C = PipeCfg<N>;
    constexpr int RG = C::ROWS_GAP;
    using PL = Planar<N, RG>;
    constexpr int BR = C::BR, G = C::G, P = C::ROWS_P, M = N / 2;
    static_assert(P >= PL::ROWS, "rows tile pitch");
    constexpr int J0 = (KIND == XF_DST) ? 1 : 0;
    constexpr int NW = C::THREADS / 32;
    static_assert(C::THREADS % 32 == 0, "whole warps");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    double* stage = sm;                               // [NSTAGE][BR * N] dense
    double* tile = stage + NSTAGE * BR * N;           // [BR][P]; DST: planar rows (Planar<N,0>)
    double* scr = tile + BR * P;
    double* SNs = scr + C::SCR;
    cd* WMs = reinterpret_cast<cd*>(SNs + (N / 2 + 2));
    double* SF1 = reinterpret_cast<double*>(WMs + N / 2);
    uint64_t* full = reinterpret_cast<uint64_t*>(SF1 + 2 * (N / 2 + 2));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int YB = 1 << a.blog, SUB = YB / BR > 0 ? YB / BR : 1;   // blk = 2: BR divides YB
    const int nyb = (a.ny + YB - 1) >> a.blog;
    const long long ntiles = (a.blk == 2) ? (long long)a.nz * nyb * SUB : (a.nrows + BR - 1) / BR;
    const double hs = 0.5 * a.scale, h2 = 0.5 * hs;

    // tile -> first input row (in rows of in_pitch), first natural row, valid rows
    auto locate = [&](long long t, long long& in_row0, long long& nat_row0, int& rows) {
        if (a.reverse) t = ntiles - 1 - t;
        if (a.blk == 2) {
            const int sub = (int)(t % SUB);
            const long long t2 = t / SUB;
            const int yb = (int)(t2 % nyb);
            const long long z = t2 / nyb;
            const int y0 = yb * YB + sub * BR;
            in_row0 = (((long long)yb * a.nz + z) << a.blog) + sub * BR;
            nat_row0 = z * a.ny + y0;
            rows = a.ny - y0 < BR ? a.ny - y0 : BR;
            if (rows < 0) rows = 0;
        } else {
            in_row0 = nat_row0 = t * BR;
            rows = (int)((a.nrows - nat_row0) < BR ? (a.nrows - nat_row0) : BR);
        }
    };
    // natural row -> output row (in rows of out_pitch)
    auto out_row = [&](long long row) -> long long {
        if (a.blk != 1) return row;
        const long long z = row / a.ny;
        const int y = (int)(row - z * a.ny);
        return ((((long long)(y >> a.blog)) * a.nz + z) << a.blog) + (y & (YB - 1));
    };

    auto issue = [&](long long t, int s) {
        long long in_row0, nat_row0; int rows;
        locate(t, in_row0, nat_row0, rows);
        const unsigned bytes = ((unsigned)rows * a.in_pitch * 8u) & ~15u;
        mbar_expect_tx(&full[s], bytes);
        if (bytes) bulk_load_1d(stage + s * BR * N, a.in + in_row0 * a.in_pitch, bytes, &full[s]);
    };

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) mbar_init(&full[s], 1);
        mbar_init_fence();
    }
    load_tables<N>(SNs, WMs, a.SN, a.WM);
    if constexpr (KIND == XF_DST) load_fold_table<N>(SF1, a.SN, hs);
    __syncthreads();
    pdl_wait();      // everything above is independent of the previous kernel's output (pdl.cuh)
    pdl_trigger();
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            long long t = blockIdx.x + (long long)s * gridDim.x;
            if (t < ntiles) issue(t, s);
        }
    }

    const int b = tid % BR, g = tid / BR;
    int it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
        const int s = it % NSTAGE;
        const unsigned parity = (it / NSTAGE) & 1;
        long long in_row0, row0; int rows;
        locate(t, in_row0, row0, rows);
        double* st = stage + s * BR * N;
        mbar_wait(&full[s], parity);
        {   // an odd tail (rows*pitch odd) leaves one double outside the 16-byte granularity of the bulk copy
            const long long cnt = (long long)rows * a.in_pitch;
            if ((cnt & 1) && tid == 0) st[cnt - 1] = a.in[in_row0 * a.in_pitch + cnt - 1];
            if (cnt & 1) __syncthreads();
        }
        // first touch: staging (dense, lanes along the row) -> compute tile; the DST fold happens here and
        // writes the planar layout (even / odd slots de-interleaved).  PARTS warps share a row when the CTA has
        // more warps than rows, so that no warp idles.
        constexpr int PARTS = NW > BR ? NW / BR : 1;
        for (int q = warp; q < BR * PARTS; q += NW) {
            const int r = q % BR, part = q / BR;
            const double* src = st + r * a.in_pitch;
            double* dst = tile + r * P;
            const bool ok = r < rows;
            if constexpr (KIND == XF_DST) {
                for (int j = 1 + lane + 32 * part; j < M; j += 32 * PARTS) {
                    double x1 = ok ? src[j - 1] : 0.0, x2 = ok ? src[N - j - 1] : 0.0;
                    double y1 = SF1[j] * (x1 + x2), y2 = h2 * (x1 - x2);
                    dst[prefold_row<N, RG, C::SWZ>(j)] = y1 + y2;
                    dst[prefold_row<N, RG, C::SWZ>(N - j)] = y1 - y2;
                }
                if (lane == 0 && part == 0) {
                    dst[prefold_row<N, RG, C::SWZ>(0)] = 0.0;
                    dst[prefold_row<N, RG, C::SWZ>(M)] = ok ? a.scale * src[M - 1] : 0.0;
                }
            } else {
                for (int j = lane + 32 * part; j < N; j += 32 * PARTS) dst[j] = ok ? src[j] : 0.0;
            }
        }
        fence_proxy_async();
        __syncthreads();
        {   // staging buffer s is free again: fetch the tile NSTAGE strides ahead
            const long long tn = t + (long long)NSTAGE * gridDim.x;
            if (tid == 0 && tn < ntiles) issue(tn, s);
        }
        if constexpr (KIND == XF_DST)
            dst_tile_fused<N, G, RG, true, C::SWZ>(tile + b * P, 1, g, hs, SNs, SF1, WMs, scr + b, BR,
                                                   OutTile<N, RG>{tile + b * P, 1});
        else
            xform_tile<N, G, KIND, true>(tile + b * P, 1, g, a.scale, SNs, WMs, scr + b, BR);
        for (int q = warp; q < BR * PARTS; q += NW) {
            const int r = q % BR, part = q / BR;
            if (r >= rows) continue;
            double* dst = a.out + out_row(row0 + r) * a.out_pitch;
            const double* src = tile + r * P;
#pragma unroll 4
            for (int x = lane + 32 * part; x < a.nvalid; x += 32 * PARTS) dst[x] = (KIND == XF_DST) ? src[PL::row(x + 1)] : src[x];
        }
        __syncthreads();   // the compute tile is rewritten by the next first touch
    }
}

// ---- host-side launchers ----------------------------------------------------------------------
// Preload mode: the launchers set up their kernel (which loads it on this device) and return without launching.
// The sharded handles run their whole launch sequence once in this mode at creation: with lazy module loading the
// first launch of a kernel may need a context synchronisation, which deadlocks behind a cross-GPU barrier kernel
// that spins until a peer driven by the same host thread arrives.
inline bool& preload_only()
{
    static thread_local bool v = false;
    return v;
}
int device_sm_count();
int current_device_slot();   // cudaGetDevice() clamped to [0, 63]

template <int N, int KIND, typename MID, int KIND2, typename OMAP = OutLinear>
inline cudaError_t launch_cols_pipe_t(const CUtensorMap& tm, const CUtensorMap& tm2, const ColsPipeArgs& a, const MID& mid,
                                      cudaStream_t st, const OMAP& omap = OMAP{})
{
    using C = PipeCfg<N, OMAP::sharded>;
    constexpr int NSTAGE = C::COLS_STAGES;
    auto kern = k_cols_pipe<N, KIND, MID, KIND2, NSTAGE, OMAP>;
    constexpr size_t smem = C::cols_smem(NSTAGE);
    static int per_sm_dev[64] = {0};     // function attributes are per device
    int& per_sm = per_sm_dev[current_device_slot()];
    if (!per_sm) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::THREADS_COLS, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
    }
    if (preload_only()) return cudaSuccess;
    const long long ntiles = (long long)((a.nb + C::B - 1) / C::B) * a.no;
    long long grid = (long long)device_sm_count() * per_sm;
    if (a.max_ctas > 0 && grid > a.max_ctas) grid = a.max_ctas;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) return cudaSuccess;
    return launch_pdl(kern, dim3((unsigned)grid), dim3(C::THREADS_COLS), smem, st, tm, tm2, a, mid, omap);
}

template <int N, int KIND>
inline cudaError_t launch_rows_pipe_t(const RowsPipeArgs& a, cudaStream_t st)
{
    using C = PipeCfg<N>;
    constexpr int NSTAGE = C::ROWS_STAGES;
    auto kern = k_rows_pipe<N, KIND, NSTAGE>;
    constexpr size_t smem = C::rows_smem(NSTAGE);
    static int per_sm_dev[64] = {0};     // function attributes are per device
    int& per_sm = per_sm_dev[current_device_slot()];
    if (!per_sm) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::THREADS, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
    }
    if (preload_only()) return cudaSuccess;
    const long long ntiles = (a.nrows + C::BR - 1) / C::BR;
    long long grid = (long long)device_sm_count() * per_sm;
    if (a.max_ctas > 0 && grid > a.max_ctas) grid = a.max_ctas;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) return cudaSuccess;
    return launch_pdl(kern, dim3((unsigned)grid), dim3(C::THREADS), smem, st, a);
}

}  // namespace fdmb
