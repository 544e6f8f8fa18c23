// Batched 1-D real transforms on shared-memory tiles (sm_100a).
//
// Transforms provided (reference: src/fft.h:47-98, src/fft.cpp; closed forms in
// src/fft_fftw3.cpp:8-64 and src/asp_fft.cpp:308-319):
//   DST   -- FFT<T>::sFFT   : S[k] = scale * sum_{j=1}^{N-1} s[j] sin(pi k j / N)
//   PFWD  -- FFT<T>::pFFT_1 : S[k] = scale * sum s[j] cos(2 pi k j/N) (k=0..N/2),
//                             S[N-k] = scale * sum s[j] sin(2 pi k j/N) (k=1..N/2-1)
//   PINV  -- FFT<T>::pFFT   : S[j] = scale * (s[0]/2 + sum_{k=1}^{N/2-1}(s[k] cos + s[N-k] sin)
//                                             + (-1)^j s[N/2]/2)
//   DCT   -- FFT<T>::cFFT   : S[k] = scale * (s[0]/2 + sum_{j=1}^{N-1} s[j] cos(pi k j / N) + (-1)^k s[N]/2),
//                             k = 0..N (src/fft.cpp:368-445, definition src/asp_fft.cpp:404-418)
//
// This is NOT the reference's Samarskii-Nikolaev recursion.  Every transform is
// mapped onto one complex FFT of length M = N/2 held in shared memory:
//   DST : fold (sin-weighted symmetric + antisymmetric parts) -> real FFT(N) via
//         complex FFT(M) -> untangle -> prefix sum of the odd outputs
//   PFWD: complex FFT(M) of the even/odd packed sequence -> untangle
//   PINV: inverse untangle -> conj -> complex FFT(M) -> conj
// The complex FFT is an in-place mixed-radix (2/4/8/16) decimation-in-frequency
// transform with radix butterflies held in registers; its output is left in
// digit-reversed order and the consumers index through pos().
//
// Thread mapping: a tile holds B independent sequences ("columns"); element j of
// column b lives at tile[j*sj + b*sb].  Thread (b, g) works on column b; the G
// threads with the same b share the butterflies of that column.  Lanes of a warp
// run over b, so every twiddle is warp-uniform and every shared-memory access is
// conflict free when one of (sj, sb) is 1 and the other is odd or >= B.
#pragma once
#ifndef FDMB_HOST_EMUL   // tests/host_emul compiles this header for the CPU with thread-barrier shims
#include <cuda_runtime.h>
#endif

#ifndef FDMB_TW_POW
#define FDMB_TW_POW 1     // fused DST passes, N >= 1024: twiddle powers by multiplication instead of table loads
                          // (-5 % at 1023^3; at N <= 512 the extra registers cost a resident CTA)
#endif

#ifndef FDMB_TAB_ROT
#define FDMB_TAB_ROT 0    // fused DST, N >= 1024: the 16 fold sines of a first-pass butterfly and the 16 untangle
                          // cos / sin of a last-pass unit are rotations by multiples of pi/8 of ONE table pair, so they
                          // cost 4 + 2 shared-memory loads and a few FMAs instead of 32 loads (every 64-bit load is two
                          // wavefronts of the shared-memory pipe, even when it is a broadcast).  Measured at 1023^3
                          // (r02e): 23.61 vs 23.44 ms -- the solve runs into the board's power cap (SM clock 1.79 of
                          // 1.97 GHz) and the extra fp64 work costs what the saved loads gain.  Off.
#endif

namespace fdmb {

enum XformKind { XF_DST = 0, XF_PFWD = 1, XF_PINV = 2, XF_DCT = 3 };

// complex value of real type T (T = double everywhere on the graded path; float for the fp32 instantiations of the
// reference, src/lapl_cube.cpp:176-177, which run the plain tile transforms below in single precision)
template <typename T> struct cx { T x, y; };
using cd = cx<double>;
template <typename T> __device__ __forceinline__ cx<T> operator+(cx<T> a, cx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> __device__ __forceinline__ cx<T> operator-(cx<T> a, cx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> __device__ __forceinline__ cx<T> cmul(cx<T> a, cx<T> w) { return {a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x}; }
template <typename T> __device__ __forceinline__ cx<T> mul_mi(cx<T> a) { return {a.y, -a.x}; }  // a * (-i)

// compile-time loop: f(std::integral_constant<int, i>) for i = 0..n-1
template <int I_> struct IntC { static constexpr int value = I_; };
template <int n, int i = 0, typename F> __device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (i < n) { f(IntC<i>{}); static_for<n, i + 1>(f); }
}

// ---- register butterflies, forward (e^{-2 pi i jk/R}), natural in / natural out ----
template <int R> struct Dft;

template <> struct Dft<2> {
    template <typename T> static __device__ __forceinline__ void run(cx<T>* v) {
        cx<T> a = v[0] + v[1], b = v[0] - v[1];
        v[0] = a; v[1] = b;
    }
};
template <> struct Dft<4> {
    template <typename T> static __device__ __forceinline__ void run(cx<T>* v) {
        cx<T> t0 = v[0] + v[2], t1 = v[0] - v[2], t2 = v[1] + v[3], t3 = mul_mi(v[1] - v[3]);
        v[0] = t0 + t2; v[2] = t0 - t2; v[1] = t1 + t3; v[3] = t1 - t3;
    }
};
template <> struct Dft<8> {
    template <typename T> static __device__ __forceinline__ void run(cx<T>* v) {
        constexpr T h = T(0.70710678118654752440);
        cx<T> u[4], w[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { u[j] = v[j] + v[j + 4]; w[j] = v[j] - v[j + 4]; }
        w[1] = {h * (w[1].x + w[1].y), h * (w[1].y - w[1].x)};
        w[2] = mul_mi(w[2]);
        w[3] = {h * (w[3].y - w[3].x), -h * (w[3].x + w[3].y)};
        Dft<4>::run(u); Dft<4>::run(w);
#pragma unroll
        for (int k = 0; k < 4; k++) { v[2 * k] = u[k]; v[2 * k + 1] = w[k]; }
    }
};
template <> struct Dft<16> {
    template <typename T> static __device__ __forceinline__ void run(cx<T>* v) {
        constexpr T h = T(0.70710678118654752440);
        constexpr T c1 = T(0.92387953251128675613);  // cos(pi/8)
        constexpr T s1 = T(0.38268343236508977173);  // sin(pi/8)
        cx<T> u[8], w[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { u[j] = v[j] + v[j + 8]; w[j] = v[j] - v[j + 8]; }
        w[1] = cmul(w[1], cx<T>{c1, -s1});
        w[2] = {h * (w[2].x + w[2].y), h * (w[2].y - w[2].x)};
        w[3] = cmul(w[3], cx<T>{s1, -c1});
        w[4] = mul_mi(w[4]);
        w[5] = cmul(w[5], cx<T>{-s1, -c1});
        w[6] = {h * (w[6].y - w[6].x), -h * (w[6].x + w[6].y)};
        w[7] = cmul(w[7], cx<T>{-c1, -s1});
        Dft<8>::run(u); Dft<8>::run(w);
#pragma unroll
        for (int k = 0; k < 8; k++) { v[2 * k] = u[k]; v[2 * k + 1] = w[k]; }
    }
};

// Twiddles of one radix-R butterfly: v[k1] *= w^k1, k1 = 1..R-1, from the single table entry w = w^1.
// A warp holds several transform positions, so every twiddle load costs up to four shared-memory wavefronts;
// the powers are cheaper as fp64 multiplies (a few ulp, far inside the 1e-12 parity bar).
template <int R>
__device__ __forceinline__ void apply_twiddle_powers(cd* v, cd w)
{
    cd p[R];
    p[1] = w;
#pragma unroll
    for (int k = 2; k < R; k++) p[k] = (k & 1) ? cmul(p[k - 1], w) : cmul(p[k / 2], p[k / 2]);
#pragma unroll
    for (int k = 1; k < R; k++) v[k] = cmul(v[k], p[k]);
}

// ---- radix plans: M = N/2 complex points, up to three passes -------------------
template <int N> struct Plan;
#define FDMB_PLAN(N_, R0_, R1_, R2_, G_)                                          \
    template <> struct Plan<N_> {                                                 \
        static constexpr int M = N_ / 2;                                          \
        static constexpr int R0 = R0_, R1 = R1_, R2 = R2_;                        \
        static constexpr int G = G_; /* default threads per column */             \
        static_assert(R0_ * R1_ * R2_ == N_ / 2, "radix plan");                   \
    };
FDMB_PLAN(4, 2, 1, 1, 1)
FDMB_PLAN(8, 4, 1, 1, 1)
FDMB_PLAN(16, 8, 1, 1, 1)
FDMB_PLAN(32, 4, 4, 1, 4)
FDMB_PLAN(64, 8, 4, 1, 4)
FDMB_PLAN(128, 8, 8, 1, 8)
// radix <= 8 keeps the butterflies near 64 registers, so 24+ warps per SM stay resident
FDMB_PLAN(256, 16, 8, 1, 8)
FDMB_PLAN(512, 8, 8, 4, 32)
FDMB_PLAN(1024, 8, 8, 8, 64)
FDMB_PLAN(2048, 16, 8, 8, 64)
#undef FDMB_PLAN

// position of output bin k after the in-place DIF passes (digit reversal)
template <int N> __device__ __forceinline__ int fft_pos(int k)
{
    using P = Plan<N>;
    int pos = 0, L = P::M;
    { int d = k % P::R0; k /= P::R0; L /= P::R0; pos += d * L; }
    if constexpr (P::R1 > 1) { int d = k % P::R1; k /= P::R1; L /= P::R1; pos += d * L; }
    if constexpr (P::R2 > 1) { int d = k % P::R2; L /= P::R2; pos += d * L; }
    return pos;
}

// One in-place DIF pass over sub-blocks of length L with radix R.
// WM[t] = (cos(2 pi t/M), -sin(2 pi t/M)), t = 0..M-1.
template <int M, int L, int R, int G, typename T>
__device__ __forceinline__ void fft_pass(T* col, int sj, int g, const cx<T>* __restrict__ WM)
{
    constexpr int S = L / R;      // distance between butterfly legs (complex elements)
    constexpr int NBF = M / R;    // butterflies per column
    constexpr int IT = (NBF + G - 1) / G;
#pragma unroll
    for (int it = 0; it < IT; it++) {
        int q = g + it * G;
        if (NBF % G != 0 && q >= NBF) break;
        int blk = q / S, n2 = q % S;
        int base = blk * L + n2;
        cx<T> v[R];
#pragma unroll
        for (int n1 = 0; n1 < R; n1++) {
            int idx = 2 * (base + n1 * S);
            v[n1].x = col[idx * sj];
            v[n1].y = col[(idx + 1) * sj];
        }
        Dft<R>::run(v);
#pragma unroll
        for (int k1 = 0; k1 < R; k1++) {
            cx<T> o = v[k1];
            if (S > 1 && k1 > 0) {
                cx<T> w = WM[n2 * k1 * (M / L)];
                o = cmul(o, w);
            }
            int idx = 2 * (base + k1 * S);
            col[idx * sj] = o.x;
            col[(idx + 1) * sj] = o.y;
        }
    }
}

// Complex FFT of length M = N/2 on (col[2m*sj], col[(2m+1)*sj]); output digit-reversed.
// Ends with __syncthreads().
template <int N, int G, typename T>
__device__ __forceinline__ void fft_inplace(T* col, int sj, int g, const cx<T>* __restrict__ WM)
{
    using P = Plan<N>;
    constexpr int M = P::M;
    fft_pass<M, M, P::R0, G>(col, sj, g, WM);
    __syncthreads();
    if constexpr (P::R1 > 1) {
        fft_pass<M, M / P::R0, P::R1, G>(col, sj, g, WM);
        __syncthreads();
    }
    if constexpr (P::R2 > 1) {
        fft_pass<M, M / P::R0 / P::R1, P::R2, G>(col, sj, g, WM);
        __syncthreads();
    }
}

// Untangle one (k, M-k) pair of the half-length FFT of a real sequence.
// Y[k] = A_k - i B_k is the length-N DFT bin; returns A_k, B_k, A_{M-k}, B_{M-k}.
// SN[j] = sin(pi j / N), j = 0..N/2.
template <int N, typename T>
__device__ __forceinline__ void untangle(const T* col, int sj, int k, const T* __restrict__ SN,
                                         T scale, T& Ak, T& Bk, T& Am, T& Bm)
{
    constexpr int M = N / 2;
    int pk = fft_pos<N>(k), pm = fft_pos<N>((M - k) & (M - 1));
    cx<T> zk = {col[(2 * pk) * sj], col[(2 * pk + 1) * sj]};
    cx<T> zm = {col[(2 * pm) * sj], col[(2 * pm + 1) * sj]};
    T c = SN[M - 2 * k], s = SN[2 * k];      // cos, sin of 2 pi k / N
    T hs = T(0.5) * scale;
    T ex = hs * (zk.x + zm.x), ey = hs * (zk.y - zm.y);
    T ox = hs * (zk.y + zm.y), oy = -hs * (zk.x - zm.x);
    T wx = c * ox + s * oy, wy = c * oy - s * ox;
    Ak = ex + wx; Bk = -(ey + wy);
    Am = ex - wx; Bm = ey - wy;
}

// ---------------------------------------------------------------------------------
// DST-I over slots 1..N-1 of every column (slot 0 is the implicit zero boundary and
// is clobbered).  All threads of the CTA must call; ends with __syncthreads().
// scr: scratch of at least (G + G/8 + 1) * scr_s doubles per CTA, column b at scr[b].
// ---------------------------------------------------------------------------------
// PREFOLD: the caller already applied the fold below while staging the tile (k_rows_pipe).
template <int N, int G, bool PREFOLD = false, typename T>
__device__ __forceinline__ void dst_tile(T* col, int sj, int g, T scale,
                                         const T* __restrict__ SN, const cx<T>* __restrict__ WM,
                                         T* scr, int scr_s)
{
    constexpr int M = N / 2;
    static_assert(G <= M / 2 || M == 2, "too many threads per column");
    if constexpr (!PREFOLD) {
        // fold: y[j] = sin(pi j/N)(x[j]+x[N-j]) + (x[j]-x[N-j])/2
#pragma unroll
        for (int j = g + 1; j < M; j += G) {
            T a = col[j * sj], c = col[(N - j) * sj];
            T y1 = SN[j] * (a + c), y2 = T(0.5) * (a - c);
            col[j * sj] = y1 + y2;
            col[(N - j) * sj] = y1 - y2;
        }
        if (g == 0) { col[0] = T(0); col[M * sj] = T(2) * col[M * sj]; }
        __syncthreads();
    }

    fft_inplace<N, G>(col, sj, g, WM);

    // untangle into natural order: even slot 2k <- S[2k] = B_k, odd slot 2k+1 <- A_k
    constexpr int HP = (M / 2 >= G) ? (M / 2) / G : 1;   // pairs per thread
    T r[HP][4];
#pragma unroll
    for (int i = 0; i < HP; i++) {
        int k = g + i * G;
        if (k < M / 2)
            untangle<N>(col, sj, k, SN, scale, r[i][0], r[i][1], r[i][2], r[i][3]);
    }
    T zh_x = 0, zh_y = 0;   // bin M/2 (self-paired), handled by g == 0
    if (g == 0) {
        int ph = fft_pos<N>(M / 2);
        zh_x = scale * col[(2 * ph) * sj];
        zh_y = scale * col[(2 * ph + 1) * sj];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < HP; i++) {
        int k = g + i * G;
        if (k == 0) {
            col[0] = T(0);
            col[sj] = T(0.5) * r[i][0];             // A_0 / 2 seeds the running sum
        } else if (k < M / 2) {
            col[(2 * k) * sj] = r[i][1];
            col[(2 * k + 1) * sj] = r[i][0];
            col[(2 * (M - k)) * sj] = r[i][3];
            col[(2 * (M - k) + 1) * sj] = r[i][2];
        }
    }
    if (g == 0) {
        col[M * sj] = zh_y;                       // S[M]   = -Im Y[M/2] = Im Z[M/2]
        col[(M + 1) * sj] = zh_x;                 // A_{M/2} = Re Z[M/2]
    }
    __syncthreads();

    // inclusive prefix sum over the odd slots: S[2k+1] = sum_{m<=k} A'_m
    constexpr int CS = M / G;     // contiguous chunk per thread
    T a[CS];
    T run = T(0);
#pragma unroll
    for (int i = 0; i < CS; i++) {
        int k = g * CS + i;
        run += col[(2 * k + 1) * sj];
        a[i] = run;
    }
    T off = T(0);
    if constexpr (G > 1) {
        scr[g * scr_s] = run;
        __syncthreads();
        if constexpr (G <= 8) {
#pragma unroll
            for (int q = 0; q < G; q++) if (q < g) off += scr[q * scr_s];
        } else {
            T* scr2 = scr + G * scr_s;
            if ((g & 7) == 0) {
                T t = T(0);
#pragma unroll
                for (int q = 0; q < 8; q++) t += scr[(g + q) * scr_s];
                scr2[(g >> 3) * scr_s] = t;
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < G / 8; q++) if (q < (g >> 3)) off += scr2[q * scr_s];
#pragma unroll
            for (int q = 0; q < 8; q++) if (q < (g & 7)) off += scr[((g & ~7) + q) * scr_s];
        }
    }
#pragma unroll
    for (int i = 0; i < CS; i++) {
        int k = g * CS + i;
        col[(2 * k + 1) * sj] = a[i] + off;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------
// Periodic forward transform (pFFT_1) over slots 0..N-1.  Ends with __syncthreads().
// ---------------------------------------------------------------------------------
template <int N, int G, typename T>
__device__ __forceinline__ void pfwd_tile(T* col, int sj, int g, T scale,
                                          const T* __restrict__ SN, const cx<T>* __restrict__ WM)
{
    constexpr int M = N / 2;
    fft_inplace<N, G>(col, sj, g, WM);
    constexpr int HP = (M / 2 >= G) ? (M / 2) / G : 1;
    T r[HP][4];
#pragma unroll
    for (int i = 0; i < HP; i++) {
        int k = g + i * G;
        if (k < M / 2)
            untangle<N>(col, sj, k, SN, scale, r[i][0], r[i][1], r[i][2], r[i][3]);
    }
    T zh_x = 0, zh_y = 0;
    if (g == 0) {
        int ph = fft_pos<N>(M / 2);
        zh_x = scale * col[(2 * ph) * sj];
        zh_y = scale * col[(2 * ph + 1) * sj];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < HP; i++) {
        int k = g + i * G;
        if (k == 0) {
            col[0] = r[i][0];                       // S[0]   = Re Y[0]
            col[M * sj] = r[i][2];                  // S[N/2] = Re Y[M]
        } else if (k < M / 2) {
            col[k * sj] = r[i][0];                  // S[k]     = A_k
            col[(N - k) * sj] = r[i][1];            // S[N-k]   = B_k
            col[(M - k) * sj] = r[i][2];            // S[M-k]   = A_{M-k}
            col[(M + k) * sj] = r[i][3];            // S[N-(M-k)] = B_{M-k}
        }
    }
    if (g == 0 && M >= 2) {
        col[(M / 2) * sj] = zh_x;                   // S[N/4]   = Re Z[M/2]
        col[(N - M / 2) * sj] = zh_y;               // S[3N/4]  = Im Z[M/2]
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------
// Periodic inverse transform (pFFT) over slots 0..N-1.  Ends with __syncthreads().
// ---------------------------------------------------------------------------------
template <int N, int G, typename T>
__device__ __forceinline__ void pinv_tile(T* col, int sj, int g, T scale,
                                          const T* __restrict__ SN, const cx<T>* __restrict__ WM)
{
    constexpr int M = N / 2;
    constexpr int HP = (M / 2 >= G) ? (M / 2) / G : 1;
    // inverse untangle: build conj(Z[k]), Z[k] = Ze[k] + i Zo[k]
    T r[HP][4];
#pragma unroll
    for (int i = 0; i < HP; i++) {
        int k = g + i * G;
        if (k == 0) {
            T a0 = col[0], aM = col[M * sj];
            r[i][0] = a0 + aM; r[i][1] = -(a0 - aM);
        } else if (k < M / 2) {
            T ak = col[k * sj], bk = col[(N - k) * sj];
            T am = col[(M - k) * sj], bm = col[(M + k) * sj];
            // X_k = ak - i bk ; conj X_{M-k} = am + i bm
            T ex = ak + am, ey = -bk + bm;          // Ze = X_k + conj X_{M-k}
            T dx_ = ak - am, dy_ = -bk - bm;        // X_k - conj X_{M-k}
            T c = SN[M - 2 * k], s = SN[2 * k];     // e^{+2 pi i k/N} = c + i s
            T ox = dx_ * c - dy_ * s, oy = dx_ * s + dy_ * c;   // Zo
            // Z[k] = Ze + i Zo = (ex - oy) + i (ey + ox); Z[M-k] = conj(Ze) + i conj(Zo) = (ex + oy) + i(-ey + ox)
            r[i][0] = ex - oy; r[i][1] = -(ey + ox);     // conj Z[k]
            r[i][2] = ex + oy; r[i][3] = -(-ey + ox);    // conj Z[M-k]
        }
    }
    T zh_x = 0, zh_y = 0;
    if (g == 0 && M >= 2) {
        T a = col[(M / 2) * sj], b = col[(N - M / 2) * sj];
        zh_x = T(2) * a; zh_y = -T(2) * b;                 // conj(2a + 2ib)
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < HP; i++) {
        int k = g + i * G;
        if (k == 0) {
            col[0] = r[i][0]; col[sj] = r[i][1];
        } else if (k < M / 2) {
            col[(2 * k) * sj] = r[i][0]; col[(2 * k + 1) * sj] = r[i][1];
            col[(2 * (M - k)) * sj] = r[i][2]; col[(2 * (M - k) + 1) * sj] = r[i][3];
        }
    }
    if (g == 0 && M >= 2) { col[M * sj] = zh_x; col[(M + 1) * sj] = zh_y; }
    __syncthreads();

    fft_inplace<N, G>(col, sj, g, WM);

    // undo the digit reversal: y[2m] = Re F[pos(m)], y[2m+1] = -Im F[pos(m)]
    constexpr int CS = M / G;
    T o[CS][2];
    T hs = T(0.5) * scale;
#pragma unroll
    for (int i = 0; i < CS; i++) {
        int m = g + i * G;
        int p = fft_pos<N>(m);
        o[i][0] = hs * col[(2 * p) * sj];
        o[i][1] = -hs * col[(2 * p + 1) * sj];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < CS; i++) {
        int m = g + i * G;
        col[(2 * m) * sj] = o[i][0];
        col[(2 * m + 1) * sj] = o[i][1];
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------
// DCT-I with halved end points (cFFT) over slots 0..N of every column (N + 1 values).
// y[j] = (x[j]+x[N-j])/2 - sin(pi j/N)(x[j]-x[N-j]) -> real FFT(N): S[2k] = Re Y[k],
// S[2k+1] = S[2k-1] - Im Y[k], seeded with S[1] = (x[0]-x[N])/2 + sum x[j] cos(pi j/N),
// which the fold accumulates on the way.  Ends with __syncthreads().
// ---------------------------------------------------------------------------------
template <int N, int G, typename T>
__device__ __forceinline__ void dct_tile(T* col, int sj, int g, T scale,
                                         const T* __restrict__ SN, const cx<T>* __restrict__ WM,
                                         T* scr, int scr_s)
{
    constexpr int M = N / 2;
    static_assert(G <= M / 2 || M == 2, "too many threads per column");
    T part = T(0);
#pragma unroll
    for (int j = g + 1; j < M; j += G) {
        T a = col[j * sj], c = col[(N - j) * sj];
        T y1 = T(0.5) * (a + c), y2 = SN[j] * (a - c);
        col[j * sj] = y1 - y2;
        col[(N - j) * sj] = y1 + y2;
        part += SN[M - j] * (a - c);              // cos(pi j/N) x[j] + cos(pi (N-j)/N) x[N-j]
    }
    if (g == 0) {
        T x0 = col[0], xn = col[N * sj];
        col[0] = T(0.5) * (x0 + xn);
        part += T(0.5) * (x0 - xn);
    }
    // S[1] / scale, parked in scratch slot G until the untangle (keeps a register free across the FFT passes)
    if constexpr (G > 1) {
        scr[g * scr_s] = part;
        __syncthreads();
        if (g == 0) {
            T c1 = T(0);
#pragma unroll 8
            for (int q = 0; q < G; q++) c1 += scr[q * scr_s];
            scr[G * scr_s] = c1;
        }
    } else {
        scr[G * scr_s] = part;
    }
    __syncthreads();

    fft_inplace<N, G>(col, sj, g, WM);

    constexpr int HP = (M / 2 >= G) ? (M / 2) / G : 1;
    T r[HP][4];
#pragma unroll
    for (int i = 0; i < HP; i++) {
        int k = g + i * G;
        if (k < M / 2)
            untangle<N>(col, sj, k, SN, scale, r[i][0], r[i][1], r[i][2], r[i][3]);
    }
    T zh_x = 0, zh_y = 0;
    if (g == 0) {
        int ph = fft_pos<N>(M / 2);
        zh_x = scale * col[(2 * ph) * sj];
        zh_y = scale * col[(2 * ph + 1) * sj];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < HP; i++) {
        int k = g + i * G;
        if (k == 0) {
            col[0] = r[i][0];                         // S[0] = Re Y[0]
            col[N * sj] = r[i][2];                    // S[N] = Re Y[M]
            col[sj] = scale * scr[G * scr_s];         // S[1] seeds the running sum
        } else if (k < M / 2) {
            col[(2 * k) * sj] = r[i][0];
            col[(2 * k + 1) * sj] = r[i][1];
            col[(2 * (M - k)) * sj] = r[i][2];
            col[(2 * (M - k) + 1) * sj] = r[i][3];
        }
    }
    if (g == 0 && M >= 2) {
        col[M * sj] = zh_x;                           // S[M]            = Re Y[M/2]
        col[(M + 1) * sj] = zh_y;                     // S[M+1] - S[M-1] = -Im Y[M/2]
    }
    __syncthreads();

    // inclusive prefix sum over the odd slots
    constexpr int CS = M / G;
    T a[CS];
    T run = T(0);
#pragma unroll
    for (int i = 0; i < CS; i++) {
        int k = g * CS + i;
        run += col[(2 * k + 1) * sj];
        a[i] = run;
    }
    T off = T(0);
    if constexpr (G > 1) {
        scr[g * scr_s] = run;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < G; q++) if (q < g) off += scr[q * scr_s];
    }
#pragma unroll
    for (int i = 0; i < CS; i++) {
        int k = g * CS + i;
        col[(2 * k + 1) * sj] = a[i] + off;
    }
    __syncthreads();
}

template <int N, int G, int KIND, bool PREFOLD = false, typename T>
__device__ __forceinline__ void xform_tile(T* col, int sj, int g, T scale,
                                           const T* __restrict__ SN, const cx<T>* __restrict__ WM,
                                           T* scr, int scr_s)
{
    if constexpr (KIND == XF_DST) dst_tile<N, G, PREFOLD>(col, sj, g, scale, SN, WM, scr, scr_s);
    else if constexpr (KIND == XF_PFWD) pfwd_tile<N, G>(col, sj, g, scale, SN, WM);
    else if constexpr (KIND == XF_DCT) dct_tile<N, G>(col, sj, g, scale, SN, WM, scr, scr_s);
    else pinv_tile<N, G>(col, sj, g, scale, SN, WM);
}

// =====================================================================================
// Shared-memory-traffic-lean DST-I ("fused" variant) on a PLANAR tile.
//
// The sweeps are bound by the shared-memory data pipe (ncu: l1tex__data_pipe_lsu_wavefronts at
// 60-75 % of peak), so this variant touches the tile as little as possible and keeps every access
// free of bank conflicts, also for 8-column tiles (64-byte rows) where two rows of equal parity
// collide:
//   layout   even slots and odd slots of a sequence live in two regions of the tile
//            (Planar<N,GAP>: O[k] = slot 2k+1 at row k, E[k] = slot 2k at row M+GAP+k), which is
//            also the layout of the half-length complex FFT (z[m] = y[2m] + i y[2m+1]: re = E[m],
//            im = O[m]).  The strided-axis kernels land their tiles in this layout with two tensor
//            maps (odd rows / even rows of the global array).
//   stage A  fold + first radix pass: every thread reads its own and the mirrored inputs straight
//            from the landed tile, folds in registers, runs the radix-R0 butterfly, stores once
//   stage B  middle radix pass (if the plan has three passes)
//   stage C  last radix pass + untangle: a thread processes a frequency block together with its
//            mirror block (k <-> M-k live in the same thread), the untangled A_k/B_k overwrite the
//            registers that held Z[k]; even output slots (S[2k] = B_k) leave through the OUT policy
//            at once, the odd-slot seeds A'_k go back to the O region
//   stage D  running sum over the odd slots, emitted through the OUT policy
// SWZ (8-column tiles): complex element s is stored at s ^ ((s / L1) & 1), L1 = M / R0, and seed k at
// k ^ ((k / CS) & 1), so that the two thread groups sharing a half-warp always touch rows of different
// parity.  Both swizzles only change per-thread base addresses.
// The fold table SF[j] = (scale/2) sin(pi j/N) carries the transform's scale, so the untangle needs no
// multiplications by it.
// OUT policies decide where a finished spectral value goes: back into the tile (rows kernel, and the
// z sweep's first transform, scaled by the spectral multiplier) or directly to global memory with
// coalesced stores (strided-axis kernels), which removes the final tile write + read.
// Requirements: G divides M / R0 (a whole number of first-pass butterflies per thread) and the last pass has
// at most G units (block pairs).
// =====================================================================================

template <int N, int GAP> struct Planar {
    static constexpr int M = N / 2;
    static constexpr int ROWS = N + GAP + 1;                        // O: [0,M), E: [M+GAP, M+GAP+M]
    static __device__ __forceinline__ int O(int k) { return k; }               // slot 2k+1
    static __device__ __forceinline__ int E(int k) { return M + GAP + k; }     // slot 2k
    static __device__ __forceinline__ int row(int j) { return (j & 1) ? (j >> 1) : (M + GAP + (j >> 1)); }
};

template <int N, int GAP> struct OutTile {
    double* col; int sj;
    __device__ __forceinline__ void emit(int j, double v) const { col[Planar<N, GAP>::row(j) * sj] = v; }
};
// direct, coalesced store: lanes run over the contiguous axis, slot j -> row j-1 of the output
struct OutGlobal {
    double* dst; long long stride; bool ok;
    __device__ __forceinline__ void emit(int j, double v) const { if (ok) dst[(long long)(j - 1) * stride] = v; }
};
template <int N, int GAP, typename MID> struct OutMidTile {
    double* col; int sj; MID mid; typename MID::Ctx ctx; bool ok;
    __device__ __forceinline__ void emit(int j, double v) const
    {
        col[Planar<N, GAP>::row(j) * sj] = ok ? mid.apply(v, j, ctx) : 0.0;
    }
};

template <int N> struct PlanInfo {
    using P = Plan<N>;
    static constexpr int M = P::M;
    static constexpr int NP = (P::R2 > 1) ? 3 : ((P::R1 > 1) ? 2 : 1);
    static constexpr int RL = (NP == 3) ? P::R2 : P::R1;   // radix of the last pass (NP >= 2)
    static constexpr int LB = M / RL;                       // frequency blocks of the last pass
    static constexpr int L1 = M / P::R0;                    // sub-block length after the first pass
};

// One in-place DIF pass on the planar tile.  SWI / SWO: input / output stored with the block swizzle.
template <int N, int GAP, int L, int R, int G, bool SWI, bool SWO>
__device__ __forceinline__ void fft_pass_planar(double* col, int sj, int g, const cd* __restrict__ WM)
{
    using PL = Planar<N, GAP>;
    constexpr int M = N / 2, L1 = PlanInfo<N>::L1;
    constexpr int S = L / R;
    constexpr int NBF = M / R;
    constexpr int IT = (NBF + G - 1) / G;
    static_assert(!(SWI || SWO) || ((L == M && S == L1) || (L == L1 && S % 2 == 0)), "swizzled pass shape");
#pragma unroll
    for (int it = 0; it < IT; it++) {
        int q = g + it * G;
        if (NBF % G != 0 && q >= NBF) break;
        const int blk = q / S, n2 = q % S;
        // swizzle bit of element base + n1*S: first pass -> n1 & 1, later passes -> blk & 1
        const int cb = (L == M) ? 0 : (blk & 1);
        cd v[R];
#pragma unroll
        for (int n1 = 0; n1 < R; n1++) {
            const int bit = SWI ? ((L == M) ? (n1 & 1) : cb) : 0;
            const int s = blk * L + (n2 ^ bit) + n1 * S;
            v[n1].x = col[PL::E(s) * sj];
            v[n1].y = col[PL::O(s) * sj];
        }
        Dft<R>::run(v);
        if constexpr (S > 1 && R > 1) {
            if constexpr (FDMB_TW_POW && N >= 1024) apply_twiddle_powers<R>(v, WM[n2 * (M / L)]);
            else {
#pragma unroll
                for (int k1 = 1; k1 < R; k1++) v[k1] = cmul(v[k1], WM[n2 * k1 * (M / L)]);
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < R; k1++) {
            cd o = v[k1];
            const int bit = SWO ? ((L == M) ? (k1 & 1) : cb) : 0;
            const int s = blk * L + (n2 ^ bit) + k1 * S;
            col[PL::E(s) * sj] = o.x;
            col[PL::O(s) * sj] = o.y;
        }
    }
}

// cos / sin of n pi/8, n = 0..8
template <int n> struct Oct {
    static constexpr double c = (n == 0) ? 1.0 : (n == 1) ? 0.92387953251128675613 : (n == 2) ? 0.70710678118654752440
                              : (n == 3) ? 0.38268343236508977173 : (n == 4) ? 0.0 : (n == 5) ? -0.38268343236508977173
                              : (n == 6) ? -0.70710678118654752440 : (n == 7) ? -0.92387953251128675613 : -1.0;
    static constexpr double s = (n == 0 || n == 8) ? 0.0 : (n == 1 || n == 7) ? 0.38268343236508977173
                              : (n == 2 || n == 6) ? 0.70710678118654752440 : (n == 3 || n == 5) ? 0.92387953251128675613 : 1.0;
};
// sin(a + n pi/8) and cos(a + n pi/8) from sa = sin a, ca = cos a (any common scale factor carries through)
template <int n> __device__ __forceinline__ double rot_sin(double sa, double ca)
{
    if constexpr (n == 0) return sa;
    else if constexpr (n == 4) return ca;
    else if constexpr (n == 8) return -sa;
    else return sa * Oct<n>::c + ca * Oct<n>::s;
}
template <int n> __device__ __forceinline__ double rot_cos(double sa, double ca)
{
    if constexpr (n == 0) return ca;
    else if constexpr (n == 4) return -sa;
    else if constexpr (n == 8) return -ca;
    else return ca * Oct<n>::c - sa * Oct<n>::s;
}

// untangle with the pair's cos / sin of 2 pi k / N given
__device__ __forceinline__ void untangle_cs(cd& zk, cd& zm, double c, double s)
{
    const double ex = zk.x + zm.x, ey = zk.y - zm.y;
    const double ox = zk.y + zm.y, oy = zm.x - zk.x;
    const double wx = c * ox + s * oy, wy = c * oy - s * ox;
    zk.x = ex + wx; zk.y = -(ey + wy);
    zm.x = ex - wx; zm.y = ey - wy;
}

// Untangle one (k, M-k) pair of the (pre-scaled) half-length FFT in place:
// zk = Z[k], zm = Z[M-k], 0 < k < M/2  ->  zk = (A_k, B_k), zm = (A_{M-k}, B_{M-k}).
template <int N>
__device__ __forceinline__ void untangle_inplace(cd& zk, cd& zm, int k, const double* __restrict__ SN)
{
    constexpr int M = N / 2;
    const double c = SN[M - 2 * k], s = SN[2 * k];      // cos, sin of 2 pi k / N
    const double ex = zk.x + zm.x, ey = zk.y - zm.y;
    const double ox = zk.y + zm.y, oy = zm.x - zk.x;
    const double wx = c * ox + s * oy, wy = c * oy - s * ox;
    zk.x = ex + wx; zk.y = -(ey + wy);
    zm.x = ex - wx; zm.y = ey - wy;
}

// Row of folded element y[j] for the PREFOLD entry of dst_tile_fused (the rows kernel's first touch).
template <int N, int GAP, bool SWZ>
__device__ __forceinline__ int prefold_row(int j)
{
    int m = j >> 1;
    if (SWZ) m ^= (m / PlanInfo<N>::L1) & 1;
    return (j & 1) ? Planar<N, GAP>::O(m) : Planar<N, GAP>::E(m);
}

// col: this thread's sequence in the planar tile (element row r at col[r * sj]); on entry the tile
// holds the raw inputs x[1..N-1] (PREFOLD = false) or the folded, (scale/2)-scaled sequence y
// (PREFOLD = true; y[0] = 0, y[M] = scale * x[M]).  SF[j] = (scale/2) * SN[j], hs = scale / 2.
// SEP: finished values never go back into the working tile -- the seeds (and, for tile OUT policies, the
// spectral values) are written to a second tile `ocol` that nobody reads during stage C, so the stage-C
// results leave the registers at once (no barrier between the last loads and the first stores, half the
// live registers).  ocol is this thread's column base in that tile (same sj); ignored when !SEP.
// Barrier policy of the fused DST: the whole CTA (__syncthreads) or one consumer group of a CTA that runs several
// tiles at once (a named barrier over the group's threads, xform_ring.cuh).
struct CtaSync {
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
// Input policy of stage A: where slot j of the sequence is read from.
//   InPlanar : the planar tile itself (the strided-axis sweeps land their tiles in this layout)
//   InDense  : a dense row in natural order, slot j at row[j - 1] (the contiguous-axis sweep lands its rows with a
//              bulk copy); the planar tile may overlap the dense rows, because stage A reads every input into
//              registers before the barrier that precedes its first store
template <int N, int GAP> struct InPlanar {
    const double* col; int sj;
    __device__ __forceinline__ double e(int m) const { return col[Planar<N, GAP>::E(m) * sj]; }   // slot 2m
    __device__ __forceinline__ double o(int m) const { return col[Planar<N, GAP>::O(m) * sj]; }   // slot 2m + 1
};
struct InDense {
    const double* row; bool ok;
    __device__ __forceinline__ double e(int m) const { return ok ? row[2 * m - 1] : 0.0; }
    __device__ __forceinline__ double o(int m) const { return ok ? row[2 * m] : 0.0; }
};

template <int N, int G, int GAP, bool PREFOLD, bool SWZ, bool SEP = false, typename OUT, typename SYNC, typename IN>
__device__ __forceinline__ void dst_tile_fused_x(double* col, int sj, int g, double hs,
                                                 const double* __restrict__ SN, const double* __restrict__ SF,
                                                 const cd* __restrict__ WM, double* scr, int scr_s, const OUT& out,
                                                 double* ocol, const SYNC& sy, const IN& in)
{
    using P = Plan<N>;
    using I = PlanInfo<N>;
    using PL = Planar<N, GAP>;
    constexpr int M = N / 2, R0 = P::R0, S0 = M / R0;
    static_assert(I::NP >= 2, "fused DST needs at least two radix passes");
    static_assert(S0 % G == 0, "fused DST: a whole number of first-pass butterflies per thread");
    static_assert(!SWZ || I::NP == 3, "the block swizzle is defined for three-pass plans");
    constexpr int NA = S0 / G;            // first-pass butterflies per thread

    // ---- stage A -------------------------------------------------------------------------
    constexpr bool TABROT = FDMB_TAB_ROT && N >= 1024 && R0 == 8 && I::RL == 8;
    if constexpr (!PREFOLD) {
        cd v[NA][R0];
        const double h2 = 0.5 * hs;
#pragma unroll
        for (int it = 0; it < NA; it++) {
            const int q = g + it * G;
            // TABROT: slot 2m = 2q + n1 N/8 sits at the angle pi 2q/N + n1 pi/8 (SF[j] = hs sin(pi j/N), cos = SF[N/2 - j])
            double sa = 0, ca = 0, sb = 0, cb = 0;
            if constexpr (TABROT) { sa = SF[2 * q]; ca = SF[N / 2 - 2 * q]; sb = SF[2 * q + 1]; cb = SF[N / 2 - 2 * q - 1]; }
            if constexpr (TABROT) {
                auto leg = [&](auto n1c) {
                    constexpr int n1 = decltype(n1c)::value;
                    const int m = q + n1 * S0;
                    double a0 = (m == 0) ? 0.0 : in.e(m), c0 = (m == 0) ? 0.0 : in.e(M - m);
                    double a1 = in.o(m), c1 = in.o(M - m - 1);
                    const double s0 = rot_sin<n1>(sa, ca), s1 = rot_sin<n1>(sb, cb);
                    v[it][n1].x = s0 * (a0 + c0) + h2 * (a0 - c0);
                    v[it][n1].y = s1 * (a1 + c1) + h2 * (a1 - c1);
                };
                static_for<R0>(leg);
            } else {
#pragma unroll
                for (int n1 = 0; n1 < R0; n1++) {
                    const int m = q + n1 * S0;
                    // y[j] = hs * (sin(pi j/N)(x[j]+x[N-j]) + (x[j]-x[N-j])/2) for every j in 1..N-1, y[0] = 0;
                    // x[2m] = E[m], x[N-2m] = E[M-m], x[2m+1] = O[m], x[N-2m-1] = O[M-m-1]
                    double a0 = (m == 0) ? 0.0 : in.e(m), c0 = (m == 0) ? 0.0 : in.e(M - m);
                    double a1 = in.o(m), c1 = in.o(M - m - 1);
                    double s0 = (n1 < R0 / 2) ? SF[2 * m] : SF[N - 2 * m];
                    double s1 = (n1 < R0 / 2) ? SF[2 * m + 1] : SF[N - 2 * m - 1];
                    v[it][n1].x = s0 * (a0 + c0) + h2 * (a0 - c0);
                    v[it][n1].y = s1 * (a1 + c1) + h2 * (a1 - c1);
                }
            }
        }
        sy.sync();     // every mirrored read is done before anyone overwrites the inputs
#pragma unroll
        for (int it = 0; it < NA; it++) {
            const int q = g + it * G;
            Dft<R0>::run(v[it]);
            if constexpr (FDMB_TW_POW && N >= 1024) apply_twiddle_powers<R0>(v[it], WM[q]);
            else {
#pragma unroll
                for (int k1 = 1; k1 < R0; k1++) v[it][k1] = cmul(v[it][k1], WM[q * k1]);
            }
#pragma unroll
            for (int k1 = 0; k1 < R0; k1++) {
                cd o = v[it][k1];
                const int s = (q ^ (SWZ ? (k1 & 1) : 0)) + k1 * S0;
                col[PL::E(s) * sj] = o.x;
                col[PL::O(s) * sj] = o.y;
            }
        }
        sy.sync();
    } else {
        // the caller stored the folded sequence at prefold_row<N,GAP,SWZ>(j): already block-swizzled, so the
        // first pass stays in place (every thread rewrites exactly the rows it read)
        fft_pass_planar<N, GAP, M, R0, G, SWZ, SWZ>(col, sj, g, WM);
        sy.sync();
    }
    // ---- stage B -------------------------------------------------------------------------
    if constexpr (I::NP == 3) {
        fft_pass_planar<N, GAP, M / R0, P::R1, G, SWZ, SWZ>(col, sj, g, WM);
        sy.sync();
    }
    // ---- stage C: last pass on a block and its mirror block, untangle in registers -------------
    constexpr int RL = I::RL, LB = I::LB;
    constexpr int NU = LB / 2;            // units: {0, LB/2} and (u, LB-u), u = 1..LB/2-1
    constexpr int CS = M / G;             // odd-slot chunk per thread in stage D
    static_assert(LB >= 2 && NU <= G, "last pass: at most one unit per thread");
    static_assert(!SWZ || (CS % 2 == 0 && (LB / CS) % 2 == 0 && LB % CS == 0), "seed swizzle shape");
    const int u = g;
    const bool active = (NU == G) || (u < NU);
    const int lo = u, hi = (u == 0) ? LB / 2 : LB - u;
    cd va[RL], vb[RL];
    if (active) {
        const int ba = fft_pos<N>(lo), bb = fft_pos<N>(hi);
        const int ca = SWZ ? ((ba / I::L1) & 1) : 0, cb = SWZ ? ((bb / I::L1) & 1) : 0;
#pragma unroll
        for (int n1 = 0; n1 < RL; n1++) {
            const int sa = ba + (n1 ^ ca), sb = bb + (n1 ^ cb);
            va[n1].x = col[PL::E(sa) * sj]; va[n1].y = col[PL::O(sa) * sj];
            vb[n1].x = col[PL::E(sb) * sj]; vb[n1].y = col[PL::O(sb) * sj];
        }
        Dft<RL>::run(va);
        Dft<RL>::run(vb);
        if (u == 0) {
            // block 0 holds k = LB*d: d = 0 (DC), d = RL/2 (k = M/2), pairs (d, RL-d);
            // block LB/2 holds k = LB/2 + LB*d, pairs (d, RL-1-d)
            va[0].x = 2.0 * (va[0].x + va[0].y);     // A_0
            va[0].y = 0.0;
            va[RL / 2].x = 2.0 * va[RL / 2].x;        // A_{M/2} = Re Z[M/2]
            va[RL / 2].y = 2.0 * va[RL / 2].y;        // S[M]    = Im Z[M/2]
#pragma unroll
            for (int d = 1; d < RL / 2; d++) untangle_inplace<N>(va[d], va[RL - d], LB * d, SN);
#pragma unroll
            for (int d = 0; d < RL / 2; d++) untangle_inplace<N>(vb[d], vb[RL - 1 - d], LB / 2 + LB * d, SN);
        } else {
            // block lo: k = lo + LB*d; its partner M-k sits in block hi at d' = RL-1-d
            if constexpr (TABROT) {
                // 2 pi k/N = b + d pi/8 for k = lo + LB d, and (8 - d) pi/8 - b for k = hi + LB (RL-1-d), b = 2 pi lo/N
                const double sb_ = SN[2 * lo], cb_ = SN[M - 2 * lo];
                auto unit = [&](auto dc) {
                    constexpr int d = decltype(dc)::value;
                    if constexpr (d < RL / 2) untangle_cs(va[d], vb[RL - 1 - d], rot_cos<d>(sb_, cb_), rot_sin<d>(sb_, cb_));
                    else untangle_cs(vb[RL - 1 - d], va[d], rot_cos<8 - d>(-sb_, cb_), rot_sin<8 - d>(-sb_, cb_));
                };
                static_for<RL>(unit);
            } else {
#pragma unroll
                for (int d = 0; d < RL; d++) {
                    if (d < RL / 2) untangle_inplace<N>(va[d], vb[RL - 1 - d], lo + LB * d, SN);           // k < M/2
                    else untangle_inplace<N>(vb[RL - 1 - d], va[d], hi + LB * (RL - 1 - d), SN);
                }
            }
        }
    }
    if constexpr (!SEP) sy.sync();
    double* scol = SEP ? ocol : col;      // where the seeds live
    if (active) {
        // seed k lives at O[k ^ ((k / CS) & 1)] when swizzled: (k / CS) & 1 is a per-thread constant
        const int za = SWZ ? ((lo / CS) & 1) : 0, zb = SWZ ? ((hi / CS) & 1) : 0;
#pragma unroll
        for (int d = 0; d < RL; d++) {
            const int ka = lo + LB * d, kb = hi + LB * d;
            if (u == 0 && d == 0) {
                scol[PL::O(0) * sj] = 0.5 * va[0].x;         // A_0 / 2 seeds the running sum
            } else {
                out.emit(2 * ka, va[d].y);
                scol[PL::O((lo ^ za) + LB * d) * sj] = va[d].x;
            }
            out.emit(2 * kb, vb[d].y);
            scol[PL::O((hi ^ zb) + LB * d) * sj] = vb[d].x;
        }
    }
    sy.sync();
    // ---- stage D: inclusive prefix sum over the odd slots: S[2k+1] = sum_{m<=k} A'_m ------------
    double a[CS];
    double run = 0.0;
    {
        const int zg = SWZ ? (g & 1) : 0;
#pragma unroll
        for (int i = 0; i < CS; i++) {
            run += scol[PL::O(g * CS + (i ^ zg)) * sj];
            a[i] = run;
        }
    }
    double off = 0.0;
    if constexpr (G > 1) {
        scr[g * scr_s] = run;
        sy.sync();
        if constexpr (G <= 8) {
#pragma unroll
            for (int q = 0; q < G; q++) if (q < g) off += scr[q * scr_s];
        } else {
            double* scr2 = scr + G * scr_s;
            if ((g & 7) == 0) {
                double t = 0.0;
#pragma unroll
                for (int q = 0; q < 8; q++) t += scr[(g + q) * scr_s];
                scr2[(g >> 3) * scr_s] = t;
            }
            sy.sync();
#pragma unroll
            for (int q = 0; q < G / 8; q++) if (q < (g >> 3)) off += scr2[q * scr_s];
#pragma unroll
            for (int q = 0; q < 8; q++) if (q < (g & 7)) off += scr[((g & ~7) + q) * scr_s];
        }
    }
#pragma unroll
    for (int i = 0; i < CS; i++) {
        int k = g * CS + i;
        out.emit(2 * k + 1, a[i] + off);
    }
    sy.sync();
}

// the whole CTA works on one tile, inputs in the planar tile (the sweeps of xform_pipe.cuh)
template <int N, int G, int GAP, bool PREFOLD, bool SWZ, bool SEP = false, typename OUT>
__device__ __forceinline__ void dst_tile_fused(double* col, int sj, int g, double hs,
                                               const double* __restrict__ SN, const double* __restrict__ SF,
                                               const cd* __restrict__ WM, double* scr, int scr_s, const OUT& out,
                                               double* ocol = nullptr)
{
    dst_tile_fused_x<N, G, GAP, PREFOLD, SWZ, SEP>(col, sj, g, hs, SN, SF, WM, scr, scr_s, out, ocol, CtaSync{},
                                                   InPlanar<N, GAP>{col, sj});
}

}  // namespace fdmb
