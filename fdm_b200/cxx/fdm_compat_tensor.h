// Minimal stand-in for the reference container header (src/tensor.h) so that the drop-in
// class headers of this directory also build OUTSIDE the reference tree (tests, examples).
// Inside the reference tree the real "tensor.h" is found first and this file is unused.
//
// Same observable contract as fdm::tensor (src/tensor.h:64-289): per-axis inclusive [lo,hi]
// index ranges, row-major with the last index fastest, public `vec`/`size`, chained operator[],
// optional bounds check (abort), optional periodic wrap per axis, use(ptr) rebinding.
#pragma once
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace fdm {

enum class tensor_flag { none = 0, periodic = 1 };
constexpr bool has_tensor_flag(tensor_flag a, tensor_flag b) { return (static_cast<int>(a) & static_cast<int>(b)) != 0; }

template <tensor_flag... flags> struct tensor_flags;
template <> struct tensor_flags<> {
    static constexpr tensor_flag head = tensor_flag::none;
    using tail = tensor_flags<>;
};
template <tensor_flag first, tensor_flag... rest> struct tensor_flags<first, rest...> {
    static constexpr tensor_flag head = first;
    using tail = tensor_flags<rest...>;
};

// src/tensor.h:34-59: trailing `none` flags dropped, so that e.g. <periodic, none> names the same type as <periodic>
template <tensor_flag... flags> struct short_flags { using value = tensor_flags<flags...>; };
template <> struct short_flags<tensor_flag::periodic, tensor_flag::none> { using value = tensor_flags<tensor_flag::periodic>; };
template <> struct short_flags<tensor_flag::none, tensor_flag::none> { using value = tensor_flags<>; };
template <> struct short_flags<tensor_flag::none, tensor_flag::none, tensor_flag::none> { using value = tensor_flags<>; };
template <> struct short_flags<tensor_flag::none> { using value = tensor_flags<>; };

template <typename T, int rank, bool check = true, typename F = tensor_flags<>>
class tensor {
    template <int level, typename FF> struct cursor {
        T* base; const int* lo; const int* len; const long long* stride;
        decltype(auto) operator[](int i) const
        {
            int l = lo[level], n = len[level];
            if constexpr (has_tensor_flag(FF::head, tensor_flag::periodic)) i = ((i - l) % n + n) % n + l;
            if constexpr (check) {
                if (i < l || i >= l + n) {
                    fprintf(stderr, "verify(index in range) failed: axis %d index %d not in [%d,%d]\n", level, i, l, l + n - 1);
                    abort();
                }
            }
            T* p = base + (long long)(i - l) * stride[level];
            if constexpr (level + 1 == rank) return static_cast<T&>(*p);
            else return cursor<level + 1, typename FF::tail>{p, lo, len, stride};
        }
    };

public:
    std::array<int, rank * 2> offsets;
    int lo_[rank], len_[rank];
    long long stride_[rank];
    long long size;
    std::vector<T> storage;
    T* vec;

    explicit tensor(const std::array<int, rank * 2>& off, T* data = nullptr) : offsets(off)
    {
        long long s = 1;
        for (int a = rank - 1; a >= 0; a--) {
            lo_[a] = off[2 * a]; len_[a] = off[2 * a + 1] - off[2 * a] + 1;
            stride_[a] = s; s *= len_[a];
        }
        size = s;
        if (!data) { storage.assign((size_t)size, T(0)); vec = storage.data(); } else vec = data;
    }
    tensor(const tensor&) = delete;

    // copies the index-range INTERSECTION of the two tensors, axis by axis (src/tensor.h:103-111,276-279); this is how
    // the reference writes `p = x` (interior of p from x, src/ns_cube.cpp:275) and `u = ns.u` (test/test_ns_cyl.cpp:95)
    tensor& operator=(const tensor& other)
    {
        int from[rank], to[rank];
        for (int a = 0; a < rank; a++) {
            const int olo = other.lo_[a], ohi = other.lo_[a] + other.len_[a] - 1, hi = lo_[a] + len_[a] - 1;
            from[a] = lo_[a] > olo ? lo_[a] : olo;
            to[a] = hi < ohi ? hi : ohi;
            if (from[a] > to[a]) return *this;
        }
        int idx[rank];
        for (int a = 0; a < rank; a++) idx[a] = from[a];
        for (;;) {
            long long d = 0, s = 0;
            for (int a = 0; a < rank; a++) { d += (long long)(idx[a] - lo_[a]) * stride_[a]; s += (long long)(idx[a] - other.lo_[a]) * other.stride_[a]; }
            vec[d] = other.vec[s];
            int a = rank - 1;
            while (a >= 0 && ++idx[a] > to[a]) { idx[a] = from[a]; a--; }
            if (a < 0) break;
        }
        return *this;
    }

    decltype(auto) operator[](int i) { return cursor<0, F>{vec, lo_, len_, stride_}[i]; }
    // offset of element {z, y, x} from vec (src/tensor.h:258-260); periodic axes wrap, like operator[]
    int index(const std::array<int, rank>& indices)
    {
        long long off = 0;
        for (int a = 0; a < rank; a++) off += (long long)(wrap(a, indices[a]) - lo_[a]) * stride_[a];
        return (int)off;
    }
    void use(T* p) { vec = p; }
    T maxabs() const
    {
        T m = 0;
        for (long long i = 0; i < size; i++) { T a = std::abs(vec[i]); if (a > m) m = a; }
        return m;
    }
    // Euclidean norm (the reference calls BLAS nrm2, src/tensor.h:272-274)
    T norm2() const
    {
        long double s = 0;
        for (long long i = 0; i < size; i++) s += (long double)vec[i] * (long double)vec[i];
        return (T)std::sqrt((double)s);
    }

private:
    template <typename FF> static constexpr bool axis_periodic(int a)
    {
        if (a == 0) return has_tensor_flag(FF::head, tensor_flag::periodic);
        else return axis_periodic<typename FF::tail>(a - 1);
    }
    int wrap(int a, int i) const
    {
        if (axis_periodic<F>(a)) { const int l = lo_[a], n = len_[a]; i = ((i - l) % n + n) % n + l; }
        return i;
    }
};

}  // namespace fdm
