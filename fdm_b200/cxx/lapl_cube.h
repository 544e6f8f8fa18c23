// Drop-in replacement for the reference's src/lapl_cube.h + src/lapl_cube.cpp.
//
// Same class name, template parameters, constructor and solve() signature as
// fdm::LaplCube<T,check,F> (reference src/lapl_cube.h:9-106); the body is the B200 path behind
// the C ABI of include/fdm_b200.h.  Callers (ut/ut_lapl_cube.cpp:95, test/nbody.cpp:76,323,
// src/ns_cube.h:35) recompile unchanged.  Errors the reference reports through verify()/abort()
// (src/verify.h:10-18, e.g. a non power-of-two transform length, src/fft.cpp:67) abort here too,
// with the C ABI's message.
//
// T = double runs natively on the device.  T = float is converted to double at the boundary
// (the device path is fp64 only in this round), so float callers keep working.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#if __has_include("tensor.h")
#include "tensor.h"               // the reference's own container header, unchanged
#else
#include "fdm_compat_tensor.h"
#endif
#include "fdm_b200.h"

namespace fdm {

inline void fdmb_verify(int rc, const char* what, const char* file, int line)
{
    if (rc != 0) {
        fprintf(stderr, "verify(%s == 0) failed at %s:%d: %s\n", what, file, line, fdmb_last_error());
        abort();
    }
}
#define FDMB_VERIFY(call) ::fdm::fdmb_verify((call), #call, __FILE__, __LINE__)

template <typename T, bool check, typename F = tensor_flags<>>
class LaplCube {
public:
    static constexpr tensor_flag zflag = F::head;
    static constexpr tensor_flag yflag = F::tail::head;
    static constexpr tensor_flag xflag = F::tail::tail::head;
    // only the all-Dirichlet and all-periodic flag sets are instantiated by the reference
    // (src/lapl_cube.cpp:174-182); mixed sets were never reachable there either
    static constexpr bool periodic = has_tensor_flag(zflag, tensor_flag::periodic);
    static_assert(has_tensor_flag(yflag, tensor_flag::periodic) == periodic &&
                  has_tensor_flag(xflag, tensor_flag::periodic) == periodic,
                  "LaplCube: all axes Dirichlet or all axes periodic");

    const double dx, dy, dz;
    const double dx2, dy2, dz2;
    const double lx, ly, lz;
    const double slx, sly, slz;
    const int z1, y1, x1;
    const int zn, yn, xn;
    const int zpoints, ypoints, xpoints;
    const int nx, ny, nz;
    const int mxdim;
    const std::array<int, 6> indices;

    LaplCube(double dx, double dy, double dz, double lx, double ly, double lz, int nx, int ny, int nz)
        : dx(dx), dy(dy), dz(dz), dx2(dx * dx), dy2(dy * dy), dz2(dz * dz), lx(lx), ly(ly), lz(lz),
          slx(std::sqrt(2. / lx)), sly(std::sqrt(2. / ly)), slz(std::sqrt(2. / lz)),
          z1(periodic ? 0 : 1), y1(periodic ? 0 : 1), x1(periodic ? 0 : 1),
          zn(periodic ? nz - 1 : nz), yn(periodic ? ny - 1 : ny), xn(periodic ? nx - 1 : nx),
          zpoints(periodic ? nz : nz + 1), ypoints(periodic ? ny : ny + 1), xpoints(periodic ? nx : nx + 1),
          nx(nx), ny(ny), nz(nz), mxdim(std::max({nx + 1, ny + 1, nz + 1})), indices({1, nz, 1, ny, 1, nx})
    {
        if constexpr (std::is_same<T, float>::value) {
            // single precision all the way: float arrays, tables and butterflies on the device (fdmb_lapl_cube_f32_*)
            FDMB_VERIFY(fdmb_lapl_cube_f32_create(&handle32, dx, dy, dz, lx, ly, lz, nx, ny, nz, periodic ? 1 : 0));
        } else {
            FDMB_VERIFY(fdmb_lapl_cube_create(&handle, dx, dy, dz, lx, ly, lz, nx, ny, nz, periodic ? 1 : 0));
        }
    }
    ~LaplCube()
    {
        if (handle) fdmb_lapl_cube_destroy(handle);
        if (handle32) fdmb_lapl_cube_f32_destroy(handle32);
    }
    LaplCube(const LaplCube&) = delete;
    LaplCube& operator=(const LaplCube&) = delete;

    // ans, rhs: HOST pointers to contiguous interior-only arrays [nz][ny][nx] (src/lapl_cube.cpp:9-11)
    void solve(T* ans, T* rhs)
    {
        if constexpr (std::is_same<T, double>::value) {
            FDMB_VERIFY(fdmb_lapl_cube_solve(handle, ans, rhs));
        } else if constexpr (std::is_same<T, float>::value) {
            FDMB_VERIFY(fdmb_lapl_cube_f32_solve(handle32, ans, rhs));
        } else {
            const size_t n = (size_t)nx * ny * nz;
            cvt_in.assign(rhs, rhs + n);
            cvt_out.resize(n);
            FDMB_VERIFY(fdmb_lapl_cube_solve(handle, cvt_out.data(), cvt_in.data()));
            for (size_t i = 0; i < n; i++) ans[i] = (T)cvt_out[i];
        }
    }

    // B200 extension: device-resident solve on a caller-provided CUDA stream (asynchronous)
    void solve_device(double* d_ans, const double* d_rhs, void* stream = nullptr)
    {
        FDMB_VERIFY(fdmb_lapl_cube_solve_device(handle, d_ans, d_rhs, stream));
    }
    // B200 extension: `count` independent solves with HOST arrays, transfers of neighbouring solves overlapped
    // (fdmb_lapl_cube_solve_batch; page-locked arrays let the copies overlap; same answers as `count` solve() calls)
    void solve_batch(int count, double* const* ans, const double* const* rhs)
    {
        FDMB_VERIFY(fdmb_lapl_cube_solve_batch(handle, count, ans, rhs));
    }
    fdmb_lapl_cube* native_handle() const { return handle; }

private:
    fdmb_lapl_cube* handle = nullptr;
    fdmb_lapl_cube_f32* handle32 = nullptr;      // T = float
    std::vector<double> cvt_in, cvt_out;         // other element types: widened to fp64 at the boundary
};

}  // namespace fdm
