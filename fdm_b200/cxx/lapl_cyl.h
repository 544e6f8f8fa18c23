// Drop-in replacement for the LaplCyl3FFT2 part of the reference's src/lapl_cyl.h + src/lapl_cyl.cpp.
//
// Same class name, template parameters, constructor, public geometry members and solve() signature
// as fdm::LaplCyl3FFT2<T,check,zflag> over fdm::LaplCyl3Data (reference src/lapl_cyl.h:12-37,172-249);
// the body calls the C ABI of include/fdm_b200.h.  The reference's LU factors (`matrices`, `ipivs`,
// src/lapl_cyl.h:235-236) do not exist here: the device path rebuilds the tridiagonal coefficients on
// the fly inside the r sweep.  The never-instantiated use_cyclic_reduction parameter
// (src/lapl_cyl.cpp:56-74,172-180) is accepted and ignored.
// T = float is converted to double at the boundary (the device path is fp64).
#pragma once
#include <array>
#include <cmath>
#include <vector>

#include "lapl_cube.h"     // tensor / compat tensor, FDMB_VERIFY

namespace fdm {

struct LaplCyl3Data {
    const double dr, dz, dphi;
    const double dr2, dz2, dphi2;
    const double r0, lr, lz;
    const double slz;
    const int zpoints;
    const int nr, nz, nphi;
    const int nrq;
    const int z1, zn;

    LaplCyl3Data(double dr, double dz, double r0, double lr, double lz, int nr, int nz, int nphi,
                 tensor_flag zflag = tensor_flag::none)
        : dr(dr), dz(dz), dphi(2 * M_PI / nphi), dr2(dr * dr), dz2(dz * dz), dphi2(dphi * dphi), r0(r0), lr(lr), lz(lz),
          slz(std::sqrt(2. / lz)), zpoints(zflag == tensor_flag::none ? nz + 1 : nz), nr(nr), nz(nz), nphi(nphi),
          nrq((int)std::ceil(std::log2(nr + 1))), z1(zflag == tensor_flag::none ? 1 : 0),
          zn(zflag == tensor_flag::none ? nz : nz - 1)
    {
    }
};

template <typename T, bool check, tensor_flag zflag = tensor_flag::none, bool use_cyclic_reduction = false>
class LaplCyl3FFT2 : public LaplCyl3Data {
public:
    constexpr static T SQRT_M_1_PI = 0.56418958354775629;
    std::array<int, 6> indices;

    LaplCyl3FFT2(double dr, double dz, double r0, double lr, double lz, int nr, int nz, int nphi)
        : LaplCyl3Data(dr, dz, r0, lr, lz, nr, nz, nphi, zflag), indices({0, nphi - 1, z1, zn, 1, nr})
    {
        FDMB_VERIFY(fdmb_lapl_cyl_create(&handle, dr, dz, r0, lr, lz, nr, nz, nphi, zflag == tensor_flag::none ? 0 : 1));
    }
    ~LaplCyl3FFT2() { if (handle) fdmb_lapl_cyl_destroy(handle); }
    LaplCyl3FFT2(const LaplCyl3FFT2&) = delete;
    LaplCyl3FFT2& operator=(const LaplCyl3FFT2&) = delete;

    // ans, rhs: HOST arrays [phi 0..nphi-1][z z1..zn][r 1..nr] (src/lapl_cyl.cpp:11-13)
    void solve(T* ans, T* rhs)
    {
        if constexpr (std::is_same<T, double>::value) {
            FDMB_VERIFY(fdmb_lapl_cyl_solve(handle, ans, rhs));
        } else {
            const size_t n = (size_t)nphi * (size_t)(zn - z1 + 1) * (size_t)nr;
            cvt_in.assign(rhs, rhs + n);
            cvt_out.resize(n);
            FDMB_VERIFY(fdmb_lapl_cyl_solve(handle, cvt_out.data(), cvt_in.data()));
            for (size_t i = 0; i < n; i++) ans[i] = (T)cvt_out[i];
        }
    }
    void solve_device(double* d_ans, const double* d_rhs, void* stream = nullptr)
    {
        FDMB_VERIFY(fdmb_lapl_cyl_solve_device(handle, d_ans, d_rhs, stream));
    }
    fdmb_lapl_cyl* native_handle() const { return handle; }

private:
    fdmb_lapl_cyl* handle = nullptr;
    std::vector<double> cvt_in, cvt_out;
};

}  // namespace fdm
