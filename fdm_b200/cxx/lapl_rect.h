// Drop-in replacement for the reference's src/lapl_rect.h + src/lapl_rect.cpp.
//
// Same class names, template parameters, constructors, public members and solve() signatures as
// fdm::LaplRect<T,check,F> (reference src/lapl_rect.h:11-78) and fdm::LaplRectFFT2<T,check,F>
// (src/lapl_rect.h:80-116); bodies call the C ABI of include/fdm_b200.h.  The public per-column scale
// vectors lm_y_scale, L_scale, U_scale (src/lapl_rect.h:57-59) stay writable host vectors: the
// cylindrical slice plotter overwrites them after construction (src/velocity_plot.h:113-127), so every
// solve() re-sends them when they changed since the previous one.
// T = float is converted to double at the boundary (the device path is fp64).
#pragma once
#include <cmath>
#include <vector>

#include "lapl_cube.h"     // tensor / compat tensor, FDMB_VERIFY

#if __has_include("asp_misc.h")
#include "asp_misc.h"      // the reference header includes it; its callers use asp::sq, asp::format through it
#endif

namespace fdm {

template <typename T, bool check, typename F = tensor_flags<>>
class LaplRect {
public:
    using matrix = tensor<T, 2, check, F>;
    const double dx, dy;
    const double dx2, dy2;
    const double lx, ly;
    const double slx, sly;
    const int nx, ny;
    const int y1, yn, ypoints;

    std::vector<T> lm_y_scale, L_scale, U_scale;

    LaplRect(double dx, double dy, double lx, double ly, int nx, int ny) : LaplRect(dx, dy, lx, ly, nx, ny, 0) {}
    ~LaplRect() { if (handle) fdmb_lapl_rect_destroy(handle); }
    LaplRect(const LaplRect&) = delete;
    LaplRect& operator=(const LaplRect&) = delete;

    // ans, rhs: HOST arrays [ny rows][nx], interior points only (src/lapl_rect.cpp:63-66)
    void solve(T* ans, T* rhs)
    {
        push_scales();
        const size_t n = (size_t)nx * (size_t)(yn - y1 + 1);
        if constexpr (std::is_same<T, double>::value) {
            FDMB_VERIFY(fdmb_lapl_rect_solve(handle, ans, rhs));
        } else {
            cvt_in.assign(rhs, rhs + n);
            cvt_out.resize(n);
            FDMB_VERIFY(fdmb_lapl_rect_solve(handle, cvt_out.data(), cvt_in.data()));
            for (size_t i = 0; i < n; i++) ans[i] = (T)cvt_out[i];
        }
    }
    fdmb_lapl_rect* native_handle() const { return handle; }

protected:
    static constexpr bool yper = has_tensor_flag(F::head, tensor_flag::periodic);
    static constexpr bool xper = has_tensor_flag(F::tail::head, tensor_flag::periodic);

    LaplRect(double dx, double dy, double lx, double ly, int nx, int ny, int kind)
        : dx(dx), dy(dy), dx2(dx * dx), dy2(dy * dy), lx(lx), ly(ly), slx(std::sqrt(2. / lx)), sly(std::sqrt(2. / ly)),
          nx(nx), ny(ny), y1(yper ? 0 : 1), yn(yper ? ny - 1 : ny), ypoints(yper ? ny : ny + 1),
          lm_y_scale(nx + 1, 1), L_scale(nx + 1, 1), U_scale(nx + 1, 1)
    {
        FDMB_VERIFY(fdmb_lapl_rect_create(&handle, kind, yper ? 1 : 0, xper ? 1 : 0, dx, dy, lx, ly, nx, ny));
        sent.assign(3 * (size_t)(nx + 1), 1.0);
    }
    void push_scales()
    {
        std::vector<double> cur(3 * (size_t)(nx + 1));
        for (int j = 0; j <= nx; j++) {
            cur[j] = (double)lm_y_scale[j]; cur[nx + 1 + j] = (double)L_scale[j]; cur[2 * (nx + 1) + j] = (double)U_scale[j];
        }
        if (cur != sent) {
            FDMB_VERIFY(fdmb_lapl_rect_set_scales(handle, cur.data(), cur.data() + nx + 1, cur.data() + 2 * (nx + 1)));
            sent.swap(cur);
        }
    }

    fdmb_lapl_rect* handle = nullptr;
    std::vector<double> sent, cvt_in, cvt_out;
};

template <typename T, bool check, typename F = tensor_flags<>>
class LaplRectFFT2 : public LaplRect<T, check, F> {
public:
    using base = LaplRect<T, check, F>;
    const int x1, xn, xpoints;

    LaplRectFFT2(double dx, double dy, double lx, double ly, int nx, int ny)
        : base(dx, dy, lx, ly, nx, ny, 1), x1(base::xper ? 0 : 1), xn(base::xper ? nx - 1 : nx),
          xpoints(base::xper ? nx : nx + 1)
    {
    }
    // solve() is the base class's: the handle was created as the two-transform kind
};

}  // namespace fdm
