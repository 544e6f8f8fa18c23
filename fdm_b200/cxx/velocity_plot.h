// Drop-in replacement for the reference's src/velocity_plot.h + src/velocity_plot.cpp.
//
// Same class name, template parameters, constructor, use()/update()/plot()/vtk_out()/set_labels() as
// fdm::velocity_plotter<T,check,F> (reference src/velocity_plot.h:11-143); bodies call the C ABI of
// include/fdm_b200.h (fdmb_vplot_*): the slices, right-hand sides, stream-function solves and the VTK cell
// velocities are computed on the device.
//
//   use(T* u, T* v, T* w)   host arrays, as in the reference (src/velocity_plot.cpp:10-14): re-read by update()
//   use(ns)                 B200 extension: an NSCube / NSCyl drop-in object; its DEVICE state is read in place,
//                           nothing but the 2-D results crosses PCIe
// The slices the reference keeps private (src/velocity_plot.h:32-44) are public host mirrors here, refreshed by
// update().  plot() needs plplot in the reference (src/matrix_plot.h); here it writes the same six panels
// (U, V, W mid-plane slices and the three stream functions, src/velocity_plot.cpp:70-114) as one binary PPM
// image next to the requested name (extension replaced by .ppm), without contour lines.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "lapl_rect.h"     // tensor / compat tensor, FDMB_VERIFY, the reference header includes it too

namespace fdm {

template <typename T, bool check, typename F = tensor_flags<>>
class velocity_plotter {
    static constexpr tensor_flag zflag = F::head;
    static constexpr tensor_flag yflag = F::tail::head;
    static constexpr bool zper = has_tensor_flag(zflag, tensor_flag::periodic);
    static constexpr bool yper = has_tensor_flag(yflag, tensor_flag::periodic);

public:
    using matrix_x = tensor<T, 2, check, typename short_flags<zflag, yflag>::value>;
    using matrix_y = tensor<T, 2, check, typename short_flags<zflag>::value>;
    using matrix_z = tensor<T, 2, check, typename short_flags<yflag>::value>;

    const int nx, ny, nz;
    const bool cyl;
    const double dx, dy, dz;
    const double xx1, yy1, zz1;
    const double xx2, yy2, zz2;
    const int y_, y0, y1, yn, ynn;
    const int z_, z0, z1, zn, znn;

    matrix_x RHS_x; matrix_y RHS_y; matrix_z RHS_z;
    matrix_x psi_x; matrix_y psi_y; matrix_z psi_z;
    matrix_x vx, wx;
    matrix_y uy, wy;
    matrix_z uz, vz;

    std::string Xlabel = "X", Ylabel = "Y", Zlabel = "Z";

    velocity_plotter(double dx, double dy, double dz, int nx, int ny, int nz, double xx1, double xx2, double yy1,
                     double yy2, double zz1, double zz2, bool cyl = false)
        : nx(nx), ny(ny), nz(nz), cyl(cyl), dx(dx), dy(dy), dz(dz), xx1(xx1), yy1(yy1), zz1(zz1), xx2(xx2), yy2(yy2),
          zz2(zz2),
          y_(yper ? 0 : -1), y0(0), y1(yper ? 0 : 1), yn(yper ? ny - 1 : ny), ynn(yper ? ny - 1 : ny + 1),
          z_(zper ? 0 : -1), z0(0), z1(zper ? 0 : 1), zn(zper ? nz - 1 : nz), znn(zper ? nz - 1 : nz + 1),
          RHS_x({z1, zn, y1, yn}), RHS_y({z1, zn, 1, nx}), RHS_z({y1, yn, 1, nx}),
          psi_x({z1, zn, y1, yn}), psi_y({z1, zn, 1, nx}), psi_z({y1, yn, 1, nx}),
          vx({z0, znn, y0, ynn}), wx({z0, znn, y0, ynn}), uy({z0, znn, 0, nx + 1}), wy({z0, znn, 0, nx + 1}),
          uz({y0, ynn, 0, nx + 1}), vz({y0, ynn, 0, nx + 1})
    {
        fdmb_vplot_params prm;
        prm.dx = dx; prm.dy = dy; prm.dz = dz; prm.nx = nx; prm.ny = ny; prm.nz = nz;
        prm.xx1 = xx1; prm.xx2 = xx2; prm.yy1 = yy1; prm.yy2 = yy2; prm.zz1 = zz1; prm.zz2 = zz2;
        prm.cyl = cyl ? 1 : 0; prm.zperiodic = zper ? 1 : 0; prm.yperiodic = yper ? 1 : 0;
        FDMB_VERIFY(fdmb_vplot_create(&handle, &prm));
        for (int f = 0; f < 3; f++) FDMB_VERIFY(fdmb_vplot_field_size(handle, f, &fsz[f]));
    }
    ~velocity_plotter() { if (handle) fdmb_vplot_destroy(handle); }
    velocity_plotter(const velocity_plotter&) = delete;
    velocity_plotter& operator=(const velocity_plotter&) = delete;

    void set_labels(const std::string& X, const std::string& Y, const std::string& Z) { Xlabel = X; Ylabel = Y; Zlabel = Z; }

    // host arrays with the extents of src/velocity_plot.h:97-99; the pointers are kept, update() re-reads them
    void use(T* u, T* v, T* w)
    {
        hu = u; hv = v; hw = w;
        if constexpr (std::is_same<T, double>::value) FDMB_VERIFY(fdmb_vplot_use_host(handle, u, v, w));
    }
    // B200 extension: read the device state of a drop-in NSCube / NSCyl object in place
    template <typename NS, typename = decltype(std::declval<NS&>().native_handle())>
    void use(NS& ns)
    {
        hu = hv = hw = nullptr;
        use_native(ns.native_handle());
    }

    void update()
    {
        if constexpr (!std::is_same<T, double>::value) {
            if (hu) {   // the device path is fp64: widen the caller's arrays
                const T* src[3] = {hu, hv, hw};
                for (int f = 0; f < 3; f++) cvt[f].assign(src[f], src[f] + fsz[f]);
                FDMB_VERIFY(fdmb_vplot_use_host(handle, cvt[0].data(), cvt[1].data(), cvt[2].data()));
            }
        }
        FDMB_VERIFY(fdmb_vplot_update(handle));
        pull(FDMB_SLICE_VX, vx); pull(FDMB_SLICE_WX, wx); pull(FDMB_SLICE_UY, uy); pull(FDMB_SLICE_WY, wy);
        pull(FDMB_SLICE_UZ, uz); pull(FDMB_SLICE_VZ, vz);
        pull(FDMB_SLICE_RHS_X, RHS_x); pull(FDMB_SLICE_RHS_Y, RHS_y); pull(FDMB_SLICE_RHS_Z, RHS_z);
        pull(FDMB_SLICE_PSI_X, psi_x); pull(FDMB_SLICE_PSI_Y, psi_y); pull(FDMB_SLICE_PSI_Z, psi_z);
    }

    void vtk_out(const std::string& name, int step) { FDMB_VERIFY(fdmb_vplot_vtk_out(handle, name.c_str(), step)); }

    // Six panels in the reference's order (src/velocity_plot.cpp:78-113): U, V, W / psi_z, psi_y, psi_x.
    void plot(const std::string& name, double t)
    {
        (void)t;
        struct panel { const T* a; int rows, cols; };
        const panel pn[6] = {{uz.vec, ynn - y0 + 1, nx + 2}, {vz.vec, ynn - y0 + 1, nx + 2}, {wy.vec, znn - z0 + 1, nx + 2},
                             {psi_z.vec, yn - y1 + 1, nx},   {psi_y.vec, zn - z1 + 1, nx},   {psi_x.vec, zn - z1 + 1, yn - y1 + 1}};
        int cw = 0, ch = 0;
        for (const panel& q : pn) { cw = std::max(cw, q.cols); ch = std::max(ch, q.rows); }
        const int gap = 4, W = 3 * cw + 4 * gap, H = 2 * ch + 3 * gap;
        std::vector<unsigned char> img((size_t)W * H * 3, 255);
        for (int s = 0; s < 6; s++) {
            const panel& q = pn[s];
            double m = 0;
            for (long long i = 0; i < (long long)q.rows * q.cols; i++) m = std::max(m, (double)std::abs(q.a[i]));
            const int ox = gap + (s % 3) * (cw + gap), oy = gap + (s / 3) * (ch + gap);
            for (int r = 0; r < q.rows; r++)
                for (int c = 0; c < q.cols; c++) {
                    // first index upwards, like a plot: row 0 at the bottom of the panel
                    const double x = m > 0 ? (double)q.a[(size_t)r * q.cols + c] / m : 0.0;   // [-1, 1]
                    unsigned char* px = &img[((size_t)(oy + q.rows - 1 - r) * W + ox + c) * 3];
                    const double a = std::min(1.0, std::abs(x));
                    px[0] = (unsigned char)(x >= 0 ? 255 : 255 * (1 - a));                   // blue - white - red
                    px[1] = (unsigned char)(255 * (1 - a));
                    px[2] = (unsigned char)(x <= 0 ? 255 : 255 * (1 - a));
                }
        }
        std::string out = name;
        const size_t dot = out.find_last_of('.');
        if (dot != std::string::npos && out.find('/', dot) == std::string::npos) out.resize(dot);
        out += ".ppm";
        if (FILE* f = fopen(out.c_str(), "wb")) {
            fprintf(f, "P6\n%d %d\n255\n", W, H);
            fwrite(img.data(), 1, img.size(), f);
            fclose(f);
        }
    }

    fdmb_vplot* native_handle() const { return handle; }

private:
    fdmb_vplot* handle = nullptr;
    long long fsz[3] = {0, 0, 0};
    T *hu = nullptr, *hv = nullptr, *hw = nullptr;
    std::vector<double> cvt[3], tmp;

    void use_native(fdmb_ns_cube* ns) { FDMB_VERIFY(fdmb_vplot_use_ns_cube(handle, ns)); }
    void use_native(fdmb_ns_cyl* ns) { FDMB_VERIFY(fdmb_vplot_use_ns_cyl(handle, ns)); }

    template <typename M>
    void pull(int id, M& m)
    {
        if constexpr (std::is_same<T, double>::value) {
            FDMB_VERIFY(fdmb_vplot_get_slice(handle, id, m.vec));
        } else {
            tmp.resize((size_t)m.size);
            FDMB_VERIFY(fdmb_vplot_get_slice(handle, id, tmp.data()));
            for (long long i = 0; i < (long long)m.size; i++) m.vec[i] = (T)tmp[i];
        }
    }
};

}  // namespace fdm
