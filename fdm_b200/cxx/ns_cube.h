// Drop-in replacement for the reference's src/ns_cube.h + src/ns_cube.cpp.
//
// Same class name, template parameters, constructor (const Config&), public state and step()
// as fdm::NSCube<T,check> (reference src/ns_cube.h:13-92, src/ns_cube.cpp:27-62).  The state
// lives on the device; the public tensors u,v,w,p,x,F,G,H,RHS are HOST mirrors with the
// reference's extents (src/ns_cube.h:66-75) so callers that read ns.u.vec or index
// ns.u[i][k][j] (test/test_ns_cube.cpp:39, src/velocity_plot.cpp:12-15) keep working.
//
// Mirror policy: with auto_sync (default) every step() first uploads the mirrors the caller changed on the
// host since they last agreed with the device (digest comparison) and ends with a download of u,v,w,p, which is
// exactly what the reference's callers can observe.  Long runs switch it off and call
// sync_to_host() before they look (plot interval), or sync_to_device() after they wrote a field.
#pragma once
#include <chrono>
#include <cmath>
#include <string>
#include <vector>

#if __has_include("config.h")
#include "config.h"               // the reference's own Config, unchanged
#else
#include "fdm_compat_config.h"
#endif
#include "lapl_cube.h"
#include "lapl_rect.h"     // the reference's ns_cube.h includes it (src/ns_cube.h:9); velocity_plot.h users rely on that

namespace fdm {

template <typename T, bool check>
class NSCube {
public:
    using tensor = fdm::tensor<T, 3, check>;

    const double x1, y1, z1;
    const double x2, y2, z2;
    const double U0;
    const double Re;
    const double dt;
    const int nx, ny, nz;
    int verbose;
    const double dx, dy, dz;
    const double dx2, dy2, dz2;

    tensor u, v, w;
    tensor p, x;
    tensor F, G, H, RHS;

    int time_index = 0;
    bool auto_sync = true;

    NSCube(const Config& c)
        : x1(c.get("ns", "x1", -M_PI)), y1(c.get("ns", "y1", -M_PI)), z1(c.get("ns", "z1", -M_PI)),
          x2(c.get("ns", "x2", M_PI)), y2(c.get("ns", "y2", M_PI)), z2(c.get("ns", "z2", M_PI)),
          U0(c.get("ns", "u0", 1.0)), Re(c.get("ns", "Re", 1.0)), dt(c.get("ns", "dt", 0.001)),
          nx(c.get("ns", "nx", 32)), ny(c.get("ns", "nx", 32) /* sic: src/ns_cube.h:58 */), nz(c.get("ns", "nz", 32)),
          verbose(c.get("ns", "verbose", 0)),
          dx((x2 - x1) / nx), dy((y2 - y1) / ny), dz((z2 - z1) / nz), dx2(dx * dx), dy2(dy * dy), dz2(dz * dz),
          u({0, nz + 1, 0, ny + 1, -1, nx + 1}), v({0, nz + 1, -1, ny + 1, 0, nx + 1}),
          w({-1, nz + 1, 0, ny + 1, 0, nx + 1}), p({0, nz + 1, 0, ny + 1, 0, nx + 1}),
          x({1, nz, 1, ny, 1, nx}), F({1, nz, 1, ny, 0, nx}), G({1, nz, 0, ny, 1, nx}), H({0, nz, 1, ny, 1, nx}),
          RHS({1, nz, 1, ny, 1, nx})
    {
        fdmb_ns_cube_params prm;
        FDMB_VERIFY(fdmb_ns_cube_default_params(&prm));
        prm.x1 = x1; prm.y1 = y1; prm.z1 = z1; prm.x2 = x2; prm.y2 = y2; prm.z2 = z2;
        prm.u0 = U0; prm.Re = Re; prm.dt = dt; prm.nx = nx; prm.nz = nz; prm.verbose = verbose;
        if constexpr (std::is_same<T, float>::value) {
            FDMB_VERIFY(fdmb_ns_cube_f32_create(&handle32, &prm));    // float fields + float pressure solve on the device
        } else {
            FDMB_VERIFY(fdmb_ns_cube_create(&handle, &prm));
        }
        tensor* all[9] = {&u, &v, &w, &p, &x, &F, &G, &H, &RHS};
        for (int id = 0; id < 9; id++) dig[id] = digest(all[id]->vec, (long long)all[id]->size);
    }
    ~NSCube()
    {
        if (handle) fdmb_ns_cube_destroy(handle);
        if (handle32) fdmb_ns_cube_f32_destroy(handle32);
    }
    NSCube(const NSCube&) = delete;
    NSCube& operator=(const NSCube&) = delete;

    void step() { steps(1); }
    // B200 extension: n steps back to back on the device, one host synchronisation at the end
    void steps(int n)
    {
        if (auto_sync) {      // upload what the caller changed on the host since the mirrors last agreed with the device
            push_if_changed(FDMB_FIELD_U, u); push_if_changed(FDMB_FIELD_V, v); push_if_changed(FDMB_FIELD_W, w);
            push_if_changed(FDMB_FIELD_P, p);
        }
        if constexpr (std::is_same<T, float>::value) {
            FDMB_VERIFY(fdmb_ns_cube_f32_step(handle32, n));
        } else {
            FDMB_VERIFY(fdmb_ns_cube_step(handle, n));
        }
        time_index += n;
        if (auto_sync) sync_to_host(false);
    }

    // device -> host mirrors; all = also the work fields x,F,G,H,RHS
    void sync_to_host(bool all = true)
    {
        pull(FDMB_FIELD_U, u); pull(FDMB_FIELD_V, v); pull(FDMB_FIELD_W, w); pull(FDMB_FIELD_P, p);
        if (all) { pull(FDMB_FIELD_X, x); pull(FDMB_FIELD_F, F); pull(FDMB_FIELD_G, G); pull(FDMB_FIELD_H, H); pull(FDMB_FIELD_RHS, RHS); }
    }
    // host mirrors -> device (after the caller assigned u,v,w,p on the host)
    void sync_to_device()
    {
        push(FDMB_FIELD_U, u); push(FDMB_FIELD_V, v); push(FDMB_FIELD_W, w); push(FDMB_FIELD_P, p);
    }
    fdmb_ns_cube* native_handle() const { return handle; }

private:
    fdmb_ns_cube* handle = nullptr;
    fdmb_ns_cube_f32* handle32 = nullptr;    // T = float
    std::vector<double> cvt;
    unsigned long long dig[9] = {};          // digest of each mirror when it last agreed with the device (by field id)

    // ---- host-mirror coherence ------------------------------------------------------------------------------
    // The reference's callers write the public tensors between steps (initial conditions, restarts: ns.u[i][k][j] = ...).  With auto_sync every step therefore starts by comparing a
    // 64-bit digest of each mirror with the digest taken when the mirror last agreed with the device, and uploads
    // the mirrors the caller changed; it ends with the download of u,v,w,p.  With auto_sync off the caller owns
    // coherence (sync_to_host / sync_to_device).
    static unsigned long long digest(const T* p, long long n)
    {
        unsigned long long h = 0x9E3779B97F4A7C15ull;
        const unsigned char* b = reinterpret_cast<const unsigned char*>(p);
        const long long bytes = n * (long long)sizeof(T), words = bytes / 8;
        const unsigned long long* q = reinterpret_cast<const unsigned long long*>(b);
        for (long long i = 0; i < words; i++) { h ^= q[i]; h *= 0x100000001B3ull; h ^= h >> 29; }
        for (long long i = words * 8; i < bytes; i++) { h ^= b[i]; h *= 0x100000001B3ull; }
        return h;
    }

    void push_if_changed(int id, tensor& t)
    {
        if (digest(t.vec, (long long)t.size) != dig[id]) push(id, t);
    }

    void pull(int id, tensor& t)
    {
        if constexpr (std::is_same<T, double>::value) {
            FDMB_VERIFY(fdmb_ns_cube_get_field(handle, id, t.vec));
        } else if constexpr (std::is_same<T, float>::value) {
            FDMB_VERIFY(fdmb_ns_cube_f32_get_field(handle32, id, t.vec));
        } else {
            cvt.resize((size_t)t.size);
            FDMB_VERIFY(fdmb_ns_cube_get_field(handle, id, cvt.data()));
            for (long long i = 0; i < (long long)t.size; i++) t.vec[i] = (T)cvt[i];
        }
        if (auto_sync) dig[id] = digest(t.vec, (long long)t.size);
    }
    void push(int id, tensor& t)
    {
        if constexpr (std::is_same<T, double>::value) {
            FDMB_VERIFY(fdmb_ns_cube_set_field(handle, id, t.vec));
        } else if constexpr (std::is_same<T, float>::value) {
            FDMB_VERIFY(fdmb_ns_cube_f32_set_field(handle32, id, t.vec));
        } else {
            cvt.assign(t.vec, t.vec + t.size);
            FDMB_VERIFY(fdmb_ns_cube_set_field(handle, id, cvt.data()));
        }
        dig[id] = digest(t.vec, (long long)t.size);
    }
};

}  // namespace fdm
