// Drop-in replacement for the reference's src/ns_cyl.h + src/ns_cyl.cpp.
//
// Same class name, template parameters, constructor (const Config&), public state, step(), L_step()
// and size() as fdm::NSCyl<T,check,zflag> (reference src/ns_cyl.h:17-132, src/ns_cyl.cpp:23-78).  The
// state lives on the device; the public tensors are HOST mirrors with the reference's extents
// (src/ns_cyl.h:80-93), so callers that alias ns.u.vec (test/test_ns_cyl_spectral.cpp:60-94 packs
// u,v,w,p into the ARPACK vector) keep working.  Mirror policy as in ns_cube.h: with auto_sync (default) a step
// first uploads the mirrors the caller changed on the host (and a changed U0) and ends with the download of
// u,v,w,p, so unmodified callers that write ns.u / ns.w0 / ns.U0 between steps behave like the reference; with
// auto_sync off the caller calls sync_to_host() / sync_to_device() itself.
// vrandom = 1 seeds v exactly like the reference (default-seeded std::default_random_engine, loop order
// of src/ns_cyl.h:99-108), on the host, and uploads it.
#pragma once
// (the system headers the reference's ns_cyl.h pulls in: its callers rely on them transitively)
#include <chrono>
#include <climits>
#include <cmath>
#include <random>
#include <string>
#include <vector>

#if __has_include("config.h")
#include "config.h"
#else
#include "fdm_compat_config.h"
#endif
#include "lapl_cyl.h"

#if __has_include("asp_misc.h")
#include "asp_misc.h"      // the reference header includes it; its callers use asp::sq, asp::format through it
#endif

namespace fdm {

template <typename T, bool check, tensor_flag zflag = tensor_flag::none>
class NSCyl {
public:
    using tensor_flags = typename fdm::short_flags<tensor_flag::periodic, zflag>::value;
    using tensor = fdm::tensor<T, 3, check, tensor_flags>;

    const double R, r0;
    const double h1, h2;
    double U0;
    const double Re;
    const double dt;
    const int nr, nz, nphi;
    const int verbose;
    const int z_, z0, z1, zn, znn;
    const double dr, dz, dphi;
    const double dr2, dz2, dphi2;

    tensor u, v, w, p;
    tensor u0, v0, w0;
    tensor x;
    tensor F, G, H, RHS;

    int time_index = 0;
    bool auto_sync = true;

    NSCyl(const Config& c)
        : R(c.get("ns", "R", M_PI)), r0(c.get("ns", "r", M_PI / 2)), h1(c.get("ns", "h1", 0)), h2(c.get("ns", "h2", 10)),
          U0(c.get("ns", "u0", 1.0)), Re(c.get("ns", "Re", 1.0)), dt(c.get("ns", "dt", 0.001)),
          nr(c.get("ns", "nr", 32)), nz(c.get("ns", "nz", 31)), nphi(c.get("ns", "nphi", 32)),
          verbose(c.get("ns", "verbose", 0)),
          z_(zflag == tensor_flag::none ? -1 : 0), z0(0), z1(zflag == tensor_flag::none ? 1 : 0),
          zn(zflag == tensor_flag::none ? nz : nz - 1), znn(zflag == tensor_flag::none ? nz + 1 : nz - 1),
          dr((R - r0) / nr), dz((h2 - h1) / nz), dphi(2 * M_PI / nphi), dr2(dr * dr), dz2(dz * dz), dphi2(dphi * dphi),
          u({0, nphi - 1, z0, znn, -1, nr + 1}), v({0, nphi - 1, z_, znn, 0, nr + 1}), w({0, nphi - 1, z0, znn, 0, nr + 1}),
          p({0, nphi - 1, z0, znn, 0, nr + 1}),
          u0({0, nphi - 1, z0, znn, -1, nr + 1}), v0({0, nphi - 1, z_, znn, 0, nr + 1}), w0({0, nphi - 1, z0, znn, 0, nr + 1}),
          x({0, nphi - 1, z1, zn, 1, nr}), F({0, nphi - 1, z1, zn, 0, nr}), G({0, nphi - 1, z0, zn, 1, nr}),
          H({0, nphi - 1, z1, zn, 1, nr}), RHS({0, nphi - 1, z1, zn, 1, nr})
    {
        fdmb_ns_cyl_params prm;
        FDMB_VERIFY(fdmb_ns_cyl_default_params(&prm));
        prm.R = R; prm.r = r0; prm.h1 = h1; prm.h2 = h2; prm.u0 = U0; prm.Re = Re; prm.dt = dt;
        prm.nr = nr; prm.nz = nz; prm.nphi = nphi; prm.verbose = verbose; prm.vrandom = 0;
        prm.zperiodic = zflag == tensor_flag::none ? 0 : 1;
        FDMB_VERIFY(fdmb_ns_cyl_create(&handle, &prm));
        if (c.get("ns", "vrandom", 0) == 1) {
            std::default_random_engine generator;
            std::uniform_real_distribution<T> distribution(-1e-3, 1e-3);
            for (int i = 0; i < nphi; i++)
                for (int k = z1; k <= zn; k++)
                    for (int j = 1; j <= nr; j++) v[i][k][j] = distribution(generator);
            push(FDMB_FIELD_V, v);
        }
        U0_dev = U0;
        tensor* all[12] = {&u, &v, &w, &p, &x, &F, &G, &H, &RHS, &u0, &v0, &w0};
        for (int id = 0; id < 12; id++) dig[id] = digest(all[id]->vec, (long long)all[id]->size);
    }
    ~NSCyl() { if (handle) fdmb_ns_cyl_destroy(handle); }
    NSCyl(const NSCyl&) = delete;
    NSCyl& operator=(const NSCyl&) = delete;

    int size() const { return (int)(u.size + v.size + w.size + p.size); }

    void step() { steps(1, false); }
    void L_step() { steps(1, true); }
    // B200 extension: n steps back to back on the device
    void steps(int n, bool linear = false)
    {
        before_step(linear);
        FDMB_VERIFY(linear ? fdmb_ns_cyl_lstep(handle, n) : fdmb_ns_cyl_step(handle, n));
        time_index += n;
        if (auto_sync) sync_to_host(false);
    }
    void sync_to_host(bool all = true)
    {
        pull(FDMB_FIELD_U, u); pull(FDMB_FIELD_V, v); pull(FDMB_FIELD_W, w); pull(FDMB_FIELD_P, p);
        if (all) { pull(FDMB_FIELD_X, x); pull(FDMB_FIELD_F, F); pull(FDMB_FIELD_G, G); pull(FDMB_FIELD_H, H); pull(FDMB_FIELD_RHS, RHS); }
    }
    // host mirrors -> device: the state u,v,w,p, the linearisation point u0,v0,w0 and the wall speed U0
    void sync_to_device()
    {
        push(FDMB_FIELD_U, u); push(FDMB_FIELD_V, v); push(FDMB_FIELD_W, w); push(FDMB_FIELD_P, p);
        push(FDMB_FIELD_U0, u0); push(FDMB_FIELD_V0, v0); push(FDMB_FIELD_W0, w0);
        FDMB_VERIFY(fdmb_ns_cyl_set_u0(handle, U0));
        U0_dev = U0;
    }
    fdmb_ns_cyl* native_handle() const { return handle; }

private:
    fdmb_ns_cyl* handle = nullptr;
    std::vector<double> cvt;
    double U0_dev = 0;                       // the wall speed the device was last given
    unsigned long long dig[12] = {};         // digest of each mirror when it last agreed with the device (by field id)

    // ---- host-mirror coherence ------------------------------------------------------------------------------
    // The reference's callers write the public tensors between steps (test/test_ns_cyl_spectral.cpp assigns
    // ns.u = u; ... and ns.w0[...] before L_step()).  With auto_sync every step therefore starts by comparing a
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
    void before_step(bool linear)
    {
        if (U0 != U0_dev) { FDMB_VERIFY(fdmb_ns_cyl_set_u0(handle, U0)); U0_dev = U0; }
        if (!auto_sync) return;
        push_if_changed(FDMB_FIELD_U, u); push_if_changed(FDMB_FIELD_V, v); push_if_changed(FDMB_FIELD_W, w);
        push_if_changed(FDMB_FIELD_P, p);
        if (linear) { push_if_changed(FDMB_FIELD_U0, u0); push_if_changed(FDMB_FIELD_V0, v0); push_if_changed(FDMB_FIELD_W0, w0); }
    }

    void pull(int id, tensor& t)
    {
        if constexpr (std::is_same<T, double>::value) {
            FDMB_VERIFY(fdmb_ns_cyl_get_field(handle, id, t.vec));
        } else {
            cvt.resize((size_t)t.size);
            FDMB_VERIFY(fdmb_ns_cyl_get_field(handle, id, cvt.data()));
            for (long long i = 0; i < (long long)t.size; i++) t.vec[i] = (T)cvt[i];
        }
        if (auto_sync) dig[id] = digest(t.vec, (long long)t.size);
    }
    void push(int id, tensor& t)
    {
        if constexpr (std::is_same<T, double>::value) {
            FDMB_VERIFY(fdmb_ns_cyl_set_field(handle, id, t.vec));
        } else {
            cvt.assign(t.vec, t.vec + t.size);
            FDMB_VERIFY(fdmb_ns_cyl_set_field(handle, id, cvt.data()));
        }
        dig[id] = digest(t.vec, (long long)t.size);
    }
};

}  // namespace fdm
