/*
 * fdm_b200 -- C ABI of the B200-native fast Poisson / Navier-Stokes projection path.
 *
 * This header is the drop-in boundary.  The reference (resetius/fdm) has no FFI
 * or plugin registry: its boundary is a set of C++ class templates explicitly
 * instantiated in libfdm.  Each group of entry points below replaces the body
 * of one of those classes; the header-compatible C++ shims in fdm_b200/cxx/
 * (namespace fdm, same class names, constructor and method signatures) call
 * these functions, so existing callers recompile unchanged.
 *
 * Conventions
 *   - plain C, no CUDA or torch types in any signature; `stream` arguments are a
 *     cudaStream_t passed as void* (NULL = the handle's own stream).
 *   - every function returns 0 on success or a negative FDMB_ERR_* code;
 *     fdmb_last_error() returns a thread-local description.  The reference
 *     aborts via verify() (src/verify.h:10-18) where we return an error; the
 *     C++ shims turn non-zero codes back into that abort.
 *   - arrays follow the reference layout: row-major, last index fastest,
 *     interior points only, contiguous (src/tensor.h:207-219).
 *   - `*_solve` / `*_step` take HOST pointers exactly like the reference
 *     methods and copy across PCIe/NVLink-C2C internally; `*_device` variants
 *     take device pointers and are asynchronous on `stream`.
 *   - a handle owns one CUDA stream and its scratch; calls on one handle are
 *     not re-entrant (same as the reference, whose solvers share member scratch).
 *   - there is no CPU fallback: without a CUDA device every create fails.
 */
#ifndef FDM_B200_H
#define FDM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FDMB_OK 0
#define FDMB_ERR_INVALID (-1)  /* bad argument / unsupported size (reference: verify((1<<n)==N), src/fft.cpp:67) */
#define FDMB_ERR_CUDA (-2)     /* CUDA runtime error */
#define FDMB_ERR_NOMEM (-3)
#define FDMB_ERR_COMM (-4)     /* multi-GPU exchange error */

/* ---- library ------------------------------------------------------------------ */
const char* fdmb_last_error(void);
int fdmb_version(void);
int fdmb_device_count(void);
int fdmb_set_device(int device);
/* number of kernels this library has launched so far in this process */
unsigned long long fdmb_launch_count(void);

/* Per-launch CUDA-event timing of this library's kernels (used by bench.py for the roofline).
 * begin() starts recording; end() synchronises and writes one line per kernel tag:
 * "<tag> <launches> <total_ms>\n".                                                   */
int fdmb_profile_begin(void);
int fdmb_profile_end(char* buf, int buflen);

/* raw device memory helpers for hosts that have no CUDA runtime of their own */
int fdmb_malloc(void** dptr, unsigned long long bytes);
int fdmb_free(void* dptr);
int fdmb_memcpy_h2d(void* dst, const void* src, unsigned long long bytes);
int fdmb_memcpy_d2h(void* dst, const void* src, unsigned long long bytes);
int fdmb_device_synchronize(void);

/* ---- 1-D transforms ------------------------------------------------------------
 * Batched fdm::FFT<double>::{sFFT,pFFT_1,pFFT,cFFT}  (src/fft.h:82-90, src/fft.cpp:109-212,294-365,368-445).
 * kind: 0 sFFT (DST-I over indices 1..N-1), 1 pFFT_1 (periodic, values->coefficients),
 *       2 pFFT (periodic, coefficients->values), 3 cFFT (DCT-I with halved end points over indices 0..N).
 * in/out: HOST arrays of `batch` rows; a row holds the N-1 (kind 0), N (kind 1,2) or N+1 (kind 3)
 * meaningful entries, contiguous.  out[k] = dx * (...) exactly as the reference.
 * fdmb_fft_batch_impl selects the kernel: 0 = the one the solvers use for this length (the persistent
 * bulk-copy-fed sweep for N >= 32, kinds 0..2), 1 = the plain tile kernel, 2 = the persistent sweep.        */
int fdmb_fft_batch(int kind, int N, long long batch, double dx, const double* in, double* out);
int fdmb_fft_batch_impl(int kind, int N, long long batch, double dx, const double* in, double* out, int impl);

/* ---- LaplCube -------------------------------------------------------------------
 * Replaces fdm::LaplCube<double,check,F> (src/lapl_cube.h:9-106, src/lapl_cube.cpp:9-172).
 * create   <-> constructor (dx,dy,dz,lx,ly,lz,nx,ny,nz)   src/lapl_cube.h:58-100
 * periodic <-> F = tensor_flags<periodic,periodic,periodic> (1) or tensor_flags<> (0),
 *              the two instantiations of src/lapl_cube.cpp:174-182
 * solve    <-> void solve(T* ans, T* rhs)                  src/lapl_cube.cpp:9-142
 * Dirichlet axes need n+1 = 2^k, periodic axes n = 2^k (src/fft.cpp:60-67).          */
typedef struct fdmb_lapl_cube fdmb_lapl_cube;
int fdmb_lapl_cube_create(fdmb_lapl_cube** h, double dx, double dy, double dz,
                          double lx, double ly, double lz, int nx, int ny, int nz, int periodic);
int fdmb_lapl_cube_solve(fdmb_lapl_cube* h, double* ans, const double* rhs);
int fdmb_lapl_cube_solve_device(fdmb_lapl_cube* h, double* d_ans, const double* d_rhs, void* stream);
/* B200 extension: `count` independent solves with HOST arrays (ans[i], rhs[i] as for solve), software-pipelined over
 * two device staging pairs: the upload of solve i+1 and the download of solve i-1 overlap solve i, so a long batch
 * costs max(upload, download, compute) per solve instead of their sum.  Page-locked host arrays are needed for the
 * copies to overlap; results are bit-identical to `count` calls of solve.  Single-GPU handles only.              */
int fdmb_lapl_cube_solve_batch(fdmb_lapl_cube* h, int count, double* const* ans, const double* const* rhs);
int fdmb_lapl_cube_destroy(fdmb_lapl_cube* h);

/* Single precision: fdm::LaplCube<float,check,F> (the instantiations of src/lapl_cube.cpp:176-177,181-182).  Same
 * arguments and meaning as above with float arrays; the device arrays, the tables and every butterfly are float (the
 * reference stores its eigenvalue tables as T too, src/lapl_cube.h:53-54).  Single GPU.                          */
typedef struct fdmb_lapl_cube_f32 fdmb_lapl_cube_f32;
int fdmb_lapl_cube_f32_create(fdmb_lapl_cube_f32** h, double dx, double dy, double dz,
                              double lx, double ly, double lz, int nx, int ny, int nz, int periodic);
int fdmb_lapl_cube_f32_solve(fdmb_lapl_cube_f32* h, float* ans, const float* rhs);
int fdmb_lapl_cube_f32_solve_device(fdmb_lapl_cube_f32* h, float* d_ans, const float* d_rhs, void* stream);
int fdmb_lapl_cube_f32_destroy(fdmb_lapl_cube_f32* h);

/* Multi-GPU LaplCube: the grid is cut into z-slabs, one rank (= one GPU, normally one process) per
 * slab.  The reference has no distributed solver; this is the slab decomposition of the same
 * solve() (src/lapl_cube.cpp:9-142): x and y transforms are local to a slab, the z transform runs
 * on y-pencils, and the two slab<->pencil transposes are fused into the stores of the y and z sweeps
 * (peer stores over NVLink into buffers the ranks map from each other).
 *   fdmb_slab_range        which 0-based interior entries of an axis of n points rank owns
 *                          (host-only arithmetic; Dirichlet axes: slot 0 is the boundary, so rank 0
 *                          owns one entry fewer)
 *   create_sharded         same arguments as create + (rank, nranks); nranks in {1,2,4,8}; call with
 *                          the rank's device current
 *   export_ipc/attach_ipc  one process per GPU: every rank exports FDMB_IPC_HANDLE_BYTES bytes, the
 *                          host exchanges them (torch.distributed / MPI all-gather), every rank
 *                          attaches the concatenation ordered by rank
 *   attach_local           all ranks in one process: pass the handles ordered by rank
 *   solve / solve_device   on a sharded handle take this rank's slab [nz_local][ny][nx]; every rank
 *                          must call them the same number of times (they meet in two device-side
 *                          barriers per solve)                                                     */
#define FDMB_IPC_HANDLE_BYTES 64
int fdmb_slab_range(int n, int periodic, int nranks, int rank, int* first, int* count);
int fdmb_lapl_cube_create_sharded(fdmb_lapl_cube** h, double dx, double dy, double dz,
                                  double lx, double ly, double lz, int nx, int ny, int nz, int periodic,
                                  int rank, int nranks);
int fdmb_lapl_cube_local_slab(fdmb_lapl_cube* h, int* z_first, int* nz_local);
int fdmb_lapl_cube_export_ipc(fdmb_lapl_cube* h, void* handle);
int fdmb_lapl_cube_attach_ipc(fdmb_lapl_cube* h, const void* handles);
int fdmb_lapl_cube_attach_local(fdmb_lapl_cube* h, fdmb_lapl_cube* const* all);

/* ---- LaplRect / LaplRectFFT2 -----------------------------------------------------
 * Replaces fdm::LaplRect<double,check,F> (src/lapl_rect.h:11-78, src/lapl_rect.cpp:44-110) and
 * fdm::LaplRectFFT2<double,check,F> (src/lapl_rect.h:80-116, src/lapl_rect.cpp:113-207).
 * create     <-> constructor (dx,dy,lx,ly,nx,ny); kind 0 = LaplRect (y transform + tridiagonal in x,
 *                x always Dirichlet), 1 = LaplRectFFT2 (transforms on both axes);
 *                yperiodic / xperiodic <-> F = tensor_flags<>, <periodic>, <periodic,periodic>
 *                (instantiations src/lapl_rect.cpp:209-238; kind 0 ignores xperiodic, :47)
 * set_scales <-> the public members lm_y_scale, L_scale, U_scale (src/lapl_rect.h:57-59; nx+1 host
 *                doubles each, entry j for column j; NULL keeps the current values) that the
 *                cylindrical slice plotter overwrites (src/velocity_plot.h:113-127)
 * solve      <-> void solve(T* ans, T* rhs); arrays are [ny rows][nx], x fastest.
 * Dirichlet axes need n+1 = 2^k, periodic axes n = 2^k on every transformed axis (2^k <= 2048); the
 * tridiagonal axis of kind 0 (x) takes any nx from 2 to about 2500 (shared-memory tile of the sweep).  */
typedef struct fdmb_lapl_rect fdmb_lapl_rect;
int fdmb_lapl_rect_create(fdmb_lapl_rect** h, int kind, int yperiodic, int xperiodic,
                          double dx, double dy, double lx, double ly, int nx, int ny);
int fdmb_lapl_rect_set_scales(fdmb_lapl_rect* h, const double* lm_y_scale, const double* L_scale,
                              const double* U_scale);
int fdmb_lapl_rect_solve(fdmb_lapl_rect* h, double* ans, const double* rhs);
int fdmb_lapl_rect_solve_device(fdmb_lapl_rect* h, double* d_ans, const double* d_rhs, void* stream);
int fdmb_lapl_rect_destroy(fdmb_lapl_rect* h);

/* ---- LaplCyl3FFT2 ---------------------------------------------------------------
 * Replaces fdm::LaplCyl3FFT2<double,check,zflag> (src/lapl_cyl.h:172-249, src/lapl_cyl.cpp:11-170):
 * Poisson equation in cylindrical coordinates, periodic in phi, Dirichlet (zperiodic=0) or
 * periodic (1) in z, Dirichlet in r.
 * create <-> constructor (dr,dz,r0,lr,lz,nr,nz,nphi)       src/lapl_cyl.h:212-245
 * solve  <-> void solve(T* ans, T* rhs)                    src/lapl_cyl.cpp:11-128
 * Arrays are [phi 0..nphi-1][z z1..zn][r 1..nr], r fastest (src/lapl_cyl.h:222).
 * nphi and the z transform length (nz+1 Dirichlet, nz periodic) must be powers of two (<= 2048);
 * nr is free, from 2 to about 2500 (shared-memory tile of the r sweep). */
typedef struct fdmb_lapl_cyl fdmb_lapl_cyl;
int fdmb_lapl_cyl_create(fdmb_lapl_cyl** h, double dr, double dz, double r0, double lr, double lz,
                         int nr, int nz, int nphi, int zperiodic);
int fdmb_lapl_cyl_solve(fdmb_lapl_cyl* h, double* ans, const double* rhs);
int fdmb_lapl_cyl_solve_device(fdmb_lapl_cyl* h, double* d_ans, const double* d_rhs, void* stream);
int fdmb_lapl_cyl_destroy(fdmb_lapl_cyl* h);
/* LaplCyl3FFT2 over 2, 4 or 8 GPUs of one node: phi-slabs (rank r owns phi in [r*nphi/P, (r+1)*nphi/P) of the
 * arrays); the commuting axis transforms are reordered to z, (transpose), phi, r, phi, (transpose), z so that only two
 * NVLink transposes are needed; both are peer stores fused into the sweeps.  Needs nphi and the z
 * transform length >= 32, an even nr and 16-byte aligned slabs.  create/attach/solve are collective.        */
int fdmb_lapl_cyl_create_sharded(fdmb_lapl_cyl** h, double dr, double dz, double r0, double lr, double lz,
                                 int nr, int nz, int nphi, int zperiodic, int rank, int nranks);
int fdmb_lapl_cyl_local_slab(fdmb_lapl_cyl* h, int* phi_first, int* nphi_local);
int fdmb_lapl_cyl_export_ipc(fdmb_lapl_cyl* h, void* handle);
int fdmb_lapl_cyl_attach_ipc(fdmb_lapl_cyl* h, const void* handles);
int fdmb_lapl_cyl_attach_local(fdmb_lapl_cyl* h, fdmb_lapl_cyl* const* all);

/* ---- NSCube ---------------------------------------------------------------------
 * Replaces fdm::NSCube<double,check> (src/ns_cube.h:13-92, src/ns_cube.cpp:27-277).
 * params   <-> the [ns] config keys read by the constructor (src/ns_cube.h:47-61);
 *              ny is taken from nx exactly like the reference (src/ns_cube.h:58).
 * step     <-> void step()  (src/ns_cube.cpp:27-62), nsteps times, state stays on the device
 * fields   <-> the public tensors u,v,w,p,x,F,G,H,RHS (src/ns_cube.h:31-33) with the
 *              reference's extents, ghosts included (src/ns_cube.h:66-75); get/set copy
 *              exactly what the reference exposes as ns.u.vec etc. (test/test_ns_cube.cpp:39). */
typedef struct fdmb_ns_cube fdmb_ns_cube;
typedef struct fdmb_ns_cube_params {
    double x1, y1, z1, x2, y2, z2;
    double u0, Re, dt;
    int nx, nz;
    int verbose;
} fdmb_ns_cube_params;
enum { FDMB_FIELD_U = 0, FDMB_FIELD_V, FDMB_FIELD_W, FDMB_FIELD_P, FDMB_FIELD_X,
       FDMB_FIELD_F, FDMB_FIELD_G, FDMB_FIELD_H, FDMB_FIELD_RHS };
int fdmb_ns_cube_default_params(fdmb_ns_cube_params* p);
int fdmb_ns_cube_create(fdmb_ns_cube** h, const fdmb_ns_cube_params* p);
int fdmb_ns_cube_step(fdmb_ns_cube* h, int nsteps);
int fdmb_ns_cube_step_async(fdmb_ns_cube* h, int nsteps, void* stream);
int fdmb_ns_cube_field_size(fdmb_ns_cube* h, int field, long long* count);
int fdmb_ns_cube_get_field(fdmb_ns_cube* h, int field, double* host);
int fdmb_ns_cube_set_field(fdmb_ns_cube* h, int field, const double* host);
int fdmb_ns_cube_field_device_ptr(fdmb_ns_cube* h, int field, void** dptr);
long long fdmb_ns_cube_time_index(fdmb_ns_cube* h);
int fdmb_ns_cube_destroy(fdmb_ns_cube* h);

/* NSCube over several GPUs of one node: the z-slabs of the sharded LaplCube above (fdmb_slab_range); stencil halo
 * planes are read from the neighbours' memory over NVLink after a device-side barrier, the pressure solve is the
 * sharded LaplCube.  The reference runs on one host only (src/ns_cube.cpp:27-62); results are identical to it.
 *   create_sharded   this rank's part; collective in the sense that every rank must create, attach and step
 *   owned_planes     which z planes (global index, e.g. -1 for w's bottom ghost) of a field a rank reports:
 *                    the ranks' ranges tile the reference array's z range (pure function, no device needed)
 *   local_planes     the same for a handle
 *   export/attach    like the sharded LaplCube; one record = 2 * FDMB_IPC_HANDLE_BYTES per rank
 *   field_size / get_field / set_field / field_device_ptr act on the OWNED planes of a sharded handle       */
int fdmb_ns_cube_create_sharded(fdmb_ns_cube** h, const fdmb_ns_cube_params* p, int rank, int nranks);
int fdmb_ns_cube_owned_planes(int nz, int field, int rank, int nranks, int* z_first, int* nplanes);
int fdmb_ns_cube_local_planes(fdmb_ns_cube* h, int field, int* z_first, int* nplanes);
int fdmb_ns_cube_export_ipc(fdmb_ns_cube* h, void* handles);
int fdmb_ns_cube_attach_ipc(fdmb_ns_cube* h, const void* handles);
int fdmb_ns_cube_attach_local(fdmb_ns_cube* h, fdmb_ns_cube* const* all);
int fdmb_ns_cube_synchronize(fdmb_ns_cube* h);

/* Single precision: fdm::NSCube<float,check> (src/ns_cube.cpp:281-282).  Float fields and a float pressure solve
 * (the member LaplCube<T,check>, src/ns_cube.h:34); the stencil arithmetic is double, like the reference's own mixed
 * float / double expressions.  Same params, field ids and extents as above; single GPU.                          */
typedef struct fdmb_ns_cube_f32 fdmb_ns_cube_f32;
int fdmb_ns_cube_f32_create(fdmb_ns_cube_f32** h, const fdmb_ns_cube_params* p);
int fdmb_ns_cube_f32_step(fdmb_ns_cube_f32* h, int nsteps);
int fdmb_ns_cube_f32_field_size(fdmb_ns_cube_f32* h, int field, long long* count);
int fdmb_ns_cube_f32_get_field(fdmb_ns_cube_f32* h, int field, float* host);
int fdmb_ns_cube_f32_set_field(fdmb_ns_cube_f32* h, int field, const float* host);
long long fdmb_ns_cube_f32_time_index(fdmb_ns_cube_f32* h);
int fdmb_ns_cube_f32_destroy(fdmb_ns_cube_f32* h);

/* ---- NSCyl ----------------------------------------------------------------------
 * Replaces fdm::NSCyl<double,check,zflag> (src/ns_cyl.h:17-132, src/ns_cyl.cpp:23-484): flow between
 * two coaxial cylinders, inner one rotating with speed u0, on a staggered grid in (phi, z, r).
 * params   <-> the [ns] config keys read by the constructor (src/ns_cyl.h:57-68); zperiodic selects
 *              zflag = tensor_flag::periodic (1) or none (0), the instantiations of
 *              src/ns_cyl.cpp:486-494; vrandom = 1 seeds v like src/ns_cyl.h:99-108
 * step     <-> void step()    (src/ns_cyl.cpp:23-63), nsteps times, state stays on the device
 * lstep    <-> void L_step()  (src/ns_cyl.cpp:66-78): the step linearised about u0, v0, w0
 * fields   <-> the public tensors u,v,w,p,x,F,G,H,RHS,u0,v0,w0 (src/ns_cyl.h:40-46) with the
 *              reference's extents [phi][z][r], ghosts included (src/ns_cyl.h:80-93).
 * The reference's verify() wall invariants inside init_bound (src/ns_cyl.cpp:136-163) are not
 * re-checked on the device.                                                                    */
typedef struct fdmb_ns_cyl fdmb_ns_cyl;
typedef struct fdmb_ns_cyl_params {
    double R, r, h1, h2;      /* outer / inner radius, z range */
    double u0, Re, dt;
    int nr, nz, nphi;
    int verbose, vrandom, zperiodic;
} fdmb_ns_cyl_params;
enum { FDMB_FIELD_U0 = 9, FDMB_FIELD_V0, FDMB_FIELD_W0 };
int fdmb_ns_cyl_default_params(fdmb_ns_cyl_params* p);
int fdmb_ns_cyl_create(fdmb_ns_cyl** h, const fdmb_ns_cyl_params* p);
int fdmb_ns_cyl_step(fdmb_ns_cyl* h, int nsteps);
int fdmb_ns_cyl_lstep(fdmb_ns_cyl* h, int nsteps);
int fdmb_ns_cyl_step_async(fdmb_ns_cyl* h, int nsteps, int linear, void* stream);
int fdmb_ns_cyl_field_size(fdmb_ns_cyl* h, int field, long long* count);
int fdmb_ns_cyl_get_field(fdmb_ns_cyl* h, int field, double* host);
int fdmb_ns_cyl_set_field(fdmb_ns_cyl* h, int field, const double* host);
int fdmb_ns_cyl_field_device_ptr(fdmb_ns_cyl* h, int field, void** dptr);
long long fdmb_ns_cyl_time_index(fdmb_ns_cyl* h);
/* the public, non-const member U0 (src/ns_cyl.h:23; test/test_ns_cyl_spectral.cpp sets it to 0): wall speed used by
 * init_bound from the next step on                                                                               */
int fdmb_ns_cyl_set_u0(fdmb_ns_cyl* h, double u0);
int fdmb_ns_cyl_destroy(fdmb_ns_cyl* h);
/* NSCyl over 2, 4 or 8 GPUs of one node: the phi-slabs of the sharded LaplCyl3FFT2; u, v, w (and u0, v0, w0 for the
 * linearised step) keep one wrap-around halo plane each side, H one below and x one above, pulled from the
 * neighbours' memory over NVLink after device-side barriers.  field_size / get_field / set_field / field_device_ptr
 * act on this rank's own phi planes [phi_first, phi_first + nphi_local).  export/attach as for NSCube.            */
int fdmb_ns_cyl_create_sharded(fdmb_ns_cyl** h, const fdmb_ns_cyl_params* p, int rank, int nranks);
int fdmb_ns_cyl_local_slab(fdmb_ns_cyl* h, int* phi_first, int* nphi_local);
int fdmb_ns_cyl_export_ipc(fdmb_ns_cyl* h, void* handles);
int fdmb_ns_cyl_attach_ipc(fdmb_ns_cyl* h, const void* handles);
int fdmb_ns_cyl_attach_local(fdmb_ns_cyl* h, fdmb_ns_cyl* const* all);
int fdmb_ns_cyl_synchronize(fdmb_ns_cyl* h);

/* ---- velocity_plotter -------------------------------------------------------------
 * Replaces fdm::velocity_plotter<double,check,F> (src/velocity_plot.h:11-143, src/velocity_plot.cpp:10-222),
 * the step AFTER the path in both reference drivers (test/test_ns_cube.cpp:24-50, test/test_ns_cyl.cpp:53-92):
 * mid-plane slices of the staggered velocity, their stream functions through LaplRectFFT2 / LaplRect
 * (cylindrical column scales, src/velocity_plot.h:113-127) and the ASCII VTK writer.  The slices, right-hand
 * sides, the three 2-D solves and the cell-centred velocities of the VTK file are computed on the device from
 * the NS state where it lives; only 2-D slices (and, for vtk_out, 3 doubles per cell) cross PCIe.
 * params     <-> the constructor (dx,dy,dz, nx,ny,nz, xx1,xx2, yy1,yy2, zz1,zz2, cyl); zperiodic / yperiodic
 *                <-> F = tensor_flags<>, <periodic>, <periodic,periodic> (src/velocity_plot.cpp:222-235).
 *                Axis names are the reference's: z slowest, x fastest (phi, z, r for cylinders).
 * use_*      <-> void use(T* u, T* v, T* w): arrays with the reference's extents
 *                u[z0..znn][y0..ynn][-1..nx+1], v[z0..znn][y_..ynn][0..nx+1], w[z_..znn][y0..ynn][0..nx+1]
 *                (src/velocity_plot.h:97-99).  use_host re-uploads the host arrays at every update(), like the
 *                reference re-reads them; use_device / use_ns_cube / use_ns_cyl read device memory in place
 *                (the NS handle's stream is synchronised first; single-GPU NS handles only).
 * update     <-> void update()  (src/velocity_plot.cpp:17-67)
 * get_slice  <-> the members vx,wx,uy,wy,uz,vz,RHS_x,RHS_y,RHS_z,psi_x,psi_y,psi_z (src/velocity_plot.h:32-44),
 *                row-major with the reference's extents; slice_dims gives rows x cols
 * cell_velocity  the "VECTORS u" block of vtk_out before formatting: for i=z1..zn, k=y1..yn, j=1..nx the three
 *                face averages (src/velocity_plot.cpp:180-182,209-213), 3 doubles per cell
 * vtk_out    <-> void vtk_out(const std::string& name, int step) (src/velocity_plot.cpp:117-220), same bytes */
typedef struct fdmb_vplot fdmb_vplot;
typedef struct fdmb_vplot_params {
    double dx, dy, dz;
    int nx, ny, nz;
    double xx1, xx2, yy1, yy2, zz1, zz2;
    int cyl;
    int zperiodic, yperiodic;
} fdmb_vplot_params;
enum { FDMB_SLICE_VX = 0, FDMB_SLICE_WX, FDMB_SLICE_UY, FDMB_SLICE_WY, FDMB_SLICE_UZ, FDMB_SLICE_VZ,
       FDMB_SLICE_RHS_X, FDMB_SLICE_RHS_Y, FDMB_SLICE_RHS_Z, FDMB_SLICE_PSI_X, FDMB_SLICE_PSI_Y, FDMB_SLICE_PSI_Z,
       FDMB_SLICE_COUNT };
int fdmb_vplot_create(fdmb_vplot** h, const fdmb_vplot_params* p);
int fdmb_vplot_field_size(fdmb_vplot* h, int field, long long* count);   /* field: FDMB_FIELD_U / _V / _W */
int fdmb_vplot_use_host(fdmb_vplot* h, const double* u, const double* v, const double* w);
int fdmb_vplot_use_device(fdmb_vplot* h, const double* d_u, const double* d_v, const double* d_w);
int fdmb_vplot_use_ns_cube(fdmb_vplot* h, fdmb_ns_cube* ns);
int fdmb_vplot_use_ns_cyl(fdmb_vplot* h, fdmb_ns_cyl* ns);
int fdmb_vplot_update(fdmb_vplot* h);
int fdmb_vplot_slice_dims(fdmb_vplot* h, int slice, int* rows, int* cols);
int fdmb_vplot_get_slice(fdmb_vplot* h, int slice, double* host);
int fdmb_vplot_cell_velocity(fdmb_vplot* h, double* host);
int fdmb_vplot_vtk_out(fdmb_vplot* h, const char* name, int time_index);
int fdmb_vplot_destroy(fdmb_vplot* h);

/* ---- particle-mesh N-body step -------------------------------------------------------
 * Replaces the step of NBody<double,check,CIC3<double>> in the reference's test/nbody.cpp (:24-596), the second
 * in-tree consumer of the periodic LaplCube, with the program's default local = 0 (mesh forces
 * only; the short-range pair correction of :344-388 is "need to check" in the reference and not built here).
 * params      <-> the constructor (x0,y0,z0,l,n,...,dt,G) :104-131; cell size h = l / n, n = 2^k
 *                 deposit_all = 0 reproduces distribute_masses (:257-272), which walks the cell lists with one
 *                 offset for all three axes and so deposits only the bodies whose cell indices are all even or all
 *                 odd; 1 deposits every body (the evident intent).  The mean density uses all bodies either way.
 * set_bodies  <-> bodies[b].x, .v, .mass (:46-58); x, v are [N][3] like Body::x[3]; a and aprev start at 0.  The
 *                 reference seeds them from std::default_random_engine (:541-587); a caller hands them over instead.
 * calc_accel  <-> calc_a_pm() (:292-421): f = -mass/l^3 + cloud-in-cell deposit, rhs = 4 pi G f / h^3,
 *                 psi = periodic LaplCube solve, E = 4-point difference of psi, a = cloud-in-cell gather of E
 * step        <-> step() (:133-142) = calc_a_pm() + move() (:469-487, velocity Verlet, periodic wrap), nsteps times;
 *                 everything stays on the device
 * get_bodies  field FDMB_PM_X / _V / _A / _APREV ([N][3]) or FDMB_PM_MASS ([N])
 * get_grid    FDMB_PM_F / _RHS / _PSI (n^3, [z][y][x]) or FDMB_PM_E (n^3 x 3) as after the last calc_accel   */
typedef struct fdmb_pm fdmb_pm;
typedef struct fdmb_pm_params {
    double x0, y0, z0, l;
    double dt, G;
    int n;
    int deposit_all;
} fdmb_pm_params;
enum { FDMB_PM_X = 0, FDMB_PM_V, FDMB_PM_A, FDMB_PM_APREV, FDMB_PM_MASS };
enum { FDMB_PM_F = 0, FDMB_PM_RHS, FDMB_PM_PSI, FDMB_PM_E };
int fdmb_pm_create(fdmb_pm** h, const fdmb_pm_params* p);
int fdmb_pm_set_bodies(fdmb_pm* h, long long N, const double* x, const double* v, const double* mass);
long long fdmb_pm_count(fdmb_pm* h);
int fdmb_pm_calc_accel(fdmb_pm* h);
int fdmb_pm_step(fdmb_pm* h, int nsteps);
int fdmb_pm_get_bodies(fdmb_pm* h, int field, double* host);
int fdmb_pm_get_grid(fdmb_pm* h, int grid, double* host);
int fdmb_pm_destroy(fdmb_pm* h);

#ifdef __cplusplus
}
#endif
#endif /* FDM_B200_H */
