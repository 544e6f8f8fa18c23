"""GPU parity tests of the single-precision instantiations, through the C ABI: fdm::LaplCube<float,check,F>
(src/lapl_cube.cpp:176-177,181-182) and fdm::NSCube<float,check> (src/ns_cube.cpp:281-282).

The bar is the reference's OWN float instantiation, compiled unmodified into oracle/_ref: (a) our float answer agrees
with the reference's float answer to a few float ulps of relative L2 (both carry their own single-precision round-off, so
they cannot agree better than either agrees with the double answer), and (b) our error against the DOUBLE reference is
not worse than a small multiple of the reference float path's own error against it."""
import math

import numpy as np
import pytest

from oracle import fdm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb():
    import fdm_b200
    assert fdm_b200.lib().fdmb_device_count() > 0, "GPU tests need a CUDA device"
    return fdm_b200


def check_cube(fb, ref, args, shape, periodic, seed):
    rhs = O.synthetic_rhs(shape, seed=seed).astype(np.float32)
    if periodic:
        rhs -= rhs.mean(dtype=np.float64).astype(np.float32)
    got = fb.LaplCubeF32(*args, periodic=periodic).solve(rhs)
    assert got.dtype == np.float32
    want32 = ref.LaplCubeF32(*args, periodic=periodic).solve(rhs)
    want64 = ref.LaplCube(*args, periodic=periodic).solve(rhs.astype(np.float64))
    e_ref = O.rel_l2(want32, want64)              # the reference float path's own round-off
    e_got = O.rel_l2(got, want64)
    assert O.rel_l2(got, want32) < 2e-5, (O.rel_l2(got, want32), e_ref, e_got)
    assert e_got < 4 * e_ref + 2e-6, (e_got, e_ref)
    return e_got, e_ref


@pytest.mark.parametrize("n", [7, 15, 31, 63, 127, 255])
def test_cube_f32_dirichlet(fb, ref, n):
    dx = 1.0 / n; l = 1 + dx
    check_cube(fb, ref, (dx, dx, dx, l, l, l, n, n, n), (n, n, n), False, n)


@pytest.mark.parametrize("shape", [(7, 15, 31), (31, 127, 63), (3, 255, 15), (15, 31, 1023), (1023, 15, 31)],
                         ids=lambda s: "x".join(map(str, s)))
def test_cube_f32_ragged(fb, ref, shape):
    nz, ny, nx = shape
    args = (0.1, 0.2, 0.3, 0.1 * (nx + 1), 0.2 * (ny + 1), 0.3 * (nz + 1), nx, ny, nz)
    check_cube(fb, ref, args, shape, False, sum(shape))


@pytest.mark.parametrize("n", [16, 64, 128])
def test_cube_f32_periodic(fb, ref, n):
    dx = 2 * math.pi / n; l = 2 * math.pi
    check_cube(fb, ref, (dx, dx, dx, l, l, l, n, n, n), (n, n, n), True, n + 1)


def test_cube_f32_device_resident_and_errors(fb):
    import torch
    n = 63; dx = 1.0 / n; l = 1 + dx
    S = fb.LaplCubeF32(dx, dx, dx, l, l, l, n, n, n)
    rhs = (torch.rand(n ** 3, dtype=torch.float32, device="cuda") - 0.5)
    ans = torch.empty_like(rhs)
    torch.cuda.synchronize()
    S.solve_device(ans.data_ptr(), rhs.data_ptr())
    fb.capi.check(fb.lib().fdmb_device_synchronize(), "sync")
    host = S.solve(rhs.cpu().numpy().reshape(n, n, n))
    assert np.array_equal(ans.cpu().numpy().reshape(n, n, n), host)
    with pytest.raises(fb.FdmB200Error):
        fb.LaplCubeF32(0.1, 0.1, 0.1, 3.3, 3.3, 3.3, 32, 32, 32)            # Dirichlet needs 2^k - 1
    with pytest.raises(fb.FdmB200Error):
        fb.LaplCubeF32(0.1, 0.1, 0.1, 204.8, 3.2, 3.2, 2047, 31, 31)       # float transforms are instantiated up to 1024
    with pytest.raises(ValueError):
        S.solve(np.zeros((n, n, n), dtype=np.float64)[:1])


@pytest.mark.parametrize("n,steps", [(15, 10), (31, 20), (63, 10), (127, 3)])
def test_ns_cube_f32(fb, ref, n, steps):
    """All nine public fields after several steps against NSCube<float> of the reference and against NSCube<double>."""
    kw = dict(nx=n, nz=n, Re=400.0, dt=0.005)
    ns = fb.NSCubeF32(**kw); r32 = ref.NSCubeF32(**kw); r64 = ref.NSCube(**kw)
    ns.step(steps); r32.step(steps); r64.step(steps)
    assert ns.time_index == steps
    scale = max(np.linalg.norm(r64.field(f)) for f in "uvwp")
    for f in ("u", "v", "w", "p", "x", "F", "G", "H", "RHS"):
        a, b32, b64 = ns.field(f), r32.field(f), r64.field(f)
        assert a.dtype == np.float32 and a.size == b32.size
        if np.linalg.norm(b64) < 1e-3 * scale and f in "vwp":
            continue                                   # negligible cross-flow fields: relative error is meaningless
        e_ref, e_got = O.rel_l2(b32, b64), O.rel_l2(a, b64)
        assert O.rel_l2(a, b32) < 1e-4, (f, O.rel_l2(a, b32), e_ref, e_got)
        assert e_got < 4 * e_ref + 2e-5, (f, e_got, e_ref)
    cat = lambda g: np.concatenate([np.asarray(g(f), dtype=np.float64).ravel() for f in "uvwp"])
    assert O.rel_l2(cat(ns.field), cat(r32.field)) < 2e-5


def test_ns_cube_f32_set_field_roundtrip(fb):
    ns = fb.NSCubeF32(nx=15, nz=15, Re=100.0, dt=0.01)
    u = np.arange(ns.field_size("u"), dtype=np.float32)
    ns.set_field("u", u)
    assert np.array_equal(ns.field("u"), u)
    with pytest.raises(ValueError):
        ns.set_field("u", u[:-1])
