"""GPU parity tests for the phi-slab decomposed LaplCyl3FFT2 solve through the C ABI.  Needs >= 2
visible B200s; skipped otherwise.  All ranks live in this process (attach_local).
Bar: relative L2 <= 1e-12 (fp64) of the gathered slabs against the compiled reference."""
import math

import numpy as np
import pytest

from oracle import fdm_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def fb():
    import fdm_b200
    if fdm_b200.lib().fdmb_device_count() < 2:
        pytest.skip("the sharded solve needs at least 2 GPUs")
    return fdm_b200


def geom(nr, nz, nphi, zperiodic):
    R, r0, h = math.pi, math.pi / 2, 10.0
    dr = (R - r0) / nr; dz = h / nz
    return (dr, dz, r0 - dr / 2, R - r0 + dr, h if zperiodic else h + dz, nr, nz, nphi)    # src/ns_cyl.h:94-97


def solve_in_process(fb, g, rhs, P, zperiodic=False, repeats=2):
    import torch
    L = fb.lib()
    parts, bufs = [], []
    for r in range(P):
        fb.capi.check(L.fdmb_set_device(r), "set_device")
        s = fb.LaplCyl3FFT2(*g, zperiodic=zperiodic, rank=r, nranks=P)
        parts.append(s)
        with torch.cuda.device(r):
            d_rhs = torch.from_numpy(np.ascontiguousarray(rhs[s.phi_first:s.phi_first + s.nphi_local])).cuda(r)
            d_ans = torch.full_like(d_rhs, float("nan"))
        bufs.append((d_rhs, d_ans))
    fb.LaplCyl3FFT2.connect_local(parts)
    for _ in range(repeats):
        for s, (d_rhs, d_ans) in zip(parts, bufs):
            s.solve_device(d_ans.data_ptr(), d_rhs.data_ptr())
    for r in range(P):
        fb.capi.check(L.fdmb_set_device(r), "set_device")
        fb.capi.check(L.fdmb_device_synchronize(), "sync")
    out = np.concatenate([b[1].cpu().numpy() for b in bufs], axis=0)
    for s in parts:
        s.close()
    fb.capi.check(L.fdmb_set_device(0), "set_device")
    return out


def ranks_available(fb):
    n = fb.lib().fdmb_device_count()
    return [p for p in (2, 4, 8) if p <= n]


@pytest.mark.parametrize("nr,nz,nphi", [(32, 31, 32), (64, 63, 32), (128, 127, 128)])
def test_sharded_cyl_dirichlet_vs_compiled_reference(fb, ref, nr, nz, nphi):
    g = geom(nr, nz, nphi, False)
    rhs = O.synthetic_rhs((nphi, nz, nr), seed=nr + nz)
    want = ref.LaplCyl3FFT2(*g).solve(rhs)
    for P in ranks_available(fb):
        if nphi // P < 2 or (nz + 1) // P < 2:
            continue
        assert O.rel_l2(solve_in_process(fb, g, rhs, P), want) < TOL, f"P={P}"


def test_sharded_cyl_zperiodic_vs_compiled_reference(fb, ref):
    nr, nz, nphi = 32, 32, 64
    g = geom(nr, nz, nphi, True)
    rhs = O.synthetic_rhs((nphi, nz, nr), seed=12)
    want = ref.LaplCyl3FFT2(*g, zperiodic=True).solve(rhs)
    for P in ranks_available(fb):
        assert O.rel_l2(solve_in_process(fb, g, rhs, P, zperiodic=True), want) < TOL, f"P={P}"


def test_sharded_cyl_host_slab_entry_point(fb, ref):
    # the host-pointer entry point takes and returns this rank's phi-slab
    nr, nz, nphi = 32, 31, 32
    g = geom(nr, nz, nphi, False)
    rhs = O.synthetic_rhs((nphi, nz, nr), seed=4)
    want = ref.LaplCyl3FFT2(*g).solve(rhs)
    import threading
    L = fb.lib()
    P = 2
    parts = []
    for r in range(P):
        fb.capi.check(L.fdmb_set_device(r), "set_device")
        parts.append(fb.LaplCyl3FFT2(*g, rank=r, nranks=P))
    fb.LaplCyl3FFT2.connect_local(parts)
    res = [None] * P

    def run(r):       # the host entry point synchronises, so every rank needs its own host thread
        res[r] = parts[r].solve(rhs[parts[r].phi_first:parts[r].phi_first + parts[r].nphi_local])
    th = [threading.Thread(target=run, args=(r,)) for r in range(P)]
    [t.start() for t in th]; [t.join() for t in th]
    assert O.rel_l2(np.concatenate(res, axis=0), want) < TOL
    for s in parts:
        s.close()
    fb.capi.check(L.fdmb_set_device(0), "set_device")


def test_sharded_cyl_rejects_bad_split(fb):
    with pytest.raises(fb.FdmB200Error):
        fb.LaplCyl3FFT2(*geom(33, 31, 32, False), rank=0, nranks=2)          # odd nr
    with pytest.raises(fb.FdmB200Error):
        fb.LaplCyl3FFT2(*geom(32, 15, 32, False), rank=0, nranks=2)          # z transform length 16 < 32
