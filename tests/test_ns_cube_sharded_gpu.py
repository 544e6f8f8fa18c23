"""GPU parity tests for the z-slab decomposed NSCube step: stencil halo planes pulled from the
neighbours over NVLink, pressure solve by the sharded LaplCube.  Needs >= 2 visible B200s; skipped otherwise.
All ranks live in this process (attach_local); the torchrun / IPC wiring is covered by tests/mp/ns_sharded_worker.py.
Bar: relative L2 <= 1e-12 (fp64) of the gathered fields against the compiled reference / the oracle."""
import numpy as np
import pytest

from oracle import fdm_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def fb():
    import fdm_b200
    if fdm_b200.lib().fdmb_device_count() < 2:
        pytest.skip("the sharded step needs at least 2 GPUs")
    return fdm_b200


def run_sharded(fb, P, steps, init=None, **kw):
    L = fb.lib()
    parts = []
    for r in range(P):
        fb.capi.check(L.fdmb_set_device(r), "set_device")
        parts.append(fb.NSCube(rank=r, nranks=P, **kw))
    fb.NSCube.connect_local(parts)
    if init is not None:
        for f, full in init.items():
            for s in parts:
                z0, npl = s.local_planes(f)
                lo = {"w": -1}.get(f, 0)
                s.set_field(f, full[z0 - lo:z0 - lo + npl])
    for _ in range(steps):
        for s in parts:
            s.step_device(1)                       # asynchronous on every rank's own stream
    for s in parts:
        s.synchronize()
    out = {}
    for f in ("u", "v", "w", "p", "x"):
        out[f] = np.concatenate([s.field(f) for s in parts])
    for s in parts:
        s.close()
    fb.capi.check(L.fdmb_set_device(0), "set_device")
    return out


def ranks_available(fb):
    n = fb.lib().fdmb_device_count()
    return [p for p in (2, 4, 8) if p <= n]


def check(got, want):
    cat_g = np.concatenate([got[f].ravel() for f in "uvwp"])
    cat_w = np.concatenate([np.asarray(want[f]).ravel() for f in "uvwp"])
    assert cat_g.size == cat_w.size
    assert O.rel_l2(cat_g, cat_w) < TOL
    for f in "uvwp":
        w = np.asarray(want[f]).ravel()
        if np.linalg.norm(w) > 1e-3 * np.linalg.norm(cat_w):
            assert O.rel_l2(got[f].ravel(), w) < TOL, f


@pytest.mark.parametrize("n,steps", [(31, 20), (63, 6)])
def test_sharded_cavity_vs_compiled_reference(fb, ref, n, steps):
    kw = dict(nx=n, nz=n, Re=250.0, dt=0.01)
    r = ref.NSCube(**kw)
    r.step(steps)
    want = {f: r.field(f) for f in "uvwp"}
    for P in ranks_available(fb):
        if (n + 1) // P < 4:
            continue
        check(run_sharded(fb, P, steps, **kw), want)


def test_sharded_perturbed_state_vs_oracle(fb):
    # a state with all three velocity components non-trivial, so that every halo plane matters
    n = 31
    kw = dict(nx=n, nz=n, Re=100.0, dt=0.005)
    po = O.NSCube(**kw)
    rng = np.random.default_rng(3)
    init = {}
    for f in "uvw":
        a = po.fields()[f]
        a[...] = 0.0
        sl = (slice(2, -2),) * 3
        a[sl] = 1e-2 * rng.uniform(-1, 1, a[sl].shape)
        init[f] = a.copy()
    for _ in range(5):
        po.step()
    want = po.fields()
    for P in ranks_available(fb):
        if (n + 1) // P < 4:
            continue
        check(run_sharded(fb, P, 5, init=init, **kw), want)


def test_sharded_rejects_thin_slabs(fb):
    with pytest.raises(fb.FdmB200Error):
        fb.NSCube(nx=7, nz=7, rank=0, nranks=4)


def test_sharded_one_process_per_gpu(fb, tmp_path):
    """torch.distributed.run, one rank per GPU, IPC handles exchanged through the process group."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "res.txt"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29741",
           os.path.join(root, "tests", "mp", "ns_sharded_worker.py"), "--size", "31", "--steps", "10", "--out", str(out)]
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert float(out.read_text()) < TOL
