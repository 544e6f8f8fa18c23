"""One full-size run of the UNMODIFIED reference: LaplCube<double> Dirichlet 1023^3 (BASELINE configs[4],
SURVEY 8d C5).  Run in the dev container (needs /root/reference, ~27 GB of host memory, about a minute):

    python tests/golden/make_golden_cube1023.py

The right-hand side is generated plane by plane from a counter-based generator (plane z uses the Philox key
(SEED, z)), so a sharded run reproduces exactly its own slab; tests/golden/cube1023.py holds the generator that
both this script and the GPU tests import.  Stored in golden_cube1023_v1.npz: the norm and the sum of the
reference's answer, a strided sample ans[::STRIDE, ::STRIDE, ::STRIDE], one full row, and the wall time of
the reference solve on this container's cores (reported by bench.py beside its bounded cpu_baseline sample).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from tests.golden.cube1023 import N, SEED, STRIDE, geometry, rhs_planes  # noqa: E402


def main():
    assert ref.build(quiet=False), "reference build failed"
    n = int(os.environ.get("CUBE_N", N))            # CUBE_N=255 for a quick rehearsal of this script
    d, l = geometry(n)
    rhs = rhs_planes(n, 0, n)
    S = ref.LaplCube(d, d, d, l, l, l, n, n, n)
    t0 = time.perf_counter()
    ans = S.solve(rhs)
    dt = time.perf_counter() - t0
    print(f"reference LaplCube {n}^3: {dt:.2f} s on {ref.num_threads()} threads = {n ** 3 / 1e9 / dt:.4f} Gpts/s")
    out = dict(n=n, seed=SEED, stride=STRIDE,
               ans_norm=float(np.linalg.norm(ans.ravel())), ans_sum=float(ans.sum(dtype=np.longdouble)),
               rhs_norm=float(np.linalg.norm(rhs.ravel())),
               sample=ans[::STRIDE, ::STRIDE, ::STRIDE].copy(), row=ans[n // 2, n // 3, :].copy(),
               ref_seconds=dt, ref_threads=ref.num_threads())
    name = "golden_cube1023_v1.npz" if n == N else f"golden_cube{n}_rehearsal.npz"
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name), **out)
    print("wrote", name)


if __name__ == "__main__":
    main()
