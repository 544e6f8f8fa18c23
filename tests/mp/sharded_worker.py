"""Worker for the one-process-per-GPU test of the sharded LaplCube solve (launched by
torch.distributed.run).  Rank 0 gathers the slabs, compares with the oracle and writes the error."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=127)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import fdm_b200
    from oracle import fdm_oracle as O

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    fdm_b200.capi.check(fdm_b200.lib().fdmb_set_device(local), "set_device")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.size
    dx = 1.0 / n; l = 1 + dx
    args = (dx, dx, dx, l, l, l, n, n, n)
    rhs = O.synthetic_rhs((n, n, n), seed=77)          # every rank generates the same field, keeps its slab
    S = fdm_b200.LaplCubeSharded(*args, rank=rank, nranks=world)
    S.connect()
    slab = rhs[S.z_first:S.z_first + S.nz_local]
    ans = None
    for _ in range(3):
        ans = S.solve(slab)                             # host-pointer entry point, slab in / slab out
    parts = [None] * world
    dist.all_gather_object(parts, ans)
    if rank == 0:
        got = np.concatenate(parts, axis=0)
        err = O.rel_l2(got, O.LaplCube(*args).solve(rhs))
        with open(a.out, "w") as f:
            f.write(repr(float(err)))
    dist.barrier()
    S.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
