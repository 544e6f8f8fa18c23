"""Worker for the one-process-per-GPU test of the sharded NSCube step (launched by torch.distributed.run):
IPC handles travel through the process group, rank 0 gathers the owned planes and compares with the oracle."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=31)
    ap.add_argument("--steps", type=int, default=10)
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
    kw = dict(nx=a.size, nz=a.size, Re=250.0, dt=0.01)
    ns = fdm_b200.NSCube(rank=rank, nranks=world, **kw)
    ns.connect()
    ns.step(a.steps)
    mine = {f: ns.field(f) for f in "uvwp"}
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    if rank == 0:
        po = O.NSCube(**kw)
        for _ in range(a.steps):
            po.step()
        got = np.concatenate([np.concatenate([p[f] for p in parts]) for f in "uvwp"])
        want = np.concatenate([po.fields()[f].ravel() for f in "uvwp"])
        with open(a.out, "w") as f:
            f.write(repr(float(O.rel_l2(got, want))))
    dist.barrier()
    ns.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
