"""gloo rehearsal of the sharded NSCube halo plan (no GPU): every rank asks the C ABI which planes it owns,
the ranks exchange their answers and check that (1) the owned ranges tile every field's global z range and
(2) every halo plane a rank pulls before FGH / update is owned by the neighbour it pulls from."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    import torch.distributed as dist
    import fdm_b200
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for nz in (31, 63, 255):
        if (nz + 1) // world < 4:
            continue
        mine = {f: fdm_b200.owned_planes(nz, f, rank, world) for f in ("u", "v", "w", "p", "x", "F", "G", "H", "RHS")}
        allp = [None] * world
        dist.all_gather_object(allp, mine)
        glob = {"u": (0, nz + 1), "v": (0, nz + 1), "w": (-1, nz + 1), "p": (0, nz + 1), "x": (1, nz),
                "F": (1, nz), "G": (1, nz), "H": (0, nz), "RHS": (1, nz)}
        for f, (lo, hi) in glob.items():
            pos = lo
            for r in range(world):
                z0, n = allp[r][f]
                ok &= (z0 == pos and n > 0)
                pos += n
            ok &= (pos == hi + 1)
        # this rank's interior slab = its owned x planes; halos: u,v one plane each side, w two below, x one above
        ilo, n = mine["x"]; ihi = ilo + n - 1
        owns = lambda r, f, z: allp[r][f][0] <= z < allp[r][f][0] + allp[r][f][1]
        if rank > 0:
            ok &= owns(rank - 1, "u", ilo - 1) and owns(rank - 1, "v", ilo - 1)
            ok &= owns(rank - 1, "w", ilo - 1) and owns(rank - 1, "w", ilo - 2)
        if rank < world - 1:
            ok &= all(owns(rank + 1, f, ihi + 1) for f in ("u", "v", "w", "x"))
    res = [None] * world
    dist.all_gather_object(res, bool(ok))
    if rank == 0:
        with open(a.out, "w") as f:
            f.write("ok" if all(res) else "bad")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
