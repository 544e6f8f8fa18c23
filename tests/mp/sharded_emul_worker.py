"""CPU (gloo) rehearsal of the multi-GPU LaplCube data path: same slab plan (fdmb_slab_range from
the C ABI), same phase order and the same two all-to-all transposes as fdm_b200/csrc/lapl_cube.cu
(solve_device_sharded), with the oracle's 1-D transforms standing in for the kernels.  Checks the
host-side partition logic against the oracle's full solve on every rank count the box can spawn."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def sharded_solve(dist, torch, slab_range, O, full, rhs, periodic):
    rank, world = dist.get_rank(), dist.get_world_size()
    nz, ny, nx = full.nz, full.ny, full.nx
    zr = [slab_range(nz, periodic, world, q) for q in range(world)]
    yr = [slab_range(ny, periodic, world, q) for q in range(world)]
    z0, nzl = zr[rank]
    y0, nyl = yr[rank]
    fwd = O.pFFT_1 if periodic else O.sFFT
    inv = O.pFFT if periodic else O.sFFT
    a = np.array(rhs[z0:z0 + nzl], dtype=np.float64)
    a = fwd(a, full.dx * full.slx, axis=2)
    a = fwd(a, full.dy * full.sly, axis=1)

    def a2a(send_blocks, recv_shapes):
        send = [torch.from_numpy(np.ascontiguousarray(b)) for b in send_blocks]
        recv = [torch.empty(s, dtype=torch.float64) for s in recv_shapes]
        dist.all_to_all(recv, send) if dist.get_backend() != "gloo" else _a2a_gloo(dist, torch, recv, send)
        return [r.numpy() for r in recv]

    # slab -> pencil: rank q receives [z of every rank][its y rows][x]
    got = a2a([a[:, yr[q][0]:yr[q][0] + yr[q][1], :] for q in range(world)],
              [(zr[q][1], nyl, nx) for q in range(world)])
    t = np.concatenate(got, axis=0)
    t = fwd(t, full.dz * full.slz, axis=0)
    k2 = full.lm_z[:, None, None] + full.lm_y[None, y0:y0 + nyl, None] + full.lm_x[None, None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = t / (-k2)
    if periodic and y0 == 0:
        t[0, 0, 0] = 0.0
    t = inv(t, full.slz, axis=0)
    # pencil -> slab
    got = a2a([t[zr[q][0]:zr[q][0] + zr[q][1]] for q in range(world)],
              [(nzl, yr[q][1], nx) for q in range(world)])
    a = np.concatenate(got, axis=1)
    a = inv(a, full.sly, axis=1)
    a = inv(a, full.slx, axis=2)
    return z0, a


def _a2a_gloo(dist, torch, recv, send):
    # gloo has no all_to_all: pairwise isend/irecv
    rank, world = dist.get_rank(), dist.get_world_size()
    reqs = []
    for q in range(world):
        if q == rank:
            recv[q].copy_(send[q])
        else:
            reqs.append(dist.isend(send[q], q))
            reqs.append(dist.irecv(recv[q], q))
    for r in reqs:
        r.wait()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import fdm_b200
    from oracle import fdm_oracle as O

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    worst = 0.0
    cases = [((31, 31, 31), False), ((63, 31, 15), False), ((7, 63, 31), False), ((32, 32, 32), True), ((16, 64, 8), True)]
    for (nz, ny, nx), periodic in cases:
        if (nz + (0 if periodic else 1)) // world < 2 or (ny + (0 if periodic else 1)) // world < 2:
            continue
        args = (0.1, 0.2, 0.3, 0.1 * (nx + 1), 0.2 * (ny + 1), 0.3 * (nz + 1), nx, ny, nz)
        full = O.LaplCube(*args, periodic=periodic)
        rhs = O.synthetic_rhs((nz, ny, nx), seed=nz + ny + nx)
        z0, part = sharded_solve(dist, torch, fdm_b200.slab_range, O, full, rhs, periodic)
        parts = [None] * world
        dist.all_gather_object(parts, (z0, part))
        if rank == 0:
            parts.sort(key=lambda p: p[0])
            got = np.concatenate([p[1] for p in parts], axis=0)
            assert got.shape == (nz, ny, nx), got.shape
            worst = max(worst, O.rel_l2(got, full.solve(rhs)))
    if rank == 0:
        with open(a.out, "w") as f:
            f.write(repr(float(worst)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
