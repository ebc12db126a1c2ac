"""CPU model of the multi-GPU row-slab path, used by tests/test_slab_gloo.py.

One process per slab over torch.distributed (gloo).  It follows the product's host logic:
the slab split of lbm_b200_decompose, ring neighbours rank-1 / rank+1, and the halo protocol
of csrc/lbm_kernels.cuh -- after every step each slab sends ONLY populations 2,5,6 of its last
row northwards and 4,7,8 of its first row southwards (12*nx bytes each) into the neighbours'
halo rows.  The per-slab stepper is the oracle's timestep on a halo'd slab.
"""
import numpy as np
import torch
import torch.distributed as dist

NORTH_PLANES = [2, 5, 6]     # leave through the top edge of a slab
SOUTH_PLANES = [4, 7, 8]     # leave through the bottom edge


def exchange_halos(cells, rank, size):
    """cells: [rows+2, nx, 9] float32 with halo rows 0 and rows+1 (numpy, modified in place)."""
    rows = cells.shape[0] - 2
    south, north = (rank - 1) % size, (rank + 1) % size
    up = torch.from_numpy(np.ascontiguousarray(cells[rows][:, NORTH_PLANES]))
    down = torch.from_numpy(np.ascontiguousarray(cells[1][:, SOUTH_PLANES]))
    from_south = torch.empty_like(up)
    from_north = torch.empty_like(down)
    ops = [dist.P2POp(dist.isend, up, north, tag=1), dist.P2POp(dist.isend, down, south, tag=2),
           dist.P2POp(dist.irecv, from_south, south, tag=1), dist.P2POp(dist.irecv, from_north, north, tag=2)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    cells[0][:, NORTH_PLANES] = from_south.numpy()
    cells[rows + 1][:, SOUTH_PLANES] = from_north.numpy()


def run_slab(pkg, oracle, rank, size, nx, ny, iters, density, accel, omega, obstacles, cells0):
    """Returns (this slab's final rows [rows, nx, 9], av_vels partial float32[iters])."""
    rows_all, first_all = pkg.decompose(ny, size)
    rows, first = int(rows_all[rank]), int(first_all[rank])
    inv = pkg.free_cells_inv(obstacles)
    # halo rows of unused populations are poisoned: the step must never read them
    a = np.full((rows + 2, nx, 9), np.nan, np.float32)
    a[1:rows + 1] = cells0[first:first + rows]
    b = np.full_like(a, np.nan)
    ob = np.zeros((rows + 2, nx), np.int32)
    ob[1:rows + 1] = obstacles[first:first + rows]
    accel_row = ny - 2 - first + 1 if first <= ny - 2 < first + rows else -1
    av = np.zeros(iters, np.float32)
    exchange_halos(a, rank, size)
    for t in range(iters):
        if accel_row > 0:
            oracle.accelerate_row(a[accel_row], ob[accel_row], density, accel)
        # NaN halos of the populations that are never pulled must not leak: only rows 1..rows are compared
        av[t] = oracle.slab_timestep(np.nan_to_num(a, nan=0.0), b, ob, 1, rows + 1, omega) * inv
        a, b = b, a
        a[0] = np.nan
        a[rows + 1] = np.nan
        exchange_halos(a, rank, size)
    return a[1:rows + 1].copy(), av


def worker(rank, size, port, args, out_queue):
    import os
    import sys
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import __graft_entry__ as entry
    import oracle_lib
    pkg = entry.load_package()
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        cells, av = run_slab(pkg, oracle_lib, rank, size, *args)
        # the final reduction of the per-rank av_vels arrays (reference d2q9-bgk.c:396)
        total = torch.from_numpy(av.copy())
        dist.reduce(total, dst=0, op=dist.ReduceOp.SUM)
        out_queue.put((rank, cells, total.numpy() if rank == 0 else None))
    finally:
        dist.destroy_process_group()
