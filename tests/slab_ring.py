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


# ---- the in-place (AA access pattern) ring: one buffer per slab, csrc/lbm_kernels.cuh kernel 4 ---------------------
CX = [0, 1, 0, -1, 0, 1, -1, -1, 1]       # lattice directions of the reference's speed order (d2q9-bgk.c:7-13)
CY = [0, 0, 1, 0, -1, 1, 1, -1, -1]
OPP = [0, 3, 4, 1, 2, 7, 8, 5, 6]


def swap_rows(up, down, rank, size):
    """Sends `up` to the northern neighbour and `down` to the southern one; returns (from_south, from_north)."""
    south, north = (rank - 1) % size, (rank + 1) % size
    up, down = torch.from_numpy(np.ascontiguousarray(up)), torch.from_numpy(np.ascontiguousarray(down))
    from_south, from_north = torch.empty_like(up), torch.empty_like(down)
    ops = [dist.P2POp(dist.isend, up, north, tag=3), dist.P2POp(dist.isend, down, south, tag=4),
           dist.P2POp(dist.irecv, from_south, south, tag=3), dist.P2POp(dist.irecv, from_north, north, tag=4)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return from_south.numpy(), from_north.numpy()


def run_slab_inplace(pkg, oracle, rank, size, nx, ny, iters, density, accel, omega, obstacles, cells0):
    """The in-place protocol of the product, step for step, with ONE array per slab:
      NEIGHBOUR step (layout L0 -> L1): pull through the halo rows; population opp(i) goes back to (slot i, cell - c_i);
          what would land in a halo row is sent into the neighbour's OWNED edge row instead;
      LOCAL step (L1 -> L0): population j of a cell is in its own slot opp(j); afterwards the usual 3-plane halo push.
    Returns (this slab's final canonical rows, av_vels partial)."""
    rows_all, first_all = pkg.decompose(ny, size)
    rows, first = int(rows_all[rank]), int(first_all[rank])
    inv = pkg.free_cells_inv(obstacles)
    buf = np.full((rows + 2, nx, 9), np.nan, np.float32)
    buf[1:rows + 1] = cells0[first:first + rows]
    tmp = np.full_like(buf, np.nan)
    ob = np.zeros((rows + 2, nx), np.int32)
    ob[1:rows + 1] = obstacles[first:first + rows]
    accel_row = ny - 2 - first + 1 if first <= ny - 2 < first + rows else -1
    av = np.zeros(iters, np.float32)
    exchange_halos(buf, rank, size)
    odd = 0

    def canonical_row(r):
        """populations of row r (an interior row) in the reference's order, whatever the layout"""
        if not odd:
            return buf[r]                                        # a view: modified in place
        return np.stack([np.roll(buf[r + CY[k], :, OPP[k]], -CX[k]) for k in range(9)], axis=1)

    def store_canonical_row(r, vals):
        if odd:
            for k in range(9):
                buf[r + CY[k], :, OPP[k]] = np.roll(vals[:, k], CX[k])

    for t in range(iters):
        if accel_row > 0:
            vals = np.ascontiguousarray(canonical_row(accel_row))
            oracle.accelerate_row(vals, ob[accel_row], density, accel)
            if odd:
                store_canonical_row(accel_row, vals)
            else:
                buf[accel_row] = vals
        if not odd:
            av[t] = oracle.slab_timestep(np.nan_to_num(buf, nan=0.0), tmp, ob, 1, rows + 1, omega) * inv
            new = np.full_like(buf, np.nan)
            for i in range(9):
                new[1 - CY[i]:rows + 1 - CY[i], :, i] = np.roll(tmp[1:rows + 1, :, OPP[i]], -CX[i], axis=1)
            # rows 0 and rows+1 of `new` belong to the neighbours' owned edge rows
            from_south, from_north = swap_rows(new[rows + 1][:, SOUTH_PLANES], new[0][:, NORTH_PLANES], rank, size)
            new[1][:, SOUTH_PLANES] = from_south             # written by the southern slab's last row
            new[rows][:, NORTH_PLANES] = from_north          # written by the northern slab's first row
            new[0] = np.nan
            new[rows + 1] = np.nan
            buf = new
        else:
            pulled = buf[1:rows + 1][:, :, OPP]                  # population j sits in slot opp(j) of its own cell
            g = np.zeros_like(buf)
            for j in range(9):                                   # lay them out so that the oracle's pull finds them
                g[1 - CY[j]:rows + 1 - CY[j], :, j] = np.roll(pulled[:, :, j], -CX[j], axis=1)
            av[t] = oracle.slab_timestep(g, tmp, ob, 1, rows + 1, omega) * inv
            buf = np.full_like(buf, np.nan)
            buf[1:rows + 1] = tmp[1:rows + 1]
            exchange_halos(buf, rank, size)
        odd ^= 1
    if odd:
        # canonical population k of an edge-row cell may live in the neighbour's owned edge row
        from_south, from_north = swap_rows(buf[rows][:, NORTH_PLANES], buf[1][:, SOUTH_PLANES], rank, size)
        ext = buf.copy()
        ext[0][:, NORTH_PLANES] = from_south                 # the southern slab's last row, slots 2,5,6
        ext[rows + 1][:, SOUTH_PLANES] = from_north          # the northern slab's first row, slots 4,7,8
        out = np.empty((rows, nx, 9), np.float32)
        for k in range(9):
            out[:, :, k] = np.roll(ext[1 + CY[k]:rows + 1 + CY[k], :, OPP[k]], -CX[k], axis=1)
        return out, av
    return buf[1:rows + 1].copy(), av


# ---- two timesteps per pass on a ring (csrc/lbm_kernels.cuh kernel 5): two halo rows per side, one exchange per pass ----
def run_slab_fused2(pkg, oracle, rank, size, nx, ny, iters, density, accel, omega, obstacles, cells0):
    """The fused protocol of the product with the oracle's stepper: per pass the slab computes the FIRST step also for
    its neighbours' edge rows (from two halo rows per side), the SECOND step for its own rows, and then pushes exactly
    the plane rows the neighbours' next pass pulls:
        north: planes 0,1,3,2,5,6 of the last row, planes 2,5,6 (+3,7) of the row below it
        south: planes 0,1,3,4,7,8 of the first row, planes 4,7,8 of the row above it
    Everything else in the halo rows stays NaN.  An odd last step is a one-step pass with the same pushes.  The slab
    north of the driven row's owner applies the body-force pre-pass to its copy of that row.
    Array rows: index i <-> slab row i-2, i.e. [row -2, row -1, rows 0..R-1, row R, row R+1]."""
    rows_all, first_all = pkg.decompose(ny, size)
    R, first = int(rows_all[rank]), int(first_all[rank])
    inv = pkg.free_cells_inv(obstacles)
    south, north = (rank - 1) % size, (rank + 1) % size
    gy = [(first + i - 2) % ny for i in range(R + 4)]
    ob = np.ascontiguousarray(obstacles[gy])                    # obstacle rows incl. the neighbours' rows (static)
    a = np.full((R + 4, nx, 9), np.nan, np.float32)
    a[2:R + 2] = cells0[first:first + R]
    accel_i = (ny - 2 - first) + 2 if first <= ny - 2 < first + R else -1      # array index of the driven row, if owned
    copy_i = 0 if first == 0 and size > 1 else -1                               # rank 0 keeps a copy of it in row -2
    av = np.zeros(iters, np.float32)

    def exchange(state):
        """pushes this slab's four edge rows into the neighbours' halo rows (only the planes their next pass pulls)"""
        def pack(row, planes):
            out = np.full((nx, 9), np.nan, np.float32)
            out[:, planes] = state[row][:, planes]
            return out
        up = np.stack([pack(R + 1, [0, 1, 3, 2, 5, 6]), pack(R, [2, 5, 6, 3, 7])])       # last row, row below it
        down = np.stack([pack(2, [0, 1, 3, 4, 7, 8]), pack(3, [4, 7, 8])])               # first row, row above it
        from_south, from_north = swap_rows(up, down, rank, size)
        state[1], state[0] = from_south[0], from_south[1]        # rows -1 and -2
        state[R + 2], state[R + 3] = from_north[0], from_north[1]   # rows R and R+1

    def force(state, i):
        vals = np.ascontiguousarray(state[i])
        oracle.accelerate_row(vals, ob[i], density, accel)
        state[i] = vals

    exchange(a)
    t = 0
    while t < iters:
        # body force of the step that comes next (pre-pass of a run, or folded into the previous store): here always
        # applied explicitly before the step, on the owner and on rank 0's copy
        if accel_i > 0:
            force(a, accel_i)
        if copy_i >= 0:
            force(a, copy_i)
        b = np.full_like(a, np.nan)
        src = a                                                  # NaN-poisoned: a pull of anything not pushed would show
        if iters - t >= 2:
            # first step: neighbours' edge rows (redundant) and own rows; only the own rows count for the average
            oracle.slab_timestep(src, b, ob, 1, 2, omega)
            av[t] = oracle.slab_timestep(src, b, ob, 2, R + 2, omega) * inv
            oracle.slab_timestep(src, b, ob, R + 2, R + 3, omega)
            if accel_i > 0:
                force(b, accel_i)                                # the second step's force on the first step's result
            c = np.full_like(a, np.nan)
            av[t + 1] = oracle.slab_timestep(b, c, ob, 2, R + 2, omega) * inv
            a = c
            t += 2
        else:
            av[t] = oracle.slab_timestep(src, b, ob, 2, R + 2, omega) * inv
            a = b
            t += 1
        a[0] = a[1] = a[R + 2] = a[R + 3] = np.nan
        exchange(a)
    return a[2:R + 2].copy(), av


# ---- K timesteps per pass on a ring (csrc/lbm_stepsk.cuh kernel 7): four halo rows per side, one exchange per pass ----
def run_slab_fusedk(pkg, oracle, rank, size, nx, ny, iters, density, accel, omega, obstacles, cells0, steps=4):
    """Kernel 7's ring protocol with the oracle's stepper: a slab keeps H = 4 halo rows per side (all nine planes); in a
    pass of k <= `steps` timesteps, step s is computed for the rows [-(k-s), R+(k-s)) -- the neighbours' rows out of the
    halo rows, redundantly --, step k for the slab's own rows only; then the slab's first and last four rows go into
    the neighbours' halo rows.  Halo rows a pass does not recompute are NaN afterwards: a pull of anything that was not
    exchanged would show.  The body force of the step that comes next is applied to the driven row by its owner and --
    while the steps still recompute it -- to the copy the slab north of the owner holds as its row -2.
    Array rows: index i <-> slab row i-4."""
    H = 4
    rows_all, first_all = pkg.decompose(ny, size)
    R, first = int(rows_all[rank]), int(first_all[rank])
    assert min(int(r) for r in rows_all) >= H + 2
    inv = pkg.free_cells_inv(obstacles)
    gy = [(first + i - H) % ny for i in range(R + 2 * H)]
    ob = np.ascontiguousarray(obstacles[gy])
    a = np.full((R + 2 * H, nx, 9), np.nan, np.float32)
    a[H:R + H] = cells0[first:first + R]
    accel_i = (ny - 2 - first) + H if first <= ny - 2 < first + R else -1      # array index of the driven row, if owned
    copy_i = H - 2 if first == 0 and size > 1 else -1                           # rank 0's copy of it: row -2
    av = np.zeros(iters, np.float32)

    def exchange(state):
        up = np.stack([state[H + R - d] for d in range(1, H + 1)])              # rows R-1 .. R-4 go north
        down = np.stack([state[H + d - 1] for d in range(1, H + 1)])            # rows 0 .. 3 go south
        from_south, from_north = swap_rows(up, down, rank, size)
        for d in range(1, H + 1):
            state[H - d] = from_south[d - 1]                                    # the southern slab's d-th row from its end
            state[H + R + d - 1] = from_north[d - 1]                            # the northern slab's d-th row

    def force(state, i):
        vals = np.ascontiguousarray(state[i])
        oracle.accelerate_row(vals, ob[i], density, accel)
        state[i] = vals

    exchange(a)
    t = 0
    while t < iters:
        k = min(steps, iters - t)
        cur = a
        for s in range(1, k + 1):
            if accel_i > 0:
                force(cur, accel_i)
            if copy_i >= 0 and s <= k - 1:                       # row -1 of step s pulls from the forced row -2
                force(cur, copy_i)
            nxt = np.full_like(cur, np.nan)
            lo, hi = H - (k - s), H + R + (k - s)
            if lo < H:
                oracle.slab_timestep(cur, nxt, ob, lo, H, omega)
            av[t + s - 1] = oracle.slab_timestep(cur, nxt, ob, H, H + R, omega) * inv   # only the own rows count
            if hi > H + R:
                oracle.slab_timestep(cur, nxt, ob, H + R, hi, omega)
            cur = nxt
        a = cur
        a[:H] = np.nan
        a[H + R:] = np.nan
        exchange(a)
        t += k
    return a[H:R + H].copy(), av


def worker(rank, size, port, args, out_queue, inplace=False):
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
        runner = {False: run_slab, True: run_slab_inplace, "fused2": run_slab_fused2,
                  "fused3": lambda *a: run_slab_fusedk(*a, steps=3), "fused4": lambda *a: run_slab_fusedk(*a, steps=4)}[inplace]
        cells, av = runner(pkg, oracle_lib, rank, size, *args)
        # the final reduction of the per-rank av_vels arrays (reference d2q9-bgk.c:396)
        total = torch.from_numpy(av.copy())
        dist.reduce(total, dst=0, op=dist.ReduceOp.SUM)
        out_queue.put((rank, cells, total.numpy() if rank == 0 else None))
    finally:
        dist.destroy_process_group()
