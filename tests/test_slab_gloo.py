"""world_size-2 (and 3) gloo runs of the row-slab path on CPU: slabs + 3-plane halo exchange
reproduce the single-domain oracle bit for bit (final_state is decomposition-invariant;
av_vels agrees to summation-order noise)."""
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import random_cells, random_obstacles

import slab_ring


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("size,nx,ny,inplace,iters", [(2, 32, 24, False, 25), (2, 40, 17, False, 25), (3, 16, 19, False, 25),
                                                      (2, 32, 24, True, 12), (2, 40, 17, True, 13), (3, 16, 19, True, 9),
                                                      (2, 32, 24, "fused2", 12), (2, 40, 17, "fused2", 13), (3, 16, 19, "fused2", 9),
                                                      (2, 32, 24, "fused3", 13), (2, 40, 17, "fused4", 14), (3, 16, 25, "fused4", 11)])
def test_slab_ring_matches_single_domain(pkg, oracle, size, nx, ny, inplace, iters):
    """inplace: the one-buffer (AA access pattern) ring protocol -- what crosses the slabs in which step flavour --
    ending in either layout.  "fused2": two timesteps per pass with two halo rows per side, only the planes the product
    pushes (everything else in the halo rows is NaN), the driven row's copy on rank 0, an odd one-step tail.  "fused3" /
    "fused4": kernel 7's protocol -- four halo rows per side, the steps before the last recomputed for the neighbours'
    rows, the driven row's copy forced while it is still recomputed, shorter last passes."""
    rng = np.random.default_rng(size * 1000 + ny)
    density, accel, omega = 0.1, 0.005, 1.85
    obstacles = random_obstacles(rng, ny, nx, 0.08, walls=(ny % 2 == 0))   # odd ny: open top/bottom, y-wrap in play
    cells0 = random_cells(rng, ny, nx, density)
    inv = pkg.free_cells_inv(obstacles)

    whole = cells0.copy()
    av_whole = oracle.run(whole, obstacles, iters, density, accel, omega, inv)

    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = free_port()
    args = (nx, ny, iters, density, accel, omega, obstacles, cells0)
    procs = [ctx.Process(target=slab_ring.worker, args=(r, size, port, args, queue, inplace)) for r in range(size)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=300) for _ in range(size)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort(key=lambda r: r[0])
    stitched = np.concatenate([r[1] for r in results], axis=0)
    assert np.array_equal(stitched.view(np.uint32), whole.view(np.uint32)), "slab run differs from the single domain"
    av = results[0][2]
    assert np.max(np.abs(av - av_whole) / np.abs(av_whole)) < 1e-5
