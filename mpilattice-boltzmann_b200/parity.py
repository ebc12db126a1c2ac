"""The multi-GPU parity case bench.py runs before its timed region (and tests/ re-use).

A ring of N row slabs, 32 rows each, on a grid that is periodic in x with period 64: walls on the first and the
last global row, small obstacles on and next to every slab boundary (so the halo rows carry bounce-back cells),
driven along row ny-2 like every deck.  The expected populations after STEPS timesteps were computed by the oracle
for the 64-wide grid (tests/golden/make_ring_parity.py -> tests/golden/ring_parity.npz); a grid whose width is a
multiple of 64 holds that solution tiled, so every rank can compare its slab BIT FOR BIT without any CPU solver
on the product path.
"""
from __future__ import annotations

import os

import numpy as np

PERIOD = 64
ROWS_PER_RANK = 32
RANK_COUNTS = (1, 2, 4, 8)
RUNS = (7, 6)                      # two back-to-back runs: pair passes + an odd tail, then an even run
STEPS = sum(RUNS)
DENSITY, ACCEL, OMEGA = 0.1, 0.005, 1.85
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ring_parity.npz")


def obstacles(nx: int, ny: int, n_ranks: int) -> np.ndarray:
    """int32 [ny, nx] obstacle map of the parity case (nx a multiple of PERIOD, ny = ROWS_PER_RANK * n_ranks)."""
    if nx % PERIOD or ny != ROWS_PER_RANK * n_ranks:
        raise ValueError("parity grid must be a multiple of 64 wide and 32 rows per rank")
    tile = np.zeros((ny, PERIOD), np.int32)
    tile[0, :] = 1
    tile[ny - 1, :] = 1
    for r in range(1, n_ranks):                      # on and around every interior slab boundary
        y = r * ROWS_PER_RANK
        tile[y - 2, 5:9] = 1                         # two rows below the boundary (second halo row of the slab above)
        tile[y - 1, 20:23] = 1                       # last row of the lower slab
        tile[y, 21:27] = 1                           # first row of the upper slab
        tile[y + 1, 40:42] = 1
        tile[y - 1, 62:64] = 1                       # touches the x-period boundary
        tile[y, 0:2] = 1
    tile[ny // 2 + 3, 30:34] = 1                     # something in the interior of a slab
    tile[ny - 3, 10:12] = 1                          # next to the driven row
    return np.tile(tile, (1, nx // PERIOD))


def expected(n_ranks: int):
    """(cells float32 [ny, 64, 9], av_vels float32 [STEPS]) from the committed fixture."""
    with np.load(GOLDEN) as z:
        return z[f"cells_n{n_ranks}"], z[f"av_vels_n{n_ranks}"]


def compare_slab(cells: np.ndarray, first_row: int, n_ranks: int) -> int:
    """Number of populations of this slab ([rows, nx, 9] float32) that differ in any bit from the expected tiling."""
    ref, _ = expected(n_ranks)
    rows, nx, _ = cells.shape
    got = np.ascontiguousarray(cells, np.float32).view(np.uint32).reshape(rows, nx // PERIOD, PERIOD, 9)
    want = np.ascontiguousarray(ref[first_row:first_row + rows]).view(np.uint32)[:, None, :, :]
    return int(np.count_nonzero(got != want))
