#!/usr/bin/env python3
"""Generates tests/golden/ring_parity.npz: the expected state of the multi-GPU parity case that bench.py steps
through the real one-rank-per-GPU ring (Simulation.slab + CUDA IPC) before its timed region, for N = 1, 2, 4, 8.

The case (see mpilattice-boltzmann_b200/parity.py, which rebuilds the same obstacle map without this script):
32 rows per rank, channel walls on the first and the last global row, and a pattern of small obstacles that is
periodic in x with period PERIOD = 64 and sits on and next to every slab boundary, so that the halo exchange
carries bounce-back cells.  Because the initial state is uniform and everything is periodic in x with period 64,
the state of the nx = 16384 wide grid bench.py runs is the 64-wide solution tiled 256 times -- which the oracle
(oracle/lbm_oracle.c, pinned to the unmodified reference) computes here in milliseconds.

The product path never imports oracle/: bench.py only reads the arrays this script committed.

    python tests/golden/make_ring_parity.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as entry  # noqa: E402
import oracle_lib  # noqa: E402


def main():
    pkg = entry.load_package()
    par = pkg.parity
    out = {}
    for n in par.RANK_COUNTS:
        ny = par.ROWS_PER_RANK * n
        ob = par.obstacles(par.PERIOD, ny, n)
        free = par.PERIOD * ny - int(ob.sum())
        inv = pkg.decks.free_cells_inv(free)
        cells = oracle_lib.init_cells(par.PERIOD, ny, par.DENSITY)
        av = oracle_lib.run(cells, ob, par.STEPS, par.DENSITY, par.ACCEL, par.OMEGA, inv)
        out[f"cells_n{n}"] = cells
        out[f"av_vels_n{n}"] = av
        out[f"obstacles_n{n}"] = ob.astype(np.uint8)
        print(f"N={n}: {par.PERIOD}x{ny}, {int(ob.sum())} blocked, av_vels[-1] = {av[-1]:.9e}")
    np.savez_compressed(os.path.join(HERE, "ring_parity.npz"), **out)
    print("wrote", os.path.join(HERE, "ring_parity.npz"), os.path.getsize(os.path.join(HERE, "ring_parity.npz")), "bytes")


if __name__ == "__main__":
    main()
