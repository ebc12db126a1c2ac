#!/usr/bin/env python3
"""Ring overhead of the fused kernels on ONE device: the 16384 x 16384 channel as one handle against the same grid
as 2 and 4 ring slabs on device 0 (the PEER instantiations, halo pushes and flags; the slabs' launches serialise on
one stream, so the inter-GPU coupling is not in this number).   tools/peer_overhead.py [fused_steps ...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
nx = ny = 16384
ob = np.zeros((ny, (nx + 31) // 32), np.uint32)
ob[0, :] = ob[-1, :] = 0xFFFFFFFF
for steps in [int(a) for a in sys.argv[1:]] or [0]:
    for n_slabs in (1, 2, 4, -2):                    # -2: weak-like, two slabs of the full 16384 rows each
        rows = ny * 2 if n_slabs < 0 else ny
        n_slabs = abs(n_slabs)
        ob = np.zeros((rows, (nx + 31) // 32), np.uint32)
        ob[0, :] = ob[-1, :] = 0xFFFFFFFF
        with pkg.Simulation(nx, rows, 0.1, 0.005, 1.85, ob, n_slabs=n_slabs, devices=[0] * n_slabs, obstacles_format="bits") as sim:
            sim.set_option("fused_steps", steps)
            sim.run(120)
            sim.run(480)
            ms = sim.elapsed_ms()
            print(json.dumps({"fused_steps": sim.get_option("fused_steps"), "kernel": sim.get_option("kernel"), "n_slabs": n_slabs, "rows": rows,
                              "band_rows": sim.get_option("band_rows"), "ms_per_100_steps": round(ms / 4.8, 3),
                              "glups": round(nx * rows * 480 / ms / 1e6, 1)}), flush=True)
