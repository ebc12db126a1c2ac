#!/usr/bin/env python3
"""Ring overhead on small slabs, one device: 16384 x 2048 as one handle against 2 ring slabs of 2048 rows each."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
nx = 16384
for rows, n_slabs, opts in [(2048, 1, {}), (4096, 2, {})]:
    ob = np.zeros((rows, (nx + 31) // 32), np.uint32)
    ob[0, :] = ob[-1, :] = 0xFFFFFFFF
    with pkg.Simulation(nx, rows, 0.1, 0.005, 1.85, ob, n_slabs=n_slabs, devices=[0] * n_slabs, obstacles_format="bits") as sim:
        for k, v in opts.items():
            sim.set_option(k, v)
        sim.run(120)
        sim.run(600)
        ms = sim.elapsed_ms()
        print(json.dumps({"rows": rows, "n_slabs": n_slabs, "opts": opts, "kernel": sim.get_option("kernel"), "band_rows": sim.get_option("band_rows"),
                          "ms_per_100_steps_per_slab": round(ms / 6 / n_slabs, 3), "glups": round(nx * rows * 600 / ms / 1e6, 1)}), flush=True)
