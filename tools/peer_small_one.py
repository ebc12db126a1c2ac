#!/usr/bin/env python3
"""One case of tools/peer_small.py for ncu: peer_small_one.py <rows> <n_slabs>"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
nx, rows, n_slabs = 16384, int(sys.argv[1]), int(sys.argv[2])
ob = np.zeros((rows, (nx + 31) // 32), np.uint32)
ob[0, :] = ob[-1, :] = 0xFFFFFFFF
with pkg.Simulation(nx, rows, 0.1, 0.005, 1.85, ob, n_slabs=n_slabs, devices=[0] * n_slabs, obstacles_format="bits") as sim:
    sim.run(60)
    print(sim.elapsed_ms())
