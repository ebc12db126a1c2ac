#!/usr/bin/env python3
"""Runs a few in-place timesteps at full size (for ncu: per-flavour durations and DRAM bytes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
nx = int(os.environ.get("N", 16384))
ny = int(os.environ.get("NY", nx))
steps = int(os.environ.get("STEPS", 6))
slabs = int(os.environ.get("SLABS", 1))
with pkg.Simulation(nx, ny, 0.1, 0.005, 1.85, pkg.decks.channel_obstacles(nx, ny), n_slabs=slabs, devices=[0] * slabs,
                    inplace=bool(int(os.environ.get("INPLACE", 1)))) as sim:
    sim.set_option("graph_steps", 0)
    if not sim.get_option("inplace"):
        sim.set_option("band_rows", int(os.environ.get("BAND", 0)))
        sim.set_option("fused2", int(os.environ.get("FUSED2", 0)))
    sim.enqueue(steps)
    sim.sync()
    sim.enqueue(steps)
    sim.sync()
    print("slabs", slabs, "inplace", sim.get_option("inplace"), "kernel", sim.get_option("kernel"), "ms per step", sim.elapsed_ms() / steps)
