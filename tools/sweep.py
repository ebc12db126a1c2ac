#!/usr/bin/env python3
"""Sweeps the step kernel's tuning knobs on one GPU and prints MLUPS / GB/s per setting.

    python tools/sweep.py [--nx 16384 --ny 16384 --timesteps 100] [--json out.json]

Every setting runs the same state forward (results do not depend on the knobs -- tests/
test_gpu_parity.py::test_launch_geometry_does_not_change_results)."""
import argparse
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=16384)
    ap.add_argument("--ny", type=int, default=16384)
    ap.add_argument("--timesteps", type=int, default=100)
    ap.add_argument("--json", default=None)
    ap.add_argument("--kernels", default="2")
    ap.add_argument("--min-ctas", default="2,3,4")
    ap.add_argument("--ctas-per-sm", default="0")
    ap.add_argument("--hints", default="0,1,2")
    ap.add_argument("--graph-steps", default="0")
    ap.add_argument("--resident", default="0")
    ap.add_argument("--fused2", action="store_true", help="two timesteps per pass over HBM (kernel 5)")
    ap.add_argument("--band-rows", default="64")
    ap.add_argument("--inplace", action="store_true", help="one population buffer (AA access pattern); kernels/min-ctas/resident are ignored")
    args = ap.parse_args()
    pkg = entry.load_package()
    obstacles = pkg.decks.channel_obstacles(args.nx, args.ny)
    results = []
    if args.inplace:
        args.kernels, args.min_ctas, args.resident = "4", "2", "0"
    with pkg.Simulation(args.nx, args.ny, 0.1, 0.005, 1.85, obstacles, inplace=args.inplace) as sim:
        lists = [[int(v) for v in s.split(",")] for s in (args.kernels, args.min_ctas, args.ctas_per_sm, args.hints, args.graph_steps, args.resident)]
        bands = [int(v) for v in args.band_rows.split(",")] if args.fused2 else [0]
        for (kernel, min_ctas, per_sm, hint, graph, resident), band in itertools.product(itertools.product(*lists), bands):
            if kernel == 1 and (min_ctas != lists[1][0] or hint != lists[3][0]):
                continue
            if not args.inplace:
                sim.set_option("resident", resident)
                sim.set_option("kernel", kernel)
                sim.set_option("min_ctas", min_ctas)
            if args.fused2:
                sim.set_option("band_rows", band)
                sim.set_option("fused2", 1)
            elif not args.inplace:
                sim.set_option("fused2", 0)
            sim.set_option("ctas_per_sm", per_sm)
            sim.set_option("cache_hint", hint)
            sim.set_option("graph_steps", graph)
            sim.enqueue(max(10, args.timesteps // 5)); sim.sync()           # warm-up
            best = None
            for _ in range(3):
                sim.enqueue(args.timesteps); sim.sync()
                ms = sim.elapsed_ms()
                best = ms if best is None else min(best, ms)
            mlups = args.nx * args.ny * args.timesteps / (best * 1e-3) / 1e6
            rec = {"band_rows": band, "kernel": sim.get_option("kernel"), "min_ctas": min_ctas, "ctas_per_sm": per_sm, "cache_hint": hint, "graph_steps": graph, "resident": resident,
                   "grid": sim.get_option("grid"), "threads": sim.get_option("threads"),
                   "us_per_step": round(best * 1e3 / args.timesteps, 2), "mlups": round(mlups, 1),
                   "gbs": round(mlups * 72e-3, 1)}
            results.append(rec)
            print(json.dumps(rec), flush=True)
    if args.json:
        with open(args.json, "w") as fh:
            json.dump({"nx": args.nx, "ny": args.ny, "timesteps": args.timesteps, "results": results}, fh, indent=1)


if __name__ == "__main__":
    main()
