#!/usr/bin/env python3
"""Pins the C restatement (oracle/lbm_oracle.c) against the compiled, unmodified reference.

TEST INFRASTRUCTURE ONLY.  For each deck it runs oracle/_ref/d2q9-bgk.strict (1 rank) in a
scratch directory, runs the restatement on the same inputs, writes the restatement's results
in the reference's file formats and requires

  * final_state.dat  byte-identical  (every printed digit of u_x, u_y, |u|, pressure),
  * av_vels.dat      byte-identical  (every step),
  * the printed Reynolds number identical.

Usage:  python oracle/pin_oracle.py [--decks 128x128 128x256 ...] [--reference /root/reference]
Needs /root/reference (inputs) -- it is run in the build container, not on the GPU box.
"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
decks = pkg.decks
sys.path.insert(0, HERE)
import oracle_lib  # noqa: E402


def pin(deck: str, reference: str, iters_override: int | None = None) -> bool:
    pfile = os.path.join(reference, f"input_{deck}.params")
    ofile = os.path.join(reference, f"obstacles_{deck}.dat")
    p = decks.read_params(pfile)
    with tempfile.TemporaryDirectory() as tmp:
        if iters_override is not None:
            p.max_iters = iters_override
            pfile = os.path.join(tmp, "in.params")
            with open(pfile, "w") as fh:
                fh.write(p.as_text())
        t0 = time.time()
        out = subprocess.run([oracle_lib.ref_binary("strict"), pfile, ofile], cwd=tmp, check=True,
                             capture_output=True, text=True).stdout
        t_ref = time.time() - t0
        ref_reynolds = [l.split()[-1] for l in out.splitlines() if l.startswith("Reynolds")][0]
        obstacles, free = decks.read_obstacles(ofile, p.nx, p.ny)
        inv = decks.free_cells_inv(free)
        t0 = time.time()
        cells = oracle_lib.init_cells(p.nx, p.ny, p.density)
        av = oracle_lib.run(cells, obstacles, p.max_iters, p.density, p.accel, p.omega, inv)
        t_ora = time.time() - t0
        ux, uy, u, pr = oracle_lib.final_state(cells, obstacles, p.density)
        decks.write_av_vels(os.path.join(tmp, "o_av.dat"), av)
        decks.write_final_state(os.path.join(tmp, "o_fs.dat"), ux, uy, u, pr, obstacles)
        ora_reynolds = "%.12E" % float(oracle_lib.reynolds(cells, obstacles, inv, p.omega, p.reynolds_dim))
        same_av = open(os.path.join(tmp, "o_av.dat"), "rb").read() == open(os.path.join(tmp, "av_vels.dat"), "rb").read()
        same_fs = open(os.path.join(tmp, "o_fs.dat"), "rb").read() == open(os.path.join(tmp, "final_state.dat"), "rb").read()
    ok = same_av and same_fs and ref_reynolds == ora_reynolds
    print(f"{deck:>10} iters={p.max_iters:<6} av_vels {'IDENTICAL' if same_av else 'DIFFER'}  "
          f"final_state {'IDENTICAL' if same_fs else 'DIFFER'}  Reynolds ref {ref_reynolds} oracle {ora_reynolds}  "
          f"[ref {t_ref:.1f}s, oracle {t_ora:.1f}s]", flush=True)
    return ok


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--decks", nargs="+", default=["128x128", "128x256", "256x256", "1024x1024"])
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--iters", type=int, default=None, help="override maxIters (quick check)")
    args = ap.parse_args()
    ok = all([pin(d, args.reference, args.iters) for d in args.decks])
    print("oracle pinned: bit-identical to the strict reference build" if ok else "ORACLE NOT PINNED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
