#!/usr/bin/env python3
"""Acceptance check for av_vels.dat / final_state.dat -- Python 3, same command line, same
metric and same exit status as the reference's check/check.py (which is Python-2-only and is
not present on the GPU box).

Contract restated from check/check.py:
  * av_vels: column 1 of every line (check.py:65); both files must have the same length (80-82)
  * final_state: columns 0, 1 (coordinates, must be identical and in the same order, 75-77)
    and column 5 (pressure) (66)
  * diff = ref - sim, percentage = 100 * diff / (ref - diff) (86-87); the largest |percentage|
    must be finite and <= --tolerance (default 1 %) for both files (134-135)
  * exit 0 and "Both tests passed!" or exit 1 (142-147)

Extras: reference files may be gzip-compressed (*.gz), so the committed fixtures are usable
directly.  tools/run_check.py runs the reference's own script instead when it is available.
"""
import argparse
import gzip
import sys

import numpy as np


def _open(path):
    return gzip.open(path, "rt") if path.endswith(".gz") else open(path, "r")


def load(av_path, fs_path):
    with _open(av_path) as fh:
        av = np.loadtxt(fh, usecols=[1], ndmin=1)
    with _open(fs_path) as fh:
        fs = np.loadtxt(fh, usecols=[0, 1, 5], ndmin=2)
    return av, fs


def compare(ref, sim):
    """check.py:84-99 -- worst relative difference, in percent of the simulated value."""
    diff = ref - sim
    with np.errstate(divide="ignore", invalid="ignore"):
        pcnt = 100.0 * (diff / (ref - diff))
    # argmax(abs) would skip NaNs silently; any non-finite entry must fail the check
    bad = ~np.isfinite(pcnt)
    worst = int(np.argmax(bad)) if bad.any() else int(np.argmax(np.abs(pcnt)))
    return {"index": worst, "diff": float(diff[worst]), "pcnt": float(pcnt[worst]),
            "sim": float(sim[worst]), "ref": float(ref[worst]), "total": float(np.sum(np.abs(diff)))}


def check(ref_av, ref_fs, sim_av, sim_fs, tolerance=1.0, out=sys.stdout):
    av_ref, fs_ref = load(ref_av, ref_fs)
    av_sim, fs_sim = load(sim_av, sim_fs)
    if fs_ref.shape != fs_sim.shape or np.any(fs_ref[:, 0:2] != fs_sim[:, 0:2]):
        print("Final state files coordinates were not the same", file=out)
        return 1
    if av_ref.size != av_sim.size:
        print("Different number of steps in av_vels files", file=out)
        return 1
    a = compare(av_ref, av_sim)
    print(f"Total difference in av_vels : {a['total']:.12E}", file=out)
    print(f"Biggest difference (at step {a['index']:d}) : {a['diff']:.12E}", file=out)
    print(f"  {a['sim']:.12E} vs. {a['ref']:.12E} = {a['pcnt']:.2g}%", file=out)
    print(file=out)
    f = compare(fs_ref[:, 2], fs_sim[:, 2])
    jj, ii = int(fs_sim[f["index"], 0]), int(fs_sim[f["index"], 1])
    print(f"Total difference in final_state : {f['total']:.12E}", file=out)
    print(f"Biggest difference (at coord ({jj:d},{ii:d})) : {f['diff']:.12E}", file=out)
    print(f"  {f['sim']:.12E} vs. {f['ref']:.12E} = {f['pcnt']:.2g}%", file=out)
    print(file=out)
    fs_failed = (not np.isfinite(f["pcnt"])) or abs(f["pcnt"]) > tolerance
    av_failed = (not np.isfinite(a["pcnt"])) or abs(a["pcnt"]) > tolerance
    if fs_failed:
        print("final state failed check", file=out)
    if av_failed:
        print("av_vels failed check", file=out)
    if fs_failed or av_failed:
        return 1
    print("Both tests passed!", file=out)
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser(description="Testing script for the D2Q9-BGK outputs (Python 3)",
                                 formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--tolerance", type=float, default=1.0, help="percentage tolerance against the reference results")
    ap.add_argument("--ref-av-vels-file", required=True, help="reference av_vels results file")
    ap.add_argument("--ref-final-state-file", required=True, help="reference final_state results file")
    ap.add_argument("--av-vels-file", required=True, help="calculated av_vels results file")
    ap.add_argument("--final-state-file", required=True, help="calculated final_state results file")
    args = ap.parse_args(argv)
    return check(args.ref_av_vels_file, args.ref_final_state_file, args.av_vels_file, args.final_state_file,
                 args.tolerance)


if __name__ == "__main__":
    sys.exit(main())
