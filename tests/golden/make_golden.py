#!/usr/bin/env python3
"""Generates the committed fixtures under tests/golden/ from /root/reference.

Run once in the build container (needs /root/reference and `make -C oracle ref`); the GPU box
has neither, so everything the -m gpu tests need is committed here:

  decks/input_*.params, decks/obstacles_*.dat   the reference's four input decks (data, verbatim)
  <deck>.av_vels.dat.gz, <deck>.final_state.dat.gz
                                                the reference's golden outputs check/*.dat
                                                (verbatim bytes, gzip'd; two final_state goldens
                                                are absent upstream -- .MISSING_LARGE_BLOBS)
  ref_strict.json + ref_strict.<deck>.av_vels.npy
                                                outputs of the UNMODIFIED reference source built
                                                strict-IEEE (oracle/_ref/d2q9-bgk.strict, 1 rank),
                                                full length: sha256 of final_state.dat, Reynolds
                                                number, av_vels as float32, and the pressure
                                                column sub-sampled every 8th row/column
"""
import concurrent.futures
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("REFERENCE", "/root/reference")
DECKS = ["128x128", "128x256", "256x256", "1024x1024"]
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "d2q9-bgk.strict")


def copy_inputs():
    os.makedirs(os.path.join(HERE, "decks"), exist_ok=True)
    for d in DECKS:
        for name in (f"input_{d}.params", f"obstacles_{d}.dat"):
            shutil.copyfile(os.path.join(REFERENCE, name), os.path.join(HERE, "decks", name))
    for name in sorted(os.listdir(os.path.join(REFERENCE, "check"))):
        if name.endswith(".dat"):
            with open(os.path.join(REFERENCE, "check", name), "rb") as src, \
                    gzip.GzipFile(os.path.join(HERE, name + ".gz"), "wb", compresslevel=9, mtime=0) as dst:
                dst.write(src.read())


def run_ref(deck):
    with tempfile.TemporaryDirectory() as tmp:
        out = subprocess.run([REF_BIN, os.path.join(REFERENCE, f"input_{deck}.params"),
                              os.path.join(REFERENCE, f"obstacles_{deck}.dat")],
                             cwd=tmp, check=True, capture_output=True, text=True).stdout
        reynolds = [l.split()[-1] for l in out.splitlines() if l.startswith("Reynolds")][0]
        fs_bytes = open(os.path.join(tmp, "final_state.dat"), "rb").read()
        av = np.loadtxt(os.path.join(tmp, "av_vels.dat"), usecols=[1]).astype(np.float32)
        fs = np.loadtxt(os.path.join(tmp, "final_state.dat"), usecols=[0, 1, 5])
    nx = int(fs[:, 0].max()) + 1
    ny = int(fs[:, 1].max()) + 1
    pressure = fs[:, 2].reshape(ny, nx).astype(np.float32)
    np.save(os.path.join(HERE, f"ref_strict.{deck}.av_vels.npy"), av)
    np.save(os.path.join(HERE, f"ref_strict.{deck}.pressure_sub8.npy"), pressure[::8, ::8].copy())
    return deck, {"final_state_sha256": hashlib.sha256(fs_bytes).hexdigest(),
                  "final_state_bytes": len(fs_bytes), "reynolds": reynolds,
                  "pressure_sha256": hashlib.sha256(pressure.tobytes()).hexdigest()}


def main():
    if not os.path.isdir(REFERENCE):
        sys.exit(f"{REFERENCE} is not present: fixtures can only be regenerated in the build container")
    copy_inputs()
    with concurrent.futures.ThreadPoolExecutor(4) as pool:
        results = dict(pool.map(run_ref, DECKS))
    meta = {"generator": "tests/golden/make_golden.py",
            "binary": "oracle/_ref/d2q9-bgk.strict (gcc -std=gnu99 -O3 -march=x86-64-v3 -ffp-contract=off, 1 rank, mpi.h shim)",
            "decks": results}
    with open(os.path.join(HERE, "ref_strict.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print(json.dumps(meta, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
