#!/usr/bin/env python3
"""Records the DRAM bytes per launch of one kernel from an `ncu --set full` capture in profiles/roofline_traffic.json.

    ncu -i X.ncu-rep --page raw --csv > X_raw.csv
    tools/ncu_to_traffic.py X_raw.csv <kernel number: 2 | 4 | 5 | 7> <nx> <ny> "<what was captured>" [timesteps per launch]

bench.py reads the record for `roofline.traffic`; the record carries the hash of the kernel sources the capture was
taken on (bench.kernel_source_hash), so a capture of older kernels is never quoted for newer ones.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    raw, kernel, nx, ny, what = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    rows = list(csv.reader(open(raw)))
    hdr, units, first = rows[0], rows[1], rows[2]

    def metric(name):
        i = hdr.index(name)
        return float(first[i]) * UNIT[units[i]]

    rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
    steps = int(sys.argv[6]) if len(sys.argv) > 6 else (2 if kernel == "5" else 1)
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    top = json.load(open(path)) if os.path.exists(path) else {}
    top.setdefault("kernels", {})[kernel] = {
        "capture": what, "kernel_name": first[hdr.index("Kernel Name")], "source_hash": kernel_source_hash(kernel),
        "nx": nx, "ny": ny, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
        "dram_bytes_per_cell_per_step": (rd + wr) / (nx * ny) / steps,
        "algorithmic_bytes_per_launch": 72 * steps * nx * ny,
        "gpu_time_under_ncu": first[hdr.index("gpu__time_duration.sum")] + " " + units[hdr.index("gpu__time_duration.sum")],
    }
    for k in ("source", "nx", "ny", "dram_bytes_read", "dram_bytes_write", "dram_bytes_per_launch", "dram_bytes_per_cell",
              "algorithmic_bytes_per_launch"):
        top.pop(k, None)                                # the round-1 top-level record (no source hash)
    json.dump(top, open(path, "w"), indent=1)
    print(json.dumps(top["kernels"][kernel]))


if __name__ == "__main__":
    main()
