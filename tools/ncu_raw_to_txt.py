#!/usr/bin/env python3
"""`ncu -i X.ncu-rep --page raw --csv` (one wide row per launch) -> one "metric<TAB>unit<TAB>value" line per metric.
    tools/ncu_raw_to_txt.py X_raw.csv "<header comment>" > profiles/NAME_ncu_metrics.txt"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
print("# " + sys.argv[2])
for name, unit, value in zip(rows[0], rows[1], rows[2]):
    print(f"{name}\t{unit}\t{value}")
