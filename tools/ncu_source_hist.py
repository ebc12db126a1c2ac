#!/usr/bin/env python3
"""Dynamic opcode histogram and stall profile of one kernel from an ncu source page
(ncu -i X.ncu-rep --page source --csv > X_source.csv).

    python tools/ncu_source_hist.py X_source.csv [--top N] [--lines]
"""
import argparse, collections, csv, re, sys

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--top", type=int, default=30)
ap.add_argument("--lines", action="store_true", help="also list the hottest SASS lines by stall samples")
args = ap.parse_args()
rows = list(csv.reader(open(args.csv)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
inst = collections.Counter(); samples = collections.Counter(); total = 0; tot_s = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stalls = collections.Counter()
lines = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[col["Source"]].strip()
    parts = src.split()
    if parts[0].startswith("@"): parts = parts[1:]
    op = parts[0].split(".")[0]
    n = int(r[col["Instructions Executed"]] or 0); s = int(r[col["# Samples"]] or 0)
    inst[op] += n; samples[op] += s; total += n; tot_s += s
    for h in stall_cols: stalls[h] += int(r[col[h]] or 0)
    lines.append((s, n, r[col["Address"]][-5:], src))
print(f"warp instructions executed: {total:,}   stall samples: {tot_s:,}")
for op, n in inst.most_common(args.top):
    print(f"  {op:10s} {n:14,d} {100*n/total:6.2f}%   samples {100*samples[op]/max(tot_s,1):6.2f}%")
print("stall reasons (all samples):")
for h, n in stalls.most_common():
    if n: print(f"  {h:24s} {100*n/max(tot_s,1):6.2f}%")
if args.lines:
    for s, n, a, src in sorted(lines, reverse=True)[:60]:
        print(f"  {s:6d} {n:12,d} {a} {src}")
