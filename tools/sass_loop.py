#!/usr/bin/env python3
"""Static estimate of the hot path of a kernel's main loop from cuobjdump -sass: the loop is the backward branch
spanning the most packed-fp32 instructions; straight-line regions that contain CALL / STL / LDL (slow paths: obstacle
rows, out-of-range operands) are left out.  Prints an opcode histogram of what remains.

    python tools/sass_loop.py <kernel-name-regex> [--list]
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mpilattice-boltzmann_b200", "lib", "liblbm_b200.so")
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
ins, on = [], False
for l in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", l)
    if m:
        if on: break
        on = bool(re.search(pat, m.group(1)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if on and m: ins.append((int(m.group(1), 16), m.group(2).strip()))
def op(i):
    p = i.split()
    if p[0].startswith("@"): p = p[1:]
    return p[0].split(".")[0]
best = None
for a, i in ins:
    mm = re.search(r"0x([0-9a-f]+)", i)
    if op(i) == "BRA" and mm and int(mm.group(1), 16) < a:
        t = int(mm.group(1), 16)
        n = sum(1 for b, j in ins if t <= b <= a and op(j) in ("FADD2", "FFMA2"))
        if best is None or n > best[0]: best = (n, t, a)
_, lo, hi = best
body = [(a, i) for a, i in ins if lo <= a <= hi]
# split into regions at branch targets and after branches
targets = set()
for a, i in body:
    mm = re.search(r"0x([0-9a-f]+)", i)
    if op(i) in ("BRA", "BSSY", "CALL") and mm: targets.add(int(mm.group(1), 16))
regions, cur = [], []
for a, i in body:
    if a in targets and cur: regions.append(cur); cur = []
    cur.append((a, i))
    if op(i) in ("BRA", "CALL", "RET", "EXIT"): regions.append(cur); cur = []
if cur: regions.append(cur)
hot = [r for r in regions if not any(op(i) in ("CALL", "STL", "LDL") or "FTZ" in i for _, i in r)]
h = collections.Counter(op(i) for r in hot for _, i in r)
print(f"loop 0x{lo:x}..0x{hi:x}: {len(body)} instructions, hot-path estimate {sum(h.values())}")
for k, n in h.most_common(40): print(f"  {n:5d} {k}")
if "--list" in sys.argv:
    for r in hot:
        for a, i in r: print(f"    {a:05x} {i[:100]}")
