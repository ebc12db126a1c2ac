#!/usr/bin/env python3
"""SASS opcode histogram of one kernel of lib/liblbm_b200.so (cuobjdump -sass), plus the packed-fp32 safety check:
every FFMA2 must be a multiply in disguise (addend = one broadcast scalar register, `Rn.F32`), and no FMUL2 may
exist -- see lbm::mul2 in csrc/lbm_cell.cuh.

    python tools/sass_hist.py [kernel-name-regex] [--lib path] [--check]
"""
import argparse, collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    name, body = None, {}
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            name = m.group(1)
            body[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and name:
            body[name].append(m.group(1).strip())
    return body


def opcode(instr):
    parts = instr.split()
    if parts[0].startswith("@"):
        parts = parts[1:]
    return parts[0].split(".")[0]


def unsafe_packed(instrs):
    """FFMA2 whose addend is not a single broadcast scalar register, and every FMUL2."""
    bad = []
    for i in instrs:
        op = opcode(i)
        if op == "FMUL2":
            bad.append(i)
        elif op == "FFMA2":
            addend = i.rsplit(",", 1)[1].strip()
            if not re.fullmatch(r"-?U?R\d+(\.reuse)?\.F32", addend):
                bad.append(i)
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("pattern", nargs="?", default="steps2_strip")
    ap.add_argument("--lib", default=os.path.join(ROOT, "mpilattice-boltzmann_b200", "lib", "liblbm_b200.so"))
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rc = 0
    for name, instrs in functions(args.lib).items():
        if not re.search(args.pattern, name):
            continue
        hist = collections.Counter(opcode(i) for i in instrs)
        bad = unsafe_packed(instrs)
        print(f"{name}: {len(instrs)} instructions, unsafe packed ops: {len(bad)}")
        if not args.check:
            for op, n in hist.most_common():
                print(f"  {n:6d} {op}")
        for b in bad[:10]:
            print("  UNSAFE", b)
        rc |= bool(bad)
    return rc


if __name__ == "__main__":
    sys.exit(main())
