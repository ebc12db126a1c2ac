#!/usr/bin/env python3
"""Runs the reference's OWN acceptance script, check/check.py, byte-for-byte as it is on disk.

check/check.py is Python-2.7-only (a version gate at lines 6-10 and print statements); no
Python 2 exists here.  This wrapper reads the file, disables the gate and turns the print
statements into calls IN MEMORY (three regular expressions), then executes it with the given
command line.  Nothing of the reference is copied into this repository.

    python tools/run_check.py [--reference /root/reference] -- --ref-av-vels-file=... \\
        --ref-final-state-file=... --av-vels-file=... --final-state-file=...

Where /root/reference is absent (the GPU box) use tools/check.py, the Python 3 restatement;
tests/test_check_tool.py proves the two give the same verdict and numbers.
"""
import os
import re
import sys


def load_reference_checker(reference="/root/reference"):
    path = os.path.join(reference, "check", "check.py")
    with open(path, "r") as fh:
        src = fh.read()
    src = src.replace("if sys.version_info[:2] != (2,7):", "if False:")
    src = re.sub(r"^([ \t]*)print[ \t]+(\S.*)$", r"\1print(\2)", src, flags=re.M)
    src = re.sub(r"^([ \t]*)print[ \t]*$", r"\1print()", src, flags=re.M)
    return compile(src, path, "exec")


def run(argv, reference="/root/reference") -> int:
    code = load_reference_checker(reference)
    old_argv = sys.argv
    sys.argv = [os.path.join(reference, "check", "check.py")] + list(argv)
    try:
        exec(code, {"__name__": "__main__", "exit": sys.exit})
    except SystemExit as e:
        return int(e.code or 0)
    finally:
        sys.argv = old_argv
    return 0


if __name__ == "__main__":
    args = sys.argv[1:]
    reference = "/root/reference"
    if args and args[0] == "--reference":
        reference, args = args[1], args[2:]
    if args and args[0] == "--":
        args = args[1:]
    sys.exit(run(args, reference))
