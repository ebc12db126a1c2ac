"""Host-side logic of the C front end that needs no GPU: the exact fast formatter and the
command-line / die() contract of the reference (d2q9-bgk.c:197-205, 1145-1157)."""
import os
import subprocess

import pytest

from conftest import ROOT, deck_paths


def test_fast_format_matches_printf(tmp_path):
    exe = str(tmp_path / "fmt_check")
    subprocess.run(["/usr/bin/gcc", "-O2", "-std=gnu99", os.path.join(ROOT, "tests", "c", "fmt_check.c"), "-o", exe, "-lm"],
                   check=True)
    for seed in (1, 2, 3):
        res = subprocess.run([exe, "400000", str(seed)], capture_output=True, text=True)
        assert res.returncode == 0 and res.stdout.strip() == "0", res.stderr[-2000:]


def test_usage_message_and_exit_status(pkg):
    res = subprocess.run([pkg.EXE_PATH], capture_output=True, text=True)
    assert res.returncode == 1
    assert res.stderr == f"Usage: {pkg.EXE_PATH} <paramfile> <obstaclefile>\n"
    res = subprocess.run([pkg.EXE_PATH, "a", "b", "c"], capture_output=True, text=True)
    assert res.returncode == 1 and res.stderr.startswith("Usage: ")


def test_exe_alias_exists(pkg):
    assert os.path.exists(pkg.EXE_PATH + ".exe")        # README says d2q9-bgk.exe, Makefile builds d2q9-bgk


def test_die_messages(pkg, tmp_path):
    pfile, ofile = deck_paths("128x128")
    res = subprocess.run([pkg.EXE_PATH, str(tmp_path / "nope.params"), ofile], capture_output=True, text=True)
    assert res.returncode == 1
    lines = res.stderr.splitlines()
    assert lines[0].startswith("Error at line ") and " of file " in lines[0]
    assert lines[1] == f"could not open input parameter file: {tmp_path / 'nope.params'}"

    bad = tmp_path / "bad.params"
    bad.write_text("128\n128\nxyz\n")
    res = subprocess.run([pkg.EXE_PATH, str(bad), ofile], capture_output=True, text=True)
    assert res.returncode == 1 and res.stderr.splitlines()[1] == "could not read param file: maxIters"

    res = subprocess.run([pkg.EXE_PATH, pfile, str(tmp_path / "nope.dat")], capture_output=True, text=True)
    assert res.returncode == 1 and res.stderr.splitlines()[1] == f"could not open input obstacles file: {tmp_path / 'nope.dat'}"

    for text, message in (("1 2\n", "expected 3 values per line in obstacle file"),
                          ("128 0 1\n", "obstacle x-coord out of range"),
                          ("0 128 1\n", "obstacle y-coord out of range"),
                          ("0 0 3\n", "obstacle blocked value should be 1")):
        ob = tmp_path / "ob.dat"
        ob.write_text(text)
        res = subprocess.run([pkg.EXE_PATH, pfile, str(ob)], capture_output=True, text=True)
        assert res.returncode == 1 and res.stderr.splitlines()[1] == message


def test_no_gpu_is_a_loud_failure_not_a_fallback(pkg, tmp_path):
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    pfile, ofile = deck_paths("128x128")
    res = subprocess.run([pkg.EXE_PATH, pfile, ofile], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 1
    assert "no CUDA device available" in res.stderr
    assert not os.path.exists(tmp_path / "av_vels.dat")


def test_mp_launcher_usage_and_no_device_failure(pkg, tmp_path):
    """d2q9-bgk-mp (one process per GPU, the reference's mpirun -np N): usage text, and -- on a box without a GPU --
    every rank dies with the library's message, the launcher reports failure and nobody is left in a barrier."""
    exe = pkg.EXE_PATH + "-mp"
    assert os.path.exists(exe)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 1 and res.stderr == f"Usage: {exe} [-np N] <paramfile> <obstaclefile>\n"
    if pkg.device_count() > 0:
        pytest.skip("a GPU is visible: the launcher's run is covered by tests/test_gpu_multi.py")
    pfile, ofile = deck_paths("128x128")
    res = subprocess.run([exe, "-np", "2", pfile, ofile], capture_output=True, text=True, cwd=tmp_path, timeout=60)
    assert res.returncode == 1
    assert "no CUDA device available (there is no CPU fallback)" in res.stderr
    assert not (tmp_path / "av_vels.dat").exists()
