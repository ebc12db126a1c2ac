"""tools/check.py restates the reference's check/check.py; where the reference is present the
two are run side by side on the same files and must print the same numbers and verdict."""
import io
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

import check as check_tool

REFERENCE = "/root/reference"


def _write(tmp_path, name, av, fs):
    av_path, fs_path = str(tmp_path / f"{name}.av"), str(tmp_path / f"{name}.fs")
    with open(av_path, "w") as fh:
        fh.write("".join("%d:\t%.12E\n" % (i, v) for i, v in enumerate(av)))
    with open(fs_path, "w") as fh:
        for (x, y, p) in fs:
            fh.write("%d %d %.12E %.12E %.12E %.12E %d\n" % (x, y, 0.0, 0.0, 0.0, p, 0))
    return av_path, fs_path


def _case(tmp_path, perturb_av=0.0, perturb_fs=0.0, zero_av=False, shuffle=False, short=False):
    rng = np.random.default_rng(1)
    av = 1e-3 + rng.random(40) * 1e-2
    fs = [(x, y, 0.03 + 0.01 * rng.random()) for y in range(5) for x in range(6)]
    ref = _write(tmp_path, "ref", av, fs)
    av2 = av * (1 + perturb_av)
    if zero_av:
        av2[7] = 0.0
    if short:
        av2 = av2[:-1]
    fs2 = [(x, y, p * (1 + perturb_fs)) for (x, y, p) in fs]
    if shuffle:
        fs2[3], fs2[4] = fs2[4], fs2[3]
    sim = _write(tmp_path, "sim", av2, fs2)
    return ref, sim


CASES = {
    "identical": ({}, 0),
    "within_tolerance": ({"perturb_av": 0.004, "perturb_fs": -0.009}, 0),
    "av_vels_off": ({"perturb_av": 0.02}, 1),
    "final_state_off": ({"perturb_fs": 0.02}, 1),
    "zero_entry_is_not_finite": ({"zero_av": True}, 1),
    "coordinates_out_of_order": ({"shuffle": True}, 1),
    "different_step_count": ({"short": True}, 1),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_check_tool_verdicts(tmp_path, name):
    kwargs, expected = CASES[name]
    ref, sim = _case(tmp_path, **kwargs)
    out = io.StringIO()
    assert check_tool.check(ref[0], ref[1], sim[0], sim[1], out=out) == expected


@pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "check", "check.py")), reason="reference not mounted")
@pytest.mark.parametrize("name", sorted(CASES))
def test_same_output_as_the_reference_checker(tmp_path, name):
    kwargs, expected = CASES[name]
    ref, sim = _case(tmp_path, **kwargs)
    argv = [f"--ref-av-vels-file={ref[0]}", f"--ref-final-state-file={ref[1]}",
            f"--av-vels-file={sim[0]}", f"--final-state-file={sim[1]}"]
    theirs = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_check.py"), "--"] + argv,
                            capture_output=True, text=True)
    ours = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check.py")] + argv,
                          capture_output=True, text=True)
    assert theirs.returncode == ours.returncode == expected
    if name != "zero_entry_is_not_finite":          # numpy prints inf/nan differently there
        assert theirs.stdout == ours.stdout


def test_gzip_references_are_accepted(tmp_path):
    ref_av = os.path.join(GOLDEN, "128x128.av_vels.dat.gz")
    ref_fs = os.path.join(GOLDEN, "128x128.final_state.dat.gz")
    out = io.StringIO()
    assert check_tool.check(ref_av, ref_fs, ref_av, ref_fs, out=out) == 0
    assert "Both tests passed!" in out.getvalue()
