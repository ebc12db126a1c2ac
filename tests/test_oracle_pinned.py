"""The oracle (oracle/lbm_oracle.c) is pinned: against the reference's golden vectors
(check/*.dat, committed gzip'd under tests/golden/), against outputs of the unmodified
reference source built strict-IEEE (tests/golden/ref_strict.*, made by make_golden.py), and --
where oracle/_ref was built -- against that binary run live."""
import hashlib
import io
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, deck_paths, load_deck

import check as check_tool

REF_STRICT = json.load(open(os.path.join(GOLDEN, "ref_strict.json")))["decks"] \
    if os.path.exists(os.path.join(GOLDEN, "ref_strict.json")) else {}


def run_oracle(pkg, oracle, name, iters=None):
    p, obstacles, free = load_deck(pkg, name)
    iters = p.max_iters if iters is None else iters
    inv = pkg.decks.free_cells_inv(free)
    cells = oracle.init_cells(p.nx, p.ny, p.density)
    av = oracle.run(cells, obstacles, iters, p.density, p.accel, p.omega, inv)
    return p, obstacles, inv, cells, av


@pytest.mark.parametrize("name", ["128x128", "128x256"])
def test_oracle_full_run_against_goldens_and_strict_reference(pkg, oracle, tmp_path, name):
    p, obstacles, inv, cells, av = run_oracle(pkg, oracle, name)
    ux, uy, u, pr = oracle.final_state(cells, obstacles, p.density)
    av_path, fs_path = str(tmp_path / "av_vels.dat"), str(tmp_path / "final_state.dat")
    pkg.decks.write_av_vels(av_path, av)
    pkg.decks.write_final_state(fs_path, ux, uy, u, pr, obstacles)
    # (1) the reference's own acceptance test against its fp64 goldens, tolerance 1 %
    out = io.StringIO()
    rc = check_tool.check(os.path.join(GOLDEN, f"{name}.av_vels.dat.gz"), os.path.join(GOLDEN, f"{name}.final_state.dat.gz"),
                          av_path, fs_path, tolerance=1.0, out=out)
    assert rc == 0, out.getvalue()
    # (2) bit-identical to the strict build of the unmodified reference source
    ref = REF_STRICT[name]
    assert hashlib.sha256(open(fs_path, "rb").read()).hexdigest() == ref["final_state_sha256"]
    assert np.array_equal(av, np.load(os.path.join(GOLDEN, f"ref_strict.{name}.av_vels.npy")))
    assert "%.12E" % float(oracle.reynolds(cells, obstacles, inv, p.omega, p.reynolds_dim)) == ref["reynolds"]


@pytest.mark.parametrize("name,iters", [("256x256", 3000), ("1024x1024", 150)])
def test_oracle_prefix_against_goldens(pkg, oracle, name, iters):
    """Larger decks: the first `iters` av_vels entries against the golden and the strict reference
    (the worst golden mismatch of an fp32 run is at step 0)."""
    p, obstacles, inv, cells, av = run_oracle(pkg, oracle, name, iters)
    golden = pkg.decks.read_av_vels(os.path.join(GOLDEN, f"{name}.av_vels.dat.gz"))[:iters]
    pcnt = 100.0 * (golden - av) / av
    assert np.all(np.isfinite(pcnt)) and np.max(np.abs(pcnt)) <= 1.0
    assert np.array_equal(av, np.load(os.path.join(GOLDEN, f"ref_strict.{name}.av_vels.npy"))[:iters])


@pytest.mark.parametrize("name,iters", [("128x128", 257), ("128x256", 100), ("256x256", 60), ("1024x1024", 12)])
def test_oracle_bitwise_against_live_reference_binary(pkg, oracle, tmp_path, name, iters):
    ref_bin = oracle.ref_binary("strict")
    if ref_bin is None:
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    p, obstacles, inv, cells, av = run_oracle(pkg, oracle, name, iters)
    p.max_iters = iters
    pfile = tmp_path / "in.params"
    pfile.write_text(p.as_text())
    out = subprocess.run([ref_bin, str(pfile), deck_paths(name)[1]], cwd=tmp_path, check=True,
                         capture_output=True, text=True).stdout
    ux, uy, u, pr = oracle.final_state(cells, obstacles, p.density)
    pkg.decks.write_av_vels(str(tmp_path / "o_av.dat"), av)
    pkg.decks.write_final_state(str(tmp_path / "o_fs.dat"), ux, uy, u, pr, obstacles)
    assert open(tmp_path / "o_fs.dat", "rb").read() == open(tmp_path / "final_state.dat", "rb").read()
    assert open(tmp_path / "o_av.dat", "rb").read() == open(tmp_path / "av_vels.dat", "rb").read()
    reynolds = [l.split()[-1] for l in out.splitlines() if l.startswith("Reynolds")][0]
    assert reynolds == "%.12E" % float(oracle.reynolds(cells, obstacles, inv, p.omega, p.reynolds_dim))


def test_reference_multi_rank_build_matches_single_rank(oracle, tmp_path):
    """The fork+shm mpi.h shim: 4 'MPI' ranks give the same cells as 1 rank (decomposition invariant)."""
    ref_bin = oracle.ref_binary("strict")
    if ref_bin is None:
        pytest.skip("oracle/_ref not built")
    pfile = tmp_path / "in.params"
    pfile.write_text("128\n128\n200\n10\n0.1\n0.005\n1.85\n")
    outs = []
    for ranks in (1, 4):
        d = tmp_path / f"r{ranks}"
        d.mkdir()
        subprocess.run([ref_bin, str(pfile), deck_paths("128x128")[1]], cwd=d, check=True, capture_output=True,
                       env={**os.environ, "MPI_SHIM_RANKS": str(ranks)})
        outs.append(d)
    assert open(outs[0] / "final_state.dat", "rb").read() == open(outs[1] / "final_state.dat", "rb").read()
    a, b = (np.loadtxt(o / "av_vels.dat", usecols=[1]) for o in outs)
    assert np.max(np.abs(a - b) / a) < 1e-5


def test_total_density_is_conserved(pkg, oracle):
    """The reference's disabled invariant (d2q9-bgk.c:132-133, 1011-1032)."""
    p, obstacles, inv, cells, av = run_oracle(pkg, oracle, "128x256", 0)
    before = oracle.total_density(cells)
    oracle.run(cells, obstacles, 500, p.density, p.accel, p.omega, inv)
    assert abs(oracle.total_density(cells) - before) / before < 1e-5


def test_timestep_of_zero_iterations_is_the_identity(pkg, oracle):
    p, obstacles, inv, cells, av = run_oracle(pkg, oracle, "128x128", 0)
    assert av.size == 0 and np.array_equal(cells, oracle.init_cells(p.nx, p.ny, p.density))
