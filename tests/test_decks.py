"""Deck parsing / writing mirrors the reference's file formats and die() messages."""
import os

import numpy as np
import pytest

from conftest import DECKS, FREE_CELLS, deck_paths, load_deck


@pytest.mark.parametrize("name", DECKS)
def test_shipped_decks_parse(pkg, name):
    p, obstacles, free = load_deck(pkg, name)
    nx, ny = (int(v) for v in name.split("x"))
    assert (p.nx, p.ny) == (nx, ny)
    assert p.reynolds_dim == 10 and p.density == float(np.float32(0.1)) and p.omega == float(np.float32(1.85))
    assert free == FREE_CELLS[name]                       # duplicates in the files are counted once
    assert obstacles.shape == (ny, nx) and set(np.unique(obstacles)) <= {0, 1}
    assert pkg.free_cells_inv(obstacles) == pkg.decks.free_cells_inv(free)


def test_param_errors_use_the_reference_messages(pkg, tmp_path):
    d = pkg.decks
    with pytest.raises(d.DeckError, match="could not open input parameter file"):
        d.read_params(str(tmp_path / "missing.params"))
    f = tmp_path / "short.params"
    f.write_text("128\n128\n100\n")
    with pytest.raises(d.DeckError, match="could not read param file: reynolds_dim"):
        d.read_params(str(f))
    f.write_text("128\n128\n100\n10\n0.1\nabc\n1.85\n")
    with pytest.raises(d.DeckError, match="could not read param file: accel"):
        d.read_params(str(f))


def test_obstacle_errors_use_the_reference_messages(pkg, tmp_path):
    d = pkg.decks
    f = tmp_path / "o.dat"
    with pytest.raises(d.DeckError, match="could not open input obstacles file"):
        d.read_obstacles(str(tmp_path / "nope.dat"), 8, 8)
    f.write_text("1 1\n")
    with pytest.raises(d.DeckError, match="expected 3 values per line"):
        d.read_obstacles(str(f), 8, 8)
    f.write_text("8 1 1\n")
    with pytest.raises(d.DeckError, match="x-coord out of range"):
        d.read_obstacles(str(f), 8, 8)
    f.write_text("1 -1 1\n")
    with pytest.raises(d.DeckError, match="y-coord out of range"):
        d.read_obstacles(str(f), 8, 8)
    f.write_text("1 1 2\n")
    with pytest.raises(d.DeckError, match="blocked value should be 1"):
        d.read_obstacles(str(f), 8, 8)
    f.write_text("")                                       # an empty obstacle file is legal
    ob, free = d.read_obstacles(str(f), 8, 8)
    assert free == 64 and ob.sum() == 0
    f.write_text("1 1 1\n1 1 1\n2 1 1\n")                  # duplicates count once
    ob, free = d.read_obstacles(str(f), 8, 8)
    assert free == 62


def test_output_formats_round_trip(pkg, tmp_path):
    d = pkg.decks
    rng = np.random.default_rng(0)
    av = rng.random(50).astype(np.float32) * 1e-2
    d.write_av_vels(str(tmp_path / "av.dat"), av)
    lines = open(tmp_path / "av.dat").read().splitlines()
    assert lines[3] == "3:\t%.12E" % float(av[3])
    assert np.array_equal(d.read_av_vels(str(tmp_path / "av.dat")).astype(np.float32), av)
    ob = np.zeros((4, 8), np.int32)
    ob[0, :] = 1
    fields = [rng.random((4, 8)).astype(np.float32) for _ in range(4)]
    d.write_final_state(str(tmp_path / "fs.dat"), *fields, ob)
    text = open(tmp_path / "fs.dat").read().splitlines()
    assert len(text) == 32
    assert text[9] == "1 1 %.12E %.12E %.12E %.12E 0" % tuple(float(f[1, 1]) for f in fields)   # y-major, x-minor
    fs = d.read_final_state(str(tmp_path / "fs.dat"))
    assert np.array_equal(fs[:, 5].astype(np.float32).reshape(4, 8), fields[3])


def test_synthetic_channel_deck(pkg, tmp_path):
    pfile, ofile = pkg.decks.write_channel_deck(str(tmp_path), 64, 32, 10)
    p = pkg.decks.read_params(pfile)
    ob, free = pkg.decks.read_obstacles(ofile, p.nx, p.ny)
    assert (p.nx, p.ny, p.max_iters) == (64, 32, 10)
    assert free == 64 * 30 and np.array_equal(ob, pkg.decks.channel_obstacles(64, 32))


def test_cylinder_array_deck_round_trips(pkg, tmp_path):
    ob = pkg.decks.cylinder_array_obstacles(256, 192)
    assert ob[0].all() and ob[-1].all() and not ob[-2].any()
    assert 0.05 < ob.mean() < 0.3
    assert np.array_equal(ob[:, :128], ob[:, 128:])               # periodic in x with period 2*pitch
    pfile, ofile = pkg.decks.write_obstacle_deck(str(tmp_path), "cyl", ob, 7)
    p = pkg.decks.read_params(pfile)
    back, free = pkg.decks.read_obstacles(ofile, p.nx, p.ny)
    assert np.array_equal(back, ob) and free == ob.size - int(ob.sum()) and p.max_iters == 7
