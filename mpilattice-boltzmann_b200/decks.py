"""Input decks and output files in the reference's formats (Python mirror of the C host).

The product CLI (host/lbm_cli.c) does its own parsing and writing in C; this module mirrors
the same contract for the Python-side tests and bench:

* ``.params``   -- 7 values ``nx ny maxIters reynolds_dim density accel omega``
                   (reference d2q9-bgk.c:781-800)
* obstacles     -- lines ``x y 1`` with range checks, duplicates counted once
                   (reference d2q9-bgk.c:933-950)
* ``av_vels.dat``     -- ``"%d:\\t%.12E\\n"``                         (d2q9-bgk.c:1136)
* ``final_state.dat`` -- ``"%d %d %.12E %.12E %.12E %.12E %d\\n"``   (d2q9-bgk.c:1115),
                         y-major / x-minor

Errors raise :class:`DeckError` carrying the reference's ``die()`` message text.
"""
from __future__ import annotations

import dataclasses
import io
import os

import numpy as np


class DeckError(ValueError):
    """A malformed deck; the message is the reference's die() text (d2q9-bgk.c:776-942)."""


@dataclasses.dataclass
class Params:
    nx: int
    ny: int
    max_iters: int
    reynolds_dim: int
    density: float
    accel: float
    omega: float

    def as_text(self) -> str:
        return (f"{self.nx}\n{self.ny}\n{self.max_iters}\n{self.reynolds_dim}\n"
                f"{self.density!r}\n{self.accel!r}\n{self.omega!r}\n")


_PARAM_FIELDS = (("nx", int), ("ny", int), ("maxIters", int), ("reynolds_dim", int),
                 ("density", float), ("accel", float), ("omega", float))


def read_params(path: str) -> Params:
    """d2q9-bgk.c:772-803."""
    try:
        with open(path, "r") as fh:
            tokens = fh.read().split()
    except OSError:
        raise DeckError(f"could not open input parameter file: {path}")
    values = []
    for i, (name, kind) in enumerate(_PARAM_FIELDS):
        try:
            values.append(kind(tokens[i]))
        except (IndexError, ValueError):
            raise DeckError(f"could not read param file: {name}")
    p = Params(*values)
    # the reference stores the three reals as C floats
    p.density = float(np.float32(p.density))
    p.accel = float(np.float32(p.accel))
    p.omega = float(np.float32(p.omega))
    return p


def read_obstacles(path: str, nx: int, ny: int):
    """d2q9-bgk.c:905-953.  Returns (obstacles int32[ny, nx], number of free cells)."""
    try:
        with open(path, "r") as fh:
            text = fh.read()
    except OSError:
        raise DeckError(f"could not open input obstacles file: {path}")
    obstacles = np.zeros((ny, nx), dtype=np.int32)
    tokens = text.split()
    if len(tokens) % 3 != 0:
        raise DeckError("expected 3 values per line in obstacle file")
    try:
        table = np.array(tokens, dtype=np.int64).reshape(-1, 3)
    except ValueError:
        raise DeckError("expected 3 values per line in obstacle file")
    if table.size:
        xs, ys, blocked = table[:, 0], table[:, 1], table[:, 2]
        # the reference reports the first offending line; checks are ordered x, y, blocked
        for row in range(table.shape[0]) if _any_bad(xs, ys, blocked, nx, ny) else ():
            if xs[row] < 0 or xs[row] > nx - 1:
                raise DeckError("obstacle x-coord out of range")
            if ys[row] < 0 or ys[row] > ny - 1:
                raise DeckError("obstacle y-coord out of range")
            if blocked[row] != 1:
                raise DeckError("obstacle blocked value should be 1")
        obstacles[ys, xs] = 1
    free_cells = nx * ny - int(obstacles.sum())       # duplicates counted once (945-946)
    return obstacles, free_cells


def _any_bad(xs, ys, blocked, nx, ny) -> bool:
    return bool(((xs < 0) | (xs > nx - 1) | (ys < 0) | (ys > ny - 1) | (blocked != 1)).any())


def free_cells_inv(free_cells: int) -> np.float32:
    """d2q9-bgk.c:950 -- ``1.0f/numOfFreeCells`` in float."""
    return np.float32(1.0) / np.float32(free_cells)


def write_av_vels(path: str, av_vels) -> None:
    """d2q9-bgk.c:1134-1137."""
    av = np.asarray(av_vels, dtype=np.float32).astype(np.float64)
    with open(path, "w") as fh:
        fh.write("".join("%d:\t%.12E\n" % (i, v) for i, v in enumerate(av)))


def write_final_state(path: str, u_x, u_y, u, pressure, obstacles) -> None:
    """d2q9-bgk.c:1071-1118 -- one line per cell, y outer, x inner."""
    ny, nx = obstacles.shape
    cols = [np.asarray(a, dtype=np.float32).astype(np.float64).reshape(ny, nx)
            for a in (u_x, u_y, u, pressure)]
    out = io.StringIO()
    for y in range(ny):
        ux, uy, uu, pp = (c[y] for c in cols)
        ob = obstacles[y]
        out.write("".join("%d %d %.12E %.12E %.12E %.12E %d\n" % (x, y, ux[x], uy[x], uu[x], pp[x], ob[x])
                          for x in range(nx)))
    with open(path, "w") as fh:
        fh.write(out.getvalue())


def read_av_vels(path: str) -> np.ndarray:
    return np.loadtxt(_open_maybe_gz(path), usecols=[1], ndmin=1)


def read_final_state(path: str) -> np.ndarray:
    """All seven columns as float64 [ncells, 7]."""
    return np.loadtxt(_open_maybe_gz(path), ndmin=2)


def _open_maybe_gz(path: str):
    if path.endswith(".gz"):
        import gzip
        return gzip.open(path, "rt")
    return open(path, "r")


def channel_obstacles(nx: int, ny: int) -> np.ndarray:
    """The synthetic deck of BASELINE.json configs[4] / SURVEY 8(d): walls on rows 0 and ny-1,
    periodic in x, flow driven along +x on row ny-2."""
    obstacles = np.zeros((ny, nx), dtype=np.int32)
    obstacles[0, :] = 1
    obstacles[ny - 1, :] = 1
    return obstacles


def write_channel_deck(directory: str, nx: int, ny: int, max_iters: int, *, reynolds_dim=10,
                       density=0.1, accel=0.005, omega=1.85):
    """Writes ``input_<nx>x<ny>.params`` and ``obstacles_<nx>x<ny>.dat`` for the synthetic
    channel; returns the two paths.  Deterministic (no RNG)."""
    os.makedirs(directory, exist_ok=True)
    pfile = os.path.join(directory, f"input_{nx}x{ny}.params")
    ofile = os.path.join(directory, f"obstacles_{nx}x{ny}.dat")
    with open(pfile, "w") as fh:
        fh.write(f"{nx}\n{ny}\n{max_iters}\n{reynolds_dim}\n{density}\n{accel}\n{omega}\n")
    with open(ofile, "w") as fh:
        for y in (0, ny - 1):
            fh.write("".join(f"{x} {y} 1\n" for x in range(nx)))
    return pfile, ofile


def cylinder_array_obstacles(nx: int, ny: int, pitch: int = 64, radius: int = 12) -> np.ndarray:
    """A richer synthetic deck: the channel walls plus a staggered array of solid discs of
    `radius` every `pitch` cells.  Periodic in x with period 2*pitch when nx is a multiple of it."""
    obstacles = channel_obstacles(nx, ny)
    yy, xx = np.mgrid[0:ny, 0:nx]
    row = yy // pitch
    cx = (xx + (row % 2) * (pitch // 2)) % pitch - pitch // 2
    cy = yy % pitch - pitch // 2
    inside = (cx * cx + cy * cy <= radius * radius) & (yy > pitch // 2) & (yy < ny - pitch // 2)
    obstacles[inside] = 1
    obstacles[ny - 2, :] = 0            # keep the driven row open
    return obstacles


def write_obstacle_deck(directory: str, name: str, obstacles: np.ndarray, max_iters: int, *, reynolds_dim=10,
                        density=0.1, accel=0.005, omega=1.85):
    """Writes ``input_<name>.params`` / ``obstacles_<name>.dat`` for an arbitrary obstacle map."""
    os.makedirs(directory, exist_ok=True)
    ny, nx = obstacles.shape
    pfile = os.path.join(directory, f"input_{name}.params")
    ofile = os.path.join(directory, f"obstacles_{name}.dat")
    with open(pfile, "w") as fh:
        fh.write(f"{nx}\n{ny}\n{max_iters}\n{reynolds_dim}\n{density}\n{accel}\n{omega}\n")
    ys, xs = np.nonzero(obstacles)
    with open(ofile, "w") as fh:
        fh.write("".join(f"{x} {y} 1\n" for x, y in zip(xs.tolist(), ys.tolist())))
    return pfile, ofile
