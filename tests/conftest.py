import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
DECK_DIR = os.path.join(GOLDEN, "decks")
DECKS = ["128x128", "128x256", "256x256", "1024x1024"]
# free cells per deck (duplicate obstacle lines counted once), SURVEY.md section 7
FREE_CELLS = {"128x128": 15876, "128x256": 32130, "256x256": 64516, "1024x1024": 1043462}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (ctypes binding of liblbm_b200.so); builds it if necessary."""
    import __graft_entry__ as entry
    p = entry.load_package()
    if not os.path.exists(p.LIB_PATH) or not os.path.exists(p.EXE_PATH):
        entry.build()
    p.library()
    return p


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement of the reference (oracle/lbm_oracle.c) -- the checker."""
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def decks(pkg):
    return pkg.decks


def deck_paths(name):
    return (os.path.join(DECK_DIR, f"input_{name}.params"), os.path.join(DECK_DIR, f"obstacles_{name}.dat"))


def load_deck(pkg, name):
    pfile, ofile = deck_paths(name)
    p = pkg.decks.read_params(pfile)
    obstacles, free = pkg.decks.read_obstacles(ofile, p.nx, p.ny)
    return p, obstacles, free


def random_obstacles(rng, ny, nx, fraction=0.05, walls=True):
    ob = (rng.random((ny, nx)) < fraction).astype(np.int32)
    if walls:
        ob[0, :] = 1
        ob[-1, :] = 1
    ob[ny // 2, :] = 0           # keep at least one open row
    return ob


def random_cells(rng, ny, nx, density=0.1):
    """A valid (positive) but non-uniform state around the reference's initial populations."""
    w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4, np.float64) * density
    cells = w[None, None, :] * (1.0 + 0.2 * (rng.random((ny, nx, 9)) - 0.5))
    return np.ascontiguousarray(cells, np.float32)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)
