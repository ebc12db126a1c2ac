"""b200-d2q9-bgk: host-side Python mirror of the reference's timestep-loop interface.

The product is the C-ABI library (include/lbm_b200.h, csrc/) and the C host program
(host/lbm_cli.c).  This package is the thin ctypes binding used by tests/ and bench.py.
"""
from . import decks, parity  # noqa: F401
from ._lib import EXE_PATH, LIB_PATH, SIGNATURES, library  # noqa: F401
from .solver import (LBMError, Simulation, decompose, device_count, free_cells_inv,  # noqa: F401
                     pack_obstacle_bits, selftest)
