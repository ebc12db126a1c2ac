"""ctypes binding of oracle/_build/liblbm_oracle.so -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs (see the
header of oracle/lbm_oracle.c).  The product package never imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liblbm_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")

_lib = None


def build() -> None:
    """Compiles the restatement (and, where /root/reference exists, the reference itself)."""
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        L.lbm_oracle_decompose.argtypes = [ctypes.c_int, ctypes.c_int, _i32p, _i32p]
        L.lbm_oracle_decompose.restype = None
        L.lbm_oracle_init.argtypes = [_f32p, ctypes.c_long, ctypes.c_float]
        L.lbm_oracle_init.restype = None
        L.lbm_oracle_accelerate_row.argtypes = [_f32p, _i32p, ctypes.c_int, ctypes.c_float, ctypes.c_float]
        L.lbm_oracle_accelerate_row.restype = None
        L.lbm_oracle_timestep_rows.argtypes = [_f32p, _f32p, _i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_float, _f64p]
        L.lbm_oracle_timestep_rows.restype = ctypes.c_float
        L.lbm_oracle_run.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                     ctypes.c_float, ctypes.c_float, _i32p, _f32p, _f32p, ctypes.c_void_p]
        L.lbm_oracle_run.restype = ctypes.c_int
        L.lbm_oracle_av_velocity.argtypes = [_f32p, _i32p, ctypes.c_int, ctypes.c_int, ctypes.c_float]
        L.lbm_oracle_av_velocity.restype = ctypes.c_float
        L.lbm_oracle_reynolds.argtypes = [_f32p, _i32p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                          ctypes.c_float, ctypes.c_int]
        L.lbm_oracle_reynolds.restype = ctypes.c_float
        L.lbm_oracle_final_state.argtypes = [_f32p, _i32p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                             _f32p, _f32p, _f32p, _f32p]
        L.lbm_oracle_final_state.restype = None
        L.lbm_oracle_total_density.argtypes = [_f32p, ctypes.c_long]
        L.lbm_oracle_total_density.restype = ctypes.c_double
        _lib = L
    return _lib


def decompose(ny: int, size: int):
    ny_local = np.zeros(size, np.int32)
    displs = np.zeros(size, np.int32)
    lib().lbm_oracle_decompose(ny, size, ny_local, displs)
    return ny_local, displs


def init_cells(nx: int, ny: int, density: float) -> np.ndarray:
    cells = np.empty((ny, nx, 9), np.float32)
    lib().lbm_oracle_init(cells, nx * ny, density)
    return cells


def run(cells: np.ndarray, obstacles: np.ndarray, iters: int, density: float, accel: float, omega: float,
        free_cells_inv: float, exact: bool = False):
    """Advances `cells` ([ny, nx, 9] float32, in place) by `iters` steps; returns av_vels float32[iters]
    (the reference's sequential fp32 accumulation).  With exact=True also returns float64[iters]: the same
    per-cell terms summed in fp64."""
    ny, nx = obstacles.shape
    assert cells.shape == (ny, nx, 9) and cells.dtype == np.float32 and cells.flags.c_contiguous
    av = np.zeros(max(iters, 1), np.float32)
    av_exact = np.zeros(max(iters, 1), np.float64) if exact else None
    rc = lib().lbm_oracle_run(nx, ny, iters, density, accel, omega, free_cells_inv,
                              np.ascontiguousarray(obstacles, np.int32), cells, av,
                              av_exact.ctypes.data if exact else None)
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    return (av[:iters], av_exact[:iters]) if exact else av[:iters]


def slab_timestep(cells: np.ndarray, tmp_cells: np.ndarray, obstacles: np.ndarray, start: int, end: int,
                  omega: float) -> np.float32:
    """reference timestep(start, end) on a halo'd slab: cells/tmp [rows+2, nx, 9], obstacles [rows+2, nx]."""
    nx = obstacles.shape[1]
    scratch = np.empty((end - start) * nx, np.float64)
    return np.float32(lib().lbm_oracle_timestep_rows(cells, tmp_cells, obstacles, nx, start, end, omega, scratch))


def accelerate_row(row_cells: np.ndarray, row_obstacles: np.ndarray, density: float, accel: float) -> None:
    lib().lbm_oracle_accelerate_row(row_cells, row_obstacles, row_obstacles.shape[0], density, accel)


def final_state(cells: np.ndarray, obstacles: np.ndarray, density: float):
    ny, nx = obstacles.shape
    out = [np.empty((ny, nx), np.float32) for _ in range(4)]
    lib().lbm_oracle_final_state(cells, np.ascontiguousarray(obstacles, np.int32), nx, ny, density, *out)
    return tuple(out)


def reynolds(cells, obstacles, free_cells_inv, omega, reynolds_dim) -> np.float32:
    ny, nx = obstacles.shape
    return np.float32(lib().lbm_oracle_reynolds(cells, np.ascontiguousarray(obstacles, np.int32), nx, ny,
                                                free_cells_inv, omega, reynolds_dim))


def av_velocity(cells, obstacles, free_cells_inv) -> np.float32:
    ny, nx = obstacles.shape
    return np.float32(lib().lbm_oracle_av_velocity(cells, np.ascontiguousarray(obstacles, np.int32), nx, ny,
                                                   free_cells_inv))


def total_density(cells) -> float:
    return float(lib().lbm_oracle_total_density(np.ascontiguousarray(cells, np.float32), cells.size // 9))


def ref_binary(kind: str = "strict") -> str | None:
    """Path of a compiled reference executable in oracle/_ref (None if it was never built)."""
    path = os.path.join(REF_DIR, f"d2q9-bgk.{kind}")
    return path if os.path.exists(path) else None
