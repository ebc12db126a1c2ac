"""`Simulation`: the Python mirror of the reference's timestep-loop interface.

Each method is a thin call through the C-ABI (include/lbm_b200.h); numpy arrays are host
buffers in the reference's layouts (AoS cells, int obstacles, float av_vels)."""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import c_float_p, c_int_p, handle_t, library


class LBMError(RuntimeError):
    """A non-zero return of the C-ABI; the text is lbm_b200_last_error()."""


def _check(rc: int) -> None:
    if rc != 0:
        raise LBMError(f"[{rc}] " + library().lbm_b200_last_error().decode(errors="replace"))


def _fp(a: np.ndarray):
    return a.ctypes.data_as(c_float_p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(c_int_p)


OBSTACLE_FORMATS = {"int32": 0, "uint8": 1, "bits": 2}      # LBM_B200_OBST_* of include/lbm_b200.h


def _obstacle_buffer(obstacles, rows: int, nx: int, fmt: str) -> np.ndarray:
    """The caller's obstacle rows as the contiguous host array the C-ABI expects for `fmt`."""
    if fmt not in OBSTACLE_FORMATS:
        raise ValueError(f"obstacles_format must be one of {sorted(OBSTACLE_FORMATS)}")
    if fmt == "bits":
        ob = np.ascontiguousarray(obstacles)
        if ob.dtype not in (np.uint32, np.int32) or ob.shape != (rows, (nx + 31) // 32):
            raise ValueError(f"bit-packed obstacles must be (u)int32 of shape ({rows}, {(nx + 31) // 32}), got {ob.dtype} {ob.shape}")
        return ob
    ob = np.ascontiguousarray(obstacles, np.int32 if fmt == "int32" else np.uint8)
    if ob.shape != (rows, nx):
        raise ValueError(f"obstacles must have shape ({rows}, {nx}), got {ob.shape}")
    return ob


def pack_obstacle_bits(obstacles: np.ndarray) -> np.ndarray:
    """[rows, nx] (non-zero = blocked) -> the "bits" format: uint32 [rows, ceil(nx / 32)], cell x = bit x & 31 of word x >> 5."""
    rows, nx = obstacles.shape
    words = (nx + 31) // 32
    padded = np.zeros((rows, words * 32), np.uint8)
    padded[:, :nx] = np.asarray(obstacles) != 0
    return np.ascontiguousarray(np.packbits(padded, axis=1, bitorder="little")).view(np.uint32).reshape(rows, words)


def decompose(ny: int, n_slabs: int):
    """Rows and first row of every slab (reference d2q9-bgk.c:834-862). Host only."""
    rows = np.zeros(n_slabs, np.int32)
    first = np.zeros(n_slabs, np.int32)
    _check(library().lbm_b200_decompose(ny, n_slabs, _ip(rows), _ip(first)))
    return rows, first


def free_cells_inv(obstacles: np.ndarray) -> np.float32:
    ob = np.ascontiguousarray(obstacles, np.int32)
    return np.float32(library().lbm_b200_free_cells_inv(_ip(ob), ob.size))


def selftest(device: int = 0):
    """(rcp mismatches, sqrt mismatches, packed-arithmetic mismatches) of lbm_b200_selftest; all must be 0."""
    counts = (ctypes.c_ulonglong * 3)()
    _check(library().lbm_b200_selftest(device, counts))
    return tuple(int(c) for c in counts)


def device_count() -> int:
    return int(library().lbm_b200_device_count())


class Simulation:
    """A D2Q9-BGK run on one or more B200s.

    Whole-domain form (one process):  Simulation(nx, ny, density, accel, omega, obstacles,
    n_slabs=1, devices=None, device=None).  inplace=True keeps ONE population buffer per slab and streams in
    place (lbm_b200_create_inplace).
    Slab form (one rank per process):  Simulation.slab(...), then export_ipc()/connect_ipc().
    """

    def __init__(self, nx, ny, density, accel, omega, obstacles, n_slabs: int = 1, devices=None, device=None,
                 inplace: bool = False, obstacles_format: str = "int32"):
        self._h = handle_t()
        self._lib = library()
        ob = _obstacle_buffer(obstacles, ny, nx, obstacles_format)
        if device is not None and devices is None:
            devices = [device] * n_slabs
        dev = None
        if devices is not None:
            dev = np.ascontiguousarray(devices, np.int32)
            if dev.shape != (n_slabs,):
                raise ValueError("devices must list one device per slab")
        if obstacles_format == "int32":                      # the reference's layout: the plain entry points
            create = self._lib.lbm_b200_create_inplace if inplace else self._lib.lbm_b200_create
            _check(create(ctypes.byref(self._h), nx, ny, density, accel, omega, _ip(ob), n_slabs,
                          _ip(dev) if dev is not None else None))
        else:
            _check(self._lib.lbm_b200_create_ex(ctypes.byref(self._h), nx, ny, density, accel, omega, ob.ctypes.data,
                                                OBSTACLE_FORMATS[obstacles_format], n_slabs,
                                                _ip(dev) if dev is not None else None, 1 if inplace else 0))
        self.nx, self.ny = nx, ny
        self._set_shape()

    @classmethod
    def slab(cls, nx, ny_global, first_row, rows, rank, n_ranks, density, accel, omega, free_cells_inv,
             obstacles_slab, device, inplace: bool = False, obstacles_format: str = "int32"):
        self = cls.__new__(cls)
        self._h = handle_t()
        self._lib = library()
        ob = _obstacle_buffer(obstacles_slab, rows, nx, obstacles_format)
        if obstacles_format == "int32":
            create = self._lib.lbm_b200_create_slab_inplace if inplace else self._lib.lbm_b200_create_slab
            _check(create(ctypes.byref(self._h), nx, ny_global, first_row, rows, rank, n_ranks,
                          density, accel, omega, free_cells_inv, _ip(ob), device))
        else:
            _check(self._lib.lbm_b200_create_slab_ex(ctypes.byref(self._h), nx, ny_global, first_row, rows, rank, n_ranks,
                                                     density, accel, omega, free_cells_inv, ob.ctypes.data,
                                                     OBSTACLE_FORMATS[obstacles_format], device, 1 if inplace else 0))
        self.nx, self.ny = nx, ny_global
        self._set_shape()
        return self

    def _set_shape(self):
        nx, rows, first = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _check(self._lib.lbm_b200_shape(self._h, ctypes.byref(nx), ctypes.byref(rows), ctypes.byref(first)))
        self.rows, self.first_row = rows.value, first.value

    # ---- ring wiring for slab handles --------------------------------------------------
    def export_ipc(self) -> bytes:
        blob = ctypes.create_string_buffer(self._lib.lbm_b200_ipc_blob_bytes())
        _check(self._lib.lbm_b200_ipc_export(self._h, blob))
        return blob.raw

    def connect_ipc(self, south_blob: bytes, north_blob: bytes) -> None:
        _check(self._lib.lbm_b200_ipc_connect(self._h, south_blob, north_blob))

    # ---- the hot path -----------------------------------------------------------------
    def run(self, iters: int) -> np.ndarray:
        """`iters` timesteps; returns av_vels float32[iters] (reference d2q9-bgk.c:315-396)."""
        av = np.zeros(max(iters, 1), np.float32)
        _check(self._lib.lbm_b200_run(self._h, iters, _fp(av)))
        return av[:iters]

    def enqueue(self, iters: int) -> None:
        _check(self._lib.lbm_b200_enqueue(self._h, iters))

    def sync(self) -> None:
        _check(self._lib.lbm_b200_sync(self._h))

    def elapsed_ms(self) -> float:
        ms = ctypes.c_float()
        _check(self._lib.lbm_b200_elapsed_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def fetch_av_vels(self, iters: int, out: np.ndarray | None = None) -> np.ndarray:
        av = out if out is not None else np.zeros(max(iters, 1), np.float32)
        _check(self._lib.lbm_b200_fetch_av_vels(self._h, iters, _fp(av)))
        return av[:iters]

    # ---- state ------------------------------------------------------------------------
    def get_cells(self, out: np.ndarray | None = None) -> np.ndarray:
        cells = out if out is not None else np.empty((self.rows, self.nx, 9), np.float32)
        _check(self._lib.lbm_b200_get_cells(self._h, _fp(cells)))
        return cells

    def set_cells(self, cells: np.ndarray) -> None:
        c = np.ascontiguousarray(cells, np.float32)
        if c.shape != (self.rows, self.nx, 9):
            raise ValueError(f"cells must have shape ({self.rows}, {self.nx}, 9)")
        _check(self._lib.lbm_b200_set_cells(self._h, _fp(c)))

    def final_state(self, out: np.ndarray | None = None):
        """(u_x, u_y, |u|, pressure), each float32[rows, nx] (reference d2q9-bgk.c:1076-1111)."""
        f = out if out is not None else np.empty((4, self.rows, self.nx), np.float32)
        _check(self._lib.lbm_b200_get_final_state(self._h, _fp(f[0]), _fp(f[1]), _fp(f[2]), _fp(f[3])))
        return f[0], f[1], f[2], f[3]

    def set_option(self, key: str, value: int) -> None:
        _check(self._lib.lbm_b200_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        v = ctypes.c_long()
        _check(self._lib.lbm_b200_get_option(self._h, key.encode(), ctypes.byref(v)))
        return v.value

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.lbm_b200_destroy(self._h)
            self._h = handle_t()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
