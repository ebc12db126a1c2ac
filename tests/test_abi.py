"""The C-ABI library loads on a CPU-only box and exports exactly what include/lbm_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_b200_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/lbm_b200.h but not exported"
    assert sorted(pkg.SIGNATURES) == names, "the ctypes binding must cover the header exactly"


def test_no_torch_or_cxx_types_in_the_header():
    text = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    assert 'extern "C"' in text
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)       # declarations only, not the prose
    for forbidden in ("torch", "at::", "std::", "cudaStream_t", "template"):
        assert forbidden not in text


def test_abi_version(pkg):
    assert pkg.library().lbm_b200_abi_version() == 2          # r02: _ex constructors, selftest


def test_oracle_is_not_linked_into_the_product(pkg):
    """The product must not route through oracle/: neither the library nor the CLI references it."""
    for path in (pkg.LIB_PATH, pkg.EXE_PATH):
        blob = open(path, "rb").read()
        assert b"lbm_oracle_" not in blob and b"liblbm_oracle" not in blob
    src_dir = os.path.join(ROOT, "mpilattice-boltzmann_b200")
    for dirpath, _, files in os.walk(src_dir):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for needle in ("oracle_lib", "liblbm_oracle", "lbm_oracle_", "import oracle"):
                    assert needle not in text, f"{f} touches the oracle ({needle})"
                assert not re.search(r"#\s*include[^\n]*oracle", text), f"{f} includes an oracle header"


@pytest.mark.parametrize("ny,n", [(128, 1), (128, 2), (128, 8), (1024, 8), (130, 4), (131, 8), (16384, 8), (24, 8), (27, 8)])
def test_decompose_matches_reference_rule(pkg, oracle, ny, n):
    rows, first = pkg.decompose(ny, n)
    ref_rows, ref_first = oracle.decompose(ny, n)     # restates d2q9-bgk.c:834-862
    assert rows.tolist() == ref_rows.tolist()
    assert first.tolist() == ref_first.tolist()
    assert rows.sum() == ny and rows[-1] >= 3 or n == 1


def test_decompose_refuses_thin_slabs(pkg):
    with pytest.raises(pkg.LBMError, match="at least 3 rows"):
        pkg.decompose(16, 8)          # the reference would create 1- and 2-row slabs here


def test_free_cells_inv_counts_blocked_once(pkg):
    ob = np.zeros((8, 8), np.int32)
    ob[0, :] = 1
    ob[3, 3] = 7                       # any non-zero value is a blocked cell
    assert pkg.free_cells_inv(ob) == np.float32(1.0) / np.float32(64 - 9)


def test_compute_fails_loudly_without_a_gpu(pkg):
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.LBMError, match="no CUDA device"):
        pkg.Simulation(16, 16, 0.1, 0.005, 1.85, np.zeros((16, 16), np.int32))


def test_bad_arguments_are_rejected_before_any_device_work(pkg):
    with pytest.raises(pkg.LBMError, match="too small"):
        pkg.Simulation(2, 2, 0.1, 0.005, 1.85, np.zeros((2, 2), np.int32))
    with pytest.raises(ValueError):
        pkg.Simulation(16, 16, 0.1, 0.005, 1.85, np.zeros((8, 16), np.int32))
    with pytest.raises(pkg.LBMError, match="nx % 4"):      # in-place streaming has no one-cell-per-thread kernel
        pkg.Simulation(30, 9, 0.1, 0.005, 1.85, np.zeros((9, 30), np.int32), inplace=True)
    with pytest.raises(pkg.LBMError, match="omega"):
        pkg.Simulation(16, 16, 0.1, 0.005, 0.0, np.zeros((16, 16), np.int32))


def test_band_plan_for_k_steps_and_ring_slabs(pkg):
    """lbm_b200_plan_bands_ex: every row in exactly one band for K = 2..4; on a ring the first and the last band hold
    the rows one work item pushes (two for kernel 5, four for kernel 7); the automatic heights of the default kernel
    (K = 3) stay inside the measured optima (profiles/r02_fused2.md, section 8)."""
    lib = pkg.library()

    def plan(rows, nx, want, steps, ring, sms=148):
        bands, per = ctypes.c_int(), ctypes.c_int()
        assert lib.lbm_b200_plan_bands_ex(rows, nx, want, sms, steps, ring, ctypes.byref(bands), ctypes.byref(per)) == 0
        return bands.value, per.value

    for steps in (2, 3, 4):
        for ring in (0, 1):
            edge = (4 if steps >= 3 else 2) if ring else (1 if steps >= 3 else 2)
            for rows in list(range(6, 70)) + [127, 128, 129, 2048, 16384]:
                for want in (0, 1, 3, 4, 5, 8, 64, 1000):
                    bands, per = plan(rows, 1024, want, steps, ring)
                    last = rows - (bands - 1) * per
                    assert bands >= 1 and per >= 1 and 0 < last <= max(per, rows), (steps, ring, rows, want)
                    if bands > 1:
                        assert per >= edge and last >= edge, (steps, ring, rows, want, bands, per)
    assert plan(16384, 16384, 0, 3, 0)[1] in (96, 128, 160, 192) and plan(2048, 16384, 0, 3, 1)[1] in (24, 48)
    assert plan(4096, 4096, 0, 3, 0)[1] in (24, 32) and plan(2048, 2048, 0, 3, 0)[1] in (24, 32)
    bands, per = ctypes.c_int(), ctypes.c_int()
    assert lib.lbm_b200_plan_bands_ex(5, 1024, 0, 148, 3, 1, ctypes.byref(bands), ctypes.byref(per)) != 0   # a ring slab that thin
    assert lib.lbm_b200_plan_bands_ex(64, 1024, 0, 148, 5, 0, ctypes.byref(bands), ctypes.byref(per)) != 0


def test_band_plan_of_the_fused_kernel(pkg):
    """Every row belongs to exactly one band, all bands but the last are equally tall, the first and the last hold
    at least two rows (ring slabs push two rows per direction from one work item); the automatic height stays inside
    the measured optima."""
    lib = pkg.library()

    def plan(rows, nx, want, sms=148):
        bands, per = ctypes.c_int(), ctypes.c_int()
        assert lib.lbm_b200_plan_bands(rows, nx, want, sms, ctypes.byref(bands), ctypes.byref(per)) == 0
        return bands.value, per.value

    for rows in list(range(2, 80)) + [127, 128, 129, 2048, 16384, 131076]:
        for want in (0, 1, 2, 3, 5, 8, 64, 1000):
            bands, per = plan(rows, 1024, want)
            last = rows - (bands - 1) * per
            assert bands >= 1 and per >= 1 and 0 < last <= max(per, rows)
            if bands > 1:
                assert per >= 2 and last >= 2, (rows, want, bands, per)
    # measured optima of the 12-warps-per-SM kernel (profiles/r02_fused2.md): 24..32 at 2048^2, 12..24 at 4096^2,
    # 32..64 at 16384^2 and on the 2048-row slabs of an eight-way split of 16384^2 (128 is 15 % slower there)
    assert plan(2048, 2048, 0)[1] in (24, 32) and plan(4096, 4096, 0)[1] in (12, 16, 24)
    assert plan(16384, 16384, 0)[1] in (32, 64) and plan(2048, 16384, 0)[1] in (16, 32, 64)
    assert lib.lbm_b200_plan_bands(1, 1024, 0, 148, None, None) != 0


def test_packed_multiplies_cannot_be_contracted(pkg):
    """ptxas 12.9 fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false, which would break bit-identity
    with the reference.  The library issues every packed multiply as fma(a, b, -0.0f) with an addend the compiler
    cannot see through (csrc/lbm_cell.cuh): in the SASS every FFMA2 must carry one broadcast scalar register as its
    addend, and no FMUL2 may exist."""
    import shutil
    import sass_hist
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not installed")
    seen = 0
    for name, instrs in sass_hist.functions(pkg.LIB_PATH).items():
        assert sass_hist.unsafe_packed(instrs) == [], name
        seen += sum(1 for i in instrs if sass_hist.opcode(i) in ("FFMA2", "FADD2"))
    assert seen > 1000                                        # the packed path is really in the library


def test_bench_parity_fixture_matches_the_oracle(pkg, oracle):
    """tests/golden/ring_parity.npz is what tests/golden/make_ring_parity.py would write today (oracle unchanged,
    obstacle generator unchanged) -- bench.py trusts it without a CPU solver on the product path."""
    par = pkg.parity
    for n in par.RANK_COUNTS:
        ny = par.ROWS_PER_RANK * n
        ob = par.obstacles(par.PERIOD, ny, n)
        cells = oracle.init_cells(par.PERIOD, ny, par.DENSITY)
        inv = pkg.decks.free_cells_inv(par.PERIOD * ny - int(ob.sum()))
        av = oracle.run(cells, ob, par.STEPS, par.DENSITY, par.ACCEL, par.OMEGA, inv)
        want_cells, want_av = par.expected(n)
        assert np.array_equal(cells.view(np.uint32), want_cells.view(np.uint32))
        assert np.array_equal(av.view(np.uint32), want_av.view(np.uint32))
        # a wider grid holds the same solution tiled: compare_slab accepts it, and notices a single flipped bit
        wide = np.tile(want_cells, (1, 4, 1))
        assert par.compare_slab(wide[par.ROWS_PER_RANK * (n - 1):], par.ROWS_PER_RANK * (n - 1), n) == 0
        wide.view(np.uint32)[3, 100, 4] ^= 1
        assert par.compare_slab(wide, 0, n) == 1


def test_obstacle_bit_packing_layout(pkg):
    ob = np.zeros((3, 70), np.int32)
    ob[1, 33] = 1
    ob[2, 69] = 5
    ob[0, 0] = 1
    bits_ = pkg.pack_obstacle_bits(ob)
    assert bits_.shape == (3, 3) and bits_.dtype == np.uint32
    assert bits_[0, 0] == 1 and bits_[1, 1] == 2 and bits_[2, 2] == 1 << 5 and int(bits_.sum()) == 1 + 2 + 32
