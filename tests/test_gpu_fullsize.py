"""BASELINE.json's full size (16384 x 16384) through size-independent properties.

The oracle cannot step 268 M cells in test time, but a grid whose obstacle map is periodic in x
with period P evolves periodically in x: every column x behaves exactly like column x mod P of
a P-wide grid (the x-wrap joins the copies seamlessly).  So the full-size GPU run is checked
BIT FOR BIT against the oracle on the narrow grid, tiled."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu

NX = NY = 16384
DENSITY, ACCEL, OMEGA = 0.1, 0.005, 1.85


def test_full_size_cylinder_array(pkg, oracle):
    """16384 x 16384 cylinder-array deck (period 128 in x) against the tiled 128-wide oracle, bit for bit."""
    period, iters = 128, 5
    narrow = pkg.decks.cylinder_array_obstacles(period, NY)
    cells = oracle.init_cells(period, NY, DENSITY)
    _, av_ref = oracle.run(cells, narrow, iters, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(narrow), exact=True)
    ref_pressure = oracle.final_state(cells, narrow, DENSITY)[3]
    obstacles = pkg.decks.cylinder_array_obstacles(NX, NY)
    assert np.array_equal(obstacles, np.tile(narrow, (1, NX // period)))
    with pkg.Simulation(NX, NY, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        av = sim.run(iters)
        pressure = sim.final_state()[3]
    assert np.array_equal(bits(pressure), bits(np.tile(ref_pressure, (1, NX // period))))
    assert np.max(np.abs(av.astype(np.float64) - av_ref) / av_ref) < 2e-6


def test_full_size_inplace(pkg, oracle):
    """One 9.66 GB buffer instead of two: 16384 x 16384 in place, an odd and an even number of steps, the state read
    back through the 256 MB staging buffer (38 chunks), against the tiled 128-wide oracle."""
    import torch
    period = 128
    rng = np.random.default_rng(43)
    narrow = narrow_pattern(period, rng)
    cells = oracle.init_cells(period, NY, DENSITY)
    obstacles = np.tile(narrow, (1, NX // period))
    free0 = torch.cuda.mem_get_info()[0]
    with pkg.Simulation(NX, NY, DENSITY, ACCEL, OMEGA, obstacles, inplace=True) as sim:
        used = free0 - torch.cuda.mem_get_info()[0]
        assert used < 1.1 * 36 * NX * NY                     # one buffer (+ mask, partials), not two
        done = 0
        for iters in (5, 3):
            _, av_ref = oracle.run(cells, narrow, iters, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(narrow), exact=True)
            av = sim.run(iters)
            done += iters
            for got, want in zip(sim.final_state(), oracle.final_state(cells, narrow, DENSITY)):
                assert np.array_equal(bits(got), bits(np.tile(want, (1, NX // period)))), f"after {done} steps"
            assert np.max(np.abs(av.astype(np.float64) - av_ref) / av_ref) < 2e-6


@pytest.mark.parametrize("steps", [2, 0])
def test_full_size_fused(pkg, oracle, steps):
    """Two timesteps per pass (kernel 5) and the automatic choice (kernel 7, three per pass) at 16384 x 16384 (137
    strips of 120 columns -- the last one 64 wide): 7 steps = three passes of two and a single-step tail, or two
    passes of three and one of one, against the tiled 128-wide oracle."""
    period, iters = 128, 7
    rng = np.random.default_rng(44)
    narrow = narrow_pattern(period, rng)
    cells = oracle.init_cells(period, NY, DENSITY)
    _, av_ref = oracle.run(cells, narrow, iters, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(narrow), exact=True)
    ref_fields = oracle.final_state(cells, narrow, DENSITY)
    obstacles = np.tile(narrow, (1, NX // period))
    with pkg.Simulation(NX, NY, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        sim.set_option("fused2", 1)
        sim.set_option("fused_steps", steps)
        assert (sim.get_option("kernel"), sim.get_option("fused_steps")) == ((5, 2) if steps == 2 else (7, 3))
        av = sim.run(iters)
        assert sim.get_option("launches") < 12
        for got, want in zip(sim.final_state(), ref_fields):
            assert np.array_equal(bits(got), bits(np.tile(want, (1, NX // period))))
        assert np.max(np.abs(av.astype(np.float64) - av_ref) / av_ref) < 2e-6


def narrow_pattern(period, rng):
    ob = (rng.random((NY, period)) < 0.02).astype(np.int32)
    ob[0, :] = ob[-1, :] = 1                     # the synthetic deck's channel walls
    ob[NY - 2, : period // 4] = 1                # some blocked cells in the accelerated row
    return ob


@pytest.mark.parametrize("n_slabs", [1, 4])
def test_full_size_matches_tiled_oracle(pkg, oracle, n_slabs):
    period, iters = 128, 6
    rng = np.random.default_rng(42)
    narrow = narrow_pattern(period, rng)
    inv_narrow = pkg.free_cells_inv(narrow)
    cells = oracle.init_cells(period, NY, DENSITY)
    _, av_ref = oracle.run(cells, narrow, iters, DENSITY, ACCEL, OMEGA, inv_narrow, exact=True)   # fp64 sum of the terms
    ref_fields = oracle.final_state(cells, narrow, DENSITY)

    obstacles = np.tile(narrow, (1, NX // period))
    with pkg.Simulation(NX, NY, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n_slabs, devices=[0] * n_slabs) as sim:
        av = sim.run(iters)
        fields = sim.final_state()
        for got, want in zip(fields, ref_fields):
            tiled = np.tile(want, (1, NX // period))
            assert np.array_equal(bits(got), bits(tiled))
        # same average: sum over 128 identical copies / (128 x free cells); 2e-6 covers the fp32 rounding of
        # free_cells_inv (two different cell counts) and of the final product
        assert np.max(np.abs(av.astype(np.float64) - av_ref) / av_ref) < 2e-6


def test_plain_channel_is_x_invariant_and_matches_an_8_wide_oracle(pkg, oracle):
    """The bench workload itself (walls on rows 0 and ny-1, nothing else)."""
    iters = 8
    narrow = pkg.decks.channel_obstacles(8, NY)
    cells = oracle.init_cells(8, NY, DENSITY)
    _, av_ref = oracle.run(cells, narrow, iters, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(narrow), exact=True)
    ref_pressure = oracle.final_state(cells, narrow, DENSITY)[3]
    with pkg.Simulation(NX, NY, DENSITY, ACCEL, OMEGA, pkg.decks.channel_obstacles(NX, NY)) as sim:
        av = sim.run(iters)
        pressure = sim.final_state()[3]
    assert np.array_equal(bits(pressure), bits(np.repeat(ref_pressure[:, :1], NX, axis=1)))
    assert np.max(np.abs(av.astype(np.float64) - av_ref) / av_ref) < 2e-6


def test_synthetic_deck_through_the_cli(pkg, oracle, tmp_path):
    """BASELINE.json configs[4] end to end: generated 16384 x 16384 channel deck (32768 obstacle lines) through
    the drop-in command line, final_state.dat switched off (it would be ~23 GB of text)."""
    import os
    import subprocess
    iters = 12
    pfile, ofile = pkg.decks.write_channel_deck(str(tmp_path), NX, NY, iters, density=DENSITY, accel=ACCEL, omega=OMEGA)
    res = subprocess.run([pkg.EXE_PATH, pfile, ofile], cwd=tmp_path, capture_output=True, text=True,
                         env={**os.environ, "LBM_FINAL_STATE": "0", "LBM_VERBOSE": "1"})
    assert res.returncode == 0, res.stderr
    assert res.stdout.splitlines()[0] == "==done==" and not os.path.exists(tmp_path / "final_state.dat")
    av = pkg.decks.read_av_vels(str(tmp_path / "av_vels.dat"))
    narrow = pkg.decks.channel_obstacles(8, NY)
    cells = oracle.init_cells(8, NY, DENSITY)
    _, av_ref = oracle.run(cells, narrow, iters, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(narrow), exact=True)
    assert av.shape == av_ref.shape and np.max(np.abs(av - av_ref) / av_ref) < 2e-6
    # Reynolds number printed from the final state: same value as the 8-wide oracle's (x-invariant flow), to fp32
    # summation noise of the 268 M-term sequential sum the reference prescribes
    reynolds = float(res.stdout.splitlines()[1].split()[-1])
    ref_re = float(oracle.reynolds(cells, narrow, pkg.free_cells_inv(narrow), OMEGA, 10))
    assert abs(reynolds - ref_re) / ref_re < 5e-2


def _enough_memory(host_gb, device_gb):
    import psutil
    import torch
    return psutil.virtual_memory().available > host_gb * 2**30 and torch.cuda.mem_get_info()[0] > device_gb * 2**30


@pytest.mark.parametrize("inplace,ny", [(False, 65536 + 4), (True, 131072 + 4)])
def test_maximum_sizes_index_arithmetic(pkg, oracle, inplace, ny):
    """Grids whose element offsets leave 32 bits: ping-pong 16384 x 65540 (9 planes of 2^30 floats, 77 GB for the
    pair) and in place 16384 x 131076 (planes of more than 2^31 floats, 77 GB) -- pressure of every cell against the
    tiled 128-wide oracle after 3 steps (in place: the shifted layout L1, read through the staging buffer)."""
    import ctypes
    host_gb = 4.0 * NX * ny * 2 / 2**30 + 8
    if not _enough_memory(host_gb, 80):
        pytest.skip("needs ~%d GB of host memory and 80 GB of device memory" % host_gb)
    period, iters = 128, 3
    narrow = np.zeros((ny, period), np.int32)
    narrow[0, :] = narrow[-1, :] = 1
    narrow[ny // 2, 5:9] = 1                       # something that is not x-invariant, far from the start of the planes
    narrow[ny - 3, 100:103] = 1
    cells = oracle.init_cells(period, ny, DENSITY)
    _, av_ref = oracle.run(cells, narrow, iters, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(narrow), exact=True)
    ref_pressure = oracle.final_state(cells, narrow, DENSITY)[3]
    obstacles = np.tile(narrow, (1, NX // period))
    with pkg.Simulation(NX, ny, DENSITY, ACCEL, OMEGA, obstacles, inplace=inplace) as sim:
        del obstacles
        av = sim.run(iters)
        pressure = np.empty((ny, NX), np.float32)
        rc = pkg.library().lbm_b200_get_final_state(sim._h, None, None, None, pressure.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        assert rc == 0
    assert np.max(np.abs(av.astype(np.float64) - av_ref) / av_ref) < 2e-6
    for y0 in range(0, ny, 8192):                  # compare in bands: no second full-size temporary
        band = slice(y0, min(ny, y0 + 8192))
        assert np.array_equal(bits(pressure[band]), bits(np.tile(ref_pressure[band], (1, NX // period)))), f"rows {band}"
