"""Round-2 additions on one B200: the arithmetic self-test, obstacle map formats, CUDA-graph replay across runs of
growing length (the graphs captured a freed av_vels pointer), the bounded halo wait with its error
report, back-to-back runs on a fused ring, and bench.py's parity case against the committed expectation."""
import numpy as np
import pytest

from conftest import bits, random_cells, random_obstacles

pytestmark = pytest.mark.gpu

DENSITY, ACCEL, OMEGA = 0.1, 0.005, 1.85


def test_arithmetic_selftest(pkg):
    """rcp_fast / sqrt_fast against __frcp_rn / __fsqrt_rn over every float of the fast range, packed fp32x2
    add / sub / mul against the scalar IEEE operations over 2^28 operand pairs of arbitrary bit patterns."""
    assert pkg.selftest(0) == (0, 0, 0)


@pytest.mark.parametrize("fmt", ["uint8", "bits"])
@pytest.mark.parametrize("n_slabs", [1, 3])
def test_obstacle_formats_give_the_same_run(pkg, fmt, n_slabs):
    """lbm_b200_create_ex: one byte per cell and one bit per cell (nx not a multiple of 32: padding bits set on
    purpose) against the reference's int per cell -- same mask, same free-cell count, same bits after 30 steps."""
    rng = np.random.default_rng(7)
    nx, ny, iters = 300, 41, 30
    obstacles = random_obstacles(rng, ny, nx, 0.08)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n_slabs, device=0) as sim:
        want_av = sim.run(iters)
        want = sim.get_cells()
    if fmt == "uint8":
        ob = (obstacles * 7).astype(np.uint8)                 # any non-zero value blocks
    else:
        ob = pkg.pack_obstacle_bits(obstacles)
        ob[:, -1] |= np.uint32(0xFFFFFFFF) << np.uint32(nx % 32)   # garbage in the padding bits past nx
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, ob, n_slabs=n_slabs, device=0, obstacles_format=fmt) as sim:
        av = sim.run(iters)
        assert np.array_equal(bits(sim.get_cells()), bits(want))
        assert np.array_equal(bits(av), bits(want_av))        # same free-cell count => same scaling, bit for bit


@pytest.mark.parametrize("n_slabs", [1, 2])
def test_graph_replay_survives_a_longer_second_run(pkg, oracle, n_slabs):
    """graph_steps=16, run(40) then run(400): the second run reallocates the device av_vels array; the graphs that
    captured the old pointer must be rebuilt (r1: they wrote the averages into freed memory)."""
    rng = np.random.default_rng(11)
    nx, ny = 128, 48
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    ref = oracle.init_cells(nx, ny, DENSITY)
    inv = pkg.free_cells_inv(obstacles)
    ref_av = oracle.run(ref, obstacles, 440, DENSITY, ACCEL, OMEGA, inv)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n_slabs, device=0) as sim:
        sim.set_option("graph_steps", 16)
        av = np.concatenate([sim.run(40), sim.run(400)])
        assert np.array_equal(bits(sim.get_cells()), bits(ref))
        assert np.max(np.abs(av - ref_av) / ref_av) < 1e-4


@pytest.mark.parametrize("fused2", [0, 1])
def test_a_missing_neighbour_times_out_instead_of_hanging(pkg, fused2):
    """One slab of a two-slab ring is never launched (test hook debug_skip_slab): its neighbour's second pass waits
    for halo rows that never come.  The wait gives up after spin_timeout_ms, the run drains, and sync reports
    LBM_B200_ERR_STATE naming the exchange -- where the reference would sit in MPI_Waitall (d2q9-bgk.c:364)."""
    nx, ny = 256, 32
    obstacles = np.zeros((ny, nx), np.int32)
    obstacles[0] = obstacles[-1] = 1
    sim = pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=2, device=0)
    try:
        sim.set_option("fused2", fused2)
        sim.set_option("spin_timeout_ms", 50)
        sim.set_option("debug_skip_slab", 1)
        with pytest.raises(pkg.LBMError) as err:
            sim.run(8)
        text = str(err.value)
        assert text.startswith("[4]") and "waiting for halo" in text and "neighbour" in text
        with pytest.raises(pkg.LBMError):                       # the handle stays failed
            sim.run(2)
    finally:
        sim.close()


def test_back_to_back_runs_on_a_fused_ring(pkg, oracle):
    """enqueue, enqueue, enqueue without a sync in between: the body-force pre-pass of every run (also the one on the
    halo copy of the driven row, ordered by the strip flags) must see the previous run's last pushes."""
    rng = np.random.default_rng(5)
    nx, ny = 480, 37
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    ref = oracle.init_cells(nx, ny, DENSITY)
    oracle.run(ref, obstacles, 7 + 6 + 5 + 2, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=4, device=0) as sim:
        sim.set_option("band_rows", 4)
        sim.set_option("fused_steps", 2)                     # kernel 5 (the automatic choice is kernel 7, K = 3)
        sim.set_option("fused2", 1)
        for it in (7, 6, 5, 2):
            sim.enqueue(it)
        sim.sync()
        assert np.array_equal(bits(sim.get_cells()), bits(ref))


@pytest.mark.parametrize("steps", [2, 0])
@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_bench_parity_case_on_one_device(pkg, n, steps):
    """What bench.py checks before its timed region, here with the N slabs of a whole-domain handle on one device:
    the committed expectation (tests/golden/ring_parity.npz) is met bit for bit by the two-steps-per-pass kernel and
    by the automatic choice (kernel 7, three steps per pass)."""
    par = pkg.parity
    nx, ny = 256, par.ROWS_PER_RANK * n
    ob = par.obstacles(nx, ny, n)
    with pkg.Simulation(nx, ny, par.DENSITY, par.ACCEL, par.OMEGA, ob, n_slabs=n, device=0) as sim:
        sim.set_option("fused_steps", steps)
        sim.set_option("fused2", 1)
        assert (sim.get_option("kernel"), sim.get_option("fused_steps")) == ((5, 2) if steps == 2 else (7, 3))
        for it in par.RUNS:
            sim.enqueue(it)
        sim.sync()
        assert par.compare_slab(sim.get_cells(), 0, n) == 0


def test_final_state_chunks_match_the_oracle(pkg, oracle):
    """get_final_state in overlapped row chunks (>= 64 rows): the same four fields as write_values computes."""
    rng = np.random.default_rng(3)
    nx, ny, iters = 128, 203, 25
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    ref = oracle.init_cells(nx, ny, DENSITY)
    oracle.run(ref, obstacles, iters, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
    want = oracle.final_state(ref, obstacles, DENSITY)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, device=0) as sim:
        sim.run(iters)
        got = sim.final_state()
    for g, w in zip(got, want):
        assert np.array_equal(bits(g), bits(w))


@pytest.mark.parametrize("iters", [20, 21])
def test_fused_deep_variant_is_bit_identical(pkg, oracle, iters):
    """Kernel 5 with two staging rows (copies two rows ahead, 3 CTAs x 4 warps per SM): same bits, ragged bands,
    obstacles, an odd tail."""
    rng = np.random.default_rng(17)
    nx, ny = 500, 77
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    cells0 = random_cells(rng, ny, nx)
    ref = cells0.copy()
    ref_av = oracle.run(ref, obstacles, iters, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
    for band in (0, 5):
        with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, device=0) as sim:
            sim.set_option("fused_steps", 2)                     # kernel 5 (the automatic choice is kernel 7, K = 3)
            sim.set_option("fused2", 1)
            sim.set_option("fused_deep", 1)
            sim.set_option("band_rows", band)
            sim.set_cells(cells0)
            av = sim.run(iters)
            assert np.array_equal(bits(sim.get_cells()), bits(ref))
            assert np.max(np.abs(av - ref_av) / ref_av) < 1e-4


@pytest.mark.parametrize("shape,rows_form", [((128, 128), 1), ((128, 256), 1), ((128, 16), 1), ((128, 48), 1),
                                             ((128, 128), 0), ((128, 256), 0), ((64, 48), 0), ((200, 32), 0)])
def test_cluster_resident_kernel_is_bit_identical(pkg, oracle, shape, rows_form):
    """Kernel 6: the grid in the shared memory of one 16-CTA cluster, halo rows over DSMEM -- the general form (any
    nx, remote loads) and the one-warp-per-row form for nx = 128 (packed pairs, pushed halo rows, 1 .. 16 rows per
    CTA).  Random obstacles on every edge, an arbitrary initial state, runs that are odd, even, longer than one
    256-step launch, and back to back."""
    nx, ny = shape
    rng = np.random.default_rng(nx + ny)
    obstacles = random_obstacles(rng, ny, nx, 0.06, walls=(rows_form == 0))
    obstacles[ny - 2, :] = rng.random(nx) < 0.2                # blocked cells in the driven row
    cells0 = random_cells(rng, ny, nx)
    cells0[ny - 2, : nx // 3, 3] = 1e-5                        # ... and cells the body force skips
    ref = cells0.copy()
    runs = (7, 300, 1, 256)
    ref_av = oracle.run(ref, obstacles, sum(runs), DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, device=0) as sim:
        sim.set_option("cluster_rows", rows_form)
        sim.set_option("cluster", 1)
        assert sim.get_option("kernel") == 6 and sim.get_option("cluster") == 1
        assert sim.get_option("cluster_rows") == rows_form
        sim.set_cells(cells0)
        av = np.concatenate([sim.run(n) for n in runs])
        assert np.array_equal(bits(sim.get_cells()), bits(ref))
        assert np.max(np.abs(av - ref_av) / ref_av) < 1e-4
        fields = sim.final_state()
    for g, w in zip(fields, oracle.final_state(ref, obstacles, DENSITY)):
        assert np.array_equal(bits(g), bits(w))


def test_cluster_kernel_is_the_default_for_the_small_decks_only(pkg):
    ob = np.zeros((128, 128), np.int32)
    with pkg.Simulation(128, 128, DENSITY, ACCEL, OMEGA, ob, device=0) as sim:
        assert sim.get_option("kernel") == 6
        sim.set_option("graph_steps", 64)                   # asking for a launch mode switches it off
        assert sim.get_option("kernel") != 6
    ob = np.zeros((256, 256), np.int32)
    with pkg.Simulation(256, 256, DENSITY, ACCEL, OMEGA, ob, device=0) as sim:   # 2 x 72 B x 4096 cells per CTA do not fit
        assert sim.get_option("kernel") != 6
    ob = np.zeros((120, 128), np.int32)                      # ny % 16 != 0
    with pkg.Simulation(128, 120, DENSITY, ACCEL, OMEGA, ob, device=0) as sim:
        assert sim.get_option("kernel") != 6


@pytest.mark.parametrize("steps,stage_rows", [(2, 1), (2, 2), (3, 1), (3, 2), (4, 1), (4, 2)])
@pytest.mark.parametrize("shape,band", [((500, 77), 0), ((500, 77), 5), ((240, 9), 0), ((1024, 40), 7), ((368, 130), 64)])
def test_k_steps_per_pass_kernel_is_bit_identical(pkg, oracle, steps, stage_rows, shape, band):
    """Kernel 7 (three / four timesteps per pass over HBM): random obstacles on every edge (x- and y-wrap), an
    arbitrary initial state, ragged last strip and band, run lengths that leave every possible shorter last pass
    (K' = 1 .. K-1), back-to-back runs, and a run longer than one 256-step chunk."""
    nx, ny = shape
    rng = np.random.default_rng(nx + ny + steps)
    obstacles = random_obstacles(rng, ny, nx, 0.05, walls=False)
    cells0 = random_cells(rng, ny, nx)
    ref = cells0.copy()
    runs = (1, 2, 3, 4, 5, 7, 260) if shape == (500, 77) and band == 0 else (7, 6, 1, 12)
    ref_av = oracle.run(ref, obstacles, sum(runs), DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, device=0) as sim:
        sim.set_option("fused2", 1)
        sim.set_option("fused_k7", 1)                      # kernel 7 also for two steps per pass (default: kernel 5)
        sim.set_option("fused_steps", steps)
        sim.set_option("fused_deep", stage_rows - 1)       # one or two staging rows (copies one / two rows ahead)
        sim.set_option("band_rows", band)
        if band == 5:
            sim.set_option("fused_ctas", 1)                # all resident warps in ONE CTA instead of one warp per CTA
        assert sim.get_option("kernel") == 7 and sim.get_option("fused_steps") == steps
        sim.set_cells(cells0)
        av = np.concatenate([sim.run(n) for n in runs])
        assert np.array_equal(bits(sim.get_cells()), bits(ref))
        assert np.max(np.abs(av - ref_av) / ref_av) < 1e-4


@pytest.mark.parametrize("iters", [3, 7, 12])
@pytest.mark.parametrize("steps", [3, 4])
@pytest.mark.parametrize("n_slabs,nx,ny,band", [(2, 256, 16, 64), (3, 360, 29, 4), (4, 244, 47, 5), (8, 600, 56, 64), (2, 1024, 12, 2)])
def test_k_steps_ring_slabs_on_one_device_bit_exact(pkg, oracle, n_slabs, nx, ny, band, steps, iters):
    """Kernel 7 on a ring: four halo rows per side, the steps before the last recomputed for the neighbours' rows,
    one push of four rows per direction and one flag handshake per pass; shorter last passes through the same strips.
    Ragged strips, bands of 4..64 rows, slabs of 6..15 rows, the driven row recomputed by the first slab (its row -2),
    blocked and non-forced cells in it.  Then a second run, kernel 5 and the one-step kernel on the same ring, and
    back (the deeper halo is fetched again)."""
    rng = np.random.default_rng(n_slabs * 1000 + nx + iters + steps)
    obstacles = random_obstacles(rng, ny, nx, 0.08, walls=(ny % 2 == 0))
    obstacles[:, 0] = rng.random(ny) < 0.5
    obstacles[:, -1] = rng.random(ny) < 0.5
    obstacles[ny - 2, :] = rng.random(nx) < 0.2
    cells0 = random_cells(rng, ny, nx)
    cells0[ny - 2, : nx // 3, 3] = 1e-5
    inv = pkg.free_cells_inv(obstacles)
    ref = cells0.copy()
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n_slabs, devices=[0] * n_slabs) as sim:
        sim.set_option("band_rows", band)
        sim.set_option("fused2", 1)
        sim.set_option("fused_steps", steps)
        assert sim.get_option("kernel") == 7 and sim.get_option("fused_steps") == steps
        sim.set_cells(cells0)
        for n, opts in ((iters, {}), (5, {}), (4, {"fused_steps": 2}), (3, {"fused2": 0}), (6, {"fused2": 1, "fused_steps": steps})):
            for key, value in opts.items():
                sim.set_option(key, value)
            ref_av = oracle.run(ref, obstacles, n, DENSITY, ACCEL, OMEGA, inv)
            av = sim.run(n)
            assert np.array_equal(bits(sim.get_cells()), bits(ref)), (n, opts)
            assert np.max(np.abs(av - ref_av) / ref_av) < 1e-4
        assert sim.get_option("kernel") == 7


def test_k_steps_kernel_is_refused_on_thin_ring_slabs(pkg):
    ob = np.zeros((15, 256), np.int32)
    with pkg.Simulation(256, 15, DENSITY, ACCEL, OMEGA, ob, n_slabs=3, device=0) as sim:   # 5 rows per slab: kernel 5 fits, 7 does not
        sim.set_option("fused2", 1)
        sim.set_option("fused_steps", 4)
        assert sim.get_option("kernel") == 5 and sim.get_option("fused_steps") == 2


def test_back_to_back_runs_on_a_k_steps_ring(pkg, oracle):
    """enqueue x 4 without a sync in between on a kernel-7 ring (the pre-pass on the halo copy of the driven row is
    ordered by the strip flags against the previous run's last pushes)."""
    rng = np.random.default_rng(6)
    nx, ny = 480, 45
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    ref = oracle.init_cells(nx, ny, DENSITY)
    oracle.run(ref, obstacles, 7 + 6 + 5 + 2, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=4, device=0) as sim:
        sim.set_option("band_rows", 4)
        sim.set_option("fused2", 1)
        sim.set_option("fused_steps", 4)
        assert sim.get_option("kernel") == 7
        for it in (7, 6, 5, 2):
            sim.enqueue(it)
        sim.sync()
        assert np.array_equal(bits(sim.get_cells()), bits(ref))
