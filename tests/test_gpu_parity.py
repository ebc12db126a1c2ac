"""GPU parity proper: the CUDA path (through the C-ABI) against the oracle on the same inputs.

Bar: every population of every cell BIT-EXACT (the kernels issue the reference's fp32
operations in the reference's order, no FMA contraction).  av_vels: the device sums the
per-cell |m|/rho terms in a fixed fp64 tree, the reference sequentially in an fp32 accumulator,
so summation order is the only difference; tolerances: 2e-6 relative against the oracle's
fp64 sum of the same terms (AV_RTOL_EXACT), 1e-4 against the reference-order fp32 sum
(AV_RTOL_REF, the reference's own accumulation error at these sizes)."""
import numpy as np
import pytest

from conftest import bits, random_cells, random_obstacles

pytestmark = pytest.mark.gpu

AV_RTOL_EXACT = 2e-6
AV_RTOL_REF = 1e-4
DENSITY, ACCEL, OMEGA = 0.1, 0.005, 1.85


def oracle_run(oracle, pkg, cells0, obstacles, iters, density=DENSITY, accel=ACCEL, omega=OMEGA):
    cells = cells0.copy()
    av, av_exact = oracle.run(cells, obstacles, iters, density, accel, omega, pkg.free_cells_inv(obstacles), exact=True)
    return cells, av, av_exact


def assert_av(av, ref_av, ref_exact):
    if len(av):
        assert np.max(np.abs(av.astype(np.float64) - ref_exact) / np.abs(ref_exact)) < AV_RTOL_EXACT
        assert np.max(np.abs(av.astype(np.float64) - ref_av) / np.abs(ref_av)) < AV_RTOL_REF


def assert_parity(sim, oracle, pkg, cells0, obstacles, iters, **kw):
    ref_cells, ref_av, ref_exact = oracle_run(oracle, pkg, cells0, obstacles, iters, **kw)
    av = sim.run(iters)
    got = sim.get_cells()
    mism = np.argwhere(bits(got) != bits(ref_cells))
    assert mism.size == 0, f"{len(mism)} populations differ, first at (y,x,k)={mism[0].tolist()}"
    assert_av(av, ref_av, ref_exact)
    return ref_cells


# nx: multiples of 128, ragged segments (136, 200), tiny (8, 12), and not a multiple of 4 (scalar kernel)
SHAPES = [(128, 128), (256, 64), (136, 20), (200, 33), (8, 8), (12, 5), (1024, 16), (30, 17), (7, 9), (129, 12)]


@pytest.mark.parametrize("nx,ny", SHAPES)
def test_uniform_start_bit_exact(pkg, oracle, nx, ny):
    rng = np.random.default_rng(nx * 131 + ny)
    obstacles = random_obstacles(rng, ny, nx, 0.06)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        # 128-bit kernel whenever nx allows it; its resident (cooperative, many steps per launch) variant
        # is the automatic choice only from 2^18 cells; grids that fit into one cluster's shared memory (ny % 16
        # == 0, 2 x 36 B per cell in 16 x 227 KB) run the cluster-resident kernel
        fits_cluster = ny % 16 == 0 and 2 * 36 * nx * (ny // 16) <= 227 * 1024
        assert sim.get_option("kernel") == (6 if fits_cluster else (2 if nx % 4 == 0 and nx >= 8 else 1))
        cells0 = oracle.init_cells(nx, ny, DENSITY)
        assert np.array_equal(bits(sim.get_cells()), bits(cells0))     # initialise(): d2q9-bgk.c:880-902
        assert_parity(sim, oracle, pkg, cells0, obstacles, 30)


@pytest.mark.parametrize("kernel", [1, 2, 3])
@pytest.mark.parametrize("nx,ny,walls", [(128, 24, True), (136, 19, False), (256, 7, False)])
def test_random_state_bit_exact(pkg, oracle, kernel, nx, ny, walls):
    """Arbitrary positive states, obstacles on every edge, open top/bottom rows (y-wrap), x-wrap."""
    rng = np.random.default_rng(kernel * 7 + nx + ny)
    obstacles = random_obstacles(rng, ny, nx, 0.10, walls=walls)
    obstacles[:, 0] = rng.random(ny) < 0.5                 # ragged side walls: x-wrap sees fluid and solid
    obstacles[:, -1] = rng.random(ny) < 0.5
    cells0 = random_cells(rng, ny, nx)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        sim.set_option("resident", 1 if kernel == 3 else 0)
        sim.set_option("kernel", min(kernel, 2))
        assert sim.get_option("kernel") == kernel
        sim.set_cells(cells0)
        assert np.array_equal(bits(sim.get_cells()), bits(cells0))
        assert_parity(sim, oracle, pkg, cells0, obstacles, 17)


@pytest.mark.parametrize("iters", [0, 1, 2, 3])
def test_short_runs_and_the_folded_accelerate(pkg, oracle, iters):
    """accelerate_flow is folded into the previous step's store; 0/1/2/3 steps pin the pre-pass,
    the fold and the un-accelerated last step."""
    rng = np.random.default_rng(5)
    nx, ny = 128, 12
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    obstacles[ny - 2, 10:20] = 1                            # blocked cells inside the accelerated row
    cells0 = random_cells(rng, ny, nx)
    cells0[ny - 2, 30:40, 3] = 1e-5                         # cells where the force must NOT be applied (f3 - w1 <= 0)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        sim.set_cells(cells0)
        assert_parity(sim, oracle, pkg, cells0, obstacles, iters)


def test_runs_compose(pkg, oracle):
    """run(7) + run(6) + run(0) + run(1) == run(14): the handle keeps the canonical state between runs."""
    rng = np.random.default_rng(9)
    nx, ny = 256, 20
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    cells0 = random_cells(rng, ny, nx)
    ref_cells, ref_av, ref_exact = oracle_run(oracle, pkg, cells0, obstacles, 14)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        sim.set_cells(cells0)
        av = np.concatenate([sim.run(7), sim.run(6), sim.run(0), sim.run(1)])
        assert np.array_equal(bits(sim.get_cells()), bits(ref_cells))
        assert_av(av, ref_av, ref_exact)


@pytest.mark.parametrize("min_ctas,ctas_per_sm,hint", [(2, 0, 0), (3, 0, 1), (2, 1, 2), (4, 7, 0), (2, 3, 4), (2, 0, 4)])
def test_launch_geometry_does_not_change_results(pkg, oracle, min_ctas, ctas_per_sm, hint):
    rng = np.random.default_rng(11)
    nx, ny = 1024, 40
    obstacles = random_obstacles(rng, ny, nx, 0.03)
    cells0 = random_cells(rng, ny, nx)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        sim.set_option("resident", 0)
        sim.set_option("min_ctas", min_ctas)
        sim.set_option("ctas_per_sm", ctas_per_sm)
        sim.set_option("cache_hint", hint)
        sim.set_cells(cells0)
        assert_parity(sim, oracle, pkg, cells0, obstacles, 9)


@pytest.mark.parametrize("graph_steps,iters", [(4, 23), (8, 8), (6, 5), (16, 300)])
def test_cuda_graph_replay_is_identical(pkg, oracle, graph_steps, iters):
    rng = np.random.default_rng(13)
    nx, ny = 128, 16
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    cells0 = random_cells(rng, ny, nx)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        sim.set_option("resident", 0)
        sim.set_option("graph_steps", graph_steps)
        sim.set_cells(cells0)
        assert_parity(sim, oracle, pkg, cells0, obstacles, iters)


@pytest.mark.parametrize("nx,ny,iters", [(128, 16, 300), (1024, 40, 513), (256, 256, 257), (8, 8, 31)])
def test_resident_kernel_many_steps_per_launch(pkg, oracle, nx, ny, iters):
    """The cooperative kernel: up to 256 steps per launch with a grid barrier in between; odd step counts
    (buffer parity), chunk boundaries (256/257/513) and a grid larger than the co-resident CTA count."""
    rng = np.random.default_rng(nx + iters)
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    cells0 = random_cells(rng, ny, nx)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        sim.set_option("resident", 1)
        assert sim.get_option("kernel") == 3
        sim.set_cells(cells0)
        assert_parity(sim, oracle, pkg, cells0, obstacles, iters)
        assert sim.get_option("launches") <= 3 * ((iters + 255) // 256) + 1


def test_av_vels_is_deterministic(pkg):
    rng = np.random.default_rng(17)
    nx, ny = 512, 64
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    runs = []
    for _ in range(2):
        with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as sim:
            runs.append(sim.run(50))
    assert np.array_equal(bits(runs[0]), bits(runs[1]))


@pytest.mark.parametrize("n_slabs,nx,ny", [(2, 128, 16), (3, 136, 19), (4, 256, 31), (8, 128, 24)])
def test_slabs_on_one_device_bit_exact(pkg, oracle, n_slabs, nx, ny):
    """The multi-GPU path (edge/interior launches, halo rows stored into the neighbour's buffer,
    flag words) with all slabs on device 0 in one stream: same bits as the single domain."""
    rng = np.random.default_rng(n_slabs * 100 + ny)
    obstacles = random_obstacles(rng, ny, nx, 0.08, walls=(ny % 2 == 0))
    cells0 = random_cells(rng, ny, nx)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n_slabs, devices=[0] * n_slabs) as sim:
        assert sim.get_option("launches_per_step") == 1      # edge rows + halo push + interior in ONE launch per slab
        sim.set_cells(cells0)
        ref = assert_parity(sim, oracle, pkg, cells0, obstacles, 21)
        ux, uy, u, pr = sim.final_state()
        rux, ruy, ru, rpr = oracle.final_state(ref, obstacles, DENSITY)
        for got, want in ((ux, rux), (uy, ruy), (u, ru), (pr, rpr)):
            assert np.array_equal(bits(got), bits(want))


@pytest.mark.parametrize("n_slabs,nx,ny", [(2, 30, 16), (3, 128, 19)])
def test_slabs_with_the_scalar_kernel(pkg, oracle, n_slabs, nx, ny):
    """nx not a multiple of 4 (or kernel 1 forced): edge rows in their own launch with a CTA-level handshake."""
    rng = np.random.default_rng(n_slabs * 31 + nx)
    obstacles = random_obstacles(rng, ny, nx, 0.08, walls=(ny % 2 == 0))
    cells0 = random_cells(rng, ny, nx)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n_slabs, devices=[0] * n_slabs) as sim:
        sim.set_option("kernel", 1)
        assert sim.get_option("launches_per_step") == 2
        sim.set_cells(cells0)
        assert_parity(sim, oracle, pkg, cells0, obstacles, 23)


def test_final_state_fields_bit_exact(pkg, oracle):
    rng = np.random.default_rng(19)
    nx, ny = 200, 33
    obstacles = random_obstacles(rng, ny, nx, 0.1)
    cells0 = random_cells(rng, ny, nx)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        sim.set_cells(cells0)
        sim.run(5)
        fields = sim.final_state()
        ref = oracle.final_state(sim.get_cells(), obstacles, DENSITY)
    for got, want in zip(fields, ref):
        assert np.array_equal(bits(got), bits(want))
    assert np.all(fields[3][obstacles == 1] == np.float32(DENSITY) * (np.float32(1.0) / np.float32(3.0)))
    assert np.all(fields[2][obstacles == 1] == 0)


def test_other_physical_parameters(pkg, oracle):
    rng = np.random.default_rng(23)
    nx, ny = 128, 20
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    cells0 = random_cells(rng, ny, nx, density=0.3)
    with pkg.Simulation(nx, ny, 0.3, 0.01, 1.2, obstacles) as sim:
        sim.set_cells(cells0)
        assert_parity(sim, oracle, pkg, cells0, obstacles, 11, density=0.3, accel=0.01, omega=1.2)


def test_total_density_is_conserved_on_device(pkg, oracle):
    rng = np.random.default_rng(29)
    nx, ny = 256, 64
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        before = oracle.total_density(sim.get_cells())
        sim.run(400)
        after = oracle.total_density(sim.get_cells())
    assert abs(after - before) / before < 1e-5


def test_cylinder_array_deck(pkg, oracle):
    """The richer synthetic deck (discs in a channel): many obstacle edges at every alignment within a segment."""
    nx, ny = 512, 192
    obstacles = pkg.decks.cylinder_array_obstacles(nx, ny)
    cells0 = oracle.init_cells(nx, ny, DENSITY)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        assert_parity(sim, oracle, pkg, cells0, obstacles, 60)


# ---- in-place streaming (one population buffer, AA access pattern) ---------------------------------

INPLACE_SHAPES = [(128, 24), (256, 7), (136, 19), (200, 33), (8, 8), (12, 5), (1024, 16), (132, 3)]


@pytest.mark.parametrize("nx,ny", INPLACE_SHAPES)
@pytest.mark.parametrize("iters", [1, 2, 17])
def test_inplace_bit_exact(pkg, oracle, nx, ny, iters):
    """Odd and even step counts: after an odd one the buffer is in the shifted layout L1, which get_cells and
    get_final_state decode; ragged side walls exercise the x-wrap, open rows the y-wrap."""
    rng = np.random.default_rng(nx * 17 + ny + iters)
    obstacles = random_obstacles(rng, ny, nx, 0.10, walls=(ny % 2 == 0))
    obstacles[:, 0] = rng.random(ny) < 0.5
    obstacles[:, -1] = rng.random(ny) < 0.5
    obstacles[ny - 2, :] = rng.random(nx) < 0.2              # blocked cells inside the accelerated row
    cells0 = random_cells(rng, ny, nx)
    cells0[ny - 2, : nx // 3, 3] = 1e-5                      # cells where the force must not be applied
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, inplace=True) as sim:
        assert sim.get_option("inplace") == 1 and sim.get_option("kernel") == 4
        assert np.array_equal(bits(sim.get_cells()), bits(oracle.init_cells(nx, ny, DENSITY)))
        sim.set_cells(cells0)
        assert np.array_equal(bits(sim.get_cells()), bits(cells0))
        ref_cells = assert_parity(sim, oracle, pkg, cells0, obstacles, iters)
        for got, want in zip(sim.final_state(), oracle.final_state(ref_cells, obstacles, DENSITY)):
            assert np.array_equal(bits(got), bits(want))


@pytest.mark.parametrize("hint", [0, 1, 2])
def test_inplace_runs_compose_across_layout_parities(pkg, oracle, hint):
    """run(3) + run(4) + run(0) + run(1) + run(5) == run(13): the body-force pre-pass and the un-accelerated last
    step work from either layout, and set_cells after an odd run returns to the canonical one."""
    rng = np.random.default_rng(21 + hint)
    nx, ny = 384, 20
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    cells0 = random_cells(rng, ny, nx)
    ref_cells, ref_av, ref_exact = oracle_run(oracle, pkg, cells0, obstacles, 13)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, inplace=True) as sim:
        sim.set_option("cache_hint", hint)
        sim.set_option("staging_bytes", 5 * nx * 36)          # five rows per chunk: the staged getters loop
        sim.run(3)
        sim.set_cells(cells0)                                # from layout L1 back to a fresh canonical state
        av = np.concatenate([sim.run(3), sim.run(4), sim.run(0), sim.run(1), sim.run(5)])
        assert np.array_equal(bits(sim.get_cells()), bits(ref_cells))
        assert_av(av, ref_av, ref_exact)
        for got, want in zip(sim.final_state(), oracle.final_state(ref_cells, obstacles, DENSITY)):
            assert np.array_equal(bits(got), bits(want))


def test_inplace_equals_ping_pong_over_a_graph_replayed_run(pkg):
    """1100 steps of a 256 x 64 grid: CUDA-graph replay of 256-step chunks on both handles; every bit equal,
    av_vels identical (same partial sums in the same tree when the launch geometry is the same)."""
    rng = np.random.default_rng(33)
    nx, ny, iters = 256, 64, 1101
    obstacles = random_obstacles(rng, ny, nx, 0.04)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as a, \
            pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, inplace=True) as b:
        a.set_option("resident", 0)
        a.set_option("cluster", 0)                           # (the cluster-resident kernel sums in another tree)
        av_a, av_b = a.run(iters), b.run(iters)
        assert np.array_equal(bits(a.get_cells()), bits(b.get_cells()))
        assert np.array_equal(bits(av_a), bits(av_b))


def test_inplace_rejects_what_it_cannot_do(pkg):
    ob = np.zeros((9, 30), np.int32)
    with pytest.raises(pkg.LBMError, match="nx % 4"):
        pkg.Simulation(30, 9, DENSITY, ACCEL, OMEGA, ob, inplace=True)
    with pkg.Simulation(32, 9, DENSITY, ACCEL, OMEGA, np.zeros((9, 32), np.int32), inplace=True) as sim:
        with pytest.raises(pkg.LBMError, match="does not apply"):
            sim.set_option("kernel", 1)


@pytest.mark.parametrize("iters", [20, 21])
@pytest.mark.parametrize("n_slabs,nx,ny", [(2, 128, 16), (3, 136, 19), (4, 256, 31), (8, 128, 24)])
def test_inplace_slabs_on_one_device_bit_exact(pkg, oracle, n_slabs, nx, ny, iters):
    """In-place slabs in a ring: the NEIGHBOUR flavour writes into the neighbours' owned edge rows, the LOCAL flavour
    pushes halo copies; after an odd number of steps the getters decode edge-row populations from the neighbours."""
    rng = np.random.default_rng(n_slabs * 100 + ny + iters)
    obstacles = random_obstacles(rng, ny, nx, 0.08, walls=(ny % 2 == 0))
    obstacles[:, 0] = rng.random(ny) < 0.5
    obstacles[:, -1] = rng.random(ny) < 0.5
    cells0 = random_cells(rng, ny, nx)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n_slabs, devices=[0] * n_slabs, inplace=True) as sim:
        assert sim.get_option("launches_per_step") == 1 and sim.get_option("inplace") == 1
        sim.set_cells(cells0)
        ref = assert_parity(sim, oracle, pkg, cells0, obstacles, iters)
        for got, want in zip(sim.final_state(), oracle.final_state(ref, obstacles, DENSITY)):
            assert np.array_equal(bits(got), bits(want))
        # and on from whichever layout the run ended in
        ref2 = ref.copy()
        ref_av, ref_exact = oracle.run(ref2, obstacles, 7, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles), exact=True)
        av = sim.run(7)
        assert np.array_equal(bits(sim.get_cells()), bits(ref2))
        assert_av(av, ref_av, ref_exact)


@pytest.mark.parametrize("inplace", [False, True])
@pytest.mark.parametrize("n_slabs,kernel", [(3, 2), (2, 1)])
def test_ring_slabs_replayed_from_graphs(pkg, oracle, n_slabs, kernel, inplace):
    """CUDA-graph replay of a ring: one graph per stream holds `graph_steps` steps of every slab on it (the flag
    handshake lives in device memory, so a replay is as good as the launches it recorded).  37 steps in chunks of 8
    plus a plain tail, then the automatic choice over 600 steps."""
    if inplace and kernel == 1:
        pytest.skip("in-place streaming has no one-cell-per-thread kernel")
    rng = np.random.default_rng(n_slabs * 7 + kernel)
    nx, ny = 136, 25
    obstacles = random_obstacles(rng, ny, nx, 0.08)
    cells0 = random_cells(rng, ny, nx)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n_slabs, devices=[0] * n_slabs, inplace=inplace) as sim:
        if not inplace:
            sim.set_option("kernel", kernel)
        sim.set_option("graph_steps", 8)
        sim.set_cells(cells0)
        ref = assert_parity(sim, oracle, pkg, cells0, obstacles, 37)
        sim.set_option("graph_steps", -1)
        ref2 = ref.copy()
        ref_av, ref_exact = oracle.run(ref2, obstacles, 600, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles), exact=True)
        av = sim.run(600)
        assert np.array_equal(bits(sim.get_cells()), bits(ref2))
        assert_av(av, ref_av, ref_exact)
        assert sim.get_option("launches") >= 600


def test_switching_kernels_on_a_ring_keeps_the_handshake_consistent(pkg, oracle):
    """The scalar kernel (slab-level flags) and the 128-bit kernel (one flag per 128-cell chunk) share the epoch."""
    rng = np.random.default_rng(77)
    nx, ny, n_slabs = 256, 22, 3
    obstacles = random_obstacles(rng, ny, nx, 0.06)
    cells0 = random_cells(rng, ny, nx)
    ref_cells, ref_av, ref_exact = oracle_run(oracle, pkg, cells0, obstacles, 15)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n_slabs, devices=[0] * n_slabs) as sim:
        sim.set_cells(cells0)
        av = [sim.run(4)]
        sim.set_option("kernel", 1)
        av.append(sim.run(5))
        sim.set_option("kernel", 2)
        av.append(sim.run(6))
        assert np.array_equal(bits(sim.get_cells()), bits(ref_cells))
        assert_av(np.concatenate(av), ref_av, ref_exact)


# ---- two timesteps per pass over HBM (kernel 5, "fused2") -----------------------------------------------------------

@pytest.mark.parametrize("nx,ny,band", [(256, 24, 64), (360, 19, 5), (244, 33, 7), (1024, 16, 4), (240, 9, 3), (2048, 40, 64),
                                        (256, 4, 0), (4096, 70, 0)])
@pytest.mark.parametrize("iters", [2, 5, 16])
def test_fused2_bit_exact(pkg, oracle, nx, ny, band, iters):
    """Pairs of timesteps fused into one pass (first step into a shared-memory ring, second step out of it): ragged last
    strips (nx not a multiple of 120), ragged last bands, bands of 3..64 rows, odd step counts (a single-step tail),
    x- and y-wrap, blocked cells and non-forced cells in the accelerated row."""
    rng = np.random.default_rng(nx + 31 * ny + iters)
    obstacles = random_obstacles(rng, ny, nx, 0.10, walls=(ny % 2 == 0))
    obstacles[:, 0] = rng.random(ny) < 0.5
    obstacles[:, -1] = rng.random(ny) < 0.5
    obstacles[ny - 2, :] = rng.random(nx) < 0.2
    cells0 = random_cells(rng, ny, nx)
    cells0[ny - 2, : nx // 3, 3] = 1e-5
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        sim.set_option("band_rows", band)                       # 0 = automatic
        sim.set_option("fused_steps", 2)                     # kernel 5 (the automatic choice is kernel 7, K = 3)
        sim.set_option("fused2", 1)
        if nx == 4096:
            sim.set_option("ctas_per_sm", 1)                    # fewer CTAs than work items: the item loop strides
        assert sim.get_option("kernel") == 5
        sim.set_cells(cells0)
        assert_parity(sim, oracle, pkg, cells0, obstacles, iters)
        assert sim.get_option("launches") <= iters // 2 + iters % 2 + 4


def test_fused2_runs_compose_and_match_the_single_step_kernel(pkg, oracle):
    rng = np.random.default_rng(91)
    nx, ny = 600, 70
    obstacles = random_obstacles(rng, ny, nx, 0.05)
    cells0 = random_cells(rng, ny, nx)
    ref_cells, ref_av, ref_exact = oracle_run(oracle, pkg, cells0, obstacles, 13)
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles) as sim:
        sim.set_option("fused_steps", 2)                     # kernel 5 (the automatic choice is kernel 7, K = 3)
        sim.set_option("fused2", 1)
        sim.set_option("band_rows", 16)
        sim.set_cells(cells0)
        av = np.concatenate([sim.run(3), sim.run(4), sim.run(0), sim.run(1), sim.run(5)])
        assert np.array_equal(bits(sim.get_cells()), bits(ref_cells))
        assert_av(av, ref_av, ref_exact)
        sim.set_option("fused2", 0)                          # and back to one step per launch on the same handle
        assert sim.get_option("kernel") in (2, 3, 6)


def test_fused2_is_refused_where_it_does_not_apply(pkg):
    ob = np.zeros((12, 128), np.int32)
    with pkg.Simulation(128, 12, DENSITY, ACCEL, OMEGA, ob) as sim:     # narrower than two strips: stays on kernel 2
        sim.set_option("fused_steps", 2)                     # kernel 5 (the automatic choice is kernel 7, K = 3)
        sim.set_option("fused2", 1)
        assert sim.get_option("kernel") != 5 and sim.get_option("fused2") == 0


@pytest.mark.parametrize("iters", [2, 7, 12])
@pytest.mark.parametrize("n_slabs,nx,ny,band", [(2, 256, 16, 64), (3, 360, 29, 4), (4, 244, 47, 5), (8, 600, 40, 64), (2, 1024, 9, 2)])
def test_fused2_ring_slabs_on_one_device_bit_exact(pkg, oracle, n_slabs, nx, ny, band, iters):
    """Two timesteps per pass on a ring: every slab recomputes the first step of its neighbours' edge rows from two halo
    rows per side, and pushes two rows per direction once per pass; odd step counts end with a one-step pass through
    the same strips.  Ragged strips, bands of 2..64 rows, slabs of 4..24 rows, the driven row's copy in the first
    slab's second halo row."""
    rng = np.random.default_rng(n_slabs * 1000 + nx + iters)
    obstacles = random_obstacles(rng, ny, nx, 0.08, walls=(ny % 2 == 0))
    obstacles[:, 0] = rng.random(ny) < 0.5
    obstacles[:, -1] = rng.random(ny) < 0.5
    obstacles[ny - 2, :] = rng.random(nx) < 0.2
    cells0 = random_cells(rng, ny, nx)
    cells0[ny - 2, : nx // 3, 3] = 1e-5
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n_slabs, devices=[0] * n_slabs) as sim:
        sim.set_option("band_rows", band)
        sim.set_option("fused_steps", 2)                     # kernel 5 (the automatic choice is kernel 7, K = 3)
        sim.set_option("fused2", 1)
        assert sim.get_option("kernel") == 5
        sim.set_cells(cells0)
        ref = assert_parity(sim, oracle, pkg, cells0, obstacles, iters)
        # a second run from the state the first one left (halo rows refreshed by its last pass), then back to the
        # one-step kernels on the same ring
        ref2 = ref.copy()
        ref_av, ref_exact = oracle.run(ref2, obstacles, 5, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles), exact=True)
        av = sim.run(5)
        assert np.array_equal(bits(sim.get_cells()), bits(ref2))
        assert_av(av, ref_av, ref_exact)
        sim.set_option("fused2", 0)
        ref3 = ref2.copy()
        oracle.run(ref3, obstacles, 3, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
        sim.run(3)
        assert np.array_equal(bits(sim.get_cells()), bits(ref3))
        sim.set_option("fused_steps", 2)                     # kernel 5 (the automatic choice is kernel 7, K = 3)
        sim.set_option("fused2", 1)                          # and on again: the two halo rows per side are fetched afresh
        oracle.run(ref3, obstacles, 4, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
        sim.run(4)
        assert np.array_equal(bits(sim.get_cells()), bits(ref3))
        sim.set_option("kernel", 1)                          # the scalar kernel switches it off implicitly ...
        assert sim.get_option("fused2") == 0
        oracle.run(ref3, obstacles, 3, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
        sim.run(3)
        sim.set_option("kernel", 0)                          # ... and leaving it brings it back: halo rows fetched again
        assert sim.get_option("fused2") == 1
        oracle.run(ref3, obstacles, 4, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
        sim.run(4)
        assert np.array_equal(bits(sim.get_cells()), bits(ref3))
