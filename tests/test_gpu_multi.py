"""Row slabs on SEVERAL B200s: halo rows travel as NVLink stores from the edge-row kernel,
flag words order the exchange.  Needs >= 2 visible GPUs (gpurun --gpus 2); skipped otherwise.
(The same code path with all slabs on one device is covered by test_gpu_parity.py.)"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, bits, random_cells, random_obstacles

pytestmark = pytest.mark.gpu

DENSITY, ACCEL, OMEGA = 0.1, 0.005, 1.85


def need_gpus(pkg, n):
    if pkg.device_count() < n:
        pytest.skip(f"needs {n} GPUs, {pkg.device_count()} visible")


@pytest.mark.parametrize("inplace", [False, True])
@pytest.mark.parametrize("n", [2, 4, 8])
def test_single_process_multi_device(pkg, oracle, n, inplace):
    need_gpus(pkg, n)
    rng = np.random.default_rng(n)
    nx, ny, iters = 256, 8 * n + 5, 201 if inplace else 200       # in place: end in the shifted layout L1
    obstacles = random_obstacles(rng, ny, nx, 0.06)
    cells0 = random_cells(rng, ny, nx)
    ref = cells0.copy()
    ref_av = oracle.run(ref, obstacles, iters, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n, inplace=inplace) as sim:
        sim.set_cells(cells0)
        av = sim.run(iters)
        assert np.array_equal(bits(sim.get_cells()), bits(ref))
        assert np.max(np.abs(av - ref_av) / ref_av) < 1e-4


@pytest.mark.parametrize("n", [2, 4, 8])
def test_fused2_across_devices(pkg, oracle, n):
    """Two timesteps per pass on a ring of real devices: two halo rows per side over NVLink, one strip-level flag
    handshake per pass, an odd tail."""
    need_gpus(pkg, n)
    rng = np.random.default_rng(60 + n)
    nx, ny, iters = 600, 9 * n + 3, 201
    obstacles = random_obstacles(rng, ny, nx, 0.06)
    cells0 = random_cells(rng, ny, nx)
    ref = cells0.copy()
    ref_av = oracle.run(ref, obstacles, iters + 30, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n) as sim:
        sim.set_option("band_rows", 4)
        sim.set_option("fused_steps", 2)                     # kernel 5 (the automatic choice is kernel 7, K = 3)
        sim.set_option("fused2", 1)
        assert sim.get_option("kernel") == 5
        sim.set_cells(cells0)
        av = np.concatenate([sim.run(iters), sim.run(30)])
        assert np.array_equal(bits(sim.get_cells()), bits(ref))
        assert np.max(np.abs(av - ref_av) / ref_av) < 1e-4


@pytest.mark.parametrize("steps", [3, 4])
@pytest.mark.parametrize("n", [2, 4, 8])
def test_k_steps_per_pass_across_devices(pkg, oracle, n, steps):
    """Kernel 7 on a ring of real devices: four halo rows per side over NVLink, one strip-level flag handshake per
    pass of three / four timesteps, shorter last passes."""
    need_gpus(pkg, n)
    rng = np.random.default_rng(70 + n + steps)
    nx, ny, iters = 600, 11 * n + 3, 202
    obstacles = random_obstacles(rng, ny, nx, 0.06)
    cells0 = random_cells(rng, ny, nx)
    ref = cells0.copy()
    ref_av = oracle.run(ref, obstacles, iters + 31, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n) as sim:
        sim.set_option("band_rows", 4)
        sim.set_option("fused2", 1)
        sim.set_option("fused_steps", steps)
        assert sim.get_option("kernel") == 7
        sim.set_cells(cells0)
        av = np.concatenate([sim.run(iters), sim.run(31)])
        assert np.array_equal(bits(sim.get_cells()), bits(ref))
        assert np.max(np.abs(av - ref_av) / ref_av) < 1e-4


@pytest.mark.parametrize("inplace", [False, True])
@pytest.mark.parametrize("n", [2, 8])
def test_graph_replay_across_devices(pkg, oracle, n, inplace):
    """One independent CUDA graph per device, ordered only by the flag words the kernels exchange over NVLink."""
    need_gpus(pkg, n)
    rng = np.random.default_rng(40 + n)
    nx, ny, iters = 384, 6 * n + 4, 1500                          # automatic choice: 256-step graphs
    obstacles = random_obstacles(rng, ny, nx, 0.06)
    cells0 = random_cells(rng, ny, nx)
    ref = cells0.copy()
    ref_av = oracle.run(ref, obstacles, iters + 45, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
    with pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=n, inplace=inplace) as sim:
        sim.set_cells(cells0)
        av = sim.run(iters)
        assert sim.get_option("launches") > iters
        sim.set_option("graph_steps", 8)
        av = np.concatenate([av, sim.run(45)])
        assert np.array_equal(bits(sim.get_cells()), bits(ref))
        assert np.max(np.abs(av - ref_av) / ref_av) < 1e-4


@pytest.mark.parametrize("inplace", [False, True, "fused2", "fused4"])
@pytest.mark.parametrize("n", [2, 4, 8])
def test_one_rank_per_gpu_over_ipc(pkg, oracle, n, inplace, tmp_path):
    """torchrun-style launch: n processes, CUDA IPC handles exchanged with torch.distributed."""
    need_gpus(pkg, n)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "result.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multi_rank_worker.py"), str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                         env={**os.environ, "LBM_TEST_INPLACE": "1" if inplace is True else "0",
                              "LBM_TEST_FUSED2": "1" if inplace in ("fused2", "fused4") else "0",
                              "LBM_TEST_FUSED_STEPS": "4" if inplace == "fused4" else ""})
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    data = np.load(out)
    obstacles, cells = data["obstacles"], data["cells"]
    iters = int(data["iters"])
    ny, nx = obstacles.shape
    ref = oracle.init_cells(nx, ny, DENSITY)
    ref_av = oracle.run(ref, obstacles, iters, DENSITY, ACCEL, OMEGA, pkg.free_cells_inv(obstacles))
    assert np.array_equal(bits(cells), bits(ref))
    assert np.max(np.abs(data["av"] - ref_av) / ref_av) < 1e-4


@pytest.mark.parametrize("deck", ["128x256", "1024x1024"])
def test_c_launcher_one_process_per_gpu(pkg, tmp_path, deck):
    """bin/d2q9-bgk-mp -np N: the C program that forks one rank per GPU and wires the ring with lbm_b200_create_slab_ex /
    ipc_export / ipc_connect -- no Python anywhere.  Its final_state.dat must be byte-identical to the strict build of
    the reference (sha256 fixture), whatever the number of ranks."""
    import hashlib
    import json
    from conftest import GOLDEN, deck_paths
    n = min(pkg.device_count(), 4)
    need_gpus(pkg, 2)
    pfile, ofile = deck_paths(deck)
    want = json.load(open(os.path.join(GOLDEN, "ref_strict.json")))["decks"][deck]
    res = subprocess.run([pkg.EXE_PATH + "-mp", "-np", str(n), pfile, ofile], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = res.stdout.splitlines()
    assert lines[0] == "==done==" and lines[1].startswith("Reynolds number:\t\t")
    assert hashlib.sha256((tmp_path / "final_state.dat").read_bytes()).hexdigest() == want["final_state_sha256"]
    assert lines[1].split()[-1] == want["reynolds"]


def test_one_rank_alone_times_out(pkg):
    """Two ranks on two GPUs, only rank 0 runs: its kernels give up waiting for rank 1's halo rows after
    spin_timeout_ms and lbm_b200_sync reports LBM_B200_ERR_STATE instead of hanging the GPU."""
    need_gpus(pkg, 2)
    nx, ny = 256, 32
    obstacles = np.zeros((ny, nx), np.int32)
    obstacles[0] = obstacles[-1] = 1
    sim = pkg.Simulation(nx, ny, DENSITY, ACCEL, OMEGA, obstacles, n_slabs=2, devices=[0, 1])
    try:
        sim.set_option("spin_timeout_ms", 100)
        sim.set_option("debug_skip_slab", 1)
        with pytest.raises(pkg.LBMError) as err:
            sim.run(6)
        assert "waiting for halo" in str(err.value)
    finally:
        sim.close()
