"""One rank of tests/test_gpu_multi.py::test_one_rank_per_gpu_over_ipc (launched by torchrun)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
from conftest import random_obstacles  # noqa: E402

DENSITY, ACCEL, OMEGA = 0.1, 0.005, 1.85


def main(out_path):
    rank, size = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    pkg = entry.load_package()
    dist.init_process_group("gloo")
    inplace = os.environ.get("LBM_TEST_INPLACE") == "1"
    nx, ny, iters = 384, 16 * size + 3, 301 if inplace else 300   # in place: end in the shifted layout L1
    if os.environ.get("LBM_TEST_FUSED2") == "1":
        iters = 303                                               # pairs of steps and a one-step tail
    obstacles = random_obstacles(np.random.default_rng(77), ny, nx, 0.06)
    rows, first = pkg.decompose(ny, size)
    r, f = int(rows[rank]), int(first[rank])
    sim = pkg.Simulation.slab(nx, ny, f, r, rank, size, DENSITY, ACCEL, OMEGA, float(pkg.free_cells_inv(obstacles)),
                              obstacles[f:f + r], device=local, inplace=inplace)
    blobs = [None] * size
    dist.all_gather_object(blobs, sim.export_ipc())
    sim.connect_ipc(blobs[(rank - 1) % size], blobs[(rank + 1) % size])
    if os.environ.get("LBM_TEST_FUSED2") == "1":
        sim.set_option("band_rows", 8)
        sim.set_option("fused_steps", 2)                     # kernel 5 (the automatic choice is kernel 7, K = 3)
        sim.set_option("fused2", 1)
        assert sim.get_option("kernel") == 5
        if os.environ.get("LBM_TEST_FUSED_STEPS"):
            sim.set_option("fused_steps", int(os.environ["LBM_TEST_FUSED_STEPS"]))
            assert sim.get_option("kernel") == 7
    dist.barrier()
    av = torch.from_numpy(sim.run(iters).copy())
    dist.barrier()
    cells = sim.get_cells()
    dist.reduce(av, dst=0, op=dist.ReduceOp.SUM)                 # reference d2q9-bgk.c:396
    gathered = [None] * size
    dist.gather_object(cells, gathered if rank == 0 else None, dst=0)
    if rank == 0:
        np.savez(out_path, obstacles=obstacles, cells=np.concatenate(gathered, axis=0), av=av.numpy(), iters=iters)
    dist.barrier()
    sim.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
