#!/usr/bin/env python3
"""bench.py -- throughput of the D2Q9-BGK timestep path on B200, in MLUPS.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[4]): the synthetic channel -- walls on rows 0 and
ny-1, periodic in x, driven on row ny-2 -- of 16384 x 16384 cells PER GPU (weak scaling: the
global grid is 16384 x 16384*N, split into N row slabs).  One bench "step" is a batch of
--timesteps LBM timesteps; K steps are enqueued as ONE run of K*timesteps timesteps with no host
synchronisation inside, exactly like the reference's timed loop (d2q9-bgk.c:278-398).

One JSON line is printed by rank 0:
  value      MLUPS = cells * timesteps / device time; CUDA events recorded by the library on the
             stream its kernels run on, maximum over ranks; state resident in HBM
  e2e        the same metric for a whole job through the C-ABI from HOST buffers: create
             (obstacle upload) + run (per-step averages copied back) + final macroscopic fields
             copied back to pinned host memory
  roofline   72 B/cell/step algorithmic traffic of the fused step kernel against the measured
             HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own CPU code (oracle/_ref, built from the unmodified source)
             timed on this box's host cores on a bounded sample of the same workload

--impl reference times only that CPU build (rank 0 alone under torchrun).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

BYTES_PER_CELL_STEP = 72.0          # 9 fp32 loads + 9 fp32 stores
FALLBACK_HBM_GBS = 6650.0           # /opt/skills/guides/B200_PROFILING.md
DENSITY, ACCEL, OMEGA = 0.1, 0.005, 1.85


def env_int(name, default):
    return int(os.environ.get(name, default))


# ----------------------------------------------------------------------------------------
# clocks during the timed region
# ----------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._proc = None
        try:
            self._proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._thread = threading.Thread(target=self._read, daemon=True)
            self._thread.start()
        except OSError:
            self._proc = None

    def _read(self):
        for line in self._proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
            except ValueError:
                continue
            for name, flag in zip(self.NAMES, parts[2:6]):
                if flag.lower().startswith("active"):
                    self.reasons.add(name)

    def stop(self):
        if self._proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self._proc.terminate()
        try:
            self._proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self._proc.kill()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------
# the reference's CPU implementation on the host cores (oracle/_ref)
# ----------------------------------------------------------------------------------------
def usable_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample_shape(nx: int):
    """A bounded sample of the workload: same row length, fewer rows, a few timesteps."""
    return nx, env_int("LBM_BENCH_CPU_ROWS", 1024), env_int("LBM_BENCH_CPU_ITERS", 16)


def run_reference_cpu(pkg, nx: int, ranks: int, kind: str = "fast.noio"):
    """Runs the compiled reference once on the sample; returns (MLUPS, description)."""
    exe = os.path.join(ROOT, "oracle", "_ref", f"d2q9-bgk.{kind}")
    if not os.path.exists(exe):
        return None, f"oracle/_ref/d2q9-bgk.{kind} is not built"
    snx, sny, iters = cpu_sample_shape(nx)
    ranks = max(1, min(ranks, 64, sny // 4))
    with tempfile.TemporaryDirectory() as tmp:
        pfile, ofile = pkg.decks.write_channel_deck(tmp, snx, sny, iters, density=DENSITY, accel=ACCEL, omega=OMEGA)
        out = subprocess.run([exe, pfile, ofile], cwd=tmp, capture_output=True, text=True,
                             env={**os.environ, "MPI_SHIM_RANKS": str(ranks), "OMP_NUM_THREADS": "1"})
    if out.returncode != 0:
        return None, f"reference exited {out.returncode}: {out.stderr[-200:]}"
    m = re.search(r"Elapsed time:\s+([0-9.]+)", out.stdout)
    elapsed = float(m.group(1))
    mlups = snx * sny * iters / elapsed / 1e6
    sample = (f"unmodified reference source (gcc -O3 -march=x86-64-v3, fork+shm mpi.h shim, {ranks} rank(s)) on a "
              f"{snx}x{sny} channel, {iters} timesteps, its own 'Elapsed time' = {elapsed:.3f} s")
    return mlups, sample


def cpu_baseline(pkg, nx: int):
    cores = usable_cores()
    ranks = min(cores, 64)
    best, sample = None, ""
    for _ in range(2):
        v, sample = run_reference_cpu(pkg, nx, ranks)
        if v is None:
            return {"value": None, "unit": "MLUPS", "cores": ranks, "kind": "reference", "sample": sample}
        best = v if best is None else max(best, v)
    serial, _ = run_reference_cpu(pkg, nx, 1)
    return {"value": round(best, 2), "unit": "MLUPS", "cores": ranks, "kind": "reference", "sample": sample,
            "serial_value": round(serial, 2) if serial else None}


def bench_reference(args, pkg):
    """--impl reference: the reference's CPU path, all host cores, rank 0 only."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    nx = args.nx
    cores = min(usable_cores(), 64)
    values, sample = [], ""
    for i in range(args.warmup + args.steps):
        v, sample = run_reference_cpu(pkg, nx, cores)
        if v is None:
            print(json.dumps({"impl": "reference", "unavailable": sample}))
            return 0
        if i >= args.warmup:
            values.append(v)
    snx, sny, iters = cpu_sample_shape(nx)
    value = statistics.mean(values)
    line = {
        "impl": "reference", "metric": "MLUPS", "value": round(value, 2), "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(snx * sny * iters / value / 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": round(value, 2), "unit": "MLUPS", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
def workload_config(args, n):
    strong = getattr(args, "scaling", "weak") == "strong"
    ny_global = args.ny if strong else args.ny * n
    per_gpu = f"{args.nx}x{ny_global // n}" if strong else f"{args.nx}x{args.ny}"
    return {
        "workload": f"synthetic channel {per_gpu} cells per GPU (global {args.nx}x{ny_global}, "
                    f"{n} row slab(s)), walls on rows 0 and ny-1, x periodic, accel on row ny-2",
        "nx": args.nx, "ny_per_gpu": ny_global // n, "ny_global": ny_global,
        "timesteps_per_step": args.timesteps,
        "density": DENSITY, "accel": ACCEL, "omega": OMEGA,
        "layout": ("fp32 SoA, 9 planes, ONE buffer streamed in place (AA access pattern); 1-bit obstacle mask"
                   if getattr(args, "inplace", False) else "fp32 SoA, 9 planes, ping-pong; 1-bit obstacle mask"),
        "l2": f"no L2 flush needed: {(1 if getattr(args, 'inplace', False) else 2) * 9 * 4 * args.nx * (ny_global // n) / 1e9:.1f} GB of state per GPU streams through the 126 MB L2 every timestep",
        "parallelism": f"row-slab x{n}" + (", halo rows as NVLink stores from the edge-row kernel" if n > 1 else ""),
    }


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, torch copy)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def measured_traffic(nx, ny, kernel):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture (profiles/roofline_traffic.json,
    one record per kernel), scaled by the cell count if the capture was taken at another size."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        rec = json.load(open(path))
        rec = rec.get("kernels", {}).get(str(kernel), rec)
        if rec["nx"] == nx and rec["ny"] == ny:
            return rec["dram_bytes_per_launch"]
        return rec["dram_bytes_per_launch"] / (rec["nx"] * rec["ny"]) * nx * ny
    except Exception:
        return None


def bench_ours(args, pkg):
    import torch
    import torch.distributed as dist

    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device -- the product has no CPU fallback")
    torch.cuda.set_device(local)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nx, n = args.nx, world
    if args.scaling == "strong":
        split_rows, split_first = pkg.decompose(args.ny, n)
        rows, first, ny_global = int(split_rows[rank]), int(split_first[rank]), args.ny
    else:
        rows, ny_global, first = args.ny, args.ny * n, rank * args.ny
    ips, K, W = args.timesteps, args.steps, args.warmup
    free = nx * (ny_global - 2)
    inv = float(np.float32(1.0) / np.float32(free))

    # host buffers (pinned): this rank's obstacle rows in, macroscopic fields + av_vels out
    obstacles_t = torch.zeros((rows, nx), dtype=torch.int32).pin_memory()
    obstacles = obstacles_t.numpy()
    if rank == 0:
        obstacles[0, :] = 1
    if rank == n - 1:
        obstacles[rows - 1, :] = 1

    def make_sim():
        if n == 1:
            sim = pkg.Simulation(nx, rows, DENSITY, ACCEL, OMEGA, obstacles, device=local, inplace=args.inplace)
        else:
            sim = pkg.Simulation.slab(nx, ny_global, first, rows, rank, n, DENSITY, ACCEL, OMEGA, inv, obstacles, device=local,
                                      inplace=args.inplace)
            blobs = [None] * n
            dist.all_gather_object(blobs, sim.export_ipc())
            sim.connect_ipc(blobs[(rank - 1) % n], blobs[(rank + 1) % n])
            dist.barrier()
        for key in ("kernel", "graph_steps", "ctas_per_sm", "min_ctas", "fused2", "band_rows"):
            v = getattr(args, key)
            if v is not None and not (args.inplace and key in ("kernel", "min_ctas", "fused2", "band_rows")):
                sim.set_option(key, v)
        return sim

    sim = make_sim()
    kernel = sim.get_option("kernel")
    kernel_name = {1: "step_scalar", 2: "step_vec4", 3: "steps_resident", 4: "step_inplace",
                   5: "steps2_strip (two timesteps per pass)"}.get(kernel, str(kernel))

    # ---- warm-up ------------------------------------------------------------------------
    for _ in range(W):
        sim.enqueue(ips)
    sim.sync()
    barrier()

    # ---- timed region: exactly K steps, no host synchronisation inside -------------------
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t0 = time.perf_counter()
    sim.enqueue(K * ips)
    sim.sync()
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    device_ms = max_over_ranks(sim.elapsed_ms())
    wall_s = max_over_ranks(wall_s)
    launches = sim.get_option("launches")
    av = sim.fetch_av_vels(K * ips)
    if distributed:                                   # the reference's final MPI_Reduce (d2q9-bgk.c:396)
        t = torch.from_numpy(av.copy()).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        av = t.cpu().numpy()
    if not (np.all(np.isfinite(av)) and av[-1] > 0):
        sys.exit("bench.py: the run produced non-finite or zero average velocities")
    sim.close()
    del sim

    cells_global = float(nx) * ny_global
    timesteps = K * ips
    value = cells_global * timesteps / (device_ms * 1e-3) / 1e6
    # the dominant kernel: one launch per timestep, or (kernel 5, "fused2") one per PAIR of timesteps
    steps_per_launch = 2 if kernel == 5 else 1
    step_launches = timesteps // steps_per_launch
    launch_ms = device_ms * steps_per_launch / timesteps
    peak, peak_src = peak_hbm()
    achieved = BYTES_PER_CELL_STEP * steps_per_launch * nx * rows / (launch_ms * 1e-3) / 1e9     # per GPU, per launch
    traffic = measured_traffic(nx, rows, kernel)

    # ---- e2e: a whole job through the C-ABI from host buffers ----------------------------
    fields_t = torch.empty((4, rows, nx), dtype=torch.float32).pin_memory()
    av_t = torch.empty(timesteps, dtype=torch.float32).pin_memory()
    e2e_s, e2e_phases, pressure_ok = None, None, True
    for _ in range(2):                                # best of two jobs: allocating 19 GB right after freeing it varies
        barrier()
        t0 = time.perf_counter()
        sim = make_sim()                              # H2D: this rank's obstacle rows (int per cell, packed on the device)
        t1 = time.perf_counter()
        sim.enqueue(timesteps)
        sim.fetch_av_vels(timesteps, av_t.numpy())    # D2H: per-step averages (synchronises)
        if args.inplace and distributed:
            dist.barrier()                            # in place, edge-row populations may live in the neighbours' buffers
        t2 = time.perf_counter()
        sim.final_state(fields_t.numpy())             # D2H: u_x, u_y, |u|, pressure
        t3 = time.perf_counter()
        barrier()
        job_s = max_over_ranks(time.perf_counter() - t0)
        pressure_ok = pressure_ok and bool(torch.isfinite(fields_t[3]).all())
        sim.close()
        if e2e_s is None or job_s < e2e_s:
            e2e_s = job_s
            e2e_phases = {"create_s": round(t1 - t0, 4), "run_and_av_vels_s": round(t2 - t1, 4), "final_state_s": round(t3 - t2, 4)}
    e2e_value = cells_global * timesteps / e2e_s / 1e6
    if not pressure_ok:
        sys.exit("bench.py: e2e run produced a non-finite pressure field")
    h2d = 4 * nx * rows
    d2h = 4 * 4 * nx * rows + 4 * timesteps

    baseline = cpu_baseline(pkg, nx) if (rank == 0 and n == 1 and not args.no_cpu_baseline) else None

    if rank == 0:
        line = {
            "metric": "MLUPS", "value": round(value, 1), "unit": "MLUPS", "n_gpus": n, "steps": K, "warmup": W,
            "ms_per_step": round(device_ms / K, 4), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, n),
            "kernel": kernel_name,
            "wall_ms_per_step": round(wall_s * 1e3 / K, 4),
            "av_vels_last": float(av[-1]),
            "e2e": {"value": round(e2e_value, 1), "unit": "MLUPS", "h2d_bytes_per_step": h2d // K,
                    "d2h_bytes_per_step": d2h // K,
                    "what": "lbm_b200_create (obstacle upload) + enqueue + fetch_av_vels + get_final_state into pinned "
                            f"host memory, {timesteps} timesteps, wall clock, max over ranks, better of two jobs",
                    "phases_rank0": e2e_phases},
            "gpu_launches": int(launches) * n,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic,
                         "peak_source": peak_src,
                         "what": (f"{kernel_name}, {BYTES_PER_CELL_STEP:.0f} B/cell/step x {steps_per_launch} timestep(s) x "
                                  f"{nx * rows} cells per launch per GPU, mean launch {launch_ms * 1e3:.1f} us over "
                                  f"{step_launches} launches"
                                  + ("; two timesteps are fused per pass over HBM, so the DRAM traffic per launch (`traffic`) is "
                                     "about half the algorithmic bytes and `frac` exceeds the one-step streaming roofline: the "
                                     "kernel is bound by instruction issue" if kernel == 5 else "")),
                         "dram_gbs": round(traffic / (launch_ms * 1e-3) / 1e9, 1) if traffic else None},
            "clocks": clocks,
        }
        if baseline is not None:
            line["cpu_baseline"] = baseline
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--nx", type=int, default=16384)
    ap.add_argument("--ny", type=int, default=16384, help="rows per GPU")
    ap.add_argument("--timesteps", type=int, default=100, help="LBM timesteps per bench step")
    ap.add_argument("--kernel", type=int, default=None)
    ap.add_argument("--graph-steps", dest="graph_steps", type=int, default=None)
    ap.add_argument("--ctas-per-sm", dest="ctas_per_sm", type=int, default=None)
    ap.add_argument("--min-ctas", dest="min_ctas", type=int, default=None)
    ap.add_argument("--fused2", type=int, default=None, choices=[-1, 0, 1],
                    help="two timesteps per pass over HBM: 1 on, 0 off, default automatic (on from 2^22 cells per GPU)")
    ap.add_argument("--band-rows", dest="band_rows", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the bit-exact multi-GPU parity check before the timed region")
    ap.add_argument("--inplace", action="store_true",
                    help="one population buffer per GPU, streamed in place (half the memory, same traffic)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak: --ny rows per GPU (default); strong: --ny rows in total, split over the GPUs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                               # timing rule: at least 3 warm-up steps
    pkg = entry.load_package()
    if args.impl == "reference":
        return bench_reference(args, pkg)
    return bench_ours(args, pkg)


if __name__ == "__main__":
    sys.exit(main())
