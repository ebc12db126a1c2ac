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
  parity     BEFORE anything is timed, every rank steps a small case through the very same path (one slab per
             rank, CUDA IPC ring, the kernel the timed region uses) and compares its slab bit for bit with the
             committed expectation (tests/golden/ring_parity.npz, made by the oracle); a mismatch ends the run
  value      MLUPS = cells * timesteps / device time; CUDA events recorded by the library on the
             stream its kernels run on, maximum over ranks; state resident in HBM
  e2e        the same metric for WHOLE jobs of the full configuration (10 000 timesteps) through the C-ABI from
             HOST buffers: create (obstacle upload) + run (per-step averages copied back) + final macroscopic
             fields copied back to pinned host memory; every job is listed, `value` is their median
  roofline   72 B/cell/step algorithmic traffic of the fused step kernel against the measured
             HBM copy bandwidth (MEASURED_PEAKS.json)
  strong     (N > 1) the 16384 x 16384 grid split over the N GPUs, timed like `value`, with rank 0's own
             one-GPU time of the same grid for the efficiency
  cpu_baseline  the reference's own CPU code (oracle/_ref, built from the unmodified source)
             timed on this box's host cores on a bounded sample of the same workload

--impl reference times only that CPU build (rank 0 alone under torchrun).
"""
from __future__ import annotations

import argparse
import glob
import hashlib
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
FULL_JOB_TIMESTEPS = 10000          # BASELINE.json configs[4]: "10k steps"


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


class CpuSample:
    """A bounded sample of the workload for the CPU arm: the same row length, fewer rows, and enough timesteps that
    the reference's own 'Elapsed time' is at least ~3 s -- so that the first-touch page faults of its tmp_cells
    (allocated before `tic`, d2q9-bgk.c:869, first written inside the loop) are a few per cent of the measurement
    and not a fifth of it, as with the 16-timestep sample of round 1."""

    MIN_ELAPSED_S = 3.2

    def __init__(self, nx: int, serial: bool = False):
        self.nx = nx
        self.rows = env_int("LBM_BENCH_CPU_ROWS", 128 if serial else 1024)
        self.iters = env_int("LBM_BENCH_CPU_ITERS", 200)
        self.calibrated = "LBM_BENCH_CPU_ITERS" in os.environ

    def calibrate(self, elapsed: float):
        """After the first (warm-up) run: more timesteps if that run was shorter than MIN_ELAPSED_S."""
        if not self.calibrated and elapsed < self.MIN_ELAPSED_S:
            want = self.iters * self.MIN_ELAPSED_S * 1.1 / max(elapsed, 1e-3)
            self.iters = int(min(4000, -(-want // 50) * 50))
        elif not self.calibrated and elapsed > 2.5 * self.MIN_ELAPSED_S and "LBM_BENCH_CPU_ROWS" not in os.environ:
            # a host with few cores: fewer rows (never fewer timesteps), so that K runs still end within minutes
            self.rows = int(max(64, min(self.rows, self.rows * 1.25 * self.MIN_ELAPSED_S / elapsed) // 16 * 16))
        self.calibrated = True

    def workload(self) -> str:
        return f"synthetic channel {self.nx}x{self.rows} cells, {self.iters} timesteps"


def run_reference_cpu(pkg, sample: CpuSample, ranks: int, kind: str = "fast.noio"):
    """Runs the compiled reference once on the sample; returns (MLUPS, elapsed seconds, description)."""
    exe = os.path.join(ROOT, "oracle", "_ref", f"d2q9-bgk.{kind}")
    if not os.path.exists(exe):
        return None, None, f"oracle/_ref/d2q9-bgk.{kind} is not built"
    snx, sny, iters = sample.nx, sample.rows, sample.iters
    ranks = max(1, min(ranks, 64, sny // 4))
    with tempfile.TemporaryDirectory() as tmp:
        pfile, ofile = pkg.decks.write_channel_deck(tmp, snx, sny, iters, density=DENSITY, accel=ACCEL, omega=OMEGA)
        out = subprocess.run([exe, pfile, ofile], cwd=tmp, capture_output=True, text=True,
                             env={**os.environ, "MPI_SHIM_RANKS": str(ranks), "OMP_NUM_THREADS": "1"})
    if out.returncode != 0:
        return None, None, f"reference exited {out.returncode}: {out.stderr[-200:]}"
    m = re.search(r"Elapsed time:\s+([0-9.]+)", out.stdout)
    elapsed = float(m.group(1))
    mlups = snx * sny * iters / elapsed / 1e6
    desc = (f"unmodified reference source (gcc -O3 -march=x86-64-v3, fork+shm mpi.h shim, {ranks} rank(s)) on a "
            f"{snx}x{sny} channel, {iters} timesteps, its own 'Elapsed time'")
    return mlups, elapsed, desc


def timed_reference_runs(pkg, nx: int, ranks: int, runs: int, warmup: int = 1, serial: bool = False):
    """`warmup` untimed runs (the first one calibrates the sample length), then `runs` timed ones."""
    sample = CpuSample(nx, serial=serial)
    values, elapsed_all, desc = [], [], ""
    for i in range(max(1, warmup) + runs):
        v, elapsed, desc = run_reference_cpu(pkg, sample, ranks)
        if v is None:
            return None, sample, desc
        if i == 0:
            sample.calibrate(elapsed)
        if i >= max(1, warmup):
            values.append(v)
            elapsed_all.append(elapsed)
    stats = {"median": statistics.median(values), "min": min(values), "max": max(values), "runs": len(values),
             "elapsed_s_median": statistics.median(elapsed_all)}
    return stats, sample, desc


def cpu_baseline(pkg, nx: int):
    cores = min(usable_cores(), 64)
    stats, sample, desc = timed_reference_runs(pkg, nx, cores, runs=3)
    if stats is None:
        return {"value": None, "unit": "MLUPS", "cores": cores, "kind": "reference", "sample": desc}
    serial, ssample, _ = timed_reference_runs(pkg, nx, 1, runs=1, serial=True)
    return {"value": round(stats["median"], 2), "unit": "MLUPS", "cores": cores, "kind": "reference",
            "sample": f"{desc} = {stats['elapsed_s_median']:.2f} s (median of {stats['runs']} runs after one warm-up run)",
            "min": round(stats["min"], 2), "max": round(stats["max"], 2),
            "spread_pct": round(100.0 * (stats["max"] - stats["min"]) / stats["median"], 2),
            "serial_value": round(serial["median"], 2) if serial else None,
            "serial_sample": ssample.workload() if serial else None}


def bench_reference(args, pkg):
    """--impl reference: the reference's CPU path, all host cores, rank 0 only.  Every bench step is one run of the
    reference on the sample (its own 'Elapsed time'); W warm-up runs, K timed ones."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    nx = args.nx
    cores = min(usable_cores(), 64)
    stats, sample, desc = timed_reference_runs(pkg, nx, cores, runs=args.steps, warmup=max(1, args.warmup))
    if stats is None:
        print(json.dumps({"impl": "reference", "unavailable": desc}))
        return 0
    value = stats["median"]
    config = workload_config(args, args.gpus)
    config["workload"] = (f"bounded CPU sample of the GPU arm's workload: {sample.workload()} (the GPU arm: "
                          f"{config['workload']})")
    config["sampled_nx"], config["sampled_ny"], config["sampled_timesteps"] = sample.nx, sample.rows, sample.iters
    line = {
        "impl": "reference", "metric": "MLUPS", "value": round(value, 2), "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(stats["elapsed_s_median"] * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": round(value, 2), "unit": "MLUPS", "cores": cores, "kind": "reference",
                         "sample": f"{desc} = {stats['elapsed_s_median']:.2f} s (median of {stats['runs']} runs)",
                         "min": round(stats["min"], 2), "max": round(stats["max"], 2),
                         "spread_pct": round(100.0 * (stats["max"] - stats["min"]) / value, 2)},
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


def kernel_source_hash(kernel=None) -> str:
    """sha256 over the CODE of one kernel's sources (comments and white space stripped, so that a comment fix does not
    disown a measurement): profiles/roofline_traffic.json records the hash its ncu captures were taken at.  Kernel 7
    lives in lbm_stepsk.cuh on top of the helpers in lbm_kernels.cuh and the arithmetic in lbm_cell.cuh; the other
    timed kernels (2, 4, 5) are in lbm_kernels.cuh."""
    h = hashlib.sha256()
    files = ["lbm_cell.cuh", "lbm_kernels.cuh"] + (["lbm_stepsk.cuh"] if str(kernel) == "7" else [])
    for name in files:
        text = open(os.path.join(ROOT, "mpilattice-boltzmann_b200", "csrc", name)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)           # block comments
        text = re.sub(r"//[^\n]*", "", text)                        # line comments (no string in these files holds "//")
        h.update(name.encode())
        h.update("".join(text.split()).encode())
    return h.hexdigest()[:16]


def measured_traffic(nx, ny, kernel):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture (profiles/roofline_traffic.json,
    one record per kernel), scaled by the cell count if the capture was taken at another size.  The record carries
    the hash of the kernel sources it was measured on: a different hash means the kernels changed since, and the
    number is withheld (returns (None, reason))."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        top = json.load(open(path))
        rec = top.get("kernels", {}).get(str(kernel), top)
        measured_at = rec.get("source_hash", top.get("source_hash"))
        now = kernel_source_hash(kernel)
        if measured_at != now:
            return None, f"ncu capture was taken at kernel sources {measured_at}, the sources are now {now}: re-capture"
        if rec["nx"] == nx and rec["ny"] == ny:
            return rec["dram_bytes_per_launch"], rec.get("capture")
        return rec["dram_bytes_per_launch"] / (rec["nx"] * rec["ny"]) * nx * ny, rec.get("capture")
    except Exception as e:                              # noqa: BLE001
        return None, f"no usable profiles/roofline_traffic.json ({e})"


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

    def sum_over_ranks(x: float) -> float:
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nx, n = args.nx, world
    ips, K, W = args.timesteps, args.steps, args.warmup
    tuning_keys = ("kernel", "graph_steps", "ctas_per_sm", "min_ctas", "fused2", "fused_steps", "band_rows", "prefetch_rows",
                   "cache_hint", "fused_deep", "fused_ctas", "fused_k7")
    pingpong_only = ("kernel", "min_ctas", "fused2", "fused_steps", "band_rows", "prefetch_rows", "fused_deep", "fused_ctas", "fused_k7")

    def make_sim(obstacles, fmt, rows, first, ny_global, ranks=n, inv=None):
        """One slab per rank on the CUDA IPC ring (ranks > 1) or the whole grid on this GPU, tuned as asked."""
        if inv is None:
            inv = float(np.float32(1.0) / np.float32(nx * (ny_global - 2)))
        if ranks == 1:
            sim = pkg.Simulation(nx, rows, DENSITY, ACCEL, OMEGA, obstacles, device=local, inplace=args.inplace,
                                 obstacles_format=fmt)
        else:
            sim = pkg.Simulation.slab(nx, ny_global, first, rows, rank, ranks, DENSITY, ACCEL, OMEGA, inv, obstacles,
                                      device=local, inplace=args.inplace, obstacles_format=fmt)
            blobs = [None] * ranks
            dist.all_gather_object(blobs, sim.export_ipc())
            sim.connect_ipc(blobs[(rank - 1) % ranks], blobs[(rank + 1) % ranks])
            dist.barrier()
        for key in tuning_keys:
            v = getattr(args, key)
            if v is not None and not (args.inplace and key in pingpong_only):
                sim.set_option(key, v)
        return sim

    def channel_bits(rows, first, ny_global):
        """This rank's rows of the channel deck, bit-packed in pinned host memory: walls on global rows 0 and ny-1."""
        words = (nx + 31) // 32
        t = torch.zeros((rows, words), dtype=torch.int32).pin_memory()
        a = t.numpy().view(np.uint32)
        if first == 0:
            a[0, :] = 0xFFFFFFFF
        if first + rows == ny_global:
            a[rows - 1, :] = 0xFFFFFFFF
        return t, a

    # ---- the timed configuration ----------------------------------------------------------
    if args.scaling == "strong":
        split_rows, split_first = pkg.decompose(args.ny, n)
        rows, first, ny_global = int(split_rows[rank]), int(split_first[rank]), args.ny
    else:
        rows, ny_global, first = args.ny, args.ny * n, rank * args.ny
    obstacles_t, obstacles = channel_bits(rows, first, ny_global)
    sim = make_sim(obstacles, "bits", rows, first, ny_global)
    kernel = sim.get_option("kernel")
    fused_steps = sim.get_option("fused_steps") if kernel in (5, 7) else 1
    kernel_name = {1: "step_scalar", 2: "step_vec4", 3: "steps_resident", 4: "step_inplace",
                   5: "steps2_strip (two timesteps per pass)",
                   7: f"steps_strip<{fused_steps}> ({fused_steps} timesteps per pass)"}.get(kernel, str(kernel))

    # ---- parity: the same path, a small case, bit for bit, before anything is timed -------
    parity = {"checked": False, "reason": "--no-parity"}
    if not args.no_parity:
        par = pkg.parity
        if nx % par.PERIOD or n not in par.RANK_COUNTS:
            parity = {"checked": False, "reason": f"the parity case needs nx % {par.PERIOD} == 0 and N in {par.RANK_COUNTS}"}
        else:
            p_ny = par.ROWS_PER_RANK * n
            p_first = par.ROWS_PER_RANK * rank
            p_ob = par.obstacles(nx, p_ny, n)
            p_inv = float(pkg.decks.free_cells_inv(nx * p_ny - int(p_ob.sum())))
            psim = make_sim(np.ascontiguousarray(p_ob[p_first:p_first + par.ROWS_PER_RANK]), "int32", par.ROWS_PER_RANK,
                            p_first, p_ny, inv=p_inv)
            if kernel in (5, 7):
                psim.set_option("fused2", 1)             # the timed kernel (automatic only from 2^22 cells per GPU)
                psim.set_option("fused_steps", fused_steps)   # ... with the timed number of steps per pass
                if distributed:
                    dist.barrier()                       # every rank has switched before any rank runs
            p_kernel = psim.get_option("kernel")
            for it in par.RUNS:                          # back to back, no host synchronisation in between
                psim.enqueue(it)
            psim.sync()
            barrier()
            p_av = psim.fetch_av_vels(par.RUNS[-1]).astype(np.float64)
            if distributed:
                t = torch.from_numpy(p_av).cuda()
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                p_av = t.cpu().numpy()
            bad = par.compare_slab(psim.get_cells(), p_first, n)
            bad_total = int(sum_over_ranks(float(bad)))
            _, av_want = par.expected(n)
            av_want = av_want[-par.RUNS[-1]:].astype(np.float64)
            av_rel = float(np.max(np.abs(p_av - av_want) / np.abs(av_want)))
            barrier()
            psim.close()
            parity = {"checked": True, "ok": bad_total == 0 and av_rel < 1e-4,
                      "cells": nx * p_ny, "values_compared": nx * p_ny * 9, "values_differing": bad_total,
                      "timesteps": par.STEPS, "runs": list(par.RUNS), "kernel": int(p_kernel),
                      "av_vels_max_rel_diff": av_rel, "av_vels_tolerance": 1e-4,
                      "what": f"{n} slab(s) of {nx}x{par.ROWS_PER_RANK} through Simulation.slab + CUDA IPC ring (N > 1), every "
                              "population compared bit for bit with tests/golden/ring_parity.npz (oracle, tiled 64-periodic)"}
            if not parity["ok"]:
                if rank == 0:
                    print(json.dumps({"metric": "MLUPS", "value": None, "n_gpus": n, "parity": parity}))
                    sys.stderr.write(f"bench.py: PARITY FAILED: {bad_total} of {nx * p_ny * 9} populations differ, "
                                     f"av_vels rel diff {av_rel:.3e}\n")
                if distributed:
                    dist.barrier()
                    dist.destroy_process_group()
                return 3

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
    # the dominant kernel: one launch per timestep, or (kernels 5 and 7) one per 2..4 timesteps; the few shorter
    # passes that finish a 256-step chunk are averaged in
    steps_per_launch = fused_steps
    step_launches = timesteps // steps_per_launch
    launch_ms = device_ms * steps_per_launch / timesteps
    peak, peak_src = peak_hbm()
    achieved = BYTES_PER_CELL_STEP * steps_per_launch * nx * rows / (launch_ms * 1e-3) / 1e9     # per GPU, per launch
    traffic, traffic_note = measured_traffic(nx, rows, kernel)

    # ---- e2e: whole jobs of the full configuration through the C-ABI from host buffers ----
    e2e = None
    if not args.no_e2e:
        job_steps = args.e2e_timesteps
        fields_t = torch.empty((4, rows, nx), dtype=torch.float32).pin_memory()
        av_t = torch.empty(job_steps, dtype=torch.float32).pin_memory()
        jobs, pressure_ok = [], True
        for _ in range(args.e2e_jobs):
            barrier()
            t0 = time.perf_counter()
            sim = make_sim(obstacles, "bits", rows, first, ny_global)   # H2D: this rank's obstacle rows, one bit per cell
            t1 = time.perf_counter()
            sim.enqueue(job_steps)
            sim.fetch_av_vels(job_steps, av_t.numpy())    # D2H: per-step averages (synchronises)
            if args.inplace and distributed:
                dist.barrier()                            # in place, edge-row populations may live in the neighbours' buffers
            t2 = time.perf_counter()
            sim.final_state(fields_t.numpy())             # D2H: u_x, u_y, |u|, pressure
            t3 = time.perf_counter()
            barrier()
            job_s = max_over_ranks(time.perf_counter() - t0)
            pressure_ok = pressure_ok and bool(torch.isfinite(fields_t[3]).all())
            sim.close()
            jobs.append({"job_s": round(job_s, 4), "create_s": round(t1 - t0, 4), "run_and_av_vels_s": round(t2 - t1, 4),
                         "final_state_s": round(t3 - t2, 4), "mlups": round(cells_global * job_steps / job_s / 1e6, 1)})
        if not pressure_ok:
            sys.exit("bench.py: e2e run produced a non-finite pressure field")
        h2d = int(obstacles.nbytes)
        d2h = 4 * 4 * nx * rows + 4 * job_steps
        job_bench_steps = max(1, job_steps // ips)
        e2e = {"value": round(statistics.median(j["mlups"] for j in jobs), 1), "unit": "MLUPS",
               "h2d_bytes_per_step": h2d // job_bench_steps, "d2h_bytes_per_step": d2h // job_bench_steps,
               "timesteps_per_job": job_steps, "h2d_bytes_per_job": h2d, "d2h_bytes_per_job": d2h,
               "what": "whole jobs of the full configuration: lbm_b200_create_[slab_]ex (bit-packed obstacle rows from pinned "
                       f"host memory) + enqueue of {job_steps} timesteps + fetch_av_vels + get_final_state into pinned host "
                       "memory; wall clock, max over ranks; `value` = median over the jobs listed (rank 0's phases)",
               "jobs": jobs}

    # ---- strong scaling: the 16384 x 16384 grid split over the N GPUs ----------------------
    strong = None
    if distributed and args.scaling == "weak" and not args.no_strong:
        s_rows_all, s_first_all = pkg.decompose(args.ny, n)
        s_rows, s_first = int(s_rows_all[rank]), int(s_first_all[rank])
        _, s_ob = channel_bits(s_rows, s_first, args.ny)
        ssim = make_sim(s_ob, "bits", s_rows, s_first, args.ny)
        for _ in range(W):
            ssim.enqueue(ips)
        ssim.sync()
        barrier()
        ssim.enqueue(K * ips)
        ssim.sync()
        barrier()
        s_ms = max_over_ranks(ssim.elapsed_ms())
        s_kernel = ssim.get_option("kernel")
        ssim.close()
        barrier()
        one_ms = None
        if rank == 0:                                  # the same grid on ONE GPU, same box, same run
            _, o_ob = channel_bits(args.ny, 0, args.ny)
            osim = make_sim(o_ob, "bits", args.ny, 0, args.ny, ranks=1)
            for _ in range(W):
                osim.enqueue(ips)
            osim.sync()
            osim.enqueue(K * ips)
            osim.sync()
            one_ms = osim.elapsed_ms()
            osim.close()
        barrier()
        s_value = float(nx) * args.ny * timesteps / (s_ms * 1e-3) / 1e6
        strong = {"value": round(s_value, 1), "unit": "MLUPS", "ms_per_step": round(s_ms / K, 4), "kernel": int(s_kernel),
                  "workload": f"synthetic channel {nx}x{args.ny} in total, {n} row slabs of {s_rows} rows (lbm_b200_decompose)"}
        if one_ms:
            one_value = float(nx) * args.ny * timesteps / (one_ms * 1e-3) / 1e6
            strong.update({"n1_value": round(one_value, 1), "n1_ms_per_step": round(one_ms / K, 4),
                           "efficiency_vs_n1": round(s_value / (n * one_value), 4),
                           "efficiency_what": "T1 / (N * TN), T1 measured by rank 0 alone on the same box right after the N-GPU run"})

    baseline = cpu_baseline(pkg, nx) if (rank == 0 and n == 1 and not args.no_cpu_baseline) else None

    if rank == 0:
        line = {
            "metric": "MLUPS", "value": round(value, 1), "unit": "MLUPS", "n_gpus": n, "steps": K, "warmup": W,
            "ms_per_step": round(device_ms / K, 4), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, n),
            "kernel": kernel_name,
            "wall_ms_per_step": round(wall_s * 1e3 / K, 4),
            "av_vels_last": float(av[-1]),
            "parity": parity,
            "gpu_launches": int(launches) * n,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_note,
                         "peak_source": peak_src,
                         "what": (f"{kernel_name}, {BYTES_PER_CELL_STEP:.0f} B/cell/step x {steps_per_launch} timestep(s) x "
                                  f"{nx * rows} cells per launch per GPU, mean launch {launch_ms * 1e3:.1f} us over "
                                  f"{step_launches} launches"
                                  + (f"; {fused_steps} timesteps are fused per pass over HBM, so the DRAM traffic per launch "
                                     f"(`traffic`) is about 1/{fused_steps} of the algorithmic bytes and `frac` exceeds the one-step "
                                     "streaming roofline; `dram_frac` = traffic / launch time / peak is the share of the HBM "
                                     "roof actually used" if kernel in (5, 7) else "")),
                         "dram_gbs": round(traffic / (launch_ms * 1e-3) / 1e9, 1) if traffic else None,
                         "dram_frac": round(traffic / (launch_ms * 1e-3) / 1e9 / peak, 4) if traffic else None},
            "clocks": clocks,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if strong is not None:
            line["strong"] = strong
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
    ap.add_argument("--fused-steps", dest="fused_steps", type=int, default=None, choices=[0, 2, 3, 4],
                    help="timesteps per pass over HBM of the fused kernel (2 = kernel 5, 3 or 4 = kernel 7)")
    ap.add_argument("--fused-k7", dest="fused_k7", type=int, default=None, help="1 = kernel 7 also for two timesteps per pass")
    ap.add_argument("--fused-ctas", dest="fused_ctas", type=int, default=None, help="kernel 7: CTAs per SM (0 = automatic)")
    ap.add_argument("--band-rows", dest="band_rows", type=int, default=None)
    ap.add_argument("--prefetch-rows", dest="prefetch_rows", type=int, default=None, help="kernel 5: L2 prefetch distance in rows")
    ap.add_argument("--cache-hint", dest="cache_hint", type=int, default=None)
    ap.add_argument("--fused-deep", dest="fused_deep", type=int, default=None, help="kernel 5 on one GPU: two staging rows, 12 warps per SM")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the bit-exact parity case before the timed region")
    ap.add_argument("--no-e2e", action="store_true", help="skip the whole-job (host buffers in, host buffers out) measurement")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling leg")
    ap.add_argument("--e2e-timesteps", dest="e2e_timesteps", type=int, default=FULL_JOB_TIMESTEPS,
                    help="timesteps of an e2e job (default: the full configuration, 10000)")
    ap.add_argument("--e2e-jobs", dest="e2e_jobs", type=int, default=2)
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
