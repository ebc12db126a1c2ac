#!/bin/bash
# 8-GPU confirmation after fused2 (two timesteps per pass) became the default on rings too:
# multi-GPU parity, weak N=8 (fused and one-step), strong N=8, the shipped decks split over 8.
set -u
OUT=${1:-gpurun_out/scale3}
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3 | tee "$OUT/pytest_multi.log"
PORT=29800
run() { PORT=$((PORT + 1)); n=$1; sc=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus "$n" --steps 3 --warmup 3 --scaling "$sc" --no-cpu-baseline "$@" 2>> "$OUT/torchrun.err"; }
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline | tee -a "$OUT/weak.jsonl"
run 8 weak | tee -a "$OUT/weak.jsonl"
run 8 strong | tee -a "$OUT/strong.jsonl"
run 8 weak --fused2 0 | tee -a "$OUT/weak_onestep.jsonl"
tools/deck_times.sh LBM_GPUS=8 | tee -a "$OUT/decks_8gpu.log"
