#!/bin/bash
# Full 1/2/4/8-GPU table with the final round-1 code (not run in round 1: no box time left after the edge-band split).
#   gpurun --gpus 8 --timeout 900 -- tools/scale_run4.sh gpurun_out/scale4
set -u
OUT=${1:-gpurun_out/scale4}
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
PORT=29900
run() { PORT=$((PORT + 1)); n=$1; sc=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus "$n" --steps 3 --warmup 3 --scaling "$sc" --no-cpu-baseline "$@" 2>> "$OUT/torchrun.err"; }
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline | tee -a "$OUT/weak.jsonl"
for n in 2 4 8; do run $n weak | tee -a "$OUT/weak.jsonl"; done
for n in 2 4 8; do run $n strong | tee -a "$OUT/strong.jsonl"; done
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3 | tee "$OUT/pytest_multi.log"
