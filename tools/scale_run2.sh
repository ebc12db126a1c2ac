#!/bin/bash
# Shorter 8-GPU confirmation: multi-GPU parity, weak N=8, strong N=4 and 8, 1024^2 deck split over 8.
set -u
OUT=${1:-gpurun_out/scale2}
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3 | tee "$OUT/pytest_multi.log"
PORT=29700
run() { PORT=$((PORT + 1)); python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus "$1" --steps 5 --warmup 3 --scaling "$2" --no-cpu-baseline 2>> "$OUT/torchrun.err"; }
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline | tee -a "$OUT/weak.jsonl"
run 8 weak | tee -a "$OUT/weak.jsonl"
run 4 strong | tee -a "$OUT/strong.jsonl"
run 8 strong | tee -a "$OUT/strong.jsonl"
D=tests/golden/decks
W=$(mktemp -d)
( cd "$W" && LBM_GPUS=8 LBM_VERBOSE=1 "$OLDPWD/mpilattice-boltzmann_b200/bin/d2q9-bgk" "$OLDPWD/$D/input_1024x1024.params" "$OLDPWD/$D/obstacles_1024x1024.dat" > run.out 2> run.err
  echo "gpus=8 $(grep 'Elapsed time' run.out) $(grep Reynolds run.out) sha256=$(sha256sum final_state.dat | cut -c1-16) $(cat run.err)" ) | tee -a "$OUT/deck1024_split.log"
