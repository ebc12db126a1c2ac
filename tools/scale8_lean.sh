#!/bin/bash
# Lean 8-GPU check (8 x box time is charged): the N = 8 bench line (parity case, e2e job, strong leg) and the
# cross-device tests.   gpurun --gpus 8 --timeout 900 -- tools/scale8_lean.sh <out-dir>
set -u
OUT=${1:-gpurun_out/scale8}; mkdir -p "$OUT"
cd "$(dirname "$0")/.."
N=$(nvidia-smi -L | wc -l)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29931 \
  bench.py --gpus "$N" --steps 5 --warmup 3 --e2e-timesteps 2000 --e2e-jobs 1 2> "$OUT/bench$N.err" | tee "$OUT/bench$N.json"
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6 | tee "$OUT/pytest_multi.log"
D=$PWD/tests/golden/decks
for n in 1 2 4 8; do
  [ "$n" -le "$N" ] || continue
  W=$(mktemp -d)
  ( cd "$W" && LBM_VERBOSE=1 timeout 120 "$OLDPWD/mpilattice-boltzmann_b200/bin/d2q9-bgk-mp" -np "$n" "$D/input_1024x1024.params" "$D/obstacles_1024x1024.dat" > run.out 2> run.err
    echo "1024x1024 -np $n $(grep 'Elapsed time' run.out | tr -s '\t' ' ') $(grep Reynolds run.out | tr -s '\t' ' ') sha256 $(sha256sum final_state.dat | cut -c1-16) $(cat run.err)" ) | tee -a "$OUT/decks.log"
  rm -rf "$W"
done
