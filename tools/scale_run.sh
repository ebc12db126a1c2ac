#!/bin/bash
# Scaling table on ONE box with as many GPUs as gpurun gave it (the N it does not have are skipped):
#   gpurun --gpus 8 --timeout 1500 -- tools/scale_run.sh <out-dir> [legs...]
#   legs (default: weak tests):
#     weak     bench.py at N = 1, 2, 4, 8 -- every N > 1 line carries the parity check, the e2e jobs and the strong leg
#     tests    tests/test_gpu_multi.py (cross-device parity, the C launcher, the halo-wait timeout)
#     decks    the 1024x1024 deck through bin/d2q9-bgk-mp -np N: sha256 of final_state.dat must not depend on N
#   STEPS / WARMUP / EXTRA (flags for bench.py) from the environment.
set -u
OUT=${1:-gpurun_out/scale}; shift || true
LEGS=${*:-weak tests}
STEPS=${STEPS:-5}; WARMUP=${WARMUP:-3}; EXTRA=${EXTRA:---e2e-timesteps 2000 --e2e-jobs 1}
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
NGPU=$(nvidia-smi -L | wc -l)
PORT=29900
for leg in $LEGS; do
  case $leg in
    weak)
      python bench.py --gpus 1 --steps "$STEPS" --warmup "$WARMUP" --no-cpu-baseline $EXTRA 2>> "$OUT/bench.err" | tee -a "$OUT/weak.jsonl"
      for n in 2 4 8; do
        [ "$n" -le "$NGPU" ] || continue
        PORT=$((PORT + 1))
        python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $PORT \
          bench.py --gpus "$n" --steps "$STEPS" --warmup "$WARMUP" $EXTRA 2>> "$OUT/bench.err" | tee -a "$OUT/weak.jsonl"
      done ;;
    tests)
      python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 | tee "$OUT/pytest_multi.log" ;;
    decks)
      D=$PWD/tests/golden/decks
      for n in 1 2 4 8; do
        [ "$n" -le "$NGPU" ] || continue
        W=$(mktemp -d)
        ( cd "$W" && LBM_VERBOSE=1 "$OLDPWD/mpilattice-boltzmann_b200/bin/d2q9-bgk-mp" -np "$n" "$D/input_1024x1024.params" "$D/obstacles_1024x1024.dat" > run.out 2> run.err
          echo "1024x1024 -np $n $(grep 'Elapsed time' run.out | tr -s '\t' ' ') $(grep Reynolds run.out | tr -s '\t' ' ') sha256 $(sha256sum final_state.dat | cut -c1-16) $(cat run.err)" ) | tee -a "$OUT/decks.log"
        rm -rf "$W"
      done ;;
  esac
done
