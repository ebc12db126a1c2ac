#!/bin/bash
# Scaling sweep on an N-GPU box: weak (16384x16384 per GPU) and strong (16384x16384 total) at 1/2/4/8 GPUs,
# the multi-GPU parity tests, and the 1024x1024 deck split over 2/4/8 GPUs through the CLI.
# usage: tools/scale_run.sh <outdir> [max_gpus]
set -u
OUT=${1:-gpurun_out/scale}
MAX=${2:-8}
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
PORT=29600
run_bench() {  # n scaling extra...
  local n=$1 scaling=$2; shift 2
  PORT=$((PORT + 1))
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 --scaling "$scaling" --no-cpu-baseline "$@"
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus "$n" --steps 5 --warmup 3 --scaling "$scaling" --no-cpu-baseline "$@" 2>> "$OUT/torchrun.err"
  fi
}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpus.csv"
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5 | tee "$OUT/pytest_multi.log"
for n in 1 2 4 8; do
  [ "$n" -le "$MAX" ] || continue
  run_bench "$n" weak | tee -a "$OUT/weak.jsonl"
done
for n in 2 4 8; do
  [ "$n" -le "$MAX" ] || continue
  run_bench "$n" strong | tee -a "$OUT/strong.jsonl"
done
# BASELINE.json configs[3]: the 1024x1024 deck row-slab split over 1/2/4/8 B200 through the drop-in CLI
D=tests/golden/decks
for n in 1 2 4 8; do
  [ "$n" -le "$MAX" ] || continue
  W=$(mktemp -d)
  ( cd "$W" && LBM_GPUS=$n LBM_VERBOSE=1 "$OLDPWD/mpilattice-boltzmann_b200/bin/d2q9-bgk" "$OLDPWD/$D/input_1024x1024.params" "$OLDPWD/$D/obstacles_1024x1024.dat" > run.out 2> run.err
    echo "gpus=$n $(grep 'Elapsed time' run.out) $(grep Reynolds run.out) sha256=$(sha256sum final_state.dat | cut -c1-16) $(cat run.err)" )  | tee -a "$OUT/deck1024_split.log"
  rm -rf "$W"
done
grep -o '"final_state_sha256": "[0-9a-f]\{16\}' tests/golden/ref_strict.json | head -1 >> "$OUT/deck1024_split.log"
