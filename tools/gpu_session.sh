#!/bin/bash
# One single-GPU measurement session on a B200 box (run under gpurun):
#   [NCU_FLAGS="--fused-steps 4"] [NCU_KERNEL=steps_strip] tools/gpu_session.sh <out-dir> [stages...]     stages: smoke tests bench benchq ref onestep inplace launches ncu ncu_onestep decks
# Each stage writes its own log under <out-dir>; a failing stage does not stop the later ones.
set -u
OUT=${1:-gpurun_out/session}; shift || true
STAGES=${*:-smoke tests bench onestep inplace launches ncu}
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
NCU_CMD="python bench.py --steps 3 --warmup 3 --timesteps 12 --no-cpu-baseline --no-parity --no-e2e ${NCU_FLAGS:-}"
for st in $STAGES; do
  case $st in
    smoke)    python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -3 | tee "$OUT/smoke.log" ;;
    tests)    python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee "$OUT/pytest_gpu.log" ;;
    bench)    python bench.py 2> "$OUT/bench.err" | tee "$OUT/bench.json" ;;
    benchq)   python bench.py --no-cpu-baseline 2> "$OUT/bench.err" | tee "$OUT/bench.json" ;;
    ref)      python bench.py --impl reference --steps 3 --warmup 1 2> "$OUT/bench_ref.err" | tee "$OUT/bench_ref.json" ;;
    onestep)  python bench.py --fused2 0 --no-cpu-baseline --steps 5 2> "$OUT/bench_onestep.err" | tee "$OUT/bench_onestep.json" ;;
    inplace)  python bench.py --inplace --no-cpu-baseline --steps 5 2> "$OUT/bench_inplace.err" | tee "$OUT/bench_inplace.json" ;;
    launches) $NCU_CMD > "$OUT/plain_for_ncu.log" 2>&1 && \
              ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" $NCU_CMD > "$OUT/ncu_launches.log" 2>&1
              tail -2 "$OUT/ncu_launches.log" ;;
    ncu)      $NCU_CMD > "$OUT/plain_for_ncu.log" 2>&1 && \
              ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-steps} -s ${NCU_SKIP:-10} -c 1 -o "$OUT/prof" $NCU_CMD > "$OUT/ncu_full.log" 2>&1
              tail -2 "$OUT/ncu_full.log" ;;
    ncu_onestep) $NCU_CMD --fused2 0 > "$OUT/plain_for_ncu_onestep.log" 2>&1 && \
              ncu --set full --clock-control none --import-source on -k regex:step_vec4 -s 10 -c 1 -o "$OUT/prof_onestep" $NCU_CMD --fused2 0 > "$OUT/ncu_full_onestep.log" 2>&1
              tail -2 "$OUT/ncu_full_onestep.log" ;;
    decks)    tools/deck_times.sh 2>&1 | tee "$OUT/decks.log" ;;
    *) echo "unknown stage $st" ;;
  esac
done
