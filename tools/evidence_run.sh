#!/bin/bash
# Single-GPU evidence pass: GPU test suite, both bench arms, ncu launch list + one full capture of the dominant kernel.
set -u
OUT=${1:-gpurun_out/evidence}
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee "$OUT/pytest_gpu.log"
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -3 | tee "$OUT/smoke.log"
python bench.py --impl reference --steps 3 --warmup 1 2> "$OUT/bench_ref.err" | tee "$OUT/bench_ref.json"
python bench.py 2> "$OUT/bench.err" | tee "$OUT/bench.json"
python bench.py --fused2 0 --no-cpu-baseline 2> "$OUT/bench_onestep.err" | tee "$OUT/bench_onestep.json"
python bench.py --inplace --no-cpu-baseline --steps 5 2> "$OUT/bench_inplace.err" | tee "$OUT/bench_inplace.json"
CMD="python bench.py --steps 3 --warmup 3 --timesteps 10 --no-cpu-baseline"
$CMD > "$OUT/plain_for_ncu.log" 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench_16384.csv" $CMD > "$OUT/ncu_launches.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:steps2_strip -s 10 -c 1 -o "$OUT/prof_steps2_strip" $CMD > "$OUT/ncu_full.log" 2>&1
tail -2 "$OUT/ncu_full.log"
