#!/bin/bash
# A/B sweep of kernel-5 knobs on one B200: tools/sweep_r02.sh <out-dir> "<flag set 1>" "<flag set 2>" ...
set -u
OUT=$1; shift
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
for flags in "$@"; do
  echo "== $flags" | tee -a "$OUT/sweep.log"
  python bench.py --no-cpu-baseline --no-parity --no-e2e --steps 5 $flags 2>> "$OUT/sweep.err" | FLAGS="$flags" python -c "
import json,os,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(json.dumps({'flags': os.environ['FLAGS'], 'value': d['value'], 'ms_per_step': d['ms_per_step'], 'clocks': d.get('clocks')}))
" | tee -a "$OUT/sweep.jsonl"
done
