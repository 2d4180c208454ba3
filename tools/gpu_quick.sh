#!/bin/bash
# Development aid: GPU tests (optional), short C2 bench, phase profile.   usage: bash tools/gpu_quick.sh <tag> [notests]
TAG=${1:-dev}
OUT=gpurun_out
mkdir -p $OUT
if [ -z "$2" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_$TAG.log
  tail -15 $OUT/pytest_$TAG.log
fi
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.loads(open('$OUT/bench_$TAG.json').read().strip().splitlines()[-1])
    print('fits/s', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['config']['mean_outer_iters'], d['config']['mean_ipm_iters'])
except Exception as e:
    print('bench parse failed', e); print(open('$OUT/bench_$TAG.err').read()[-2000:])
PY
timeout 600 python tools/phase_profile.py 1332 > $OUT/phase_$TAG.txt 2>&1
cat $OUT/phase_$TAG.txt
