#!/bin/bash
# Round-2 diagnostics: throughput vs resident CTAs per SM, per-phase clock64 profile.
OUT=gpurun_out
mkdir -p $OUT
for occ in 1 2 3; do
  HDRT_DEBUG_OCC=$occ timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/occ_$occ.json 2> $OUT/occ_$occ.err
  python - <<PY
import json
try:
    d = json.loads(open('$OUT/occ_$occ.json').read().strip().splitlines()[-1])
    print('occ', $occ, 'fits/s', d['value'], 'frac', d['roofline']['frac'])
except Exception as e:
    print('occ', $occ, 'failed', e)
PY
done
timeout 600 python tools/phase_profile.py 1332 > $OUT/phase_r2_base.txt 2>&1
cat $OUT/phase_r2_base.txt
