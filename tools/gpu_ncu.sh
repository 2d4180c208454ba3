#!/bin/bash
# Development aid: one ncu --set full capture (with source) of the fit kernel on a small C2 batch.  usage: bash tools/gpu_ncu.sh <tag> [batch]
TAG=${1:-dev}
B=${2:-1332}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qphb -s 3 -c 1 -o $OUT/prof_$TAG -f \
  python bench.py --batch $B --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full exit $?"
tail -3 $OUT/ncu_full_$TAG.log
ls -la $OUT/prof_$TAG.ncu-rep
