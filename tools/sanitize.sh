#!/bin/bash
# compute-sanitizer over the reduced case set; logs land in gpurun_out/ (copy the summaries to profiles/).
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > $OUT/sanitizer_${tool}_$TAG.txt 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Toeplitz" $OUT/sanitizer_${tool}_$TAG.txt | tail -4
done
