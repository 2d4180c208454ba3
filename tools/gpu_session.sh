#!/bin/bash
# Development aid: one gpurun call = GPU tests + bench + ncu launch list + ncu full capture of the QPHB kernel.
# usage (on the GPU box, from the repo root): bash tools/gpu_session.sh <tag> [skip_tests]
TAG=${1:-dev}
OUT=gpurun_out
mkdir -p $OUT
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_$TAG.log
  tail -5 $OUT/pytest_$TAG.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
cat $OUT/bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qphb -s 1 -c 1 -o $OUT/prof_$TAG -f \
  python bench.py --batch 2368 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full exit $?"
ls -la $OUT
# secondary kernels: one full capture each (matrix builder, chrono filter)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:impedance_interp_kernel -s 3 -c 1 -o $OUT/prof_interp_$TAG -f \
  python tools/interp_bench.py > $OUT/ncu_interp_$TAG.log 2>&1; echo "ncu interp exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_gather -s 4 -c 1 -o $OUT/prof_filter_$TAG -f \
  python bench.py --batch 512 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_filter_$TAG.log 2>&1; echo "ncu filter exit $?"
