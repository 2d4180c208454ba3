#!/bin/bash
# Round-2 profile set: bench line, launch list, ncu --set full of the fit kernels (C2 warp kernel, C3 hybrid, C4 DOP) and
# of the matrix builder.  Everything lands in gpurun_out/; the summaries are copied to profiles/ afterwards.
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_r02.json 2> $OUT/bench_r02.err; echo "bench exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r02.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/launches_r02.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qphb_warp -s 2 -c 1 -o $OUT/prof_r02_c2 -f \
  python tools/config_run.py c2 5920 > $OUT/ncu_r02_c2.log 2>&1; echo "ncu c2 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qphb -s 2 -c 1 -o $OUT/prof_r02_c3 -f \
  python tools/config_run.py c3 592 > $OUT/ncu_r02_c3.log 2>&1; echo "ncu c3 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qphb -s 2 -c 1 -o $OUT/prof_r02_c4 -f \
  python tools/config_run.py c4 296 > $OUT/ncu_r02_c4.log 2>&1; echo "ncu c4 exit $?"
timeout 600 ncu --set full --clock-control none -k regex:impedance_interp -s 3 -c 1 -o $OUT/prof_r02_interp -f \
  python tools/interp_bench.py > $OUT/ncu_r02_interp.log 2>&1; echo "ncu interp exit $?"
ls -la $OUT/prof_r02_*.ncu-rep
