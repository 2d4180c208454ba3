#!/bin/bash
# Round-2 profile set: bench line, launch list, ncu --set full of the fit kernels (C2 warp kernel, C3 hybrid, C4 DOP).
# The summaries are produced on the box (gpurun brings back at most 64 MiB: the reports themselves stay there, except C2's);
# everything lands in gpurun_out/ and is copied to profiles/ afterwards.
OUT=gpurun_out
LIB=$PWD/hybrid-drt_b200/_lib/libhybdrt_b200.so
mkdir -p $OUT
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_r02.json 2> $OUT/bench_r02.err; echo "bench exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r02.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/launches_r02.log 2>&1; echo "launch list exit $?"
for cfg in "c2 5920 qphb_warp" "c3 592 qphb" "c4 296 qphb"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s 2 -c 1 -o $OUT/prof_r02_$1 -f \
    python tools/config_run.py $1 $2 > $OUT/ncu_r02_$1.log 2>&1; echo "ncu $1 exit $?"
  python tools/ncu_summary.py $OUT/prof_r02_$1.ncu-rep "round 2, $1" > $OUT/ncu_qphb_r02_$1.txt 2>&1
done
python tools/ncu_functions.py $OUT/prof_r02_c2.ncu-rep $LIB > $OUT/ncu_functions_r02_c2.txt 2>&1
python tools/ncu_lines.py $OUT/prof_r02_c2.ncu-rep qphb_warp_kernel $LIB > $OUT/ncu_lines_r02_c2.txt 2>&1
python tools/ncu_lines.py $OUT/prof_r02_c3.ncu-rep qphb_kernelINS_3CfgILi13ELi2ELi2ELb0 $LIB > $OUT/ncu_lines_r02_c3.txt 2>&1
python tools/ncu_lines.py $OUT/prof_r02_c4.ncu-rep qphb_kernelINS_3CfgILi20ELi3ELi1ELb0 $LIB > $OUT/ncu_lines_r02_c4.txt 2>&1
rm -f $OUT/prof_r02_c3.ncu-rep $OUT/prof_r02_c4.ncu-rep
for c in c2 c3 c4; do timeout 300 python tools/phase_profile.py --nobuild $c 592 > $OUT/phase_r02_$c.txt 2>&1; done
ls -la $OUT
