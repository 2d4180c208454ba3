#!/bin/bash
# Scaling runs on N GPUs of one box: C2 (weak, no data-path collective) and C5 (strong, NCCL gather timed) + the 2-rank GPU test.
# usage (under gpurun --gpus N): bash tools/gpu_scale.sh N
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-configs > $OUT/scale_c2_n$N.json 2> $OUT/scale_c2_n$N.err; echo "c2 exit $?"
timeout 600 $RUN --master-port 29512 bench.py --config c5 --gpus $N --steps 2 --warmup 1 > $OUT/scale_c5_n$N.json 2> $OUT/scale_c5_n$N.err; echo "c5 exit $?"
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k two_gpus > $OUT/pytest_2gpu.log 2>&1; tail -3 $OUT/pytest_2gpu.log; fi
tail -c 600 $OUT/scale_c2_n$N.json; echo; tail -c 900 $OUT/scale_c5_n$N.json
