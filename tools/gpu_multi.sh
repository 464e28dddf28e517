#!/bin/bash
# N-GPU bench as the driver launches it (+ the reference arm): tools/gpu_multi.sh <N> [steps] [warmup] [extra flags]
N=${1:-4}; S=${2:-20}; W=${3:-5}; shift 3
OUT=gpurun_out/n$N; mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt
DCORA_SETUP_TIMING=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps $S --warmup $W "$@" > $OUT/bench.json 2> $OUT/bench.log
grep -v "dcora setup" $OUT/bench.log | tail -8 | cut -c1-300; grep "dcora setup" $OUT/bench.log | sort | uniq -c | sort -rn | head -4; cut -c1-300 $OUT/bench.json
