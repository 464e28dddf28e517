#!/bin/bash
N=${1:-4}
OUT=gpurun_out/n$N; mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --ttc-grid 0 > $OUT/bench.json 2> $OUT/bench.log
tail -4 $OUT/bench.log | cut -c1-300; cut -c1-300 $OUT/bench.json
