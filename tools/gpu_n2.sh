#!/bin/bash
OUT=gpurun_out/n2; mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_n2.json 2> $OUT/bench_n2.log
tail -5 $OUT/bench_n2.log; cat $OUT/bench_n2.json
