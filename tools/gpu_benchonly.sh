#!/bin/bash
# bench only: tools/gpu_benchonly.sh <tag> <steps> <warmup> [extra bench flags]
OUT=gpurun_out/${1:-benchonly}; mkdir -p $OUT
S=${2:-20}; W=${3:-5}; shift 3
DCORA_SETUP_TIMING=1 timeout 1200 python bench.py --steps $S --warmup $W "$@" > $OUT/bench.json 2> $OUT/bench.log; tail -25 $OUT/bench.log | cut -c1-250; wc -l $OUT/bench.json; cut -c1-300 $OUT/bench.json
