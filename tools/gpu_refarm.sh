#!/bin/bash
# the reference arm (CPU, all host cores) as the driver runs it
OUT=gpurun_out/${1:-refarm}; mkdir -p $OUT
nproc > $OUT/nproc.txt
( time timeout 900 python bench.py --impl reference --steps ${2:-20} --warmup ${3:-5} > $OUT/bench_ref.json 2> $OUT/bench_ref.log ); tail -6 $OUT/bench_ref.log | cut -c1-200; cut -c1-700 $OUT/bench_ref.json; cat $OUT/nproc.txt
