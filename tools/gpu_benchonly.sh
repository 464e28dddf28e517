#!/bin/bash
OUT=gpurun_out/benchonly; mkdir -p $OUT
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.log; tail -3 $OUT/bench.log | cut -c1-200; wc -l $OUT/bench.json; cut -c1-200 $OUT/bench.json
