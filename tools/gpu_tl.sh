#!/bin/bash
OUT=gpurun_out/tl; mkdir -p $OUT
timeout 900 python tools/timeline.py --steps 2 > $OUT/timeline.txt 2> $OUT/timeline.log; cat $OUT/timeline.txt; tail -3 $OUT/timeline.log
