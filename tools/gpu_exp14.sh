#!/bin/bash
OUT=gpurun_out/exp14; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.log; tail -4 $OUT/bench.log; cut -c1-300 $OUT/bench.json
timeout 900 python tools/timeline.py --steps 2 > $OUT/timeline.txt 2> $OUT/timeline.log; head -3 $OUT/timeline.txt; grep -A12 "device time by" $OUT/timeline.txt
