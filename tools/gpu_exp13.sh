#!/bin/bash
OUT=gpurun_out/exp13; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 900 python tools/timeline.py --steps 2 > $OUT/timeline.txt 2> $OUT/timeline.log; head -3 $OUT/timeline.txt; grep -A8 "device time by" $OUT/timeline.txt
timeout 600 python tools/kprofile.py precond 2>/dev/null
python tools/trsv_trace.py > $OUT/trace.txt 2>&1; head -8 $OUT/trace.txt | cut -c1-400
