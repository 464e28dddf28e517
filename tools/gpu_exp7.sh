#!/bin/bash
OUT=gpurun_out/exp7; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.log; cat $OUT/bench.json
DCORA_TIMING=1 timeout 600 python tools/kprofile.py --rbcd-steps 2 --split spmm 2>$OUT/kt.log >/dev/null; grep "dcora timing" $OUT/kt.log | tail -3; grep "colour\|monitor" $OUT/kt.log | tail -10
