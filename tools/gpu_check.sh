#!/bin/bash
# full GPU parity suite + K5 slot trace + bench
OUT=gpurun_out/${1:-check}; mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
python tools/trsv_trace.py > $OUT/trace.txt 2>&1; grep "total" $OUT/trace.txt
timeout 900 python bench.py --steps ${2:-10} --warmup ${3:-3} > $OUT/bench.json 2> $OUT/bench.log; tail -12 $OUT/bench.log; cut -c1-600 $OUT/bench.json
