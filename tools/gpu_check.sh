#!/bin/bash
# full GPU parity suite + smoke + bench (as the driver runs it)
OUT=gpurun_out/${1:-check}; mkdir -p $OUT
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1; tail -4 $OUT/smoke.log | cut -c1-300
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest.log 2>&1; tail -6 $OUT/pytest.log | cut -c1-250
DCORA_SETUP_TIMING=1 timeout 900 python bench.py --steps ${2:-20} --warmup ${3:-5} > $OUT/bench.json 2> $OUT/bench.log; tail -22 $OUT/bench.log | cut -c1-250; cut -c1-300 $OUT/bench.json
