#!/bin/bash
OUT=gpurun_out/exp17; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -6 $OUT/pytest.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.log; tail -4 $OUT/bench.log; cut -c1-400 $OUT/bench.json
