#!/bin/bash
# parity tests + bench (no profiling)
OUT=gpurun_out/check; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.log; tail -4 $OUT/bench.log; cut -c1-300 $OUT/bench.json
