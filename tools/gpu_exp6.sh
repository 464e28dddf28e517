#!/bin/bash
OUT=gpurun_out/exp6; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
for v in 0 1; do
  DCORA_SPMM_PF=$v timeout 600 python tools/kprofile.py --rbcd-steps 4 spmm 2>$OUT/k$v.log >$OUT/k$v.json; echo "PF=$v"; cat $OUT/k$v.json; grep rbcd $OUT/k$v.log | tail -2
done
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.log; cat $OUT/bench.json
