#!/bin/bash
OUT=gpurun_out/exp4; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_problem.py tests/test_gpu_rbcd.py -x -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
for v in 1 8 10; do
  DCORA_SPMM_MINB=$v timeout 600 python tools/kprofile.py --rbcd-steps 4 spmm 2>$OUT/k$v.log >$OUT/k$v.json; echo "MINB=$v"; cat $OUT/k$v.json; grep rbcd $OUT/k$v.log | tail -2
done
