#!/bin/bash
OUT=gpurun_out/exp8; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_problem.py tests/test_gpu_solver.py -x -q > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
for v in 0 2; do
  DCORA_SPMM_XV=$v timeout 600 python tools/kprofile.py --rbcd-steps 4 spmm 2>$OUT/k$v.log >$OUT/k$v.json; echo "XV=$v"; cat $OUT/k$v.json; grep rbcd $OUT/k$v.log | tail -2
done
DCORA_SPMM_XV=2 timeout 600 python -m pytest tests/test_gpu_problem.py -x -q 2>&1 | tail -2
timeout 600 python tools/ttc.py --grid 20 > $OUT/ttc20.json 2> $OUT/ttc20.log; cut -c1-900 $OUT/ttc20.json; tail -3 $OUT/ttc20.log
