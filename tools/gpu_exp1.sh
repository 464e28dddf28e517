#!/bin/bash
OUT=gpurun_out/exp1; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_problem.py tests/test_gpu_rbcd.py -x -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
for v in 5 6 8; do
  DCORA_SPMM_MINB=$v timeout 600 python tools/kprofile.py --rbcd-steps 3 spmm 2>$OUT/k$v.log >$OUT/k$v.json; echo "MINB=$v"; cat $OUT/k$v.json; grep rbcd $OUT/k$v.log | tail -1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bsr_spmm_kernel|trsv_tape_kernel' \
  --launch-skip 1 -c 6 -o $OUT/full_batch python tools/kprofile.py --iters 2 spmm precond > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
