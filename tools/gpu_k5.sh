#!/bin/bash
# K5 check: parity of the preconditioner path + slot trace + full-batch timing + RBCD steps
OUT=gpurun_out/${1:-k5}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_problem.py tests/test_gpu_solver.py tests/test_gpu_rbcd.py -q -x > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
python tools/trsv_trace.py > $OUT/trace.txt 2>&1; grep "total" $OUT/trace.txt
timeout 600 python tools/kprofile.py --rbcd-steps 6 precond spmm 2>$OUT/k.log >$OUT/k.json; grep rbcd $OUT/k.log | tail -3; cat $OUT/k.json
