#!/bin/bash
OUT=gpurun_out/${1:-last}; mkdir -p $OUT
( time timeout 170 python -m pytest tests/test_gpu_problem.py tests/test_gpu_solver.py tests/test_gpu_rbcd.py -m gpu -x -q ) > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log | cut -c1-200
DCORA_SETUP_TIMING=1 timeout 60 python tools/ttc_multi.py --grid 47 --agents 16 --nesterov-blocks colours --restart-interval auto --check-every 10 > $OUT/ttc.json 2> $OUT/ttc.log; grep "dcora setup" $OUT/ttc.log | cut -c1-100; cut -c1-700 $OUT/ttc.json
