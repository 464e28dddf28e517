#!/bin/bash
OUT=gpurun_out/${1:-stable}; mkdir -p $OUT
( timeout 300 python -m pytest tests/test_gpu_solver.py -m gpu -q -k "stable_decrease" ) > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log | cut -c1-200
timeout 420 python tools/ttc_multi.py --grid 100 --agents 256 --max-iters 1500 --check-every 20 --nesterov-blocks colours --restart-interval 1000000 --adaptive-restart --stable-decrease --certify-max-iters 1500 > $OUT/ttc.json 2> $OUT/ttc.log
grep -E "sweep [0-9]*00[: ]|lambda_min" $OUT/ttc.log | cut -c1-120; cut -c1-500 $OUT/ttc.json
