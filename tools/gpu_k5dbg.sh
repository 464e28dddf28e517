#!/bin/bash
# K5 experiments: slot trace under the DCORA_K5_DBG switches (1: no panel gathers, 2: ring depth 4,
# 4: no L2 prefetch, 8: per-iteration cycle stamps of one warp)
OUT=gpurun_out/${1:-k5dbg}; mkdir -p $OUT
for d in ${2:-8}; do
  DCORA_K5_DBG=$d python tools/trsv_trace.py > $OUT/trace_$d.txt 2>&1
  echo "== dbg $d"; grep "total" $OUT/trace_$d.txt
done
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_rbcd.py -q -x > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 600 python tools/kprofile.py --rbcd-steps 6 precond 2>$OUT/k.log >$OUT/k.json; grep rbcd $OUT/k.log | tail -3; cat $OUT/k.json
