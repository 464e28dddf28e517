#!/bin/bash
# K5 experiments: slot trace under the DCORA_K5_DBG switches (1: no panel gathers, 2: ring depth 4, 4: no L2 prefetch)
OUT=gpurun_out/${1:-k5dbg}; mkdir -p $OUT
for d in 0 1 2 4 6; do
  DCORA_K5_DBG=$d python tools/trsv_trace.py > $OUT/trace_$d.txt 2>&1
  echo "== dbg $d"; grep "total" $OUT/trace_$d.txt
done
