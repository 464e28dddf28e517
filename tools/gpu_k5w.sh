#!/bin/bash
# K5 slot trace with tapes of 64 / 96 / 128 / 256 warps per agent
OUT=gpurun_out/${1:-k5w}; mkdir -p $OUT
for w in 64 96 128 256; do
  DCORA_TRSV_WARPS=$w timeout 300 python tools/trsv_trace.py > $OUT/trace_w$w.txt 2>&1; echo "== $w warps"; grep total $OUT/trace_w$w.txt; tail -2 $OUT/trace_w$w.txt | cut -c1-200
done
