#!/bin/bash
# K5 with 96 / 128 warps per agent: 16-warp CTAs vs 8-warp CTAs
OUT=gpurun_out/${1:-k5w}; mkdir -p $OUT
for w in 96 128; do
  DCORA_TRSV_WARPS=$w python tools/trsv_trace.py > $OUT/trace_w${w}_nw16.txt 2>&1; echo "== $w warps, 16-warp CTAs"; grep total $OUT/trace_w${w}_nw16.txt
  DCORA_K5_CL8_BIG=1 DCORA_TRSV_WARPS=$w python tools/trsv_trace.py > $OUT/trace_w${w}_nw8.txt 2>&1; echo "== $w warps, 8-warp CTAs"; grep total $OUT/trace_w${w}_nw8.txt; tail -2 $OUT/trace_w${w}_nw8.txt | cut -c1-200
done
