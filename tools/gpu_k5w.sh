#!/bin/bash
# K5 slot trace with 64- and 128-warp tapes + ncu --set full of the 64-warp kernel
OUT=gpurun_out/${1:-k5w}; mkdir -p $OUT
python tools/trsv_trace.py > $OUT/trace64.txt 2>&1; grep total $OUT/trace64.txt
DCORA_TRSV_WARPS=128 python tools/trsv_trace.py > $OUT/trace128.txt 2>&1; grep total $OUT/trace128.txt; tail -3 $OUT/trace128.txt | cut -c1-300
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trsv_tape_kernel \
  -o $OUT/k5 -f python tools/trsv_trace.py > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log; ls -la $OUT
