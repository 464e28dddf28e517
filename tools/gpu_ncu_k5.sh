#!/bin/bash
# ncu --set full capture of the K5 kernel: alone / colour class / whole batch (launches 3, 6, 9 of trsv_trace.py)
OUT=gpurun_out/${1:-ncu_k5}; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trsv_tape_kernel \
  -o $OUT/k5 -f python tools/trsv_trace.py > $OUT/ncu.log 2>&1
tail -5 $OUT/ncu.log
ls -la $OUT
