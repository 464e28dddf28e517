#!/bin/bash
# K5: 64 warps per agent as 4 CTAs x 16 warps (default) vs 8 CTAs x 8 warps; then ncu --set full of the shipped kernel
OUT=gpurun_out/${1:-k5w}; mkdir -p $OUT
python tools/trsv_trace.py > $OUT/trace_4x16.txt 2>&1; echo "== 4x16"; grep total $OUT/trace_4x16.txt
DCORA_K5_CL8=1 python tools/trsv_trace.py > $OUT/trace_8x8.txt 2>&1; echo "== 8x8"; grep total $OUT/trace_8x8.txt
for v in "" 1; do
  DCORA_K5_CL8=$v timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --ttc-grid 0 > $OUT/bench_cl8_$v.json 2> $OUT/bench_cl8_$v.log; grep -E "value leg|kernels" $OUT/bench_cl8_$v.log | cut -c1-250
done
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trsv_tape_kernel \
  -o $OUT/k5 -f python tools/trsv_trace.py > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
