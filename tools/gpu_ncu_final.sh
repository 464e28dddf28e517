#!/bin/bash
# round-end ncu evidence: launch list of two bench iterations + ncu --set full of K5 and K1
OUT=gpurun_out/${1:-ncu_final}; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --kernel-iters 1 --ttc-grid 0 \
  > $OUT/ncu_bench.log 2>&1
python tools/launch_summary.py $OUT/launches.csv | head -30
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trsv_tape_kernel \
  -o $OUT/k5 -f python tools/trsv_trace.py > $OUT/ncu_k5.log 2>&1; tail -2 $OUT/ncu_k5.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bsr_spmm_kernel' --launch-skip 30 -c 3 \
  -o $OUT/k1 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --kernel-iters 2 --ttc-grid 0 > $OUT/ncu_k1.log 2>&1; tail -2 $OUT/ncu_k1.log
ls -la $OUT
