#!/bin/bash
OUT=gpurun_out/exp9; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.log; tail -4 $OUT/bench.log; cat $OUT/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --kernel-iters 2 --ttc-grid 0 \
  > $OUT/ncu_bench.log 2>&1
for k in bsr_spmm_kernel trsv_tape_kernel; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k \
  --launch-skip 1 -c 1 -o $OUT/$k python tools/kprofile.py --iters 2 spmm precond > $OUT/ncu_$k.log 2>&1
ncu -i $OUT/$k.ncu-rep --page raw --csv > $OUT/$k.raw.csv 2>/dev/null
ncu -i $OUT/$k.ncu-rep --page details > $OUT/$k.details.txt 2>/dev/null
ncu -i $OUT/$k.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/$k.source.csv.gz
rm -f $OUT/$k.ncu-rep
done
ls -la $OUT
