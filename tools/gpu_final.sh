#!/bin/bash
# Round-end evidence run: parity tests, smoke, bench, ncu launch list, ncu --set full of K1/K5, CUPTI timeline
TAG=${1:-r01_final}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.log; tail -4 $OUT/bench.log; cut -c1-300 $OUT/bench.json
timeout 900 python tools/timeline.py --steps 2 > $OUT/timeline.txt 2> $OUT/timeline.log; head -3 $OUT/timeline.txt
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
python tools/trsv_trace.py > $OUT/trsv_trace.txt 2>&1
ls -la $OUT
