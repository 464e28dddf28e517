#!/bin/bash
# One GPU box visit: parity tests, bench, ncu launch list, ncu --set full of the top kernels.
# usage: tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest.log 2>&1
tail -5 $OUT/pytest.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.log
tail -3 $OUT/bench.log; cat $OUT/bench.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --kernel-iters 2 \
  > $OUT/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bsr_spmm_kernel|trsv_tape_kernel' \
  --launch-skip 40 -c 4 -o $OUT/top_kernels python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --kernel-iters 2 \
  > $OUT/ncu_full.log 2>&1
ls -la $OUT
fi
