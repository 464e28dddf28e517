#!/bin/bash
# time-to-certified experiments of the multi-agent staircase: tools/gpu_ttc.sh <tag> "<ttc_multi flags>" ...
OUT=gpurun_out/${1:-ttc}; mkdir -p $OUT
shift 1
i=0
for cfg in "$@"; do
  i=$((i+1))
  timeout ${TTC_TIMEOUT:-300} python tools/ttc_multi.py $cfg > $OUT/ttc_$i.json 2> $OUT/ttc_$i.log
  echo "== $cfg"; grep -E "sweep [0-9]*00[: ]|lambda_min" $OUT/ttc_$i.log | tail -30 | cut -c1-120; cut -c1-600 $OUT/ttc_$i.json
done
