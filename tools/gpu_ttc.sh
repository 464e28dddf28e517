#!/bin/bash
# tests named in $2 + time-to-certified experiments of the multi-agent staircase
OUT=gpurun_out/${1:-ttc}; mkdir -p $OUT
( time timeout 900 python -m pytest $2 -m gpu -q --durations=5 ) > $OUT/pytest.log 2>&1; tail -25 $OUT/pytest.log
shift 2
i=0
for cfg in "$@"; do
  i=$((i+1))
  timeout 420 python tools/ttc_multi.py $cfg > $OUT/ttc_$i.json 2> $OUT/ttc_$i.log
  echo "== $cfg"; grep -E "sweep (5|10|20|40|60|80|100|150|200|300|400|500)[: ]|lambda_min" $OUT/ttc_$i.log | tail -14; cut -c1-400 $OUT/ttc_$i.json
done
