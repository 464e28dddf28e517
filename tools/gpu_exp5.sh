#!/bin/bash
OUT=gpurun_out/exp5; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 600 python tools/kprofile.py --rbcd-steps 4 --split spmm 2>$OUT/k.log >$OUT/k.json; cat $OUT/k.json; grep -v graph $OUT/k.log | tail -16
DCORA_TIMING=1 timeout 600 python tools/kprofile.py --rbcd-steps 2 spmm 2>$OUT/kt.log >/dev/null; grep "dcora timing" $OUT/kt.log | tail -9
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bsr_spmm_kernel' \
  --launch-skip 1 -c 1 -o $OUT/spmm_full python tools/kprofile.py --iters 2 spmm > $OUT/ncu_spmm.log 2>&1
ncu -i $OUT/spmm_full.ncu-rep --page raw --csv > $OUT/spmm_full.raw.csv 2>/dev/null
ncu -i $OUT/spmm_full.ncu-rep --page details > $OUT/spmm_full.details.txt 2>/dev/null
ncu -i $OUT/spmm_full.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/spmm_full.source.csv.gz
rm -f $OUT/spmm_full.ncu-rep
