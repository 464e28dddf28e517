#!/bin/bash
OUT=gpurun_out/exp2; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bsr_spmm_kernel' \
  --launch-skip 1 -c 1 -o $OUT/spmm_full python tools/kprofile.py --iters 2 spmm > $OUT/ncu_spmm.log 2>&1
tail -2 $OUT/ncu_spmm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'trsv_tape_kernel' \
  --launch-skip 1 -c 1 -o $OUT/trsv_full python tools/kprofile.py --iters 2 precond > $OUT/ncu_trsv.log 2>&1
tail -2 $OUT/ncu_trsv.log
for k in spmm_full trsv_full; do
  ncu -i $OUT/$k.ncu-rep --page raw --csv > $OUT/$k.raw.csv 2>/dev/null
  ncu -i $OUT/$k.ncu-rep --page source --csv > $OUT/$k.source.csv 2>/dev/null
  ncu -i $OUT/$k.ncu-rep --page details > $OUT/$k.details.txt 2>/dev/null
  sz=$(stat -c %s $OUT/$k.ncu-rep); if [ $sz -gt 20000000 ]; then rm $OUT/$k.ncu-rep; fi
  gzip -f $OUT/$k.source.csv
done
ls -la $OUT
