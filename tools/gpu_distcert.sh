#!/bin/bash
OUT=gpurun_out/distcert; mkdir -p $OUT
timeout 600 python tools/dist_certify_check.py 2> $OUT/n1.log; grep lambda_min $OUT/n1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/dist_certify_check.py 2> $OUT/n2.log; grep lambda_min $OUT/n2.log || tail -20 $OUT/n2.log
