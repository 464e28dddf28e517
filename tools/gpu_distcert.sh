#!/bin/bash
OUT=gpurun_out/distcert; mkdir -p $OUT
for pc in 0 1; do
CERT_PRECOND=$pc CERT_ITERS=400 timeout 600 python tools/dist_certify_check.py 2> $OUT/n1_pc$pc.log; grep lambda_min $OUT/n1_pc$pc.log
done
CERT_ITERS=400 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/dist_certify_check.py 2> $OUT/n2.log; grep lambda_min $OUT/n2.log || tail -20 $OUT/n2.log
timeout 600 python -m pytest tests/test_gpu_rbcd.py -q -k "distributed_certificate" 2>&1 | tail -2
