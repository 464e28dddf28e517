#!/bin/bash
OUT=gpurun_out/exp12; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_solver.py -q > $OUT/pytest.log 2>&1; tail -12 $OUT/pytest.log
