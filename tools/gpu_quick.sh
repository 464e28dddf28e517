#!/bin/bash
# quick GPU check: selected tests   usage: tools/gpu_quick.sh <pytest args>
OUT=gpurun_out/quick; mkdir -p $OUT
timeout 900 python -m pytest "$@" -q > $OUT/pytest.log 2>&1; tail -15 $OUT/pytest.log
