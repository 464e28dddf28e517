#!/bin/bash
OUT=gpurun_out/${1:-tests}; mkdir -p $OUT
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1; grep "smoke ok" $OUT/smoke.log | cut -c1-200
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log | cut -c1-200
