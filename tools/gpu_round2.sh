#!/bin/bash
# smoke + selected tests (default: the whole GPU suite)
OUT=gpurun_out/${1:-round2}; mkdir -p $OUT
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log | cut -c1-300
( time timeout ${3:-1500} python -m pytest ${2:-tests} -m gpu -q --durations=8 ${4:-} ) > $OUT/pytest.log 2>&1; tail -40 $OUT/pytest.log | cut -c1-250
