#!/bin/bash
# round 2, last session (1 GPU): the full GPU suite and smoke() on the final tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/final_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/final_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final_smoke.log 2>&1
tail -6 gpurun_out/final_pytest.log; tail -4 gpurun_out/final_smoke.log
