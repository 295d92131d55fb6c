#!/bin/bash
# round 2, session E (2 GPUs): the whole GPU test suite on a 2-GPU box (single-GPU tests + world-2 tests)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 ) > gpurun_out/e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/e_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/e_smoke.log 2>&1
tail -8 gpurun_out/e_pytest.log; tail -2 gpurun_out/e_smoke.log
