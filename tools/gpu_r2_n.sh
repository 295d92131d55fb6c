#!/bin/bash
# round 2, session N (1 GPU): several ranks sharing one device (LJMD_SHARE_DEVICES=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -k "sharing or argument" --maxfail=3 -p no:cacheprovider ) > gpurun_out/n_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/n_pytest.log
tail -15 gpurun_out/n_pytest.log
