#!/bin/bash
# round 2, last session (1 GPU): sizes beyond the tested ones
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tests/large_n_check.py 4194304 > gpurun_out/zz_large_n.log 2>&1
timeout 200 python tests/large_n_check.py 8388608 >> gpurun_out/zz_large_n.log 2>&1
cat gpurun_out/zz_large_n.log | tail -5
