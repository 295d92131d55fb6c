#!/bin/bash
# round 2, session Q (1 GPU): compute-sanitizer memcheck + racecheck of the round-2 paths, the 2M-particle test
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LJMD_SHARE_DEVICES=1 CUDA_DEVICE_MAX_CONNECTIONS=32 CUDA_MODULE_LOADING=EAGER
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanity_r2.py > gpurun_out/q_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/q_sanitizer_memcheck.log
unset LJMD_SHARE_DEVICES
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanity_r2.py > gpurun_out/q_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/q_sanitizer_racecheck.log
( time timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "twice_the_benchmark" -p no:cacheprovider ) > gpurun_out/q_pytest_2M.log 2>&1
tail -4 gpurun_out/q_sanitizer_memcheck.log; tail -4 gpurun_out/q_sanitizer_racecheck.log; tail -5 gpurun_out/q_pytest_2M.log
