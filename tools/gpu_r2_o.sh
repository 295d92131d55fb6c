#!/bin/bash
# round 2, session O (4 GPUs): the default bench as the driver launches it, and the world-4 tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 4 --steps 10 --warmup 3 ) > gpurun_out/o_bench_C5_4gpu.json 2> gpurun_out/o_bench_C5_4gpu.err
( time LJMD_TEST_WORLDS=4 timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -k "not sharing" --maxfail=5 -p no:cacheprovider ) > gpurun_out/o_pytest_w4.log 2>&1
echo "pytest exit $?" >> gpurun_out/o_pytest_w4.log
tail -6 gpurun_out/o_pytest_w4.log; tail -5 gpurun_out/o_bench_C5_4gpu.err; cut -c1-400 gpurun_out/o_bench_C5_4gpu.json
