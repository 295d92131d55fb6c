#!/bin/bash
# round 2, session X (8 GPUs): the default bench as the driver launches it (C5 + the C3/C4 sweep with the step
# breakdown), and the world-8 tests (one process per GPU, both transports, and the single-process handle)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 \
    bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/x_bench_C5_8gpu.json 2> gpurun_out/x_bench_C5_8gpu.err
( time LJMD_TEST_WORLDS=8 timeout 500 python -m pytest tests/test_multigpu.py -m gpu -q -k "not sharing" --maxfail=5 -p no:cacheprovider ) > gpurun_out/x_pytest_w8.log 2>&1
echo "pytest exit $?" >> gpurun_out/x_pytest_w8.log
tail -6 gpurun_out/x_pytest_w8.log; tail -5 gpurun_out/x_bench_C5_8gpu.err; cut -c1-300 gpurun_out/x_bench_C5_8gpu.json
