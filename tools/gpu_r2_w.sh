#!/bin/bash
# round 2, session W (2 GPUs): the world-2 tests (one process per GPU, both transports; the single-process handle;
# the reference driver on two GPUs) and the default bench under torchrun with the batched gather / reduce kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time LJMD_TEST_WORLDS=2 timeout 900 python -m pytest tests/test_multigpu.py tests/test_dropin_gpu.py -m gpu -q -k "not sharing" --maxfail=5 -p no:cacheprovider ) > gpurun_out/w_pytest_w2.log 2>&1
echo "pytest exit $?" >> gpurun_out/w_pytest_w2.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 \
    bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/w_bench_C5_2gpu.json 2> gpurun_out/w_bench_C5_2gpu.err
tail -6 gpurun_out/w_pytest_w2.log; tail -4 gpurun_out/w_bench_C5_2gpu.err; cut -c1-250 gpurun_out/w_bench_C5_2gpu.json
