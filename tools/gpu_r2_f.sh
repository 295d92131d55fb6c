#!/bin/bash
# round 2, session F (8 GPUs): parity at world 4 and 8 (both ways to shard), the 8-GPU bench line with its sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/f_smi.log 2>&1
( time LJMD_TEST_WORLDS=4,8 timeout 1500 python -m pytest tests/test_multigpu.py -m gpu -q --maxfail=10 ) > gpurun_out/f_pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/f_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/f_bench_C5_8gpu.json 2> gpurun_out/f_bench_C5_8gpu.err
tail -6 gpurun_out/f_pytest_multi.log
cut -c1-400 gpurun_out/f_bench_C5_8gpu.json
