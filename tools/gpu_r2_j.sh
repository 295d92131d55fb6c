#!/bin/bash
# round 2, session J (2 GPUs): sharded runs with sorted records + warp frames and the fused last barrier
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/j_smi.log 2>&1
( time timeout 1200 python -m pytest tests/test_multigpu.py tests/test_dropin_gpu.py -m gpu -q --maxfail=10 -p no:cacheprovider ) > gpurun_out/j_pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/j_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench_C5_2gpu.json 2> gpurun_out/j_bench_C5_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 2 --config C3 --steps 40 --warmup 5 --no-cpu-baseline --no-sweep > gpurun_out/j_bench_C3_2gpu.json 2> gpurun_out/j_bench_C3_2gpu.err
tail -8 gpurun_out/j_pytest_multi.log
for f in gpurun_out/j_bench_*.json; do echo "== $f"; cut -c1-200 $f; done
