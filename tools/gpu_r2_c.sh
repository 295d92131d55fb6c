#!/bin/bash
# round 2, session C (2 GPUs): multi-GPU parity (one process per GPU and single-process handle), benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c_smi.log 2>&1
( time timeout 1500 python -m pytest tests/test_multigpu.py -m gpu -x -q ) > gpurun_out/c_pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest_multi.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c_bench_C5_1gpu.json 2> gpurun_out/c_bench_C5_1gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c_bench_C5_2gpu.json 2> gpurun_out/c_bench_C5_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --config C4 --steps 10 --warmup 3 > gpurun_out/c_bench_C4_2gpu.json 2> gpurun_out/c_bench_C4_2gpu.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_smoke.log 2>&1
tail -5 gpurun_out/c_pytest_multi.log
tail -2 gpurun_out/c_smoke.log
tail -c 600 gpurun_out/c_bench_C5_1gpu.err
