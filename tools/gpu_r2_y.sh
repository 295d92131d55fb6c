#!/bin/bash
# round 2, session Y (1 GPU): k_gather after the batched loads under ncu; C1 and C2 bench lines of the final state
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gather -s 6 -c 1 -o gpurun_out/y_gather_C3 \
    python bench.py --config C3 --steps 6 --warmup 3 --no-cpu-baseline --no-parity --no-sweep > gpurun_out/y_ncu_gather.log 2>&1
timeout 300 python bench.py --config C1 --steps 400 --warmup 20 --no-sweep > gpurun_out/y_bench_C1.json 2> gpurun_out/y_bench_C1.err
timeout 300 python bench.py --config C2 --steps 100 --warmup 10 --no-sweep > gpurun_out/y_bench_C2.json 2> gpurun_out/y_bench_C2.err
timeout 200 python tools/step_latency.py > gpurun_out/y_step_latency.log 2>&1
cut -c1-200 gpurun_out/y_bench_C1.json; cut -c1-200 gpurun_out/y_bench_C2.json; tail -8 gpurun_out/y_step_latency.log
