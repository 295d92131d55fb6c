#!/bin/bash
# round 2, session S (1 GPU): ncu --set full of k_gather and of the FRAMES force kernel at C3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gather -s 6 -c 1 -o gpurun_out/s_gather_C3 \
    python bench.py --config C3 --steps 6 --warmup 3 --no-cpu-baseline --no-parity --no-sweep > gpurun_out/s_ncu_gather.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_sym -s 6 -c 1 -o gpurun_out/s_force_C3 \
    python bench.py --config C3 --steps 6 --warmup 3 --no-cpu-baseline --no-parity --no-sweep > gpurun_out/s_ncu_force.log 2>&1
ls -la gpurun_out/*.ncu-rep
