#!/bin/bash
# round 2, session A: GPU test suite + super-tile shape scan + benches (logs under gpurun_out/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/a_smi.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
for n in 65536 262144; do
  timeout 600 tools/tune_force $n 5 super > gpurun_out/a_super_$n.log 2>&1
done
timeout 900 tools/tune_force 1048576 2 super > gpurun_out/a_super_1048576.log 2>&1
for c in C5 C4 C3 C2; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/a_bench_$c.json 2> gpurun_out/a_bench_$c.err
done
tail -3 gpurun_out/a_pytest.log
cat gpurun_out/a_bench_C5.json
