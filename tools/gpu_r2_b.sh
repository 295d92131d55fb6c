#!/bin/bash
# round 2, session B: work-list kernel: tests, shape scan, benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest.log
timeout 600 tools/tune_force 65536 5 super > gpurun_out/b_super_65536.log 2>&1
timeout 600 tools/tune_force 262144 3 super quick > gpurun_out/b_super_262144.log 2>&1
timeout 900 tools/tune_force 1048576 2 super quick > gpurun_out/b_super_1048576.log 2>&1
for c in C5 C4 C3 C2; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/b_bench_$c.json 2> gpurun_out/b_bench_$c.err
done
tail -3 gpurun_out/b_pytest.log
cut -c1-200 gpurun_out/b_super_1048576.log
