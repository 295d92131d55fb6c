#!/bin/bash
# round 2, session I (1 GPU): record-major gather; tests + benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider ) > gpurun_out/i_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/i_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_C5.json 2> gpurun_out/i_bench_C5.err
timeout 600 python bench.py --config C3 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/i_bench_C3.json 2> gpurun_out/i_bench_C3.err
timeout 600 python bench.py --config C2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/i_bench_C2.json 2> gpurun_out/i_bench_C2.err
tail -8 gpurun_out/i_pytest.log
for f in gpurun_out/i_bench_*.json; do echo "== $f"; cut -c1-200 $f; done
