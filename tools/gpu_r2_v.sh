#!/bin/bash
# round 2, session V (1 GPU): final state -- full GPU suite, smoke, both bench arms as the driver runs them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider ) > gpurun_out/v_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/v_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/v_smoke.log 2>&1
( time timeout 600 python bench.py --impl reference ) > gpurun_out/v_bench_reference.json 2> gpurun_out/v_bench_reference.err
( time timeout 900 python bench.py ) > gpurun_out/v_bench_C5.json 2> gpurun_out/v_bench_C5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v_launches_C3.csv \
    python bench.py --config C3 --steps 12 --warmup 3 --no-cpu-baseline --no-parity --no-sweep > gpurun_out/v_ncu_launches.log 2>&1
tail -6 gpurun_out/v_pytest.log; tail -3 gpurun_out/v_smoke.log
tail -4 gpurun_out/v_bench_reference.err; tail -4 gpurun_out/v_bench_C5.err
cut -c1-300 gpurun_out/v_bench_C5.json
