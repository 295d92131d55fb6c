#!/bin/bash
# round 2, session D (1 GPU): full GPU test suite, smoke, planner scans, ncu evidence, benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 ) > gpurun_out/d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/d_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d_smoke.log 2>&1
timeout 300 tools/tune_force 262144 5 shardscan 8 > gpurun_out/d_shardscan_262144_8.log 2>&1
timeout 300 tools/tune_force 262144 5 shardscan 4 > gpurun_out/d_shardscan_262144_4.log 2>&1
timeout 600 tools/tune_force 1048576 2 shardscan 8 > gpurun_out/d_shardscan_1048576_8.log 2>&1
for n in 8192 16384 32768 65536 131072; do
  echo "== N=$n" >> gpurun_out/d_split_scan.log
  timeout 600 tools/tune_force $n 5 scan >> gpurun_out/d_split_scan.log 2>&1
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/d_bench_C5.json 2> gpurun_out/d_bench_C5.err
timeout 600 python bench.py --config C2 --steps 50 --warmup 5 > gpurun_out/d_bench_C2.json 2> gpurun_out/d_bench_C2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/d_bench_reference.json 2> gpurun_out/d_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/d_launches_C3.csv \
    python bench.py --config C3 --steps 12 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/d_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force_sym -s 3 -c 1 -o gpurun_out/d_sym_C3 \
    python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/d_ncu_sym_C3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_force_sym -s 4 -c 1 -o gpurun_out/d_sym_C5 \
    python bench.py --config C5 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-sweep > gpurun_out/d_ncu_sym_C5.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_gather -s 4 -c 1 -o gpurun_out/d_gather_C5 \
    python bench.py --config C5 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-sweep > gpurun_out/d_ncu_gather_C5.log 2>&1
tail -6 gpurun_out/d_pytest.log; tail -2 gpurun_out/d_smoke.log; ls -la gpurun_out | grep " d_"
