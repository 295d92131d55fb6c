#!/bin/bash
# One gpurun session: correctness first, then tuning, bench and profiles.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt
STAGE=${1:-all}
if [[ $STAGE == all || $STAGE == test ]]; then
  timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  tail -5 gpurun_out/pytest_gpu.log
fi
if [[ $STAGE == all || $STAGE == tune ]]; then
  timeout 400 tools/tune_force 65536 5 > gpurun_out/tune_65536.log 2>&1
  timeout 200 tools/tune_force 16384 5 > gpurun_out/tune_16384.log 2>&1
  cat gpurun_out/tune_65536.log
fi
if [[ $STAGE == all || $STAGE == bench ]]; then
  timeout 600 python bench.py --config C3 --steps 20 --warmup 3 > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err
  timeout 600 python bench.py --config C2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_C2.json 2> gpurun_out/bench_C2.err
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C5.json 2> gpurun_out/bench_C5.err
  cat gpurun_out/bench_C3.json gpurun_out/bench_C2.json gpurun_out/bench_C5.json; for f in gpurun_out/bench_*.err; do tail -n 3 $f; done
fi
if [[ $STAGE == all || $STAGE == prof ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_C3.csv \
      python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 1 -c 2 -o gpurun_out/force_prof \
      tools/tune_force 65536 1 > gpurun_out/ncu_force.log 2>&1
  ls -la gpurun_out
fi
