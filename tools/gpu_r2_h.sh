#!/bin/bash
# round 2, session H (1 GPU): sorted records + warp frames in the library: tests, benches, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider ) > gpurun_out/h_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/h_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/h_bench_C5.json 2> gpurun_out/h_bench_C5.err
timeout 600 python bench.py --config C3 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/h_bench_C3.json 2> gpurun_out/h_bench_C3.err
timeout 600 python bench.py --config C2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/h_bench_C2.json 2> gpurun_out/h_bench_C2.err
LJMD_FRAMES=0 timeout 600 python bench.py --config C3 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/h_bench_C3_noframes.json 2> gpurun_out/h_bench_C3_noframes.err
TUNE_ORDER=hilbert TUNE_RHO=1.1 timeout 300 tools/tune_force 65536 5 frames > gpurun_out/h_frames_65536_hilbert.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force_sym -s 4 -c 1 -o gpurun_out/h_sym_C5 \
    python bench.py --config C5 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-sweep > gpurun_out/h_ncu_sym_C5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/h_launches_C3.csv \
    python bench.py --config C3 --steps 12 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/h_ncu_launches.log 2>&1
tail -12 gpurun_out/h_pytest.log; tail -2 gpurun_out/h_smoke.log
for f in gpurun_out/h_bench_*.json; do echo "== $f"; cut -c1-330 $f; done
cut -c7-50,100-135,268-330 gpurun_out/h_frames_65536_hilbert.log
