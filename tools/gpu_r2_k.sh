#!/bin/bash
# round 2, session K (1 GPU): per-chunk RDF pruning in the FRAMES kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TUNE_ORDER=hilbert
TUNE_RHO=1.1 timeout 300 tools/tune_force 65536 5 frames > gpurun_out/k_frames_65536_hilbert.log 2>&1
TUNE_RHO=0.85 timeout 300 tools/tune_force 16384 10 frames > gpurun_out/k_frames_16384_hilbert.log 2>&1
TUNE_RHO=0.3 timeout 300 tools/tune_force 262144 3 frames > gpurun_out/k_frames_262144_hilbert.log 2>&1
unset TUNE_ORDER
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider ) > gpurun_out/k_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/k_pytest.log
timeout 600 python bench.py --config C3 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/k_bench_C3.json 2> gpurun_out/k_bench_C3.err
tail -8 gpurun_out/k_pytest.log
for f in gpurun_out/k_frames_*.log; do echo "== $f"; grep RDF $f | cut -c7-58,120-160,290-360; done
cut -c1-200 gpurun_out/k_bench_C3.json
