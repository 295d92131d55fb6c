#!/bin/bash
# round 2, session M (1 GPU): HEAD after the container was re-created -- full GPU suite, smoke, both bench arms timed.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider ) > gpurun_out/m_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/m_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/m_smoke.log 2>&1
( time timeout 600 python bench.py --impl reference ) > gpurun_out/m_bench_reference.json 2> gpurun_out/m_bench_reference.err
( time timeout 900 python bench.py ) > gpurun_out/m_bench_C5.json 2> gpurun_out/m_bench_C5.err
tail -8 gpurun_out/m_pytest.log; tail -5 gpurun_out/m_smoke.log
tail -4 gpurun_out/m_bench_reference.err; tail -4 gpurun_out/m_bench_C5.err
cut -c1-300 gpurun_out/m_bench_C5.json
