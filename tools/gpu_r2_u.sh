#!/bin/bash
# round 2, session U (1 GPU): k_gather / k_reduce_reaction with batched loads and the two-run reaction walk
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in C3 C2 C4; do
  timeout 300 python bench.py --config $cfg --steps 40 --warmup 5 --no-cpu-baseline --no-sweep > gpurun_out/u_bench_${cfg}.json 2> gpurun_out/u_bench_${cfg}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/u_bench_${cfg}.json"))
print("$cfg: step %.4f ms force %.4f gather %.4f parity %.2e" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_hbm"]["kernel_ms"], d["parity"]["max_err"]))
PY
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sweep > gpurun_out/u_bench_C5.json 2> gpurun_out/u_bench_C5.err
python - <<PY
import json
d=json.load(open("gpurun_out/u_bench_C5.json"))
print("C5: step %.4f ms force %.4f gather %.4f parity %.2e" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_hbm"]["kernel_ms"], d["parity"]["max_err"]))
PY
( time timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_multigpu.py tests/test_frames_gpu.py -m gpu -q -x -k "golden or sharing or batched or determin or fuzz or ragged or graph" -p no:cacheprovider ) > gpurun_out/u_pytest.log 2>&1
tail -5 gpurun_out/u_pytest.log
