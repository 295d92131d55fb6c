#!/bin/bash
# round 2, session Z (1 GPU): constant-stride reaction walk in k_gather / k_reduce_reaction
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in C3 C4; do
  timeout 300 python bench.py --config $cfg --steps 40 --warmup 5 --no-cpu-baseline --no-sweep > gpurun_out/z_bench_${cfg}.json 2> gpurun_out/z_bench_${cfg}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/z_bench_${cfg}.json"))
print("$cfg: step %.4f ms force %.4f gather %.4f parity %.2e" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_hbm"]["kernel_ms"], d["parity"]["max_err"]))
PY
done
( time timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_multigpu.py tests/test_frames_gpu.py -m gpu -q -x -k "golden or sharing or batched or determin or fuzz or ragged or graph" -p no:cacheprovider ) > gpurun_out/z_pytest.log 2>&1
tail -3 gpurun_out/z_pytest.log
