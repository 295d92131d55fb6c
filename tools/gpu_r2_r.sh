#!/bin/bash
# round 2, session R (1 GPU): racecheck after the RDF staging fix; k_gather lanes per particle A/B (LJMD_GATHER_SHIFT)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanity_r2.py > gpurun_out/r_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r_sanitizer_racecheck.log
tail -3 gpurun_out/r_sanitizer_racecheck.log
for cfg in C3 C2; do
  for sh in default 1 2 3; do
    if [ $sh = default ]; then unset LJMD_GATHER_SHIFT; else export LJMD_GATHER_SHIFT=$sh; fi
    timeout 300 python bench.py --config $cfg --steps 40 --warmup 5 --no-cpu-baseline --no-parity --no-sweep > gpurun_out/r_bench_${cfg}_shift_$sh.json 2> gpurun_out/r_bench_${cfg}_shift_$sh.err
    python - <<PY
import json
d=json.load(open("gpurun_out/r_bench_${cfg}_shift_$sh.json"))
print("$cfg shift $sh: step %.4f ms force %.4f gather %.4f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline_hbm"]["kernel_ms"]))
PY
  done
done
unset LJMD_GATHER_SHIFT
TUNE_ORDER=hilbert TUNE_RHO=1.1 timeout 300 tools/tune_force 65536 5 frames > gpurun_out/r_frames_65536_hilbert.log 2>&1
grep -E "RDF|frames on " gpurun_out/r_frames_65536_hilbert.log | cut -c1-160
