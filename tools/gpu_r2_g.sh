#!/bin/bash
# round 2, session G: warp frames in the Newton-3 kernel — A/B on sorted, lattice and shuffled particle orders
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "65536 1.1 5" "16384 0.85 10" "262144 0.3 3"; do
  set -- $cfg
  for ord in hilbert lattice random; do
    if [ $ord = lattice ]; then unset TUNE_ORDER; else export TUNE_ORDER=$ord; fi
    TUNE_RHO=$2 timeout 300 tools/tune_force $1 $3 frames > gpurun_out/g_frames_$1_$ord.log 2>&1
  done
done
export TUNE_ORDER=hilbert
TUNE_RHO=0.3 timeout 600 tools/tune_force 1048576 2 frames 16 16 > gpurun_out/g_frames_1048576_hilbert.log 2>&1
for f in gpurun_out/g_frames_*_hilbert.log; do echo "== $f"; cut -c1-60,118-175,230-330 $f; done
( time timeout 1500 python -m pytest tests -m gpu -q -x --maxfail=5 -p no:cacheprovider ) > gpurun_out/g_pytest.log 2>&1
tail -15 gpurun_out/g_pytest.log
