#!/bin/bash
# Wall time of the fluctuation driver: device-resident CLI (tasks/bin) vs the reference's driver source on the
# MDSystem class (oracle/_ref, one upload + download per step).  Usage: tests/task_timing.sh [N] [rho] [steps]
N=${1:-400}; RHO=${2:-0.05}; STEPS=${3:-2000}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d)
TFIN=$(python3 -c "print(2.0 + 0.004*$STEPS + 0.002)")
printf "N $N\nT* 1.4\nrho* $RHO\nteq 2.\ntfin $TFIN\ndt* 0.004\ncanonical 1\nsubvolume_spacing 0.05\nuseCUDA 1\n" > $W/in
cd $W
# warm the page cache / driver first: whichever program runs first on a fresh box pays ~1.5 s of cold start
LJMD_SEED=1 $ROOT/lennard-jones-cuda_b200/tasks/bin/semiGCEfluctuations 1 64 > /dev/null 2>&1
for exe in $ROOT/lennard-jones-cuda_b200/tasks/bin/run-fluctuations $ROOT/oracle/_ref/run-fluctuations; do
  [ -x $exe ] || continue
  s=$(date +%s.%N)
  LJMD_SEED=1 $exe $W/in > $W/out.txt 2>&1
  e=$(date +%s.%N)
  echo "N=$N rho*=$RHO steps=500+$STEPS $(echo $exe | sed "s|$ROOT/||"): $(python3 -c "print(f'{$e-$s:.2f}')") s; last row: $(grep -E '^ +[0-9]' $W/out.txt | tail -1)"
done
rm -rf $W
