#!/bin/bash
# Round-end evidence on one B200: benches (C1-C5), ncu launch list of the C3 step, ncu --set full of the shipped
# Newton-3 kernel and of k_gather, compute-sanitizer memcheck of the new paths.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
for c in C1 C2 C3; do
  timeout 600 python bench.py --config $c --steps $([ $c = C3 ] && echo 40 || echo 200) --warmup 10 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_C5.json 2> gpurun_out/bench_C5.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_C3.csv \
    python bench.py --config C3 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force_sym -s 3 -c 1 -o gpurun_out/sym_final \
    python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_sym.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_gather -s 3 -c 1 -o gpurun_out/gather_final \
    python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gather.log 2>&1
cat > /tmp/sanity.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from ljpkg import load
pkg = load()
for N, rho, bc in ((700, 0.3, 0), (5000, 0.8, 0), (5000, 0.05, 1)):
    pos, vel = pkg.snapshots.lattice(N, rho, jitter=0.05, seed=5), pkg.snapshots.velocities(N, 1.0, seed=5)
    with pkg.ljmd.LJSystem(N, T0=1.0, rho=rho, canonical=True, bc=bc) as s:
        s.set_state(pos, vel)
        s.trace_begin([(0, 0.05), (3, 0.05), (6, 0.05, 3.0)], 64)
        s.step(0.004, 40, rdf_every=3)
        tr = s.trace_read()
        s.trace_end()
        s.step(0.004, 60, rdf_every=5)          # graph replay + RDF pruning (Newton-3 for N = 5000)
        print(N, bc, tr["counts"].shape, int(s.rdf_counts().sum()), s.scalars()["T"])
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/sanity.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -3 gpurun_out/sanitizer_memcheck.log
cat gpurun_out/bench_C5.json | cut -c1-200; ls -la gpurun_out | head -40
