#!/bin/bash
# round 2, session L (1 GPU): where does an RDF launch of sorted records spend its time? (ncu, source view)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TUNE_ORDER=hilbert TUNE_RHO=1.1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force_sym -s 8 -c 1 -o gpurun_out/l_rdf_C3 \
    tools/tune_force 65536 1 frames > gpurun_out/l_ncu_rdf_C3.log 2>&1
tail -3 gpurun_out/l_ncu_rdf_C3.log
