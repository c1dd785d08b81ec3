#!/bin/bash
# A/B two library builds on the same GPU, alternating to cancel thermal / power drift.
# usage: gpu_ab.sh <libA> <libB> [sweep cases...]
A=$1; B=$2; shift 2
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
for round in 1 2 3; do
  for lib in $A $B; do
    echo "== round $round lib $lib" >> gpurun_out/ab.log
    LIS_LIB=$lib timeout 600 python scripts/gpu_ab_case.py >> gpurun_out/ab.log 2>&1
  done
done
cat gpurun_out/ab.log
