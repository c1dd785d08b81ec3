#!/bin/bash
# A/B several library builds on the pair kernel (5, 6 and 16 query tiles), alternating.
mkdir -p gpurun_out; rm -f gpurun_out/ab3.log
for round in 1 2; do
  for lib in "$@"; do
    echo "== round $round lib $lib" >> gpurun_out/ab3.log
    LIS_LIB=multi-modal_colpali_b200/_lib/$lib timeout 300 python scripts/gpu_ab3_case.py >> gpurun_out/ab3.log 2>&1
  done
done
cat gpurun_out/ab3.log
