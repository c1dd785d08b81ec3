#!/bin/bash
# ncu full capture (with source correlation) of the 5-tile CTA-pair kernel on a 20 000-page slice.
mkdir -p gpurun_out
TAG=${TAG:-r2a}
timeout 300 python scripts/gpu_pair_prof_case.py > gpurun_out/prof_case_plain.log 2>&1
echo "plain exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:maxsim_pair_kernel -s 2 -c 1 \
    -o gpurun_out/prof_k1pair_$TAG -f python scripts/gpu_pair_prof_case.py > gpurun_out/prof_ncu_full_$TAG.log 2>&1
echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep
