#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_round2.py -q -p no:cacheprovider > gpurun_out/r2s4_new_tests.log 2>&1
echo "new tests exit $?" | tee gpurun_out/status4.txt
TAG=r2b bash scripts/gpu_r2_prof_pair.sh
tail -15 gpurun_out/r2s4_new_tests.log
