#!/bin/bash
# Iteration loop: GPU tests, sweep, bench.
mkdir -p gpurun_out
rm -f gpurun_out/sweep.jsonl gpurun_out/status.txt
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/s3_pytest_gpu.log 2>&1
echo "pytest_gpu exit $?" | tee -a gpurun_out/status.txt
timeout 1200 python scripts/gpu_sweep.py $SWEEP > gpurun_out/s3_sweep.log 2>&1
echo "sweep exit $?" | tee -a gpurun_out/status.txt
timeout 900 python bench.py --no-cpu > gpurun_out/s3_bench.log 2>&1
echo "bench exit $?" | tee -a gpurun_out/status.txt
tail -3 gpurun_out/s3_pytest_gpu.log; tail -1 gpurun_out/s3_bench.log | cut -c1-600
