#!/bin/bash
# First contact with the GPU: raw MMA path, then parity, then a short bench.  Each stage in its own
# process under `timeout` so that a trap in one does not take the others down.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_debug_tile.py -q -s -p no:cacheprovider > gpurun_out/s1_debug_tile.log 2>&1
echo "debug_tile exit $?" | tee -a gpurun_out/status.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider > gpurun_out/s1_parity.log 2>&1
echo "parity exit $?" | tee -a gpurun_out/status.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s1_bench.log 2>&1
echo "bench exit $?" | tee -a gpurun_out/status.txt
tail -5 gpurun_out/s1_debug_tile.log
tail -30 gpurun_out/s1_parity.log
tail -5 gpurun_out/s1_bench.log
