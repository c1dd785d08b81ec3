#!/bin/bash
# Round 2, first GPU contact: new tests, old tests, smoke, micro-benchmark, short bench.  Each stage under its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_round2.py -q -x -p no:cacheprovider > gpurun_out/r2_new_tests.log 2>&1
echo "new tests exit $?" | tee -a gpurun_out/status.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_debug_tile.py -q -p no:cacheprovider > gpurun_out/r2_old_tests.log 2>&1
echo "old tests exit $?" | tee -a gpurun_out/status.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/status.txt
timeout 300 scripts/micro/mma_pair_shapes > gpurun_out/r2_mma_pair_shapes.txt 2>&1
echo "micro exit $?" | tee -a gpurun_out/status.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_short.log 2> gpurun_out/r2_bench_short.err
echo "bench exit $?" | tee -a gpurun_out/status.txt
tail -25 gpurun_out/r2_new_tests.log
tail -8 gpurun_out/r2_old_tests.log
tail -3 gpurun_out/r2_smoke.log
cat gpurun_out/r2_mma_pair_shapes.txt
tail -c 3000 gpurun_out/r2_bench_short.log
tail -5 gpurun_out/r2_bench_short.err
