#!/bin/bash
# Regression pass after a change: the whole GPU suite, smoke, the two fuzz scripts, the default bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c_pytest_gpu.log 2>&1; echo "pytest_gpu exit $?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_smoke.log 2>&1; echo "smoke exit $?"
timeout 300 python scripts/gpu_fuzz_search.py 30 > gpurun_out/c_fuzz_search.log 2>&1; echo "fuzz_search exit $?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c_bench.log 2> gpurun_out/c_bench.err; echo "bench exit $?"
tail -3 gpurun_out/c_pytest_gpu.log; tail -1 gpurun_out/c_smoke.log; tail -1 gpurun_out/c_fuzz_search.log; tail -1 gpurun_out/c_bench.log | cut -c1-5000; tail -3 gpurun_out/c_bench.err
