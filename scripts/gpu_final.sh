#!/bin/bash
# Final evidence run for a round: GPU tests, smoke, default bench, sweep, launch list + full ncu capture of K1.
mkdir -p gpurun_out; rm -f gpurun_out/sweep.jsonl gpurun_out/status.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/f_pytest_gpu.log 2>&1; echo "pytest_gpu exit $?" | tee -a gpurun_out/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/status.txt
timeout 900 python bench.py > gpurun_out/f_bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/status.txt
timeout 600 python bench.py --impl reference > gpurun_out/f_bench_ref.log 2>&1; echo "bench_ref exit $?" | tee -a gpurun_out/status.txt
timeout 900 python scripts/gpu_sweep.py c2 hbm c5 c3 > gpurun_out/f_sweep.log 2>&1; echo "sweep exit $?" | tee -a gpurun_out/status.txt
TAG=${TAG:-final} bash scripts/gpu_prof.sh | tee -a gpurun_out/status.txt
tail -2 gpurun_out/f_pytest_gpu.log; tail -1 gpurun_out/f_smoke.log; tail -1 gpurun_out/f_bench.log | cut -c1-400; tail -1 gpurun_out/f_bench_ref.log | cut -c1-300
