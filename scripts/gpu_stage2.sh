#!/bin/bash
# Full GPU test pass, default bench, tiling sweep, ncu launch list + one full capture of K1.
mkdir -p gpurun_out
rm -f gpurun_out/sweep.jsonl gpurun_out/status.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/s2_pytest_gpu.log 2>&1
echo "pytest_gpu exit $?" | tee -a gpurun_out/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2_smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/status.txt
timeout 900 python bench.py > gpurun_out/s2_bench.log 2>&1
echo "bench exit $?" | tee -a gpurun_out/status.txt
timeout 900 python scripts/gpu_sweep.py > gpurun_out/s2_sweep.log 2>&1
echo "sweep exit $?" | tee -a gpurun_out/status.txt
BENCH_SMALL="python bench.py --steps 3 --warmup 3 --pages 20000 --no-cpu --search-iters 5"
timeout 600 $BENCH_SMALL > gpurun_out/s2_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r1.csv $BENCH_SMALL > gpurun_out/s2_ncu_launches.log 2>&1
echo "ncu_launches exit $?" | tee -a gpurun_out/status.txt
timeout 600 $BENCH_SMALL > gpurun_out/s2_plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:maxsim_kernel -s 4 -c 1 \
    -o gpurun_out/prof_k1_r1 -f $BENCH_SMALL > gpurun_out/s2_ncu_full.log 2>&1
echo "ncu_full exit $?" | tee -a gpurun_out/status.txt
tail -3 gpurun_out/s2_pytest_gpu.log; tail -2 gpurun_out/s2_smoke.log; tail -1 gpurun_out/s2_bench.log; cat gpurun_out/s2_sweep.log | cut -c1-400
