#!/bin/bash
# Round-2 evidence, final kernel build: GPU suite, smoke, fuzz, default bench, pass costs, cycle counters, ncu launch list and
# full captures of the 5- and 6-tile pair kernel.
mkdir -p gpurun_out; S=gpurun_out/status_final2.txt; rm -f $S
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/g_pytest_gpu.log 2>&1; echo "pytest_gpu exit $?" | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g_smoke.log 2>&1; echo "smoke exit $?" | tee -a $S
timeout 300 python scripts/gpu_fuzz_pair.py > gpurun_out/g_fuzz_pair.log 2>&1; echo "fuzz_pair exit $?" | tee -a $S
timeout 300 python scripts/gpu_fuzz_search.py 40 > gpurun_out/g_fuzz_search.log 2>&1; echo "fuzz_search exit $?" | tee -a $S
timeout 120 python scripts/sanitize_case.py > gpurun_out/g_small_case.log 2>&1; echo "small_case exit $?" | tee -a $S
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench.log 2> gpurun_out/g_bench.err; echo "bench exit $?" | tee -a $S
ROUNDS=2 timeout 400 python scripts/gpu_pass_costs.py > gpurun_out/g_pass_costs.jsonl 2>&1; echo "pass_costs exit $?" | tee -a $S
python -c "
import importlib,sys
sys.path.insert(0,'.')
b=importlib.import_module('multi-modal_colpali_b200.build'); print(b.build_variant('stats',['LIS_K1_STATS']))" > gpurun_out/g_build_stats.log 2>&1
LIS_LIB=$PWD/multi-modal_colpali_b200/_lib/liblis_stats.so timeout 300 python scripts/gpu_pair_stats.py > gpurun_out/g_pair_stats.jsonl 2> gpurun_out/g_pair_stats.err; echo "stats exit $?" | tee -a $S
BENCH_SMALL="python bench.py --steps 3 --warmup 3 --pages 20000 --search-pages 20000 --no-cpu --search-iters 5 --tensor-pages 1000 --host-pages 1000 --ragged-pages 20000"
timeout 300 $BENCH_SMALL > gpurun_out/g_prof_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_r2.csv $BENCH_SMALL > gpurun_out/g_prof_ncu_launches.log 2>&1
echo "ncu_launches exit $?" | tee -a $S
timeout 200 python scripts/gpu_pair_prof_case.py > gpurun_out/g_prof_case_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_pair_kernel -s 2 -c 1 \
    -o gpurun_out/prof_k1pair_r2 -f python scripts/gpu_pair_prof_case.py > gpurun_out/g_prof_ncu_pair.log 2>&1
echo "ncu_full_pair5 exit $?" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_pair_kernel -s 5 -c 1 \
    -o gpurun_out/prof_k1pair6_r2 -f python scripts/gpu_pair_prof_case.py > gpurun_out/g_prof_ncu_pair6.log 2>&1
echo "ncu_full_pair6 exit $?" | tee -a $S
tail -2 gpurun_out/g_pytest_gpu.log; tail -1 gpurun_out/g_smoke.log; tail -1 gpurun_out/g_fuzz_pair.log; tail -1 gpurun_out/g_fuzz_search.log
tail -1 gpurun_out/g_bench.log | cut -c1-2200
