#!/bin/bash
# Round-2 evidence run on one B200: GPU test suite, smoke, fuzz, default bench + reference arm, fp32 / K3 sweep,
# ncu launch list of a short bench and full captures (source-correlated) of the two K1 forms.  Every stage under a timeout.
mkdir -p gpurun_out; rm -f gpurun_out/sweep.jsonl gpurun_out/status_final.txt
S=gpurun_out/status_final.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/f_pytest_gpu.log 2>&1; echo "pytest_gpu exit $?" | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke exit $?" | tee -a $S
timeout 300 python scripts/gpu_fuzz_pair.py > gpurun_out/f_fuzz_pair.log 2>&1; echo "fuzz_pair exit $?" | tee -a $S
timeout 300 python scripts/gpu_fuzz_search.py 40 > gpurun_out/f_fuzz_search.log 2>&1; echo "fuzz_search exit $?" | tee -a $S
timeout 120 python scripts/sanitize_case.py > gpurun_out/f_small_case.log 2>&1; echo "small_case exit $?" | tee -a $S
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench.log 2> gpurun_out/f_bench.err; echo "bench exit $?" | tee -a $S
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/f_bench_ref.log 2>&1; echo "bench_ref exit $?" | tee -a $S
timeout 400 python scripts/gpu_sweep.py f32 k3 > gpurun_out/f_sweep.log 2>&1; echo "sweep exit $?" | tee -a $S
ROUNDS=2 timeout 400 python scripts/gpu_pass_costs.py > gpurun_out/f_pass_costs.jsonl 2>&1; echo "pass_costs exit $?" | tee -a $S
BENCH_SMALL="python bench.py --steps 3 --warmup 3 --pages 20000 --search-pages 20000 --no-cpu --search-iters 5 --tensor-pages 2000 --host-pages 1000 --ragged-pages 20000"
timeout 300 $BENCH_SMALL > gpurun_out/f_prof_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_r2.csv $BENCH_SMALL > gpurun_out/f_prof_ncu_launches.log 2>&1
echo "ncu_launches exit $?" | tee -a $S
timeout 200 python scripts/gpu_pair_prof_case.py > gpurun_out/f_prof_case_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_pair_kernel -s 2 -c 1 \
    -o gpurun_out/prof_k1pair_r2 -f python scripts/gpu_pair_prof_case.py > gpurun_out/f_prof_ncu_pair.log 2>&1
echo "ncu_full_pair5 exit $?" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_pair_kernel -s 5 -c 1 \
    -o gpurun_out/prof_k1pair6_r2 -f python scripts/gpu_pair_prof_case.py > gpurun_out/f_prof_ncu_pair6.log 2>&1
echo "ncu_full_pair6 exit $?" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxsim_kernel -s 2 -c 1 \
    -o gpurun_out/prof_k1single_r2 -f $BENCH_SMALL > gpurun_out/f_prof_ncu_single.log 2>&1
echo "ncu_full_single exit $?" | tee -a $S
tail -2 gpurun_out/f_pytest_gpu.log; tail -1 gpurun_out/f_smoke.log; tail -1 gpurun_out/f_fuzz_pair.log; tail -1 gpurun_out/f_fuzz_search.log
tail -1 gpurun_out/f_bench.log | cut -c1-1200; tail -1 gpurun_out/f_bench_ref.log | cut -c1-300; cat gpurun_out/f_sweep.log | cut -c1-300
