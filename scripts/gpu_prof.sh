#!/bin/bash
# ncu: launch list + full captures of K1 (the CTA-pair kernel of a bench step, the single-CTA kernel of a search),
# on a short bench (same command plain first).
mkdir -p gpurun_out
TAG=${TAG:-r1}
BENCH_SMALL="python bench.py --steps 3 --warmup 3 --pages 20000 --no-cpu --search-iters 5"
timeout 600 $BENCH_SMALL > gpurun_out/prof_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv $BENCH_SMALL > gpurun_out/prof_ncu_launches.log 2>&1
echo "ncu_launches exit $?"
timeout 600 $BENCH_SMALL > gpurun_out/prof_plain2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:maxsim_pair_kernel -s 4 -c 1 \
    -o gpurun_out/prof_k1pair_$TAG -f $BENCH_SMALL > gpurun_out/prof_ncu_full.log 2>&1
echo "ncu_full_pair exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:maxsim_kernel -s 2 -c 1 \
    -o gpurun_out/prof_k1single_$TAG -f $BENCH_SMALL > gpurun_out/prof_ncu_full2.log 2>&1
echo "ncu_full_single exit $?"
