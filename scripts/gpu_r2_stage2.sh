#!/bin/bash
# Round 2, stage 2: parity of the restructured pair kernel, then pass costs and cycle counters.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_debug_tile.py tests/test_gpu_round2.py -q -p no:cacheprovider > gpurun_out/r2s2_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/status2.txt
timeout 600 python scripts/gpu_fuzz_pair.py > gpurun_out/r2s2_fuzz_pair.log 2>&1
echo "fuzz exit $?" | tee -a gpurun_out/status2.txt
timeout 900 python scripts/gpu_pass_costs.py > gpurun_out/r2s2_pass_costs.jsonl 2> gpurun_out/r2s2_pass_costs.err
echo "pass costs exit $?" | tee -a gpurun_out/status2.txt
python -c "
import importlib,sys
sys.path.insert(0,'.')
b=importlib.import_module('multi-modal_colpali_b200.build'); print(b.build_variant('stats',['LIS_K1_STATS']))" > gpurun_out/r2s2_build_stats.log 2>&1
LIS_LIB=$PWD/multi-modal_colpali_b200/_lib/liblis_stats.so timeout 600 python scripts/gpu_pair_stats.py > gpurun_out/r2s2_pair_stats.jsonl 2> gpurun_out/r2s2_pair_stats.err
echo "stats exit $?" | tee -a gpurun_out/status2.txt
timeout 600 python bench.py --steps 10 --warmup 5 --no-extra --search-pages 100000 > gpurun_out/r2s2_bench.log 2> gpurun_out/r2s2_bench.err
echo "bench exit $?" | tee -a gpurun_out/status2.txt
tail -30 gpurun_out/r2s2_tests.log
tail -5 gpurun_out/r2s2_fuzz_pair.log
cat gpurun_out/r2s2_pass_costs.jsonl
grep '"pair"' gpurun_out/r2s2_pair_stats.jsonl
tail -c 1500 gpurun_out/r2s2_bench.log
