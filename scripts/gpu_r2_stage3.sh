#!/bin/bash
# Round 2, stage 3: K1P iteration loop -- parity of the pair kernel, pass costs, cycle counters.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_debug_tile.py -q -x -p no:cacheprovider -k "pair or tiling or debug or config1 or ragged or boundary or randomized" > gpurun_out/r2s3_tests.log 2>&1
echo "tests exit $?" | tee gpurun_out/status3.txt
timeout 600 python scripts/gpu_fuzz_pair.py > gpurun_out/r2s3_fuzz_pair.log 2>&1
echo "fuzz exit $?" | tee -a gpurun_out/status3.txt
timeout 900 python scripts/gpu_pass_costs.py > gpurun_out/r2s3_pass_costs.jsonl 2> gpurun_out/r2s3_pass_costs.err
echo "pass costs exit $?" | tee -a gpurun_out/status3.txt
python -c "
import importlib,sys
sys.path.insert(0,'.')
b=importlib.import_module('multi-modal_colpali_b200.build'); print(b.build_variant('stats',['LIS_K1_STATS']))" > gpurun_out/r2s3_build_stats.log 2>&1
LIS_LIB=$PWD/multi-modal_colpali_b200/_lib/liblis_stats.so timeout 600 python scripts/gpu_pair_stats.py > gpurun_out/r2s3_pair_stats.jsonl 2> gpurun_out/r2s3_pair_stats.err
echo "stats exit $?" | tee -a gpurun_out/status3.txt
tail -5 gpurun_out/r2s3_tests.log
tail -2 gpurun_out/r2s3_fuzz_pair.log
cat gpurun_out/r2s3_pass_costs.jsonl
grep '"pair"' gpurun_out/r2s3_pair_stats.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: d[k] for k in ('rows','mma_loop_per_use','mma_wait_tiles_per_use','mma_wait_acc_per_use','mma_issue_per_use','all_use_hold_w0_w4_per_own_use','all_use_wait_w0_w4_per_own_use','slow_tiles_w0','slow_tiles_w4')})"
