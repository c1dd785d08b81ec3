#!/bin/bash
# quick K1P iteration: pair parity subset + pass costs (one round) + stats for 640 rows
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_debug_tile.py -q -x -p no:cacheprovider -k "pair or debug or boundary or randomized" > gpurun_out/q_tests.log 2>&1
echo "tests exit $?"
timeout 600 python scripts/gpu_fuzz_pair.py > gpurun_out/q_fuzz.log 2>&1
echo "fuzz exit $?"
if [ -z "$NOCOST" ]; then ROUNDS=1 timeout 600 python scripts/gpu_pass_costs.py > gpurun_out/q_pass_costs.jsonl 2> gpurun_out/q_pass_costs.err; fi
python -c "
import importlib,sys
sys.path.insert(0,'.')
b=importlib.import_module('multi-modal_colpali_b200.build'); print(b.build_variant('stats',['LIS_K1_STATS']))" > gpurun_out/q_build_stats.log 2>&1
LIS_LIB=$PWD/multi-modal_colpali_b200/_lib/liblis_stats.so timeout 600 python scripts/gpu_pair_stats.py > gpurun_out/q_pair_stats.jsonl 2> gpurun_out/q_pair_stats.err
tail -2 gpurun_out/q_tests.log; tail -1 gpurun_out/q_fuzz.log
grep pair gpurun_out/q_pass_costs.jsonl
grep '"pair"' gpurun_out/q_pair_stats.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: d[k] for k in ('rows','mma_loop_per_use','mma_wait_acc_per_use','all_use_hold_w0_w4_per_own_use','all_use_wait_w0_w4_per_own_use','slow_tiles_w0','spe_per_slow_tile_w0','spe_per_slow_tile_w4')})"
