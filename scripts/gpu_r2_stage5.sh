#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep.jsonl
timeout 300 python scripts/gpu_load_bw.py > gpurun_out/load_bw.jsonl 2> gpurun_out/load_bw.err; echo "load bw exit $?"; cat gpurun_out/load_bw.jsonl; tail -3 gpurun_out/load_bw.err
for occ in 1 2 0; do echo "== LIS_K3_OCC=$occ"; LIS_K3_OCC=$occ timeout 200 python scripts/gpu_sweep.py k3 2>&1 | grep '^{' | cut -c1-200; done | tee gpurun_out/k3_occ.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -p no:cacheprovider -k "projection or head or ingestion or fp32 or roundtrip or sharded_directory" > gpurun_out/s5_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/s5_tests.log
