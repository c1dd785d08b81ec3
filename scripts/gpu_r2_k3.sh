#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep.jsonl
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -p no:cacheprovider -k "projection or head or ingestion" > gpurun_out/k3_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/k3_tests.log
for occ in 1 2 3 0; do echo "== LIS_K3_OCC=$occ"; LIS_K3_OCC=$occ timeout 200 python scripts/gpu_sweep.py k3 2>&1 | grep '^{' | cut -c1-200; done | tee gpurun_out/k3_occ.log
for occ in 1 2 3; do LIS_K3_OCC=$occ timeout 200 python -m pytest tests/test_gpu_round2.py -q -p no:cacheprovider -k "projection_head_reference or fused_ingestion" > gpurun_out/k3_tests_occ$occ.log 2>&1; echo "tests occ $occ exit $?"; done
