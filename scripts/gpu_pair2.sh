#!/bin/bash
# pair kernel iteration: full GPU test suite, cycle stats, A/B
mkdir -p gpurun_out; rm -f gpurun_out/status.txt gpurun_out/pair_ab.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/p_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/status.txt
LIS_LIB=multi-modal_colpali_b200/_lib/liblis_stats.so timeout 300 python scripts/gpu_pair_stats.py > gpurun_out/pair_stats.log 2>&1
echo "stats exit $?" | tee -a gpurun_out/status.txt
timeout 500 python scripts/gpu_pair_ab.py 60000 > gpurun_out/p_ab.log 2>&1
echo "ab exit $?" | tee -a gpurun_out/status.txt
tail -8 gpurun_out/p_pytest.log; cat gpurun_out/pair_stats.log | cut -c1-420; tail -30 gpurun_out/p_ab.log
