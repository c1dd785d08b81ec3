#!/bin/bash
# First contact of the CTA-pair kernel with the GPU: raw accumulator dump, parity, then A/B against the single-CTA form.
mkdir -p gpurun_out; rm -f gpurun_out/status.txt gpurun_out/pair_ab.jsonl
timeout 300 python -m pytest tests/test_gpu_debug_tile.py -q -s -x -p no:cacheprovider -k pair > gpurun_out/p_debug.log 2>&1
echo "pair_debug exit $?" | tee -a gpurun_out/status.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "pair" > gpurun_out/p_parity.log 2>&1
echo "pair_parity exit $?" | tee -a gpurun_out/status.txt
timeout 600 python scripts/gpu_pair_ab.py 60000 > gpurun_out/p_ab.log 2>&1
echo "pair_ab exit $?" | tee -a gpurun_out/status.txt
tail -25 gpurun_out/p_debug.log; tail -25 gpurun_out/p_parity.log; tail -40 gpurun_out/p_ab.log
