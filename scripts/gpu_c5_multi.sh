#!/bin/bash
# BASELINE configs[4] (1024 queries x 32 tokens vs 200 000 pages) on N GPUs
N=${N:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 \
    scripts/gpu_c5_sharded.py > gpurun_out/multi_c5_$N.log 2>&1
echo "c5_sharded exit $?"; grep '^{' gpurun_out/multi_c5_$N.log | cut -c1-500
