#!/bin/bash
# BASELINE configs[2] (1 M ragged pages, top-100) on N GPUs
N=${N:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
    scripts/gpu_c3_sharded.py > gpurun_out/multi_c3_$N.log 2>&1
echo "c3_sharded exit $?"; grep '^{' gpurun_out/multi_c3_$N.log | cut -c1-420
